#!/bin/bash
# Final bench lines of the round (both arms) and the U-Net / guided kernel breakdowns; the ncu raster captures of
# tools/gpu_round_final.sh are not repeated when the rasterizer has not changed since they were taken.
R=${1:-r02g}
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
timeout 900 python bench.py > gpurun_out/${R}_bench_ours.json 2> gpurun_out/${R}_bench_ours.err
timeout 600 python tools/profile_guided.py unet > gpurun_out/${R}_unet_kernel_breakdown.txt 2>&1
timeout 600 python tools/profile_guided.py guided host > gpurun_out/${R}_guided_kernel_breakdown.txt 2>&1
python -c "
import json
for f in ('${R}_bench_reference','${R}_bench_ours'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], (d.get('denoise') or {}).get('value'), (d.get('guided') or {}).get('value'), (d.get('train_step') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e, open('gpurun_out/%s.err'%f).read()[-600:])
"
grep -A12 "^# unet" gpurun_out/${R}_unet_kernel_breakdown.txt | cut -c1-130
