#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_raster_gpu.py tests/test_render_dropin_gpu.py -q -x -p no:cacheprovider ) > gpurun_out/r2n_raster_pytest.log 2>&1
tail -5 gpurun_out/r2n_raster_pytest.log
timeout 300 python bench.py --steps 200 --warmup 10 --no-denoise --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
python -c "
import json
for f in ('r2n_bench',):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], d.get('roofline',{}).get('stage_ms'))
    except Exception as e: print(f, 'ERR', e, open('gpurun_out/%s.err'%f).read()[-1500:])
"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2n_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-denoise --no-cpu-baseline > gpurun_out/r2n_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r2n_launches.csv | head -30
