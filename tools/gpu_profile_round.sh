#!/bin/bash
# Round artefacts: both bench arms, the ncu launch list of the bench command and one full ncu capture of the render kernels.
mkdir -p gpurun_out
timeout 400 python bench.py --impl reference --steps 100 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 400 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-denoise --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'render_(forward|backward)_kernel' --launch-skip 20 -c 2 \
    -o gpurun_out/render_full -f python bench.py --steps 5 --warmup 3 --no-denoise --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
    --clock-control none -k regex:flash_attn -c 3 --csv --log-file gpurun_out/ncu_flash.csv python tools/bench_attn.py > gpurun_out/ncu_flash.log 2>&1
timeout 120 python tools/bench_attn.py > gpurun_out/bench_attn.log 2>&1
timeout 300 python tools/profile_unet.py 25 72 128 ours > gpurun_out/unet_profile.txt 2>&1
python -c "
import json
for f in ('bench_reference','bench_ours'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], d.get('denoise',{}).get('value'))
    except Exception as e: print(f, 'ERR', e, open('gpurun_out/%s.err'%f).read()[-600:])
"
