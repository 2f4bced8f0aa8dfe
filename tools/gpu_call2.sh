#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
( GVD_FLASH_V1=1 timeout 120 python tools/bench_attn.py; GVD_FLASH_POLY=0 timeout 120 python tools/bench_attn.py; timeout 120 python tools/bench_attn.py ) > gpurun_out/bench_attn.log 2>&1
cat gpurun_out/bench_attn.log
timeout 400 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
timeout 120 python bench.py --workload small --steps 300 --warmup 10 --no-denoise --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
GVD_SPECULATE=sync timeout 120 python bench.py --steps 200 --warmup 10 --no-denoise --no-cpu-baseline > gpurun_out/bench_sync.json 2> gpurun_out/bench_sync.err
timeout 200 python -m cProfile -o gpurun_out/prof.out bench.py --workload small --steps 1000 --warmup 10 --no-denoise --no-cpu-baseline > /dev/null 2>&1
python -c "
import pstats; p=pstats.Stats('gpurun_out/prof.out'); p.sort_stats('tottime').print_stats(45)" > gpurun_out/prof.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
  --clock-control none -k regex:flash_attn -c 4 --csv --log-file gpurun_out/ncu_flash.csv python tools/bench_attn.py > gpurun_out/ncu_flash.log 2>&1
python -c "
import json
for f in ('bench_ours','bench_small','bench_sync'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], d.get('roofline',{}).get('stage_ms'), d.get('denoise',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
"
