#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 40 --warmup 3 > gpurun_out/r3p_bench_ref.json 2> gpurun_out/r3p_bench_ref.err
python -c "
import json
d=json.loads(open('gpurun_out/r3p_bench_ref.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'])
print('  denoise', {k:v for k,v in (d.get('denoise') or {}).items() if k in ('value','ms_per_step','error')})
print('  guided', {k:v for k,v in (d.get('guided') or {}).items() if k in ('value','ms_per_step','error','peak_mem_gb')})
print('  train', {k:v for k,v in (d.get('train_step') or {}).items() if k in ('value','ms_per_iteration','error')})
"
tail -3 gpurun_out/r3p_bench_ref.err | cut -c1-200
