#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/r2f_bench_ours.json 2> gpurun_out/r2f_bench_ours.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
python -c "
import json
for f in ('r2f_bench_ours','r2f_bench_ref'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['e2e']['value'], (d.get('roofline') or {}).get('stage_ms'))
        print('  denoise', {k:v for k,v in (d.get('denoise') or {}).items() if k in ('value','ms_per_step','tflops_per_s','error','cpu_baseline')})
        print('  guided', {k:v for k,v in (d.get('guided') or {}).items() if k in ('value','ms_per_step','tflops_per_s','error','peak_mem_gb')})
        print('  cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
    print(open('gpurun_out/%s.err'%f).read()[-700:])
"
