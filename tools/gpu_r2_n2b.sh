#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N ) > gpurun_out/r2l_bench_ours_n$N.json 2> gpurun_out/r2l_bench_ours_n$N.err
( time GVD_UNET_GRAPH=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --no-c5 --no-guided --steps 50 ) > gpurun_out/r2l_bench_nograph_n$N.json 2> gpurun_out/r2l_bench_nograph_n$N.err
python -c "
import json,sys
N=sys.argv[1]
for f in ('r2l_bench_ours_n'+N,'r2l_bench_nograph_n'+N):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('exchange'))
        print('  check', d.get('exchange_check'))
        print('  denoise', {k:v for k,v in (d.get('denoise') or {}).items() if k in ('value','ms_per_step','tflops_per_s','error')})
        print('  guided', {k:v for k,v in (d.get('guided') or {}).items() if k in ('value','ms_per_step','tflops_per_s','error','peak_mem_gb')})
        print('  c5', {k:v for k,v in (d.get('c5') or {}).items() if k in ('value','ms_per_step','error')})
    except Exception as e: print(f, 'ERR', e)
    print(open('gpurun_out/%s.err'%f).read()[-1200:])
" $N
