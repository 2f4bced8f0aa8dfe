#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
( timeout 400 python -m pytest tests/test_exchange_gpu.py -q -x -p no:cacheprovider ) > gpurun_out/r2g_exchange_pytest_n$N.log 2>&1; tail -3 gpurun_out/r2g_exchange_pytest_n$N.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N ) > gpurun_out/r2g_bench_ours_n$N.json 2> gpurun_out/r2g_bench_ours_n$N.err
( time GVD_EXCHANGE=peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --no-denoise --no-c5 ) > gpurun_out/r2g_bench_peer_n$N.json 2> gpurun_out/r2g_bench_peer_n$N.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --impl reference --no-denoise ) > gpurun_out/r2g_bench_ref_n$N.json 2> gpurun_out/r2g_bench_ref_n$N.err
python -c "
import json,sys
N=sys.argv[1]
for f in ('r2g_bench_ours_n'+N,'r2g_bench_peer_n'+N,'r2g_bench_ref_n'+N):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('exchange'))
        print('  check', d.get('exchange_check'))
        print('  denoise', {k:v for k,v in (d.get('denoise') or {}).items() if k in ('value','ms_per_step','tflops_per_s','error')})
        print('  guided', {k:v for k,v in (d.get('guided') or {}).items() if k in ('value','ms_per_step','tflops_per_s','error','peak_mem_gb')})
        print('  c5', {k:v for k,v in (d.get('c5') or {}).items() if k in ('value','ms_per_step','error','config')})
    except Exception as e: print(f, 'ERR', e)
    print(open('gpurun_out/%s.err'%f).read()[-900:])
" $N
