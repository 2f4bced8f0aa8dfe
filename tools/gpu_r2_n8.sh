#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N ) > gpurun_out/r2m_bench_ours_n$N.json 2> gpurun_out/r2m_bench_ours_n$N.err
( time GVD_EXCHANGE=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus $N --no-denoise --no-c5 --steps 100 ) > gpurun_out/r2m_bench_peer_n$N.json 2> gpurun_out/r2m_bench_peer_n$N.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29633 tools/probe_exchange.py ) > gpurun_out/r2m_probe_n$N.json 2> gpurun_out/r2m_probe_n$N.err
( time GVD_EXCHANGE=peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29634 tools/probe_exchange.py ) > gpurun_out/r2m_probe_peer_n$N.json 2> gpurun_out/r2m_probe_peer_n$N.err
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29635 bench.py --gpus $N --impl reference --no-denoise --no-c5 --steps 60 ) > gpurun_out/r2m_bench_ref_n$N.json 2> gpurun_out/r2m_bench_ref_n$N.err
python -c "
import json,sys
N=sys.argv[1]
for f in ('r2m_bench_ours_n'+N,'r2m_bench_peer_n'+N,'r2m_bench_ref_n'+N):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('exchange'))
        print('  check', d.get('exchange_check'))
        print('  denoise', {k:v for k,v in (d.get('denoise') or {}).items() if k in ('value','ms_per_step','tflops_per_s','error')})
        print('  guided', {k:v for k,v in (d.get('guided') or {}).items() if k in ('value','ms_per_step','tflops_per_s','error','peak_mem_gb')})
        print('  c5', {k:v for k,v in (d.get('c5') or {}).items() if k in ('value','ms_per_step','error')})
    except Exception as e: print(f, 'ERR', e)
    print(open('gpurun_out/%s.err'%f).read()[-600:])
for f in ('r2m_probe_n'+N, 'r2m_probe_peer_n'+N):
    print(f, open('gpurun_out/%s.json'%f).read()[-1500:]); print(open('gpurun_out/%s.err'%f).read()[-400:])
" $N
