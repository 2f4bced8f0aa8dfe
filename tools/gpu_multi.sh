#!/bin/bash
# N GPUs of one box (gpurun --gpus N -- ./tools/gpu_multi.sh N [ref]): the cross-GPU gradient-sum tests, our bench arm
# (exchange_check, denoise, guided, c5 blocks included) and, with "ref", the reference arm (+ NCCL) beside it.
mkdir -p gpurun_out
N=${1:-2}
TAG=r02_n$N
( timeout 400 python -m pytest tests/test_exchange_gpu.py -q -x -p no:cacheprovider ) > gpurun_out/${TAG}_exchange_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_exchange_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N ) > gpurun_out/${TAG}_bench_ours.json 2> gpurun_out/${TAG}_bench_ours.err
if [ "$2" = "ref" ]; then
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --impl reference --no-denoise ) > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
fi
python -c "
import json,sys,os
tag=sys.argv[1]
for f in (tag+'_bench_ours', tag+'_bench_ref'):
    if not os.path.exists('gpurun_out/%s.json'%f): continue
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('exchange'))
        print('  check', d.get('exchange_check'))
        print('  denoise', {k:v for k,v in (d.get('denoise') or {}).items() if k in ('value','ms_per_step','tflops_per_s','error')})
        print('  guided', {k:v for k,v in (d.get('guided') or {}).items() if k in ('value','ms_per_step','tflops_per_s','error','peak_mem_gb')})
        print('  c5', {k:v for k,v in (d.get('c5') or {}).items() if k in ('value','ms_per_step','error')})
    except Exception as e: print(f, 'ERR', e)
    print(open('gpurun_out/%s.err'%f).read()[-700:])
" $TAG
