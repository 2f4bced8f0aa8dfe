#!/bin/bash
# four-GPU check: exchange kernel at world 4, frame-sharded denoiser (cfg2 x frames2), bench at N=4
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
( timeout 300 python -m pytest tests/test_exchange_gpu.py -x -q ) > gpurun_out/pytest_n4.log 2>&1
tail -2 gpurun_out/pytest_n4.log
timeout 300 $TR --master-port 29533 tools/run_frame_parallel.py --h 40 --w 64 --reps 1 > gpurun_out/frame_parallel_n4.json 2> gpurun_out/frame_parallel_n4.err
tail -c 700 gpurun_out/frame_parallel_n4.json; tail -c 400 gpurun_out/frame_parallel_n4.err
timeout 600 $TR --master-port 29511 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
python -c "
import json
try:
    d=json.loads(open('gpurun_out/bench_n4.json').read().strip().splitlines()[-1]); print('bench_n4', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('denoise'))
except Exception as e: print('ERR', e, open('gpurun_out/bench_n4.err').read()[-1500:])
"
