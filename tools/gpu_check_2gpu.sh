#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_exchange_gpu.py -x -q ) > gpurun_out/pytest_n2.log 2>&1
tail -2 gpurun_out/pytest_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tools/probe_exchange.py > gpurun_out/probe_exchange.json 2> gpurun_out/probe_exchange.err
tail -c 1200 gpurun_out/probe_exchange.json
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --no-denoise > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print('bench_n2', d['value'], d['ms_per_step'], d['e2e']['value'])
"
