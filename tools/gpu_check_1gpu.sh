#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python -c "
import json
for f in ('bench_ours',):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], d.get('roofline',{}).get('stage_ms'), d.get('denoise',{}).get('value'))
    except Exception as e: print(f, 'ERR', e, open('gpurun_out/%s.err'%f).read()[-600:])
"
