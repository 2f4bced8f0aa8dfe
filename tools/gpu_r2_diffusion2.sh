#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_unet_gpu.py -q -s -p no:cacheprovider -k "graph" ) > gpurun_out/r2i_unet_pytest.log 2>&1
grep -E "passed|failed|Error" gpurun_out/r2i_unet_pytest.log | tail -5
timeout 600 python tools/profile_copies.py > gpurun_out/r2i_copies.txt 2>&1; tail -62 gpurun_out/r2i_copies.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-guided 2>gpurun_out/r2i_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('denoise with graph', {k:v for k,v in d.get('denoise',{}).items() if k in ('value','ms_per_step','error')})"
tail -3 gpurun_out/r2i_bench.err
