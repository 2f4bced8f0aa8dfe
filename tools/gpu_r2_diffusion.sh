#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_unet_gpu.py -q -s -p no:cacheprovider ) > gpurun_out/r2h_unet_pytest.log 2>&1
grep -E "rel L2|C3 \[|C4 latent|chained DDIM|passed|failed|Error" gpurun_out/r2h_unet_pytest.log | tail -20
timeout 600 python tools/profile_guided.py guided > gpurun_out/r2h_guided_profile.txt 2>&1; head -45 gpurun_out/r2h_guided_profile.txt
timeout 600 python tools/profile_guided.py unet > gpurun_out/r2h_unet_profile.txt 2>&1; head -30 gpurun_out/r2h_unet_profile.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-guided 2>gpurun_out/r2h_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('denoise with graph', d.get('denoise'))"
GVD_UNET_GRAPH=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-guided 2>>gpurun_out/r2h_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('denoise eager', {k:v for k,v in d.get('denoise',{}).items() if k in ('value','ms_per_step')})"
tail -5 gpurun_out/r2h_bench.err
