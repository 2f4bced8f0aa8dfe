#!/bin/bash
# after the GroupNorm / column split / upsample-route changes: parity of the touched operators and models, guided step timing + breakdown
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_nn_ops_gpu.py tests/test_gemm_gpu.py tests/test_zz_guided_gpu.py tests/test_vae_gpu.py -q -p no:cacheprovider -x ) > gpurun_out/r4b_pytest.log 2>&1
grep -E "passed|failed|Error|^E " gpurun_out/r4b_pytest.log | tail -8
timeout 600 python tools/bench_guided.py --arm ours --steps 4 2>&1 | tail -2 | cut -c1-260
timeout 600 python tools/profile_guided.py guided host > gpurun_out/r4b_guided_kernel_breakdown.txt 2>&1
grep -A28 "^# guided" gpurun_out/r4b_guided_kernel_breakdown.txt | cut -c1-130
