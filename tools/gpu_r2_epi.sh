#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_nn_ops_gpu.py -q -p no:cacheprovider -x ) > gpurun_out/r3h_pytest.log 2>&1
grep -E "passed|failed|Error|^E " gpurun_out/r3h_pytest.log | tail -6
timeout 300 python tools/bench_gemm_pair.py 2>&1 | tail -20
timeout 600 python tools/profile_unet.py 25 72 128 > gpurun_out/r3h_unet_profile.txt 2>&1
grep -E "Self CUDA time total|gemm_bf16|flash" gpurun_out/r3h_unet_profile.txt | cut -c1-75,150-230
