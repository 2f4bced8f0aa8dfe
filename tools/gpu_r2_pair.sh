#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gemm_gpu.py -q -p no:cacheprovider -x ) > gpurun_out/r2t_gemm_pytest.log 2>&1
grep -E "passed|failed|Error|^E " gpurun_out/r2t_gemm_pytest.log | tail -8
( timeout 300 python -m pytest tests/test_nn_ops_gpu.py -q -p no:cacheprovider -x ) > gpurun_out/r2t_ops_pytest.log 2>&1
grep -E "passed|failed|Error|^E " gpurun_out/r2t_ops_pytest.log | tail -8
GVD_GEMM_PAIR=0 timeout 300 python tools/bench_gemm.py 2>&1 | tail -12
timeout 300 python tools/bench_gemm.py 2>&1 | tail -12
timeout 600 python tools/profile_unet.py 25 72 128 > gpurun_out/r2t_unet_profile.txt 2>&1
grep -E "Self CUDA time total|gemm_bf16" gpurun_out/r2t_unet_profile.txt | cut -c1-75,150-230
