#!/bin/bash
# implicit-GEMM convolution: parity first, then the kernel breakdown of one C3 forward with and without it
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_nn_ops_gpu.py tests/test_gemm_gpu.py -q -p no:cacheprovider -k "conv or gemm" ) > gpurun_out/r2p_conv_pytest.log 2>&1
grep -E "passed|failed|Error|^E " gpurun_out/r2p_conv_pytest.log | tail -12
GVD_IMPLICIT_CONV=0 timeout 600 python tools/profile_unet.py 25 72 128 > gpurun_out/r2p_unet_profile_im2col.txt 2>&1
timeout 600 python tools/profile_unet.py 25 72 128 > gpurun_out/r2p_unet_profile_implicit.txt 2>&1
grep -E "Self CUDA time total|gemm_bf16|im2col" gpurun_out/r2p_unet_profile_im2col.txt | cut -c1-75,150-230
grep -E "Self CUDA time total|gemm_bf16|im2col" gpurun_out/r2p_unet_profile_implicit.txt | cut -c1-75,150-230
