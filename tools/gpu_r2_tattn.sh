#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_nn_ops_gpu.py tests/test_zz_nn_fast_gpu.py -q -p no:cacheprovider ) > gpurun_out/r2q_pytest.log 2>&1
grep -E "passed|failed|Error|^E " gpurun_out/r2q_pytest.log | tail -12
GVD_TATTN_MMA=0 timeout 600 python tools/profile_unet.py 25 72 128 2>&1 | grep -E "Self CUDA time total|temporal_attn" | cut -c1-75,150-230
timeout 600 python tools/profile_unet.py 25 72 128 > gpurun_out/r2q_unet_profile.txt 2>&1
grep -E "Self CUDA time total|temporal_attn" gpurun_out/r2q_unet_profile.txt | cut -c1-75,150-230
