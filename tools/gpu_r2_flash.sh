#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_nn_ops_gpu.py -q -p no:cacheprovider ) > gpurun_out/r2s_pytest.log 2>&1
grep -E "passed|failed|Error|^E " gpurun_out/r2s_pytest.log | tail -8
GVD_FLASH=v2 timeout 300 python tools/bench_attn.py 2>&1 | tail -3
GVD_FLASH=v3 timeout 300 python tools/bench_attn.py 2>&1 | tail -3
timeout 600 python tools/profile_unet.py 25 72 128 > gpurun_out/r2s_unet_profile.txt 2>&1
grep -E "Self CUDA time total|flash_attn|gn_apply|gn_partial|layernorm" gpurun_out/r2s_unet_profile.txt | cut -c1-75,150-230
