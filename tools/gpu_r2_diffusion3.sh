#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_zz_guided_gpu.py tests/test_render_dropin_gpu.py -q -p no:cacheprovider -k "mn_major or attention_bwd or unet_input_gradient or guided_step or vae_decoder or trajectory or outgrows" ) > gpurun_out/r2j_pytest.log 2>&1
grep -E "passed|failed|Error|^E " gpurun_out/r2j_pytest.log | tail -12
timeout 600 python tools/profile_guided.py guided > gpurun_out/r2j_guided_profile.txt 2>&1; head -14 gpurun_out/r2j_guided_profile.txt
