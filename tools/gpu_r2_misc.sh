#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_zz_guided_gpu.py tests/test_render_dropin_gpu.py tests/test_raster_gpu.py -q -p no:cacheprovider ) > gpurun_out/r2k_pytest.log 2>&1
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2k_pytest.log | tail -12
bash tools/gpu_profile_round.sh r02
