#!/bin/bash
# compute-sanitizer over the round's new kernels at small sizes: memcheck (out-of-bounds, misaligned) and racecheck
# (shared-memory hazards) of the implicit-GEMM convolution, the CTA-pair GEMM, the fused GEGLU epilogue, flash attention
# generation 7 (and the selectable ones), the mma temporal attention, and the rasterizer chain (smoke).
mkdir -p gpurun_out
TAG=${1:-r02}
CS=/usr/local/cuda/bin/compute-sanitizer
K="conv3x3_implicit or conv_t3_implicit or temporal_attention or attention_paths or fused_geglu or linear_shapes"
( timeout 1500 $CS --tool memcheck --launch-timeout 0 python -m pytest tests/test_nn_ops_gpu.py tests/test_gemm_gpu.py -q -x -p no:cacheprovider -k "$K" ) > gpurun_out/${TAG}_memcheck_nn.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_memcheck_nn.log | tail -3
( timeout 900 $CS --tool racecheck python -m pytest tests/test_nn_ops_gpu.py -q -x -p no:cacheprovider -k "temporal_attention or attention_paths" ) > gpurun_out/${TAG}_racecheck_nn.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${TAG}_racecheck_nn.log | tail -3
( timeout 600 $CS --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_memcheck_smoke.log 2>&1
grep -E "ERROR SUMMARY|smoke ok" gpurun_out/${TAG}_memcheck_smoke.log | tail -2
( timeout 600 $CS --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_racecheck_smoke.log 2>&1
grep -E "RACECHECK SUMMARY|smoke ok" gpurun_out/${TAG}_racecheck_smoke.log | tail -2
