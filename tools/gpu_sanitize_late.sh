#!/bin/bash
# compute-sanitizer over the kernels added at the end of round 2: fused attention adjoint (tcgen05 / TMEM), mma temporal attention
# adjoint, GroupNorm forward / backward with the cp.async ring, upsample route, column-split GEMM / convolution.
mkdir -p gpurun_out
TAG=${1:-r02late}
CS=/usr/local/cuda/bin/compute-sanitizer
( timeout 1500 $CS --tool memcheck --launch-timeout 0 python -m pytest tests/test_zz_guided_gpu.py -q -x -p no:cacheprovider -k "flash_attention or temporal_attention_bwd or groupnorm_bwd or conv3x3_dx" ) > gpurun_out/${TAG}_memcheck_bwd.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_memcheck_bwd.log | tail -3
( timeout 1500 $CS --tool memcheck --launch-timeout 0 python -m pytest tests/test_nn_ops_gpu.py -q -x -p no:cacheprovider -k "upsample or groupnorm or conv3x3_implicit" ) > gpurun_out/${TAG}_memcheck_fwd.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_memcheck_fwd.log | tail -3
( timeout 1200 $CS --tool racecheck python -m pytest tests/test_zz_guided_gpu.py -q -x -p no:cacheprovider -k "groupnorm_bwd or temporal_attention_bwd or (flash_attention_bwd_fused and not 2560)" ) > gpurun_out/${TAG}_racecheck_bwd.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${TAG}_racecheck_bwd.log | tail -3
( timeout 900 $CS --tool racecheck python -m pytest tests/test_nn_ops_gpu.py -q -x -p no:cacheprovider -k "groupnorm" ) > gpurun_out/${TAG}_racecheck_fwd.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${TAG}_racecheck_fwd.log | tail -3
