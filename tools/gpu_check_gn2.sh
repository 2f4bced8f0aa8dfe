#!/bin/bash
# GroupNorm forward / backward with the cp.async ring: parity (operator + model level), then the micro-benchmark
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_nn_ops_gpu.py tests/test_zz_guided_gpu.py tests/test_unet_gpu.py tests/test_zz_nn_fast_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -6
timeout 300 python tools/bench_norm_bwd.py gn 2>&1 | tee gpurun_out/norm_bwd_bench_after.txt
