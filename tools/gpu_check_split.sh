#!/bin/bash
# column split of the N = 320 / 640 GEMMs: parity (GEMM + convolution tests), then the A/B sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_nn_ops_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -4
for m in 0 1; do GVD_GEMM_SPLIT=$m timeout 300 python tools/bench_gemm_split.py 2>&1 | tee -a gpurun_out/gemm_split_sweep.txt; done
