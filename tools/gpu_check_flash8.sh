#!/bin/bash
# generation 8 flash attention: parity in its own process, then v7 vs v8 timing on the C3 self-attention shapes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nn_ops_gpu.py -q -x -k "variants" 2>&1 | tail -5
for v in v7 v8; do GVD_FLASH=$v timeout 300 python tools/bench_attn.py 2>&1 | tee -a gpurun_out/flash8_bench.txt; done
