#!/bin/bash
# fused attention adjoint: parity (operator level + autograd route) in both forms, then timing against the materialised backward
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_guided_gpu.py -q -x -k "flash_attention or attention_bwd" 2>&1 | tail -4
GVD_FLASH_BWD_CTAS=1 timeout 900 python -m pytest tests/test_zz_guided_gpu.py -q -x -k "flash_attention" 2>&1 | tail -2
rm -f gpurun_out/flash_bwd_bench.txt
for f in 2 1; do echo "# GVD_FLASH_BWD_CTAS=$f" | tee -a gpurun_out/flash_bwd_bench.txt; GVD_FLASH_BWD_CTAS=$f timeout 300 python tools/bench_attn_bwd.py 2>&1 | tee -a gpurun_out/flash_bwd_bench.txt; done
