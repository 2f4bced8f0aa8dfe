#!/bin/bash
# fused attention adjoint: parity (operator level + autograd route), then timing against the materialised backward
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_guided_gpu.py -q -x -k "flash_attention or attention_bwd" 2>&1 | tail -15
timeout 300 python tools/bench_attn_bwd.py 2>&1 | tee gpurun_out/flash_bwd_bench.txt
