#!/bin/bash
# temporal attention backward on mma tiles: parity; GroupNorm / temporal attention backward timings; ncu of the GroupNorm backward pair
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_guided_gpu.py -q -x -k "temporal_attention_bwd or groupnorm_bwd" 2>&1 | tail -5
timeout 300 python tools/bench_norm_bwd.py 2>&1 | tee gpurun_out/norm_bwd_bench.txt
GVD_TATTN_MMA=0 timeout 300 python tools/bench_norm_bwd.py tattn 2>&1 | tee -a gpurun_out/norm_bwd_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_bwd -s 4 -c 2 -f -o gpurun_out/gn_bwd python tools/bench_norm_bwd.py gn > gpurun_out/gn_bwd_ncu.log 2>&1
ncu -i gpurun_out/gn_bwd.ncu-rep --page raw --csv > gpurun_out/gn_bwd_raw.csv 2>/dev/null
ncu -i gpurun_out/gn_bwd.ncu-rep --page source --csv > gpurun_out/gn_bwd_source.csv 2>/dev/null
ls -la gpurun_out/gn_bwd*
