#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_bwd2?_kernel -s 2 -c 2 -f -o gpurun_out/attn_bwd python tools/one_attn_bwd.py > gpurun_out/attn_bwd_ncu.log 2>&1
ncu -i gpurun_out/attn_bwd.ncu-rep --page raw --csv > gpurun_out/attn_bwd_raw.csv 2>/dev/null
ncu -i gpurun_out/attn_bwd.ncu-rep --page source --csv > gpurun_out/attn_bwd_source.csv 2>/dev/null
ls -la gpurun_out/attn_bwd*
