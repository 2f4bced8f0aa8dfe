#!/bin/bash
mkdir -p gpurun_out
for v in ${1:-v7}; do
GVD_FLASH=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 1 -c 1 -f -o gpurun_out/r3k_attn_$v python tools/one_attn.py > gpurun_out/r3k_ncu_$v.log 2>&1
ncu -i gpurun_out/r3k_attn_$v.ncu-rep --page raw --csv > gpurun_out/r3k_attn_${v}_raw.csv 2>/dev/null
ncu -i gpurun_out/r3k_attn_$v.ncu-rep --page source --csv > gpurun_out/r3k_attn_${v}_source.csv 2>/dev/null
done
