#!/bin/bash
mkdir -p gpurun_out
M=${1:-230400}; N=${2:-320}; K=${3:-1280}
for mode in 0 1; do
GVD_GEMM_PAIR=$mode timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -f -o gpurun_out/r2u_gemm_pair${mode} python tools/one_gemm.py $M $N $K > gpurun_out/r2u_ncu_pair${mode}.log 2>&1
ncu -i gpurun_out/r2u_gemm_pair${mode}.ncu-rep --page raw --csv > gpurun_out/r2u_gemm_pair${mode}_raw.csv 2>/dev/null
done
ls -la gpurun_out/r2u*
