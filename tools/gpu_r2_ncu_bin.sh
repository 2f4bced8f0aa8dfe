#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'sort_pass_kernel|compact_kernel|bin_fill_kernel' --launch-skip 60 -c 6 \
    -o gpurun_out/r2b_binning_full -f python bench.py --steps 5 --warmup 3 --no-denoise --no-cpu-baseline > gpurun_out/r2b_ncu_full.log 2>&1
tail -3 gpurun_out/r2b_ncu_full.log; ls -la gpurun_out/*.ncu-rep
