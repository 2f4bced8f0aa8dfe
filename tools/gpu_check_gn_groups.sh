#!/bin/bash
# frame-grouped GroupNorm launches (L2-resident second pass): parity, then graph-timed A/B over the group budget
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nn_ops_gpu.py tests/test_zz_guided_gpu.py -q -x -p no:cacheprovider -k "groupnorm" 2>&1 | tail -3
rm -f gpurun_out/gn_groups.txt
for mb in 0 24 48 72; do echo "# GVD_GN_GROUP_MB=$mb" | tee -a gpurun_out/gn_groups.txt; GVD_GN_GROUP_MB=$mb timeout 300 python tools/bench_norm_bwd.py gn 2>&1 | tee -a gpurun_out/gn_groups.txt; done
