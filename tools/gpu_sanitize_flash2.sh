#!/bin/bash
# memcheck / racecheck of the two-CTA form of the fused attention adjoint (the default) and the forward that writes its statistic;
# the racecheck pass runs three times: its timing once exposed a parity-aliased wait in the forward's epilogue (fixed)
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
( timeout 600 $CS --tool memcheck --launch-timeout 0 python -m pytest tests/test_zz_guided_gpu.py -q -x -p no:cacheprovider -k "flash_attention" ) > gpurun_out/flash2_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/flash2_memcheck.log | tail -2
for i in 1 2 3; do
( timeout 600 $CS --tool racecheck python -m pytest tests/test_zz_guided_gpu.py tests/test_nn_ops_gpu.py -q -x -p no:cacheprovider -k "(flash_attention and not 2560) or attention_paths" ) > gpurun_out/flash2_racecheck_$i.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/flash2_racecheck_$i.log | tail -2
done
