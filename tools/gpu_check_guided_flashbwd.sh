#!/bin/bash
# guided step with the fused attention adjoint: parity suite, then the step with GVD_FLASH_BWD = 0 / 1, then its kernel breakdown
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_zz_guided_gpu.py -q -p no:cacheprovider -x ) > gpurun_out/r4a_pytest.log 2>&1
grep -E "passed|failed|Error|^E |Warning: vc_b200|capture" gpurun_out/r4a_pytest.log | tail -8
for f in 0 1; do GVD_FLASH_BWD=$f timeout 600 python tools/bench_guided.py --arm ours --steps 4 2>&1 | tail -3 | cut -c1-300 | sed "s/^/flash_bwd=$f /"; done
timeout 600 python tools/profile_guided.py guided host > gpurun_out/r4a_guided_kernel_breakdown.txt 2>&1
head -16 gpurun_out/r4a_guided_kernel_breakdown.txt | cut -c1-130
