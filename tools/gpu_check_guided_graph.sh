#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_zz_guided_gpu.py -q -p no:cacheprovider -x ) > gpurun_out/r3n_pytest.log 2>&1
grep -E "passed|failed|Error|^E |Warning: vc_b200|capture" gpurun_out/r3n_pytest.log | tail -8
for g in 0 1; do GVD_GUIDED_GRAPH=$g timeout 600 python tools/bench_guided.py --arm ours --steps 4 2>&1 | tail -3 | cut -c1-300 | sed "s/^/graph=$g /"; done
