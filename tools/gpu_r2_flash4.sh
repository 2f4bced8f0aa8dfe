#!/bin/bash
mkdir -p gpurun_out
( GVD_FLASH=v7 timeout 300 python -m pytest tests/test_nn_ops_gpu.py -q -p no:cacheprovider -k "attention" ) > gpurun_out/r3j_pytest.log 2>&1
grep -E "passed|failed|Error|^E " gpurun_out/r3j_pytest.log | tail -8
GVD_FLASH=v2 timeout 200 python tools/bench_attn.py 2>&1 | tail -3
GVD_FLASH=v7 timeout 200 python tools/bench_attn.py 2>&1 | tail -3
