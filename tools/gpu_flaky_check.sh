#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4; do timeout 300 python -m pytest tests/test_render_dropin_gpu.py -q -x -p no:cacheprovider -k "outgrows" 2>&1 | tail -1; done
./tools/gpu_full_suite.sh
