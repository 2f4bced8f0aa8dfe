#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_render_dropin_gpu.py tests/test_raster_gpu.py -q -p no:cacheprovider -x ) > gpurun_out/r3m_pytest.log 2>&1
grep -E "passed|failed|Error|^E " gpurun_out/r3m_pytest.log | tail -12
