#!/bin/bash
# the whole GPU suite, as the driver runs it at round end
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -q -m gpu -p no:cacheprovider ) > gpurun_out/full_suite.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|real" gpurun_out/full_suite.log | tail -15
