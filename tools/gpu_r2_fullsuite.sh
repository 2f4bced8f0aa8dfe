#!/bin/bash
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests/ -q -m gpu -p no:cacheprovider -x ) > gpurun_out/r3f_full_pytest.log 2>&1
tail -5 gpurun_out/r3f_full_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
