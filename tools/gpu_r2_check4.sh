#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_raster_gpu.py tests/test_render_dropin_gpu.py -q -p no:cacheprovider ) > gpurun_out/r2e_raster_pytest.log 2>&1
tail -4 gpurun_out/r2e_raster_pytest.log; grep -n "AssertionError: (" gpurun_out/r2e_raster_pytest.log | head
for m in sync exact defer; do
GVD_SPECULATE=$m timeout 300 python bench.py --steps 300 --warmup 10 --no-denoise --no-cpu-baseline > gpurun_out/r2e_bench_$m.json 2> gpurun_out/r2e_bench_$m.err
done
python -c "
import json
for f in ('r2e_bench_sync','r2e_bench_exact','r2e_bench_defer'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], d.get('roofline',{}).get('stage_ms'))
    except Exception as e: print(f, 'ERR', e, open('gpurun_out/%s.err'%f).read()[-1500:])
"
nproc; lscpu | grep -E "Model name|MHz" | head -3
