#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_raster_gpu.py tests/test_render_dropin_gpu.py -q -p no:cacheprovider ) > gpurun_out/r2d_raster_pytest.log 2>&1
tail -4 gpurun_out/r2d_raster_pytest.log
true
tail -3 gpurun_out/r2d_rcp_pytest.log; grep -n "AssertionError: (" gpurun_out/r2d_rcp_pytest.log gpurun_out/r2d_raster_pytest.log | head
for v in 0; do
GVD_BWD_EXACT_DIV=$v timeout 300 python bench.py --steps 200 --warmup 10 --no-denoise --no-cpu-baseline > gpurun_out/r2d_bench_div$v.json 2> gpurun_out/r2d_bench_div$v.err
done
python -c "
import json
for f in ('r2d_bench_div0',):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], d.get('roofline',{}).get('stage_ms'))
    except Exception as e: print(f, 'ERR', e, open('gpurun_out/%s.err'%f).read()[-1500:])
"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2d_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-denoise --no-cpu-baseline > gpurun_out/r2d_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r2d_launches.csv | head -16
