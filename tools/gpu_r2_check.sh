#!/bin/bash
# Round 2 check of the restructured rasterizer (compaction + own radix sort + mask-ranked fill + exact early-R sizing),
# the reference-Python drop-in tests and the guided path after the transposed-weight fix.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_raster_gpu.py tests/test_render_dropin_gpu.py tests/test_knn_gpu.py -q -rA -p no:cacheprovider ) > gpurun_out/r2_raster_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_raster_pytest.log
grep -E "^(PASSED|FAILED|ERROR)|passed|failed|exit" gpurun_out/r2_raster_pytest.log | tail -40
( timeout 300 python -m pytest tests/test_zz_guided_gpu.py tests/test_zz_nn_fast_gpu.py -q -p no:cacheprovider -k "not groupnorm_bwd" ) > gpurun_out/r2_guided_pytest.log 2>&1
tail -5 gpurun_out/r2_guided_pytest.log; grep -E "rel L2|VAE decoder" gpurun_out/r2_guided_pytest.log
timeout 300 python bench.py --steps 200 --warmup 10 --no-denoise --no-cpu-baseline > gpurun_out/r2_bench_exact.json 2> gpurun_out/r2_bench_exact.err
GVD_SPECULATE=defer timeout 300 python bench.py --steps 200 --warmup 10 --no-denoise --no-cpu-baseline > gpurun_out/r2_bench_defer.json 2> gpurun_out/r2_bench_defer.err
python -c "
import json
for f in ('r2_bench_exact','r2_bench_defer'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], d.get('roofline',{}).get('stage_ms'))
    except Exception as e: print(f, 'ERR', e, open('gpurun_out/%s.err'%f).read()[-1500:])
"
timeout 400 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_memcheck_smoke.log 2>&1
grep -E "ERROR SUMMARY|smoke ok" gpurun_out/r2_memcheck_smoke.log | tail -3
timeout 500 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_racecheck_smoke.log 2>&1
grep -E "RACECHECK SUMMARY|smoke ok" gpurun_out/r2_racecheck_smoke.log | tail -3
