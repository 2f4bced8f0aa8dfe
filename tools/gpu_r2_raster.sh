#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2w}
( timeout 1200 python -m pytest tests/test_raster_gpu.py tests/test_render_dropin_gpu.py -q -p no:cacheprovider -x ) > gpurun_out/${TAG}_raster_pytest.log 2>&1
grep -E "passed|failed|Error|^E " gpurun_out/${TAG}_raster_pytest.log | tail -8
for mom in 0 1; do
GVD_BWD_MOMENTS=$mom timeout 600 python bench.py --no-cpu-baseline --no-denoise > gpurun_out/${TAG}_bench_mom${mom}.json 2> gpurun_out/${TAG}_bench_mom${mom}.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_mom${mom}.json').read().strip().splitlines()[-1])
print('moments=${mom}', d['value'], d['e2e']['value'], d['roofline']['stage_ms'])
"
done
timeout 300 python tools/profile_host.py > gpurun_out/${TAG}_host_profile.txt 2>&1; head -45 gpurun_out/${TAG}_host_profile.txt
