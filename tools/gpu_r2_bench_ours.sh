#!/bin/bash
# ours arm only: headline + denoise + guided blocks, plus the guided parity tests that cover the implicit dgrad
mkdir -p gpurun_out
TAG=${1:-r2r}
( timeout 900 python -m pytest tests/test_zz_guided_gpu.py tests/test_unet_gpu.py -q -p no:cacheprovider -x ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|Error|^E " gpurun_out/${TAG}_pytest.log | tail -8
( time timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/${TAG}_bench_ours.json 2> gpurun_out/${TAG}_bench_ours.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_ours.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], (d.get('roofline') or {}).get('stage_ms'))
print('  denoise', {k:v for k,v in (d.get('denoise') or {}).items() if k in ('value','ms_per_step','tflops_per_s','error')})
print('  guided', {k:v for k,v in (d.get('guided') or {}).items() if k in ('value','ms_per_step','tflops_per_s','error','peak_mem_gb')})
"
tail -5 gpurun_out/${TAG}_bench_ours.err
