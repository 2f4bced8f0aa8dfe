#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2x}
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['stage_ms'])
print('  denoise', {k:v for k,v in (d.get('denoise') or {}).items() if k in ('value','ms_per_step','tflops_per_s','error')})
print('  guided', {k:v for k,v in (d.get('guided') or {}).items() if k in ('value','ms_per_step','tflops_per_s','error','peak_mem_gb')})
"
tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python tools/profile_host.py > gpurun_out/${TAG}_host_profile.txt 2>&1; head -24 gpurun_out/${TAG}_host_profile.txt
