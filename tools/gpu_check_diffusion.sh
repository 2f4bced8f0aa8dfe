#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2z}
( timeout 1500 python -m pytest tests/test_nn_ops_gpu.py tests/test_zz_guided_gpu.py tests/test_unet_gpu.py tests/test_zz_nn_fast_gpu.py -q -p no:cacheprovider ) > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "passed|failed|Error|^E " gpurun_out/${TAG}_pytest.log | tail -8
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python -c "import json; d=json.loads(open(\"gpurun_out/${TAG}_bench.json\").read().strip().splitlines()[-1]); print(d[\"value\"], d[\"e2e\"][\"value\"], (d.get(\"denoise\") or {}).get(\"value\"), (d.get(\"guided\") or {}).get(\"value\"))"
timeout 600 python tools/profile_guided.py guided > gpurun_out/${TAG}_guided_profile.txt 2>&1; grep -E "# guided|gn_|elementwise" gpurun_out/${TAG}_guided_profile.txt | cut -c1-100 | head -12
