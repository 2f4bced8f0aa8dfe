#!/bin/bash
# First hardware run of the guided-sampler path (csrc/nn_backward.cu, vc_b200.grad / guided / vae): parity, then timing.
# usage (from the repo root, one GPU):  gpurun --timeout 1500 -- 'bash tools/gpu_check_guided.sh'
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_zz_guided_gpu.py -q -rxXs --runxfail -p no:cacheprovider ) > gpurun_out/guided_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/guided_pytest.log
tail -15 gpurun_out/guided_pytest.log
# memory-checked run of the smallest cases (catches out-of-bounds accesses the parity asserts may not see)
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_zz_guided_gpu.py -q --runxfail -k "groupnorm_bwd or temporal_attention_bwd or conv3x3_dx or layernorm" > gpurun_out/guided_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/guided_memcheck.log | tail -3
( timeout 300 python -m pytest tests/test_zz_pcd2img_gpu.py -q -rxXs --runxfail -p no:cacheprovider ) > gpurun_out/pcd2img_pytest.log 2>&1; tail -3 gpurun_out/pcd2img_pytest.log
( timeout 300 python -m pytest tests/test_zz_train_ops_gpu.py -q -rxXs --runxfail -p no:cacheprovider ) > gpurun_out/train_ops_pytest.log 2>&1; tail -3 gpurun_out/train_ops_pytest.log
timeout 900 python tools/bench_guided.py --arm both > gpurun_out/guided_bench.jsonl 2> gpurun_out/guided_bench.err
cat gpurun_out/guided_bench.jsonl; tail -3 gpurun_out/guided_bench.err
timeout 600 python tools/bench_train_step.py --arm both --iters 100 > gpurun_out/train_step_bench.jsonl 2> gpurun_out/train_step_bench.err
cat gpurun_out/train_step_bench.jsonl; tail -2 gpurun_out/train_step_bench.err
timeout 300 python tools/bench_nn_fast.py > gpurun_out/nn_fast_bench.json 2> gpurun_out/nn_fast_bench.err; cat gpurun_out/nn_fast_bench.json; ( timeout 300 python -m pytest tests/test_zz_nn_fast_gpu.py -q -rxXs --runxfail -p no:cacheprovider ) 2>&1 | tail -2
