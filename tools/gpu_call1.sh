#!/bin/bash
# One-GPU round check: GPU parity tests, smoke, both bench arms, launch list, one full ncu capture of the render kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=20 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
( time timeout 400 python bench.py --impl reference --steps 100 --warmup 5 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
( time timeout 400 python bench.py --steps 200 --warmup 10 ) > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
timeout 120 python bench.py --workload small --steps 300 --warmup 10 --no-denoise --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-denoise --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'render_(forward|backward)_kernel' --launch-skip 20 -c 2 \
    -o gpurun_out/render_full -f python bench.py --steps 5 --warmup 3 --no-denoise --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/bench_ours.json | head -c 3000
