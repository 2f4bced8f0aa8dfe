#!/bin/bash
# Round artefacts: both bench arms, the ncu launch list of the bench command and one full ncu capture of the rasterizer's kernels.
R=${1:-r02}
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
timeout 600 python bench.py > gpurun_out/${R}_bench_ours.json 2> gpurun_out/${R}_bench_ours.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 8 --warmup 3 --no-denoise --no-cpu-baseline > gpurun_out/${R}_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/${R}_launches.csv 30 > gpurun_out/${R}_launches_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'preprocess_kernel|compact_kernel|sort_pass_kernel|bin_|render_(forward|backward)_kernel|gaussian_backward' --launch-skip 150 -c 28 \
    -o gpurun_out/${R}_raster_full -f python bench.py --steps 8 --warmup 3 --no-denoise --no-cpu-baseline > gpurun_out/${R}_ncu_full.log 2>&1
python tools/ncu_traffic.py gpurun_out/${R}_raster_full.ncu-rep C2 > gpurun_out/${R}_ncu_traffic.json 2> gpurun_out/${R}_ncu_traffic.err
ncu -i gpurun_out/${R}_raster_full.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_active.avg','sm__cycles_elapsed.max','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct']
idx=[hdr.index(w) for w in want if w in hdr]
w=csv.writer(sys.stdout)
for r in rows: w.writerow([r[i][:100] for i in idx])
" > gpurun_out/${R}_ncu_raster_summary.csv
python -c "
import json
for f in ('${R}_bench_reference','${R}_bench_ours'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], (d.get('denoise') or {}).get('value'), (d.get('guided') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e, open('gpurun_out/%s.err'%f).read()[-600:])
"
cat gpurun_out/${R}_launches_summary.txt | head -20
