#!/usr/bin/env python
"""profiles/r02_ncu_traffic.json from an `ncu --set full` report: dram__bytes_read.sum + dram__bytes_write.sum per launch of
the rasterizer's kernels (mean over the captured launches), keyed by bench.py's stage names.  bench.py reads this file for
`roofline.traffic` instead of carrying pasted constants.
usage: python tools/ncu_traffic.py gpurun_out/<report>.ncu-rep [workload] > profiles/r02_ncu_traffic.json"""
import csv
import io
import json
import subprocess
import sys

STAGE = {"preprocess_kernel": "preprocess", "compact_kernel": "depth_sort", "sort_pass_kernel": "depth_sort", "bin_count_kernel": "bin_count",
         "bin_prefix_kernel": "bin_count", "bin_ranges_kernel": "bin_count", "bin_fill_kernel": "bin_fill",
         "render_forward_kernel": "render_fwd", "render_backward_kernel": "render_bwd", "gaussian_backward_kernel": "gaussian_bwd"}


def main(rep, workload="C2"):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki, ri, wi, ti = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per_kernel = {}
    for r in rows[2:]:
        name = next((k for k in STAGE if k in r[ki]), None)
        if name is None:
            continue
        b = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
        per_kernel.setdefault(name, []).append((b, float(r[ti])))
    stage = {}
    launches = {"sort_pass_kernel": 4}
    for k, v in per_kernel.items():
        mean_b = sum(x[0] for x in v) / len(v)
        stage[STAGE[k]] = stage.get(STAGE[k], 0.0) + mean_b * launches.get(k, 1)
    out = {"source": f"profiles/{rep.split('/')[-1].replace('.ncu-rep', '')} (ncu --set full --clock-control none; dram__bytes_read.sum + "
                     "dram__bytes_write.sum, mean per launch, summed over the kernels of a stage)",
           workload: {k: round(v) for k, v in stage.items()},
           "kernels": {k: {"launches_captured": len(v), "dram_bytes_per_launch": round(sum(x[0] for x in v) / len(v)),
                           "gpu_time_us": round(sum(x[1] for x in v) / len(v), 2)} for k, v in per_kernel.items()}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:])
