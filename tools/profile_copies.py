"""Which torch copy kernels does one native guided step launch?  Groups aten::copy_/contiguous/pad by input shape."""
import os, sys, collections
import torch
from torch.profiler import ProfilerActivity, profile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.argv = [sys.argv[0], "guided"]
src = open(os.path.join(ROOT, "tools", "profile_guided.py")).read().split("step()\ntorch.cuda.synchronize()\nwith profile")[0]
exec(compile(src, "profile_guided_setup", "exec"))
step(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof:
    step(); torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=6):
    if e.key in ("aten::copy_", "aten::contiguous", "aten::pad", "aten::clone", "aten::_to_copy", "aten::add", "aten::add_", "aten::cat", "aten::zeros"):
        rows.append((e.device_time_total, e.count, e.key, str(e.input_shapes)[:90], [s for s in e.stack if "vc_b200" in s or "tests" in s][:3]))
rows.sort(reverse=True)
for r in rows[:30]:
    print(f"{r[0] / 1e3:8.2f} ms x{r[1]:4d} {r[2]:16s} {r[3]}\n      {r[4]}")
