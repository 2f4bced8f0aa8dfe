import sys
sys.path.insert(0, "tests"); sys.path.insert(0, "guidedvd-3dgs_b200")
import torch, unet_ref
from torch.profiler import profile, ProfilerActivity
from vc_b200.unet import UNetB200
t, h, w = (int(a) for a in sys.argv[1:4])
ref, cfg = unet_ref.build_reference_unet(model_channels=320)
ours = UNetB200(ref.state_dict(), device="cuda", **cfg)
which = sys.argv[4] if len(sys.argv) > 4 else "ours"
x, cc, ctx, _ = unet_ref.synth_inputs(t, h, w)
xin = torch.cat([x, cc], 1); ts = torch.tensor([481], device="cuda"); fs = torch.tensor([10], device="cuda")
def run():
    if which == "ours":
        return ours(xin, ts, ctx, fs=fs)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return ref(xin, ts, context=ctx, fs=fs)
run(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
