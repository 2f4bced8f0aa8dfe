import sys
sys.path.insert(0, "tests"); sys.path.insert(0, "guidedvd-3dgs_b200")
import torch, unet_ref
from vc_b200.unet import UNetB200
mc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
t, h, w = 5, 16, 16
ref, cfg = unet_ref.build_reference_unet(model_channels=mc)
ours = UNetB200(ref.state_dict(), device="cuda", **cfg)
x, cc, ctx, _ = unet_ref.synth_inputs(t, h, w)
xin = torch.cat([x, cc], 1); ts = torch.tensor([481], device="cuda"); fs = torch.tensor([10], device="cuda")
acts = {}
def hook(name):
    def f(m, i, o): acts[name] = o.detach().float()
    return f
for i, m in enumerate(ref.input_blocks): m.register_forward_hook(hook(f"input_blocks.{i}"))
ref.init_attn.register_forward_hook(hook("init_attn"))
ref.middle_block.register_forward_hook(hook("middle_block"))
for i, m in enumerate(ref.output_blocks): m.register_forward_hook(hook(f"output_blocks.{i}"))
# finer: submodules of input_blocks.1
for n, m in ref.input_blocks[1].named_children(): m.register_forward_hook(hook(f"ib1.{n}"))
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    y_ref = ref(xin, ts, context=ctx, fs=fs)
ours.trace = []
y = ours(xin, ts, ctx, fs=fs)
for name, a in ours.trace:
    r = acts[name]
    err = ((a - r).norm() / r.norm()).item()
    print(f"{name:20s} rel {err:.3e}  |ref| {r.abs().mean().item():.3e} nan={bool(torch.isnan(a).any())}")
print("final", ((y.float() - y_ref.float()).norm() / y_ref.float().norm()).item())
