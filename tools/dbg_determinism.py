"""Run the same U-Net forward twice and report whether the outputs are bit-identical."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "guidedvd-3dgs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import unet_ref  # noqa: E402
from vc_b200.unet import UNetB200  # noqa: E402

h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (40, 64)
ref, cfg = unet_ref.build_reference_unet(model_channels=320, device="cpu")
unet = UNetB200(ref.state_dict(), device="cuda", **cfg)
del ref
x, cc, ctx, _ = unet_ref.synth_inputs(25, h, w, device="cuda")
xc = torch.cat([x, cc], 1)
ts, fs = torch.tensor([999], device="cuda"), torch.tensor([10], device="cuda")
a = unet(xc, ts, ctx, fs=fs).float()
b = unet(xc, ts, ctx, fs=fs).float()
print("latent", h, w, "bit-identical:", bool(torch.equal(a, b)), "rel_l2:", float((a - b).norm() / a.norm()))
