"""One fused attention backward at the guided step's self-attention shape (25 x 5 heads x 2560 x 2560), for ncu."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "guidedvd-3dgs_b200"))
import torch
from vc_b200 import ops

B, N, H = 25, 2560, 5
g = torch.Generator(device="cuda").manual_seed(3)
q, k, v, do = (torch.randn(B, N, H * 64, device="cuda", generator=g).bfloat16() for _ in range(4))
out, lse = ops.flash_attention_lse(q, k, v, B, N, N, H, 0.125)
for _ in range(3):
    ops.flash_attention_bwd(q, k, v, out, lse, do, B, N, N, H, 0.125)
torch.cuda.synchronize()
