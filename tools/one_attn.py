import sys
sys.path.insert(0, "guidedvd-3dgs_b200")
import torch
from vc_b200 import ops
B, N, H = 25, 9216, 5
q, k, v = (torch.randn(B, N, H * 64, device="cuda").bfloat16() for _ in range(3))
for _ in range(2):
    ops.flash_attention(q, k, v, B, N, N, H, 0.125)
A = torch.randn(57600, 640, device="cuda").bfloat16(); W = torch.randn(5120, 640, device="cuda").bfloat16(); b = torch.randn(5120, device="cuda")
for _ in range(2):
    ops.linear(A, W, bias=b)
A = torch.randn(14400, 11520, device="cuda").bfloat16(); W = torch.randn(1280, 11520, device="cuda").bfloat16(); b = torch.randn(1280, device="cuda")
for _ in range(2):
    ops.linear(A, W, bias=b)
torch.cuda.synchronize()
