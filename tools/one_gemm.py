import sys
sys.path.insert(0, "guidedvd-3dgs_b200")
import torch
from vc_b200 import ops
M, N, K = (int(a) for a in sys.argv[1:4])
A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16(); bias = torch.randn(N, device="cuda")
for _ in range(3):
    ops.linear(A, B, bias=bias)
torch.cuda.synchronize()
