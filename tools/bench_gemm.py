import sys, time
sys.path.insert(0, "guidedvd-3dgs_b200")
import torch
from vc_b200 import ops
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
for (M,N,K) in [(230400,320,320),(230400,320,2880),(57600,640,5760),(14400,1280,11520),(230400,2560,320),(230400,320,1280),(8192,8192,8192)]:
    A=torch.randn(M,K,device="cuda").bfloat16(); B=torch.randn(N,K,device="cuda").bfloat16(); bias=torch.randn(N,device="cuda")
    ms=t(lambda: ops.linear(A,B,bias=bias)); ms_t=t(lambda: torch.nn.functional.linear(A,B,bias.bfloat16()))
    fl=2*M*N*K
    print(f"M={M} N={N} K={K}: ours {ms:.3f} ms {fl/ms/1e9:.0f} TF/s | torch {ms_t:.3f} ms {fl/ms_t/1e9:.0f} TF/s")
