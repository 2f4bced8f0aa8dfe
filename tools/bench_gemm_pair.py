"""A/B sweep of the CTA-pair GEMM kernel against the one-CTA kernels (GVD_GEMM_PAIR=0|1 picks the mode for the whole run)."""
import os, sys
sys.path.insert(0, "guidedvd-3dgs_b200")
import torch
from vc_b200 import ops
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
tag = "pair" if os.environ.get("GVD_GEMM_PAIR", "1") != "0" else "one "
shapes = [(57600,1280,k) for k in (320,640,1280,2560,5120)] + [(57600,640,k) for k in (640,1280,2560,5760)] + \
         [(230400,320,k) for k in (320,1280,2880)] + [(230400,2560,320),(230400,1280,320),(14400,1280,11520),(14400,2560,1280),(14400,10240,1280),(3600,1280,11520)]
for (M,N,K) in shapes:
    A=torch.randn(M,K,device="cuda").bfloat16(); B=torch.randn(N,K,device="cuda").bfloat16(); bias=torch.randn(N,device="cuda")
    ms=t(lambda: ops.linear(A,B,bias=bias))
    print(f"[{tag}] M={M} N={N} K={K}: {ms:.3f} ms {2*M*N*K/ms/1e9:.0f} TF/s", flush=True)
