"""A/B of the column split for N = 320 / 640 outputs (GVD_GEMM_SPLIT=0|1 for the whole run): linears and implicit 3x3 convolutions
on the U-Net's shapes at C3 (230400 = 25 x 72 x 128 rows at ds0, 57600 at ds1) and C4 (64000 / 16000)."""
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
tag = "split" if os.environ.get("GVD_GEMM_SPLIT", "1") != "0" else "whole"
shapes = [(230400,320,k) for k in (320,1280)] + [(57600,640,k) for k in (640,2560)] + [(64000,320,320),(64000,320,1280),(16000,640,640),(230400,960,320)]
for (M,N,K) in shapes:
    A=torch.randn(M,K,device="cuda").bfloat16(); B=torch.randn(N,K,device="cuda").bfloat16(); bias=torch.randn(N,device="cuda")
    ms=t(lambda: ops.linear(A,B,bias=bias))
    print(f"[{tag}] linear M={M} N={N} K={K}: {ms:.3f} ms {2*M*N*K/ms/1e9:.0f} TF/s", flush=True)
for (F,H,W,Cin,Cout) in ((25,72,128,320,320),(25,72,128,640,320),(25,36,64,640,640),(25,36,64,1280,640),(25,40,64,320,320)):
    x=torch.randn(F,H*W,Cin,device="cuda").bfloat16(); w=(torch.randn(Cout,9*Cin,device="cuda")*0.05).bfloat16(); b=torch.randn(Cout,device="cuda")
    ms=t(lambda: ops.conv3x3(x,F,H,W,w,b), n=10)
    print(f"[{tag}] conv3x3 F={F} {H}x{W} {Cin}->{Cout}: {ms:.3f} ms {2*F*H*W*Cout*Cin*9/ms/1e9:.0f} TF/s", flush=True)
