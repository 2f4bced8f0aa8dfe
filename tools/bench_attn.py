"""Times gvd_flash_attention on the U-Net's self-attention shapes at C3 (env GVD_FLASH=v1|v2 and GVD_FLASH_POLY=0|1 pick the variant)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "guidedvd-3dgs_b200"))
import torch
from vc_b200 import ops

tag = f"{os.environ.get('GVD_FLASH', 'v2')} poly={os.environ.get('GVD_FLASH_POLY', '0')}"
for (B, N, H) in ((25, 9216, 5), (25, 2304, 10), (25, 576, 20)):
    g = torch.Generator(device="cuda").manual_seed(3)
    q, k, v = (torch.randn(B, N, H * 64, device="cuda", generator=g).bfloat16() for _ in range(3))
    for _ in range(3):
        o = ops.flash_attention(q, k, v, B, N, N, H, 0.125)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        o = ops.flash_attention(q, k, v, B, N, N, H, 0.125)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 4.0 * B * H * N * N * 64
    # error against fp32 softmax attention on one (batch, head), first 512 queries
    qh, kh, vh = (t[0].float().view(N, H, 64)[:, 0] for t in (q, k, v))
    ref = torch.softmax(qh[:512] @ kh.T * 0.125, -1) @ vh
    err = (o[0].float().view(N, H, 64)[:512, 0] - ref).abs().max().item() / ref.abs().max().item()
    print(f"[{tag}] B={B} N={N} H={H}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s  rel-max-err {err:.2e}", flush=True)
