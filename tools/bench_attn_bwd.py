"""Times the attention adjoint on the guided sampler's self-attention shapes (C4 latent 40 x 64: 2560 tokens, 25 frames):
fused (gvd_flash_attention_bwd) against the first backward (scores materialised, ops.attention_bwd)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "guidedvd-3dgs_b200"))
import torch
from vc_b200 import ops


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (B, N, Nk, H, shared) in ((25, 2560, 2560, 5, False), (25, 640, 640, 10, False), (25, 160, 160, 20, False), (25, 2560, 77, 5, True)):
    g = torch.Generator(device="cuda").manual_seed(3)
    q, do = (torch.randn(B, N, H * 64, device="cuda", generator=g).bfloat16() for _ in range(2))
    k, v = (torch.randn(1 if shared else B, Nk, H * 64, device="cuda", generator=g).bfloat16() for _ in range(2))
    out, lse = ops.flash_attention_lse(q, k, v, B, N, Nk, H, 0.125, shared_kv=shared)
    t_f = timed(lambda: ops.flash_attention_lse(q, k, v, B, N, Nk, H, 0.125, shared_kv=shared))
    t_b = timed(lambda: ops.flash_attention_bwd(q, k, v, out, lse, do, B, N, Nk, H, 0.125, shared_kv=shared, need_kv=not shared))
    t_m = timed(lambda: ops.attention_bwd(q, k, v, do, B, N, Nk, H, 0.125, shared_kv=shared, need_kv=not shared), n=3)
    fl = 2.0 * B * H * N * Nk * 64 * (3 if shared else 7)  # MMAs the fused kernels execute
    print(f"B={B} Nq={N} Nk={Nk} H={H} shared={shared}: forward+lse {t_f:.3f} ms | fused bwd {t_b:.3f} ms ({fl / t_b / 1e9:.0f} TFLOP/s executed) | "
          f"materialised bwd {t_m:.3f} ms  ({t_m / t_b:.1f}x)", flush=True)
