"""Times the memory-bound adjoints of the guided step on its C4 shapes (25 frames, latent 40 x 64): GroupNorm(+SiLU)
forward / backward in GB/s of algorithmic traffic, and the temporal attention backward (GVD_TATTN_MMA=0: first kernel)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "guidedvd-3dgs_b200"))
import torch
from vc_b200 import ops

BF = torch.bfloat16


def timed(fn, n=20):
    """ms per call, GPU time only: the n calls are captured into one CUDA graph (as the U-Net runs them) and replayed."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(n):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * n)


only = sys.argv[1] if len(sys.argv) > 1 else "all"
if only in ("all", "gn"):
    # rotate over 4 buffer sets so that a 41 MB tensor is not served from the 126 MB L2 between repetitions
    for (F, S, C, silu) in ((25, 2560, 320, 1), (25, 2560, 640, 1), (25, 640, 1280, 1), (25, 160, 1280, 1), (25, 2560, 320, 0)):
        sets = []
        for i in range(4):
            g = torch.Generator(device="cuda").manual_seed(i)
            sets.append((torch.randn(F, S, C, device="cuda", generator=g).to(BF), torch.randn(F, S, C, device="cuda", generator=g).to(BF)))
        gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
        stats = [ops.groupnorm_with_stats(x, gamma, beta, F, S, silu=silu)[1] for x, _ in sets]
        it = [0]

        def fwd():
            x, _ = sets[it[0] % 4]
            it[0] += 1
            ops.groupnorm_with_stats(x, gamma, beta, F, S, silu=silu)

        def bwd():
            i = it[0] % 4
            it[0] += 1
            ops.groupnorm_bwd(sets[i][0], sets[i][1], gamma, beta, F, S, silu=silu, stats=stats[i])

        nb = F * S * C * 2
        tf, tb = timed(fwd), timed(bwd)
        print(f"GroupNorm F={F} S={S} C={C} silu={silu}: fwd {tf * 1e3:.0f} us ({3 * nb / tf / 1e6:.0f} GB/s of 3 passes) | "
              f"bwd {tb * 1e3:.0f} us ({5 * nb / tb / 1e6:.0f} GB/s of 5 passes; minimum 3 passes = {3 * nb / 6.5e12 * 1e6:.0f} us at 6.5 TB/s)", flush=True)
if only in ("all", "tattn"):
    for (B, T, S, H) in ((1, 25, 2560, 5), (1, 25, 640, 10), (1, 25, 160, 20)):
        g = torch.Generator(device="cuda").manual_seed(1)
        q, k, v, do = (torch.randn(B, T, S, H * 64, device="cuda", generator=g).to(BF) for _ in range(4))
        t = timed(lambda: ops.temporal_attention_bwd(q, k, v, do, B, T, S, H, 0.125))
        nb = 7 * q.numel() * 2
        print(f"temporal attention bwd T={T} S={S} H={H} (GVD_TATTN_MMA={os.environ.get('GVD_TATTN_MMA', '1')}): {t * 1e3:.0f} us "
              f"({nb / t / 1e6:.0f} GB/s of 7 tensors)", flush=True)
