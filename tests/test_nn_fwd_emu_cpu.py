"""The denoiser's memory-bound FORWARD kernels (csrc/nn_kernels.cu: GroupNorm in its three activation modes and its
sharded stats/apply split, LayerNorm, GEGLU, row softmax, the two im2col layouts incl. stride 2 and the fused 2x
upsampling, temporal attention, the fused DDIM update) executed on the host (tests/cuda_emu) through the product's
bindings, against
  (1) plain fp32 PyTorch statements of the reference layers, and
  (2) the closed forms of tests/fake_nn_lib.py.
(2) closes the chain of trust of the network-level CPU tests: the reference U-Net / VAE / guided sampler are compared
with vc_b200 running over fake_nn_lib, and fake_nn_lib is compared here with the kernels' own source."""
import ctypes as C
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "guidedvd-3dgs_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "cuda_emu")):
    if p not in sys.path:
        sys.path.insert(0, p)

from test_unet_grad_cpu import _rel, install_fake  # noqa: E402

BF = torch.bfloat16
FWD = ("gvd_groupnorm_tmp_floats", "gvd_groupnorm_cl", "gvd_groupnorm_cl_stats", "gvd_groupnorm_cl_apply", "gvd_layernorm", "gvd_geglu",
       "gvd_softmax_rows", "gvd_im2col3x3_cl", "gvd_im2col_t3_cl", "gvd_temporal_attention", "gvd_ddim_step",
       "gvd_groupnorm_cl_keep_stats", "gvd_upsample2x_cl")


@pytest.fixture(scope="module")
def emu_lib():
    import build_emu
    import gvd_native

    return gvd_native.bind_nn(C.CDLL(build_emu.build("nn_kernels", ["nn_kernels.cu", "nn_fast.cu"])), partial=True)


def _both(monkeypatch, lib, fn):
    import gvd_native

    fake = install_fake(monkeypatch, BF)
    for name in FWD:
        setattr(fake, name, getattr(lib, name))
    monkeypatch.setattr(gvd_native, "nn", lambda: fake)
    a = fn()
    install_fake(monkeypatch, BF)
    return a, fn()


def _bf(*shape, seed=0, scale=1.0):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale).to(BF)


@pytest.mark.parametrize("silu", [0, 1, 2])
@pytest.mark.parametrize("F,S,C", [(2, 40, 64), (1, 70, 320), (1, 19, 2560), (3, 700, 320)])
def test_groupnorm_kernels(monkeypatch, emu_lib, F, S, C, silu):
    from vc_b200 import ops

    x = _bf(F, S, C, seed=1, scale=2.0) + 0.5
    g = torch.Generator().manual_seed(3)
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.groupnorm(x, gamma, beta, F, S, 32, 1e-5, silu))
    z = torch.nn.functional.group_norm(x.float().permute(0, 2, 1), 32, gamma, beta, 1e-5).permute(0, 2, 1)
    ref = torch.nn.functional.silu(z.to(BF).float() if silu == 1 else z) if silu else z
    assert _rel(emu, closed) < 4e-3 and _rel(emu, ref) < 6e-3   # bf16 output rounding (and the extra rounding of mode 1)


@pytest.mark.parametrize("F,S,C", [(2, 200, 64), (1, 70, 320)])
def test_groupnorm_keep_stats_equals_fused_and_split(monkeypatch, emu_lib, F, S, C):
    """gvd_groupnorm_cl_keep_stats: the output bits of the fused call and the statistic bits of the split one (the guided
    tape's forward must not differ from the plain sampler's, and its backward reads exactly these sums)."""
    import gvd_native
    from vc_b200 import ops

    x = _bf(F, S, C, seed=7, scale=2.0) + 0.25
    g = torch.Generator().manual_seed(8)
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    fake = install_fake(monkeypatch, BF)
    for name in FWD:
        setattr(fake, name, getattr(emu_lib, name))
    monkeypatch.setattr(gvd_native, "nn", lambda: fake)
    y_fused = ops.groupnorm(x, gamma, beta, F, S, 32, 1e-5, 1)
    y_keep, stats = ops.groupnorm_with_stats(x, gamma, beta, F, S, 32, 1e-5, 1)
    nfl = int(emu_lib.gvd_groupnorm_tmp_floats(F, S, 32))
    tmp, st2 = torch.empty(nfl), torch.empty(F * 32 * 2)
    assert emu_lib.gvd_groupnorm_cl_stats(x.data_ptr(), st2.data_ptr(), F, S, C, 32, tmp.data_ptr(), nfl, None) == 0
    assert torch.equal(y_fused, y_keep) and torch.equal(stats, st2)
    xs = x.float().view(F, S, 32, C // 32)
    assert torch.allclose(stats.view(F, 32, 2)[..., 0], xs.sum(dim=(1, 3)), rtol=1e-5, atol=1e-3)


def test_upsample2x_kernel(monkeypatch, emu_lib):
    from vc_b200 import ops

    F, H, W, C = 2, 5, 7, 24
    x = _bf(F, H * W, C, seed=9)
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.upsample2x(x, F, H, W))
    assert torch.equal(emu, closed)
    assert torch.equal(emu.view(F, 2 * H, 2 * W, C), x.view(F, H, W, C).repeat_interleave(2, dim=1).repeat_interleave(2, dim=2))


def test_groupnorm_sharded_split(monkeypatch, emu_lib):
    """stats -> (sum over shards) -> apply with stat_rows = total rows: two row shards reproduce the one-piece norm."""
    from vc_b200 import ops

    F, S, C, cut = 2, 64, 64, 23
    x = _bf(F, S, C, seed=5, scale=1.5)
    g = torch.Generator().manual_seed(6)
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    import gvd_native
    fake = install_fake(monkeypatch, BF)
    for name in FWD:
        setattr(fake, name, getattr(emu_lib, name))
    monkeypatch.setattr(gvd_native, "nn", lambda: fake)
    whole = ops.groupnorm(x, gamma, beta, F, S, 32, 1e-5, 2)
    shards = [x[:, :cut].contiguous(), x[:, cut:].contiguous()]
    totals = {}

    class Part:
        def __init__(self, sid):
            self.sid = sid

        def sum_stats(self, t):
            if "total" in totals:
                t.copy_(totals["total"])
            else:
                totals[self.sid] = t.clone()
            return t

    for sid, xs in enumerate(shards):
        ops.groupnorm_sharded(xs, gamma, beta, F, xs.shape[1], S, Part(sid), 32, 1e-5, 2)
    totals["total"] = totals[0] + totals[1]
    parts = [ops.groupnorm_sharded(xs, gamma, beta, F, xs.shape[1], S, Part(sid), 32, 1e-5, 2) for sid, xs in enumerate(shards)]
    assert _rel(torch.cat(parts, dim=1), whole) < 4e-3


def test_layernorm_geglu_softmax_kernels(monkeypatch, emu_lib):
    from vc_b200 import ops

    x = _bf(37, 320, seed=4, scale=3.0)
    g = torch.Generator().manual_seed(6)
    gamma, beta = 1 + 0.1 * torch.randn(320, generator=g), 0.1 * torch.randn(320, generator=g)
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.layernorm(x, gamma, beta, 1e-5))
    assert _rel(emu, closed) < 4e-3 and _rel(emu, torch.nn.functional.layer_norm(x.float(), (320,), gamma, beta, 1e-5)) < 5e-3
    h = _bf(21, 2 * 128, seed=7, scale=1.5)
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.geglu(h))
    assert _rel(emu, closed) < 4e-3 and _rel(emu, h.float()[:, :128] * torch.nn.functional.gelu(h.float()[:, 128:])) < 8e-3
    for dtype in (BF, torch.float32):
        s = (torch.randn(19, 80, generator=torch.Generator().manual_seed(9)) * 2).to(dtype)
        emu, closed = _both(monkeypatch, emu_lib, lambda: ops.softmax_rows(s, 77, 80))
        assert _rel(emu, closed) < 4e-3 and float(emu[:, 77:].abs().max()) == 0.0
        assert _rel(emu[:, :77], torch.softmax(s.float()[:, :77], -1)) < 5e-3


@pytest.mark.parametrize("stride,up", [(1, False), (2, False), (1, True)])
@pytest.mark.parametrize("H,W", [(6, 5), (7, 8)])
def test_im2col_kernels(monkeypatch, emu_lib, stride, up, H, W):
    from vc_b200 import ops

    F_, Cin, Cout = 2, 16, 8
    x = _bf(F_, H * W, Cin, seed=11)
    w4 = (torch.randn(Cout, Cin, 3, 3, generator=torch.Generator().manual_seed(12)) / (9 * Cin) ** 0.5).to(BF)
    w = w4.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.conv3x3(x, F_, H, W, w, None, stride=stride, upsample=up)[0])
    img = x.float().view(F_, H, W, Cin).permute(0, 3, 1, 2)
    if up:
        img = torch.nn.functional.interpolate(img, scale_factor=2, mode="nearest")
    ref = torch.nn.functional.conv2d(img, w4.float(), stride=stride, padding=1)
    assert torch.equal(emu, closed)   # the gather is exact; the GEMM stand-in is the same in both runs
    assert _rel(emu, ref.permute(0, 2, 3, 1).reshape(emu.shape)) < 8e-3
    if not up and stride == 1:
        B, T, S = 1, F_ * 2, H
        xt = _bf(B * T, S, Cin, seed=13)
        w5 = (torch.randn(Cout, Cin, 3, 1, 1, generator=torch.Generator().manual_seed(14)) / (3 * Cin) ** 0.5).to(BF)
        wt = w5[:, :, :, 0, 0].permute(0, 2, 1).reshape(Cout, -1).contiguous()
        emu, closed = _both(monkeypatch, emu_lib, lambda: ops.conv_t3(xt, B, T, S, wt, None))
        vol = xt.float().view(B, T, S, 1, Cin).permute(0, 4, 1, 2, 3)
        ref = torch.nn.functional.conv3d(vol, w5.float(), padding=(1, 0, 0)).permute(0, 2, 3, 4, 1).reshape(B * T, S, Cout)
        assert torch.equal(emu, closed) and _rel(emu, ref) < 8e-3


@pytest.mark.parametrize("T", [25, 32, 3, 1])
def test_temporal_attention_kernel(monkeypatch, emu_lib, T):
    from vc_b200 import ops

    B, S, H = 1, 5, 2
    q, k, v = (_bf(B * T, S, H * 64, seed=20 + i) for i in range(3))
    scale = 64 ** -0.5
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.temporal_attention(q, k, v, B, T, S, H, scale))
    sp = lambda t: t.float().view(B, T, S, H, 64).permute(0, 3, 2, 1, 4)  # noqa: E731
    p = torch.softmax(torch.einsum("bhsid,bhsjd->bhsij", sp(q), sp(k)) * scale, -1)
    ref = torch.einsum("bhsij,bhsjd->bhsid", p, sp(v)).permute(0, 3, 2, 1, 4).reshape(B * T, S, H * 64)
    assert _rel(emu, closed) < 4e-3 and _rel(emu, ref) < 1e-2


def test_ddim_step_kernels_match_reference_golden(monkeypatch, emu_lib):
    """The fused DDIM update executed on the host against the golden steps recorded from the REFERENCE sampler
    (tests/golden/ddim_steps.npz, tests/make_golden_ddim.py) -- the same check tests/test_ddim_gpu.py makes on hardware."""
    import gvd_native
    import numpy as np
    from vc_b200 import ops
    from vc_b200.schedule import DdimSchedule, ModelSchedule

    fake = install_fake(monkeypatch, BF)
    fake.gvd_ddim_step = emu_lib.gvd_ddim_step
    monkeypatch.setattr(gvd_native, "nn", lambda: fake)
    g = np.load(os.path.join(ROOT, "tests", "golden", "ddim_steps.npz"))
    sched = DdimSchedule(ModelSchedule(), 50, "uniform_trailing", 1.0)
    for k, idx in enumerate(g["picks"]):
        coef = sched.coefficients(int(idx), float(g["cfg"]), float(g["guidance_rescale"]), 1.0)
        t = lambda name: torch.from_numpy(g[f"{name}_{k}"]).float().contiguous()  # noqa: E731
        x_prev, pred_x0 = ops.ddim_step(t("x"), t("e_c"), t("e_u"), t("noise"), coef)
        assert _rel(x_prev, t("x_prev")) < 5e-6 and _rel(pred_x0, t("pred_x0")) < 5e-6


def test_fast_variants_are_bit_identical(monkeypatch, emu_lib):
    """csrc/nn_fast.cu (GVD_NN_FAST=1: vectorised GEGLU, 32-bit-indexed im2col kernels) against the kernels they replace:
    the same bits, on shapes that exercise row / vector boundaries, stride 2 and the fused upsampling."""
    import gvd_native
    from vc_b200 import ops

    fake = install_fake(monkeypatch, BF)
    for name in FWD + ("gvd_nn_set_fast",):
        setattr(fake, name, getattr(emu_lib, name))
    monkeypatch.setattr(gvd_native, "nn", lambda: fake)
    lib = emu_lib

    def both(fn):
        was = lib.gvd_nn_set_fast(0)
        a = fn()
        lib.gvd_nn_set_fast(2)
        b = fn()
        lib.gvd_nn_set_fast(was)
        return a, b

    for rows, D in ((21, 128), (7, 1280), (130, 8)):
        h = _bf(rows, 2 * D, seed=rows, scale=1.5)
        a, b = both(lambda: ops.geglu(h))
        assert torch.equal(a, b), (rows, D)
    for (F_, H, W, Cin, stride, up) in ((2, 6, 5, 16, 1, 0), (1, 7, 8, 8, 2, 0), (3, 5, 4, 24, 1, 1)):
        x = _bf(F_, H * W, Cin, seed=H)
        Hin, Win = (2 * H, 2 * W) if up else (H, W)
        Ho, Wo = (Hin - 1) // stride + 1, (Win - 1) // stride + 1
        outs = []
        for fast in (0, 1):
            lib.gvd_nn_set_fast(fast)
            c = torch.full((F_ * Ho * Wo, 9 * Cin), 3.0, dtype=BF)
            assert lib.gvd_im2col3x3_cl(x.data_ptr(), c.data_ptr(), F_, H, W, Cin, stride, up, None) == 0
            outs.append(c)
        assert torch.equal(outs[0], outs[1]), (F_, H, W, Cin, stride, up)
    for (B, T, S, Cin) in ((1, 5, 7, 16), (2, 3, 9, 8)):
        x = _bf(B * T, S, Cin, seed=T)
        outs = []
        for fast in (0, 1):
            lib.gvd_nn_set_fast(fast)
            c = torch.full((B * T * S, 3 * Cin), 3.0, dtype=BF)
            assert lib.gvd_im2col_t3_cl(x.data_ptr(), c.data_ptr(), B, T, S, Cin, None) == 0
            outs.append(c)
        assert torch.equal(outs[0], outs[1])
    for T in (25, 32, 3, 1):
        q, k, v = (_bf(T, 6, 2 * 64, seed=50 + i + T) for i in range(3))
        a, b = both(lambda: ops.temporal_attention(q, k, v, 1, T, 6, 2, 0.125))
        assert torch.equal(a, b), T
    lib.gvd_nn_set_fast(0)
    assert lib.gvd_nn_set_fast(7) == 0   # any other value only queries
    lib.gvd_nn_set_fast(1)
