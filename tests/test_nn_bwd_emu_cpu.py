"""Runs the CUDA SOURCE of csrc/nn_backward.cu on the host (tests/cuda_emu: every CUDA thread of a block is an OS thread,
barriers / shuffles / shared-memory atomics keep their meaning) and compares each kernel, through its C-ABI entry point
and the product's own ctypes bindings (vc_b200.ops), with
  (1) torch.autograd over a plain fp32 statement of the layer, and
  (2) the closed forms of tests/fake_nn_lib.py that the network-level CPU tests rely on.
This is what stands in for a GPU run of kernels written without GPU access: index math, shared-memory layout, barrier
placement and reduction logic are executed for real; alignment faults, resource limits and intrinsic accuracy are not
modelled (tests/test_zz_guided_gpu.py covers the hardware)."""
import ctypes as C
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "guidedvd-3dgs_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "cuda_emu")):
    if p not in sys.path:
        sys.path.insert(0, p)

from fake_nn_lib import FakeNN  # noqa: E402
from test_unet_grad_cpu import _rel, install_fake  # noqa: E402

BF = torch.bfloat16
EMU_FUNCS = ("gvd_groupnorm_bwd_tmp_bytes", "gvd_groupnorm_cl_bwd", "gvd_layernorm_bwd", "gvd_geglu_bwd", "gvd_softmax_bwd_rows",
             "gvd_col2im3x3_cl", "gvd_col2im_t3_cl", "gvd_temporal_attention_bwd", "gvd_ddim_pred_x0_vjp", "gvd_upsample2x_bwd_cl")


@pytest.fixture(scope="module")
def emu_lib():
    import build_emu
    import gvd_native

    lib = C.CDLL(build_emu.build("nn_backward"))
    vp, ll, i32, f32 = C.c_void_p, C.c_longlong, C.c_int, C.c_float
    lib.gvd_groupnorm_bwd_tmp_bytes.restype = C.c_size_t
    lib.gvd_groupnorm_bwd_tmp_bytes.argtypes = [i32, ll, i32]
    lib.gvd_groupnorm_cl_bwd.argtypes = [vp, vp, vp, vp, vp, vp, i32, ll, i32, i32, f32, i32, vp, C.c_size_t, vp]
    lib.gvd_layernorm_bwd.argtypes = [vp, vp, vp, vp, ll, i32, f32, vp]
    lib.gvd_geglu_bwd.argtypes = [vp, vp, vp, ll, i32, vp]
    lib.gvd_softmax_bwd_rows.argtypes = [vp, vp, vp, ll, ll, i32, vp]
    lib.gvd_col2im3x3_cl.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.gvd_col2im_t3_cl.argtypes = [vp, vp, i32, i32, ll, i32, vp]
    lib.gvd_upsample2x_bwd_cl.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.gvd_temporal_attention_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, ll, i32, f32, vp]
    lib.gvd_ddim_pred_x0_vjp.argtypes = [C.POINTER(gvd_native.DdimVjpArgs), vp]
    return lib


class EmuNN(FakeNN):
    """FakeNN (bf16 storage) whose input-gradient entry points are the emulated CUDA kernels."""

    def __init__(self, lib):
        super().__init__(BF)
        for name in EMU_FUNCS:
            setattr(self, name, getattr(lib, name))


def _install(monkeypatch, lib, emulated):
    fake = install_fake(monkeypatch, BF)
    if emulated:
        import gvd_native
        emu = EmuNN(lib)
        monkeypatch.setattr(gvd_native, "nn", lambda: emu)
        return emu
    return fake


def _both(monkeypatch, lib, fn):
    """Evaluate fn() once over the emulated kernels and once over the closed forms."""
    _install(monkeypatch, lib, True)
    a = fn()
    _install(monkeypatch, lib, False)
    b = fn()
    return a, b


def _bf(*shape, seed=0, scale=1.0):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale).to(BF)


@pytest.mark.parametrize("silu", [0, 1, 2])
@pytest.mark.parametrize("F,S,C", [(2, 40, 64), (1, 70, 320), (1, 19, 2560), (3, 700, 320)])
def test_groupnorm_bwd_kernels(monkeypatch, emu_lib, F, S, C, silu):
    from vc_b200 import ops

    x, dy = _bf(F, S, C, seed=1, scale=2.0) + 0.5, _bf(F, S, C, seed=2)
    g = torch.Generator().manual_seed(3)
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.groupnorm_bwd(x, dy, gamma, beta, F, S, 32, 1e-5, silu))
    xf = x.float().requires_grad_(True)
    z = torch.nn.functional.group_norm(xf.permute(0, 2, 1), 32, gamma, beta, 1e-5).permute(0, 2, 1)
    if silu == 1:
        z = z + (z.to(BF).float() - z).detach()
    (torch.nn.functional.silu(z) if silu else z).backward(dy.float())
    assert _rel(emu, closed) < 4e-3          # both round their result to bf16
    assert _rel(emu, xf.grad) < 1e-2


def test_layernorm_geglu_softmax_bwd_kernels(monkeypatch, emu_lib):
    from vc_b200 import ops

    x, dy = _bf(37, 320, seed=4, scale=3.0), _bf(37, 320, seed=5)
    g = torch.Generator().manual_seed(6)
    gamma, beta = 1 + 0.1 * torch.randn(320, generator=g), 0.1 * torch.randn(320, generator=g)
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.layernorm_bwd(x, dy, gamma, 1e-5))
    xf = x.float().requires_grad_(True)
    torch.nn.functional.layer_norm(xf, (320,), gamma, beta, 1e-5).backward(dy.float())
    assert _rel(emu, closed) < 4e-3 and _rel(emu, xf.grad) < 1e-2

    h, do = _bf(21, 2 * 128, seed=7, scale=1.5), _bf(21, 128, seed=8)
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.geglu_bwd(h, do))
    hf = h.float().requires_grad_(True)
    (hf[:, :128] * torch.nn.functional.gelu(hf[:, 128:])).backward(do.float())
    assert _rel(emu, closed) < 4e-3 and _rel(emu, hf.grad) < 1e-2

    rows, cols, ld = 19, 77, 80
    p = torch.softmax(_bf(rows, cols, seed=9, scale=2.0).float(), -1)
    pp = torch.zeros(rows, ld, dtype=BF)
    pp[:, :cols] = p.to(BF)
    dp = _bf(rows, ld, seed=10)
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.softmax_bwd_rows(pp, dp.clone(), cols))
    ref = pp.float()[:, :cols] * (dp.float()[:, :cols] - (pp.float()[:, :cols] * dp.float()[:, :cols]).sum(-1, keepdim=True))
    assert _rel(emu, closed) < 4e-3 and _rel(emu[:, :cols], ref) < 1e-2 and float(emu[:, cols:].abs().max()) == 0.0


@pytest.mark.parametrize("stride,up", [(1, False), (2, False), (1, True)])
@pytest.mark.parametrize("H,W", [(6, 5), (7, 8)])
def test_col2im3x3_kernel(monkeypatch, emu_lib, stride, up, H, W):
    from vc_b200 import ops

    F_, Cin, Cout = 2, 16, 8
    w4 = (torch.randn(Cout, Cin, 3, 3, generator=torch.Generator().manual_seed(12)) / (9 * Cin) ** 0.5).to(BF)
    w = w4.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
    xf = _bf(F_, H * W, Cin, seed=11).float().requires_grad_(True)
    img = xf.view(F_, H, W, Cin).permute(0, 3, 1, 2)
    if up:
        img = torch.nn.functional.interpolate(img, scale_factor=2, mode="nearest")
    y_ref = torch.nn.functional.conv2d(img, w4.float(), stride=stride, padding=1)
    Ho, Wo = y_ref.shape[2:]
    dy = _bf(F_, Ho * Wo, Cout, seed=13)
    y_ref.backward(dy.float().view(F_, Ho, Wo, Cout).permute(0, 3, 1, 2))
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.conv3x3_dx(dy, F_, H, W, Cin, w, stride, up))
    assert _rel(emu, closed) < 1e-3          # same bf16 dcol, fp32 sums of <= 36 taps in a different order
    assert _rel(emu, xf.grad) < 1.5e-2


def test_groupnorm_frame_groups_in_a_subprocess():
    """GVD_GN_GROUP_MB (read once per process) splits the frames of one GroupNorm call into L2-sized launch pairs; with a 1 MB
    budget the (3, 700, 320) cases above run as three groups, each with its own chunking, scratch window and statistics
    offset -- forward (fused + kept statistics) and backward, same tolerances."""
    import subprocess

    env = dict(os.environ, GVD_GN_GROUP_MB="1")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider",
                        os.path.join(here, "test_nn_fwd_emu_cpu.py"), os.path.join(here, "test_nn_bwd_emu_cpu.py"),
                        "-k", "(test_groupnorm_kernels or test_groupnorm_bwd_kernels) and 700"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_upsample2x_bwd_kernel(monkeypatch, emu_lib):
    from vc_b200 import ops

    F_, H, W, Cc = 2, 5, 7, 24
    dy = _bf(F_, 4 * H * W, Cc, seed=17)
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.upsample2x_bwd(dy, F_, H, W))
    want = dy.float().view(F_, H, 2, W, 2, Cc).sum(dim=(2, 4)).reshape(F_, H * W, Cc)
    assert _rel(emu, closed) < 4e-3 and _rel(emu, want) < 4e-3  # bf16 output rounding; the order of the four adds differs


def test_col2im_t3_kernel(monkeypatch, emu_lib):
    from vc_b200 import ops

    B, T, S, Cin, Cout = 2, 5, 9, 16, 8
    w5 = (torch.randn(Cout, Cin, 3, 1, 1, generator=torch.Generator().manual_seed(14)) / (3 * Cin) ** 0.5).to(BF)
    w = w5[:, :, :, 0, 0].permute(0, 2, 1).reshape(Cout, -1).contiguous()
    xf = _bf(B * T, S, Cin, seed=15).float().requires_grad_(True)
    vol = xf.view(B, T, S, 1, Cin).permute(0, 4, 1, 2, 3)
    y_ref = torch.nn.functional.conv3d(vol, w5.float(), padding=(1, 0, 0))
    dy = _bf(B * T, S, Cout, seed=16)
    y_ref.backward(dy.float().view(B, T, S, 1, Cout).permute(0, 4, 1, 2, 3))
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.conv_t3_dx(dy, B, T, S, Cin, w))
    assert _rel(emu, closed) < 4e-3 and _rel(emu, xf.grad) < 1.5e-2


@pytest.mark.parametrize("T", [25, 32, 3, 1])
def test_temporal_attention_bwd_kernel(monkeypatch, emu_lib, T):
    from vc_b200 import ops

    B, S, H = 1, 5, 2
    q, k, v = (_bf(B * T, S, H * 64, seed=20 + i) for i in range(3))
    do = _bf(B * T, S, H * 64, seed=24)
    scale = 64 ** -0.5
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.temporal_attention_bwd(q, k, v, do, B, T, S, H, scale))
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    sp = lambda t: t.view(B, T, S, H, 64).permute(0, 3, 2, 1, 4)  # noqa: E731
    p = torch.softmax(torch.einsum("bhsid,bhsjd->bhsij", sp(qf), sp(kf)) * scale, -1)
    torch.einsum("bhsij,bhsjd->bhsid", p, sp(vf)).permute(0, 3, 2, 1, 4).reshape(B * T, S, H * 64).backward(do.float())
    for name, a, b, ref in zip("qkv", emu, closed, (qf.grad, kf.grad, vf.grad)):
        if T == 1 and name in "qk":  # one key: the softmax is constant, dq = dk = 0 exactly
            assert float(a.abs().max()) == 0.0 and float(ref.abs().max()) < 1e-6
            continue
        assert _rel(a, b) < 4e-3, name
        assert _rel(a, ref) < 2e-2, name   # the forward rounds its logits to bf16 twice (attention.py:103)


def test_pred_x0_vjp_kernels(monkeypatch, emu_lib):
    from vc_b200 import ops
    from vc_b200.schedule import DdimSchedule, ModelSchedule

    g = torch.Generator().manual_seed(4)
    shape = (1, 4, 5, 12, 10)
    coef = DdimSchedule(ModelSchedule(), 50, "uniform_trailing", 1.0).coefficients(30, 7.5, 0.7, 1.0)
    e_c = torch.randn(shape, generator=g)
    e_u = e_c + 0.3 * torch.randn(shape, generator=g)
    G = torch.randn(shape, generator=g)
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.ddim_pred_x0_vjp(e_c, e_u, G, coef))
    for a, b in zip(emu, closed):
        assert _rel(a, b) < 1e-5
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.ddim_pred_x0_vjp(e_c, None, G, coef))
    assert emu[2] is None and _rel(emu[0], closed[0]) < 1e-6 and _rel(emu[1], closed[1]) < 1e-6


def test_emulated_kernels_reject_bad_arguments(emu_lib):
    """Error behaviour of the C ABI (return 2 + message, never a crash) -- exercised on the host build."""
    buf = torch.zeros(1024, dtype=BF)
    p = buf.data_ptr()
    assert emu_lib.gvd_groupnorm_cl_bwd(p, p, p, p, p, p, 1, 4, 60, 32, 1e-5, 0, p, 1 << 20, None) == 2   # C % groups
    assert emu_lib.gvd_groupnorm_cl_bwd(p, p, p, p, p, p, 1, 4, 64, 32, 1e-5, 3, p, 1 << 20, None) == 2   # do_silu range
    assert emu_lib.gvd_groupnorm_cl_bwd(p, p, p, p, p, p, 1, 4, 64, 32, 1e-5, 0, p, 8, None) == 2         # scratch too small
    assert emu_lib.gvd_layernorm_bwd(p, p, p, p, 4, 63, 1e-5, None) == 2
    assert emu_lib.gvd_layernorm_bwd(p, None, p, p, 4, 64, 1e-5, None) == 2
    assert emu_lib.gvd_softmax_bwd_rows(p, p, p, 8, 4, 9, None) == 2
    assert emu_lib.gvd_col2im3x3_cl(p, p, 1, 4, 4, 12, 1, 0, None) == 2
    assert emu_lib.gvd_col2im3x3_cl(p, p, 1, 4, 4, 16, 2, 1, None) == 2
    assert emu_lib.gvd_temporal_attention_bwd(p, p, p, p, p, p, p, 1, 33, 4, 1, 0.125, None) == 2
    assert emu_lib.gvd_layernorm_bwd(p, p, p, p, 0, 64, 1e-5, None) == 0   # empty input is a no-op


@pytest.mark.parametrize("H,W", [(8, 6), (9, 7), (2, 2)])
def test_encoder_downsample_im2col_kernel(monkeypatch, H, W):
    """csrc/nn_vae.cu executed on the host: the right/bottom-padded stride-2 im2col of the VAE encoder's Downsample
    (ae_modules.py:93-106) through ops.conv3x3_down vs F.pad(x, (0,1,0,1)) + conv2d(stride=2, padding=0)."""
    import build_emu
    import gvd_native
    from vc_b200 import ops

    lib = C.CDLL(build_emu.build("nn_vae"))
    lib.gvd_im2col3x3_down_cl.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    fake = install_fake(monkeypatch, BF)
    fake.gvd_im2col3x3_down_cl = lib.gvd_im2col3x3_down_cl
    monkeypatch.setattr(gvd_native, "nn", lambda: fake)
    F_, Cin, Cout = 2, 16, 8
    x = _bf(F_, H * W, Cin, seed=40)
    w4 = (torch.randn(Cout, Cin, 3, 3, generator=torch.Generator().manual_seed(41)) / (9 * Cin) ** 0.5).to(BF)
    w = w4.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
    y, Ho, Wo = ops.conv3x3_down(x, F_, H, W, w, None)
    img = torch.nn.functional.pad(x.float().view(F_, H, W, Cin).permute(0, 3, 1, 2), (0, 1, 0, 1))
    y_ref = torch.nn.functional.conv2d(img, w4.float(), stride=2, padding=0)
    assert (Ho, Wo) == tuple(y_ref.shape[2:])
    assert _rel(y, y_ref.permute(0, 2, 3, 1).reshape(F_, Ho * Wo, Cout)) < 1e-2
    assert lib.gvd_im2col3x3_down_cl(x.data_ptr(), x.data_ptr(), 1, 1, 4, 16, None) == 2   # H < 2


@pytest.mark.parametrize("silu", [0, 2])
def test_groupnorm_bwd_split_at_its_reduction(monkeypatch, emu_lib, silu):
    """gvd_groupnorm_cl_bwd_sums / _apply (rows of a group sharded over GPUs) executed on the host: two row shards with
    their statistics and sums added up by hand reproduce the one-piece backward; an empty shard contributes zeros."""
    import gvd_native
    from vc_b200 import ops

    vp, ll, i32, f32 = C.c_void_p, C.c_longlong, C.c_int, C.c_float
    emu_lib.gvd_groupnorm_cl_bwd_sums.argtypes = [vp, vp, vp, vp, vp, vp, i32, ll, ll, i32, i32, f32, i32, vp, C.c_size_t, vp]
    emu_lib.gvd_groupnorm_cl_bwd_apply.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, ll, ll, i32, i32, f32, i32, vp]
    fake = install_fake(monkeypatch, BF)
    for name in ("gvd_groupnorm_bwd_tmp_bytes", "gvd_groupnorm_cl_bwd", "gvd_groupnorm_cl_bwd_sums", "gvd_groupnorm_cl_bwd_apply"):
        setattr(fake, name, getattr(emu_lib, name))
    monkeypatch.setattr(gvd_native, "nn", lambda: fake)
    F, S, Cc = 2, 90, 64
    x, dy = _bf(F, S, Cc, seed=1, scale=2.0) + 0.5, _bf(F, S, Cc, seed=2)
    g = torch.Generator().manual_seed(3)
    gamma, beta = 1 + 0.1 * torch.randn(Cc, generator=g), 0.1 * torch.randn(Cc, generator=g)
    whole = ops.groupnorm_bwd(x, dy, gamma, beta, F, S, 32, 1e-5, silu)

    cut = 37
    shards = [(x[:, :cut].contiguous(), dy[:, :cut].contiguous()), (x[:, cut:].contiguous(), dy[:, cut:].contiguous())]
    # FramePartition.sum_stats stand-in: call k of a backward is the k-th all-reduce (0: statistics, 1: backward sums).
    # Run the shards once to collect their local statistics, add them; again for the sums; a third time for dx.
    totals = {}

    class Collect:
        def __init__(self, sid):
            self.sid, self.calls = sid, 0

        def sum_stats(self, t):
            k = self.calls
            self.calls += 1
            if (k, "total") in totals:
                t.copy_(totals[(k, "total")])
            else:
                totals[(k, self.sid)] = t.clone()
            return t

    for sid, (xs, ds) in enumerate(shards):   # statistics of each shard
        ops.groupnorm_sharded_bwd(xs, ds, gamma, beta, F, xs.shape[1], S, Collect(sid), 32, 1e-5, silu)
    totals[(0, "total")] = totals[(0, 0)] + totals[(0, 1)]
    for k in [key for key in totals if key[0] == 1]:   # sums were formed with partial statistics: recompute
        del totals[k]
    for sid, (xs, ds) in enumerate(shards):
        ops.groupnorm_sharded_bwd(xs, ds, gamma, beta, F, xs.shape[1], S, Collect(sid), 32, 1e-5, silu)
    totals[(1, "total")] = totals[(1, 0)] + totals[(1, 1)]
    parts = [ops.groupnorm_sharded_bwd(xs, ds, gamma, beta, F, xs.shape[1], S, Collect(sid), 32, 1e-5, silu) for sid, (xs, ds) in enumerate(shards)]
    assert _rel(torch.cat(parts, dim=1), whole) < 4e-3
    # a shard without rows writes zero sums (what a rank with no pixel of a coarse level contributes)
    sums = torch.full((F * 32 * 2,), 5.0, dtype=torch.float64)
    stats = torch.zeros(F * 32 * 2)
    assert emu_lib.gvd_groupnorm_cl_bwd_sums(None, None, None, None, stats.data_ptr(), sums.data_ptr(), F, 0, S, Cc, 32, 1e-5, silu, None, 0, None) == 0
    assert float(sums.abs().max()) == 0.0
