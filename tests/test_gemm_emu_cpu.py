"""Runs the CUDA SOURCE of csrc/gemm_tc.cu -- the tcgen05 GEMM and the implicit-GEMM convolutions -- on the host
(tests/cuda_emu + tc_emu.h): the persistent one-CTA kernels of every tile width (64 / 128 / 256, double-buffered TMEM
accumulator, burst-fed operand ring, the eight-warp epilogue with its swizzled staging), the first-generation kernel for
odd shapes, the MN-major B mode, the fused GEGLU epilogue, and the convolutions whose A tiles TMA
fetches at shifted coordinates of a 4-D tensor map (zero padding = out-of-bounds fill; temporal taps as row shifts).  CTA
pairs are not emulated (one block runs at a time) and the column split needs >= 148 row tiles at K >= 2048: both stay
GPU-only tests.  Each case is evaluated once
over the emulated kernels and once over tests/fake_nn_lib.py's closed forms, and against torch."""
import ctypes as C
import os
import sys

import pytest
import torch
import torch.nn.functional as Fn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "cuda_emu"))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "guidedvd-3dgs_b200"))
sys.path.insert(0, HERE)

from test_unet_grad_cpu import _rel, install_fake  # noqa: E402

BF = torch.bfloat16
EMU = ("gvd_gemm_bf16", "gvd_conv_bf16", "gvd_conv_bf16_supported")


@pytest.fixture(scope="module")
def emu_lib():
    import build_emu
    import gvd_native

    L = C.CDLL(build_emu.build("gemm_tc"))
    L.gvd_gemm_bf16.argtypes = [C.POINTER(gvd_native.GemmArgs), C.c_void_p]
    L.gvd_conv_bf16.argtypes = [C.POINTER(gvd_native.ConvArgs), C.c_void_p]
    L.gvd_conv_bf16_supported.argtypes = [C.c_int] * 5
    L.gvd_nn_last_error.restype = C.c_char_p
    return L


def _both(monkeypatch, lib, fn):
    fake = install_fake(monkeypatch, BF)
    for name in EMU:
        setattr(fake, name, getattr(lib, name))
    fake.gvd_nn_last_error = lib.gvd_nn_last_error
    a = fn()
    install_fake(monkeypatch, BF)
    return a, fn()


def _bf(*shape, seed, scale=1.0):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale).to(BF)


@pytest.mark.parametrize("M,N,K", [(200, 320, 192), (130, 64, 64), (129, 512, 128), (300, 136, 72), (128, 640, 64)])
@pytest.mark.parametrize("epi", ["plain", "full"])
def test_gemm_tile_widths_and_epilogues(monkeypatch, emu_lib, M, N, K, epi):
    from vc_b200 import ops

    x, w = _bf(M, K, seed=1), _bf(N, K, seed=2, scale=K ** -0.5)
    g = torch.Generator().manual_seed(3)
    bias, bias2 = torch.randn(N, generator=g), torch.randn(N, generator=g)
    res = _bf(M, N, seed=4)
    if epi == "plain":
        fn = lambda: ops.linear(x, w, bias=bias)  # noqa: E731
        ref = x.float() @ w.float().T + bias
    else:
        fn = lambda: ops.linear(x, w, bias=bias, act="silu", residual=res, bias2=bias2, alpha=0.5)  # noqa: E731
        ref = Fn.silu(0.5 * (x.float() @ w.float().T) + bias) + bias2 + res.float()
    emu, closed = _both(monkeypatch, emu_lib, fn)
    assert _rel(emu, closed) < 4e-3 and _rel(emu, ref) < 8e-3


def test_gemm_odd_output_goes_through_the_first_generation_kernel(monkeypatch, emu_lib):
    from vc_b200 import ops

    x, w = _bf(70, 72, seed=5), _bf(70, 72, seed=6, scale=0.1)   # N = 70: not a multiple of 8
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.linear(x, w, out_dtype=torch.float32))
    assert emu.dtype == torch.float32 and _rel(emu, closed) < 1e-5


def test_gemm_batched_heads_and_mn_major_b(monkeypatch, emu_lib):
    """The two batch levels (attention heads / items addressed by strides) and the MN-major B operand of dK^T = Q^T dS."""
    from vc_b200 import ops

    H, nb, M, N, D = 2, 2, 136, 72, 64
    q, k = _bf(nb, M, H * D, seed=7), _bf(nb, N, H * D, seed=8)

    def scores():
        sim = torch.zeros(nb, H, M, N, dtype=BF)
        ops.gemm_raw(q, k, sim, M, N, D, H * D, H * D, N, batch_h=H, batch_b=nb, a_strides=(D, M * H * D), b_strides=(D, N * H * D),
                     c_strides=(M * N, H * M * N), alpha=0.125, act="round_scale")
        return sim
    emu, closed = _both(monkeypatch, emu_lib, scores)
    ref = torch.einsum("bihd,bjhd->bhij", q.float().view(nb, M, H, D), k.float().view(nb, N, H, D)) * 0.125
    assert _rel(emu, closed) < 4e-3 and _rel(emu, ref) < 8e-3
    a, b = _bf(64, 136, seed=9), _bf(136, 72, seed=10)   # C[64, 72] = A[64, 136] B[136, 72], B stored k-major rows of n

    def mn():
        c = torch.zeros(64, 72, dtype=BF)
        ops.gemm_raw(a, b, c, 64, 72, 136, 136, 72, 72, b_mn_major=True)
        return c
    emu, closed = _both(monkeypatch, emu_lib, mn)
    assert _rel(emu, closed) < 4e-3 and _rel(emu, a.float() @ b.float()) < 8e-3


def test_fused_geglu_epilogue(monkeypatch, emu_lib):
    from vc_b200 import ops

    x, w = _bf(150, 64, seed=11), _bf(128, 64, seed=12, scale=0.2)   # 2 D = 128 projection rows
    bias = torch.randn(128, generator=torch.Generator().manual_seed(13))
    w_il, b_il = ops.geglu_weight(w, bias)
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.linear_geglu(x, w_il, b_il))
    h = x.float() @ w.float().T + bias
    ref = h[:, :64] * Fn.gelu(h[:, 64:])
    assert _rel(emu, closed) < 8e-3 and _rel(emu, ref) < 1.5e-2


@pytest.mark.parametrize("F,H,W,Cin,Cout", [(2, 8, 16, 64, 72), (1, 3, 128, 64, 64), (2, 16, 8, 128, 320)])
def test_implicit_gemm_conv3x3(monkeypatch, emu_lib, F, H, W, Cin, Cout):
    from vc_b200 import ops

    x4 = _bf(F, Cin, H, W, seed=20)
    w4 = _bf(Cout, Cin, 3, 3, seed=21, scale=(9 * Cin) ** -0.5)
    bias = torch.randn(Cout, generator=torch.Generator().manual_seed(22))
    res = _bf(F, H * W, Cout, seed=23)
    x_cl = x4.permute(0, 2, 3, 1).reshape(F, H * W, Cin).contiguous()
    w_cl = w4.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.conv3x3(x_cl, F, H, W, w_cl, bias, residual=res)[0])
    ref = Fn.conv2d(x4.float(), w4.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(F, H * W, Cout) + res.float()
    assert _rel(emu, closed) < 4e-3 and _rel(emu, ref) < 8e-3


def test_implicit_gemm_temporal_conv(monkeypatch, emu_lib):
    from vc_b200 import ops

    B, T, S, Cin, Cout = 2, 5, 50, 64, 72
    x = _bf(B * T, S, Cin, seed=30)
    w5 = _bf(Cout, Cin, 3, seed=31, scale=(3 * Cin) ** -0.5)
    w = w5.permute(0, 2, 1).reshape(Cout, -1).contiguous()
    emu, closed = _both(monkeypatch, emu_lib, lambda: ops.conv_t3(x, B, T, S, w))
    vol = x.float().view(B, T, S, Cin).permute(0, 3, 1, 2)                    # [B, Cin, T, S]
    ref = Fn.conv2d(vol, w5.float().unsqueeze(-1), padding=(1, 0)).permute(0, 2, 3, 1).reshape(B * T, S, Cout)
    assert _rel(emu, closed) < 4e-3 and _rel(emu, ref) < 8e-3
