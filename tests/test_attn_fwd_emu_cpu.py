"""Runs the CUDA SOURCE of csrc/attn_tc.cu -- the tcgen05 flash-attention forward, generation 7 by default -- on the host
(tests/cuda_emu + tc_emu.h, see tests/test_attn_bwd_emu_cpu.py for what the emulation covers).  Executed for real: the Q /
K / V tensor-map loads with zero fill beyond the sequence, the S / P double buffers in tensor memory, the lazy in-TMEM
rescale of O, every parity expression, the single-use barrier that closes the accumulation, the row statistic the
adjoint needs.  Other generations run in processes of their own (GVD_FLASH is read once); generations 5 and 8 use named
barriers, which have no host stand-in."""
import ctypes as C
import os
import subprocess
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "cuda_emu"))

BF = torch.bfloat16
LOG2E = 1.4426950408889634


def _lib():
    import build_emu

    L = C.CDLL(build_emu.build("attn_tc"))
    vp, ll, i32, f32 = C.c_void_p, C.c_longlong, C.c_int, C.c_float
    L.gvd_flash_attention.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, ll, ll, f32, vp]
    L.gvd_flash_attention_lse.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, ll, ll, f32, vp]
    return L


def run_case(B, Nq, Nk, H, spread, climb, with_lse=True, seed=0):
    L = _lib()
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, Nq, H * 64, generator=g).to(BF)
    k = (torch.randn(B, Nk, H * 64, generator=g) * spread).to(BF)
    v = torch.randn(B, Nk, H * 64, generator=g).to(BF)
    if climb:  # logits that climb by ~35 across the keys: the running maximum outgrows its 2^8 slack, O is rescaled in TMEM
        q.view(B, Nq, H, 64)[..., 0] = 4.0
        k.view(B, Nk, H, 64)[..., 0] = (climb * 70.0 / Nk * torch.arange(Nk))[None, :, None].to(BF)
    scale = 0.125
    s = torch.einsum("bihd,bjhd->bhij", q.float().view(B, Nq, H, 64), k.float().view(B, Nk, H, 64)) * scale
    ref = torch.einsum("bhij,bjhd->bihd", torch.softmax(s, -1), v.float().view(B, Nk, H, 64)).reshape(B, Nq, H * 64)
    out = torch.full_like(q, float("nan"))
    ldl = (Nq + 127) // 128 * 128
    lse = torch.full((B, H, ldl), float("nan"))
    if with_lse:
        rc = L.gvd_flash_attention_lse(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(), B, Nq, Nk, H, Nq * H * 64,
                                       Nk * H * 64, scale, None)
    else:
        rc = L.gvd_flash_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, Nq, Nk, H, Nq * H * 64, Nk * H * 64, scale, None)
    assert rc == 0
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1.5e-2, err
    if with_lse:
        want = torch.logsumexp(s, -1) * LOG2E
        assert (lse[:, :, :Nq] - want).abs().max().item() < 2e-3
        assert torch.isfinite(lse).all()  # the rows beyond Nq hold finite filler
    return out


CASES = [(1, 200, 150, 1, 1.0, 0.0), (2, 128, 64, 2, 3.0, 0.0), (1, 130, 77, 1, 1.0, 0.0), (1, 64, 300, 1, 2.0, 1.0), (1, 64, 300, 1, 1.0, -1.0),
         (1, 257, 1, 1, 1.0, 0.0), (1, 100, 129, 2, 1.0, 0.0)]


@pytest.mark.parametrize("B,Nq,Nk,H,spread,climb", CASES)
def test_flash_attention_generation7_on_the_host(B, Nq, Nk, H, spread, climb):
    a = run_case(B, Nq, Nk, H, spread, climb, with_lse=True)
    b = run_case(B, Nq, Nk, H, spread, climb, with_lse=False)
    assert torch.equal(a, b)  # writing the statistic does not change the output


@pytest.mark.parametrize("variant", ["v1", "v2", "v6"])
def test_selectable_generations_on_the_host(variant):
    code = ("import sys; sys.path.insert(0, %r); import test_attn_fwd_emu_cpu as t\n"
            "for c in t.CASES[:5]: t.run_case(*c, with_lse=False)\nprint('ok')" % HERE)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GVD_FLASH=variant), capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
