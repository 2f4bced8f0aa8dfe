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


ASYNC_CODE = r'''
import ctypes as C, sys, torch
sys.path.insert(0, %(emu)r)
import build_emu
old = %(old)d
L = C.CDLL(build_emu.build("attn_tc_oldwait", ["attn_tc.cu"], extra_flags=("-DGVD_EMU_OLD_EPILOGUE_WAIT",)) if old else build_emu.build("attn_tc"))
vp, ll, i32, f32 = C.c_void_p, C.c_longlong, C.c_int, C.c_float
L.gvd_flash_attention.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, ll, ll, f32, vp]
bad = 0
for seed in range(10):
    g = torch.Generator().manual_seed(seed)
    q, k, v = (torch.randn(1, n, 64, generator=g).to(torch.bfloat16) for n in (128, 256, 256))
    out = torch.zeros_like(q)
    assert L.gvd_flash_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), 1, 128, 256, 1, 128 * 64, 256 * 64, 0.125, None) == 0
    ref = torch.softmax((q.float()[0] @ k.float()[0].T) * 0.125, -1) @ v.float()[0]
    bad += int((out.float()[0] - ref).abs().max().item() / ref.abs().max().item() > 1.5e-2)
print("wrong", bad)
'''


def test_asynchronous_engine_on_the_real_source_and_the_wait_it_used_to_have():
    """tc_emu.h with GVD_EMU_ASYNC: MMAs and commits execute from an in-order queue, copies from an unordered one, whenever a
    spinning waiter lets them (here: 3 % of the spins).  The kernel as it is: every output right.  The same source compiled
    with its FORMER epilogue wait (-DGVD_EMU_OLD_EPILOGUE_WAIT: parity wait on the per-sub-block barrier): the accumulator is
    read while PVs are still queued -- what compute-sanitizer's timing showed once on the GPU shows here at will."""
    emu = os.path.join(HERE, "cuda_emu")
    res = {}
    for old in (0, 1):
        r = subprocess.run([sys.executable, "-c", ASYNC_CODE % {"emu": emu, "old": old}], env=dict(os.environ, GVD_EMU_ASYNC="97"),
                           capture_output=True, text=True, timeout=1500, cwd=os.path.dirname(HERE))
        assert r.returncode == 0 and "wrong" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
        res[old] = int(r.stdout.strip().split()[-1])
    assert res[0] == 0 and res[1] > 0, res


@pytest.mark.parametrize("lag", ["90"])
def test_tensor_core_kernels_under_the_asynchronous_engine(lag):
    """The forward, both forms of the adjoint and the GEMM / convolution cases again, with deferred MMA execution."""
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", os.path.join(HERE, "test_attn_fwd_emu_cpu.py"),
                        os.path.join(HERE, "test_attn_bwd_emu_cpu.py"), os.path.join(HERE, "test_gemm_emu_cpu.py"), "-k",
                        "generation7 or two_ctas or feeds or gemm_tile or implicit or geglu or batched"],
                       env=dict(os.environ, GVD_EMU_ASYNC=lag), capture_output=True, text=True, timeout=2400)
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
