"""Runs the CUDA SOURCE of csrc/attn_bwd_tc.cu -- the fused tcgen05 attention adjoint -- on the host (tests/cuda_emu +
tc_emu.h: mbarrier phases and transaction counts, tensor-map box loads with the 128-byte swizzle and zero fill, tensor
memory, tcgen05.mma with both operands in shared memory or A in tensor memory, tcgen05.ld / st).  What executes for real:
the descriptor start addresses and their k / sub-block offsets (K-major and MN-major), the TMEM column plan of both forms
including P written over S and dS over dP in place, the packed-bf16 A operands, the per-column statistics staged with the
tile, the ragged-column mask, every parity expression, the epilogue.  What does not: asynchrony (an MMA or a copy is
complete when the call returns) -- the orderings only that can break are tests/test_barrier_protocol_cpu.py's business.
Compared with autograd over fp32 attention, as tests/test_zz_guided_gpu.py does on the GPU."""
import ctypes as C
import os
import subprocess
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "cuda_emu"))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "guidedvd-3dgs_b200"))

BF = torch.bfloat16
LOG2E = 1.4426950408889634


def _lib():
    import build_emu
    import gvd_native

    L = C.CDLL(build_emu.build("attn_bwd_tc"))
    L.gvd_flash_attention_bwd.argtypes = [C.POINTER(gvd_native.FlashBwdArgs), C.c_void_p]
    return L, gvd_native


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def run_case(B, Nq, Nk, H, need_kv, spread=1.0, seed=0):
    L, gvd_native = _lib()
    g = torch.Generator().manual_seed(seed)
    q, do = (torch.randn(B, Nq, H * 64, generator=g).to(BF) for _ in range(2))
    k = (torch.randn(B, Nk, H * 64, generator=g) * spread).to(BF)
    v = torch.randn(B, Nk, H * 64, generator=g).to(BF)
    scale = 0.125
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    s = torch.einsum("bihd,bjhd->bhij", qf.view(B, Nq, H, 64), kf.view(B, Nk, H, 64)) * scale
    o = torch.einsum("bhij,bjhd->bihd", torch.softmax(s, -1), vf.view(B, Nk, H, 64)).reshape(B, Nq, H * 64)
    o.backward(do.float())
    out = o.detach().to(BF)
    ldl = (Nq + 127) // 128 * 128
    lse = torch.zeros(B, H, ldl)
    lse[:, :, :Nq] = torch.logsumexp(s.detach(), -1) * LOG2E
    delta = torch.full((B, H, ldl), float("nan"))
    dq = torch.full_like(q, float("nan"))
    dk, dv = (torch.full_like(k, float("nan")) for _ in range(2))
    a = gvd_native.FlashBwdArgs(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), delta.data_ptr(),
                                dq.data_ptr(), dk.data_ptr() if need_kv else None, dv.data_ptr() if need_kv else None, B, Nq, Nk, H,
                                Nq * H * 64, Nk * H * 64, scale)
    assert L.gvd_flash_attention_bwd(C.byref(a), None) == 0
    assert torch.isfinite(dq.float()).all() and _rel(dq, qf.grad) < 2e-2, _rel(dq, qf.grad)
    d_ref = (do.float() * out.float()).view(B, Nq, H, 64).sum(-1).permute(0, 2, 1)
    assert torch.allclose(delta[:, :, :Nq], d_ref, rtol=1e-5, atol=1e-5) and (delta[:, :, Nq:] == 0).all()
    if need_kv:
        assert torch.isfinite(dk.float()).all() and torch.isfinite(dv.float()).all()
        assert _rel(dk, kf.grad) < 2e-2 and _rel(dv, vf.grad) < 2e-2, (_rel(dk, kf.grad), _rel(dv, vf.grad))


CASES = [(1, 200, 150, 1, True, 1.0), (2, 128, 64, 2, True, 3.0), (1, 130, 77, 1, False, 1.0), (1, 64, 300, 1, True, 2.0), (1, 257, 1, 1, True, 1.0)]


@pytest.mark.parametrize("B,Nq,Nk,H,need_kv,spread", CASES)
def test_fused_attention_adjoint_on_the_host_two_ctas_per_sm_form(B, Nq, Nk, H, need_kv, spread):
    if Nk == 1:
        pytest.skip("dQ = dK = 0 exactly: covered on the GPU with absolute bounds")
    run_case(B, Nq, Nk, H, need_kv, spread)


def test_fused_attention_adjoint_on_the_host_first_form():
    """GVD_FLASH_BWD_CTAS is read once per process: the one-CTA-per-SM form (512 tensor-memory columns, double-buffered
    sub-blocks, three-stage ring) runs in a process of its own."""
    code = ("import sys; sys.path.insert(0, %r); import test_attn_bwd_emu_cpu as t\n"
            "for c in t.CASES[:4]: t.run_case(*c)\nprint('ok')" % HERE)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GVD_FLASH_BWD_CTAS="1"), capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_forward_statistic_feeds_the_adjoint_on_the_host():
    """The two kernels together, as vc_b200.grad.FlashAttention chains them: out and lse from the emulated forward
    (attn_tc.cu, generation 7) go into the emulated adjoint; gradients against autograd over fp32 attention."""
    import build_emu

    Lb, gvd_native = _lib()
    Lf = C.CDLL(build_emu.build("attn_tc"))
    vp, ll, i32, f32 = C.c_void_p, C.c_longlong, C.c_int, C.c_float
    Lf.gvd_flash_attention_lse.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, ll, ll, f32, vp]
    B, Nq, Nk, H, scale = 1, 190, 140, 2, 0.125
    g = torch.Generator().manual_seed(5)
    q, do = (torch.randn(B, Nq, H * 64, generator=g).to(BF) for _ in range(2))
    k = (torch.randn(B, Nk, H * 64, generator=g) * 2.5).to(BF)
    v = torch.randn(B, Nk, H * 64, generator=g).to(BF)
    out = torch.empty_like(q)
    ldl = (Nq + 127) // 128 * 128
    lse, delta = torch.empty(B, H, ldl), torch.empty(B, H, ldl)
    assert Lf.gvd_flash_attention_lse(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(), B, Nq, Nk, H, Nq * H * 64,
                                      Nk * H * 64, scale, None) == 0
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    a = gvd_native.FlashBwdArgs(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), delta.data_ptr(),
                                dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, Nq, Nk, H, Nq * H * 64, Nk * H * 64, scale)
    assert Lb.gvd_flash_attention_bwd(C.byref(a), None) == 0
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    s = torch.einsum("bihd,bjhd->bhij", qf.view(B, Nq, H, 64), kf.view(B, Nk, H, 64)) * scale
    torch.einsum("bhij,bjhd->bihd", torch.softmax(s, -1), vf.view(B, Nk, H, 64)).reshape(B, Nq, H * 64).backward(do.float())
    assert _rel(dq, qf.grad) < 2e-2 and _rel(dk, kf.grad) < 2e-2 and _rel(dv, vf.grad) < 2e-2
