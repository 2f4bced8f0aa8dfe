"""Operator-level Python bindings of include/gvd_nn.h (torch is used only for memory and streams)."""
import ctypes as C
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import gvd_native as _n  # noqa: E402

ACT = {"none": 0, "silu": 1, "gelu": 2}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(rc, lib, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed: " + (lib.gvd_nn_last_error() or b"").decode())


def gemm_raw(A, B, Cout, M, N, K, lda, ldb, ldc, batch_h=1, batch_b=1, a_strides=(0, 0), b_strides=(0, 0),
             c_strides=(0, 0), bias=None, residual=None, alpha=1.0, act="none"):
    """C[b,h,m,n] = act(alpha * sum_k A[b,h,m,k] B[b,h,n,k] + bias[n]) + residual; strides in elements."""
    lib = _n.nn()
    a = _n.GemmArgs()
    a.M, a.N, a.K, a.batch_h, a.batch_b = int(M), int(N), int(K), int(batch_h), int(batch_b)
    a.A, a.lda, a.a_stride_h, a.a_stride_b = A.data_ptr(), int(lda), int(a_strides[0]), int(a_strides[1])
    a.B, a.ldb, a.b_stride_h, a.b_stride_b = B.data_ptr(), int(ldb), int(b_strides[0]), int(b_strides[1])
    a.C, a.ldc, a.c_stride_h, a.c_stride_b = Cout.data_ptr(), int(ldc), int(c_strides[0]), int(c_strides[1])
    a.bias = bias.data_ptr() if bias is not None else None
    a.residual = residual.data_ptr() if residual is not None else None
    a.alpha, a.act, a.out_fp32 = float(alpha), ACT[act], int(Cout.dtype == torch.float32)
    with torch.cuda.device(A.device):
        _check(lib.gvd_gemm_bf16(C.byref(a), _stream()), lib, "gvd_gemm_bf16")
    return Cout


def linear(x, weight, bias=None, act="none", residual=None, out_dtype=torch.bfloat16, alpha=1.0):
    """y = act(x @ weight^T + bias) + residual.  x [..., K] bf16 contiguous, weight [N, K] bf16, bias fp32 [N]."""
    K = x.shape[-1]
    N = weight.shape[0]
    x2 = x.reshape(-1, K)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    M = x2.shape[0]
    out = torch.empty(M, N, dtype=out_dtype, device=x.device)
    res2 = residual.reshape(M, N) if residual is not None else None
    gemm_raw(x2, weight, out, M, N, K, K, K, N, bias=bias, residual=res2, alpha=alpha, act=act)
    return out.reshape(*x.shape[:-1], N)
