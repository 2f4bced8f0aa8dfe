"""Operator-level Python bindings of include/gvd_nn.h.

torch is used for device memory, streams and pure layout plumbing (views, one head-transpose of V); every
arithmetic operator of the denoiser goes through the sm_100a library.  There is no fallback path: a missing
library raises in gvd_native.nn().

Activation layout: channels-last bf16, x[F, S, C] with F = batch*frames, S = h*w pixels, C channels (the K-major
operand layout of the tensor-core GEMM).
"""
import ctypes as C
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import gvd_native as _n  # noqa: E402

ACT = {"none": 0, "silu": 1, "gelu": 2, "round_scale": 3, "geglu": 4}
BF16 = torch.bfloat16


def _wants_grad(*tensors):
    """True when a call must be recorded for the guided sampler's backward (vc_b200.grad): autograd is on and one of
    the activations carries a graph.  Parameters never require grad here, and every U-Net call of the plain sampler
    runs under torch.no_grad(), so the inference path never takes this branch."""
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _grad():
    from . import grad
    return grad


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """Current stream of the current device as a cudaStream_t (the raw-handle query: torch.cuda.current_stream() costs
    ~26 us per call, and an eager forward makes ~1 300 launches)."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch._C._cuda_getDevice()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _on_device:
    """`with _on_device(dev)` only when dev is not already current (the guard costs ~10 us per use, and the eager
    guided step makes ~2 000 launches)."""
    __slots__ = ("guard",)

    def __init__(self, dev):
        self.guard = None
        if dev.type == "cuda" and dev.index is not None and torch.cuda.current_device() != dev.index:
            self.guard = torch.cuda.device(dev)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *exc):
        if self.guard is not None:
            self.guard.__exit__(*exc)
        return False


def _check(rc, lib, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed: " + (lib.gvd_nn_last_error() or b"").decode())


def _p(t):
    return None if t is None else t.data_ptr()


def gemm_raw(A, B, Cout, M, N, K, lda, ldb, ldc, batch_h=1, batch_b=1, a_strides=(0, 0), b_strides=(0, 0),
             c_strides=(0, 0), bias=None, bias2=None, residual=None, alpha=1.0, act="none", b_mn_major=False):
    """C[b,h,m,n] = act(alpha * sum_k A[b,h,m,k] B[b,h,n,k] + bias[n]) (+ bias2[n]) + residual; strides in elements.
    b_mn_major: B is stored B[b,h,k,n] (n contiguous, ldb between consecutive k) -- no transposed copy needed."""
    lib = _n.nn()
    # positional construction in field order (include/gvd_nn.h::GvdGemmArgs): one C-level call instead of 24 attribute
    # stores -- the eager guided step makes ~2 400 GEMM calls
    a = _n.GemmArgs(int(M), int(N), int(K), int(batch_h), int(batch_b),
                    A.data_ptr(), int(lda), int(a_strides[0]), int(a_strides[1]),
                    B.data_ptr(), int(ldb), int(b_strides[0]), int(b_strides[1]),
                    Cout.data_ptr(), int(ldc), int(c_strides[0]), int(c_strides[1]),
                    _p(bias), _p(bias2), _p(residual), float(alpha), ACT[act], int(Cout.dtype == torch.float32),
                    int(bool(b_mn_major)))
    with _on_device(A.device):
        _check(lib.gvd_gemm_bf16(C.byref(a), _stream()), lib, "gvd_gemm_bf16")
    return Cout


def linear(x, weight, bias=None, act="none", residual=None, out_dtype=None, alpha=1.0, bias2=None):
    """y = act(x @ weight^T + bias) (+ bias2) + residual.  x [..., K] bf16, weight [N, K] bf16, biases fp32 [N]."""
    if _wants_grad(x, residual):
        return _grad().Linear.apply(x, weight, bias, act, residual, out_dtype, alpha, bias2)
    out_dtype = out_dtype or BF16
    K = x.shape[-1]
    N = weight.shape[0]
    x2 = x.reshape(-1, K)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    M = x2.shape[0]
    out = torch.empty(M, N, dtype=out_dtype, device=x.device)
    res2 = residual.reshape(M, N) if residual is not None else None
    gemm_raw(x2, weight, out, M, N, K, K, K, N, bias=bias, bias2=bias2, residual=res2, alpha=alpha, act=act)
    return out.reshape(*x.shape[:-1], N)


_gn_tmp = {}


def groupnorm(x, gamma, beta, F, S, groups=32, eps=1e-5, silu=False):
    """x viewed as [F, S, C] channels-last bf16; statistics over S x (C/groups)."""
    if _wants_grad(x):
        return _grad().GroupNorm.apply(x, gamma, beta, F, S, groups, eps, int(silu))
    lib = _n.nn()
    Cc = x.shape[-1]
    x = x.contiguous()
    y = torch.empty_like(x)
    nfl = int(lib.gvd_groupnorm_tmp_floats(int(F), int(S), int(groups)))
    key = (x.device, nfl)
    tmp = _gn_tmp.get(key)
    if tmp is None:
        tmp = _gn_tmp[key] = torch.empty(nfl, dtype=torch.float32, device=x.device)
    _check(lib.gvd_groupnorm_cl(x.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), int(F), int(S), int(Cc),
                                int(groups), float(eps), int(silu), tmp.data_ptr(), nfl, _stream()), lib, "gvd_groupnorm_cl")
    return y


def groupnorm_with_stats(x, gamma, beta, F, S, groups=32, eps=1e-5, silu=False):
    """`groupnorm` that also returns stats[F, groups, 2] = (sum x, sum x^2) (gvd_groupnorm_cl_keep_stats: two launches, the
    same bits as the split entry points), so the guided sampler's backward does not have to read x a second time for the
    statistics (vc_b200.grad.GroupNorm)."""
    lib = _n.nn()
    Cc = x.shape[-1]
    x = x.contiguous()
    y = torch.empty_like(x)
    nfl = int(lib.gvd_groupnorm_tmp_floats(int(F), int(S), int(groups)))
    key = (x.device, nfl)
    tmp = _gn_tmp.get(key)
    if tmp is None:
        tmp = _gn_tmp[key] = torch.empty(nfl, dtype=torch.float32, device=x.device)
    stats = torch.empty(F * groups * 2, dtype=torch.float32, device=x.device)
    _check(lib.gvd_groupnorm_cl_keep_stats(x.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), stats.data_ptr(), int(F), int(S),
                                           int(Cc), int(groups), float(eps), int(silu), tmp.data_ptr(), nfl, _stream()), lib,
           "gvd_groupnorm_cl_keep_stats")
    return y, stats


def groupnorm_sharded(x, gamma, beta, F, S_local, S_total, part, groups=32, eps=1e-5, silu=False):
    """GroupNorm whose rows are spread over the ranks of `part` (vc_b200.frame_parallel.FramePartition): local
    (sum, sumsq) -> one all-reduce of F*groups*2 floats -> local normalisation with the global statistics."""
    if _wants_grad(x):
        return _grad().GroupNormSharded.apply(x, gamma, beta, F, S_local, S_total, part, groups, eps, int(silu))
    lib = _n.nn()
    Cc = x.shape[-1]
    x = x.contiguous()
    y = torch.empty_like(x)
    nfl = int(lib.gvd_groupnorm_tmp_floats(int(F), int(max(S_local, 1)), int(groups)))
    key = (x.device, nfl)
    tmp = _gn_tmp.get(key)
    if tmp is None:
        tmp = _gn_tmp[key] = torch.empty(nfl, dtype=torch.float32, device=x.device)
    stats = torch.empty(F * groups * 2, dtype=torch.float32, device=x.device)
    _check(lib.gvd_groupnorm_cl_stats(x.data_ptr(), stats.data_ptr(), int(F), int(S_local), int(Cc), int(groups), tmp.data_ptr(),
                                      nfl, _stream()), lib, "gvd_groupnorm_cl_stats")
    part.sum_stats(stats)
    _check(lib.gvd_groupnorm_cl_apply(x.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), stats.data_ptr(), int(F),
                                      int(S_local), int(S_total), int(Cc), int(groups), float(eps), int(silu), _stream()), lib,
           "gvd_groupnorm_cl_apply")
    return y


def layernorm(x, gamma, beta, eps=1e-5):
    if _wants_grad(x):
        return _grad().LayerNorm.apply(x, gamma, beta, eps)
    lib = _n.nn()
    x = x.contiguous()
    y = torch.empty_like(x)
    Cc = x.shape[-1]
    _check(lib.gvd_layernorm(x.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), x.numel() // Cc, int(Cc),
                             float(eps), _stream()), lib, "gvd_layernorm")
    return y


FUSED_GEGLU = os.environ.get("GVD_FUSED_GEGLU", "1") != "0"  # A/B switch: 0 = projection GEMM, then the GEGLU kernel


def geglu_weight(weight, bias):
    """GEGLU projection `proj` [2D, K] (rows 0..D-1 values, D..2D-1 gates; attention.py:415-423) -> the row order the fused
    epilogue wants: blocks of 32 rows = 16 value rows followed by the 16 gate rows of the same outputs (GVD_ACT_GEGLU).
    Returns (weight', bias') or None when D is not a multiple of 16."""
    N, K = weight.shape
    D = N // 2
    if N % 32 or D % 16:
        return None
    w = torch.stack([weight[:D].view(D // 16, 16, K), weight[D:].view(D // 16, 16, K)], dim=1).reshape(N, K).contiguous()
    b = None
    if bias is not None:
        b = torch.stack([bias[:D].view(D // 16, 16), bias[D:].view(D // 16, 16)], dim=1).reshape(N).contiguous()
    return w, b


def linear_geglu(x, weight_il, bias_il):
    """GEGLU(x) = (x W_v^T + b_v) * gelu(x W_g^T + b_g) as ONE tensor-core GEMM whose epilogue applies the gate: the 2D-wide
    projection never reaches memory.  weight_il / bias_il from `geglu_weight`.  Inference only (the guided sampler's tape
    keeps the projection for its backward and goes through `linear` + `geglu`)."""
    K = x.shape[-1]
    N = weight_il.shape[0]
    x2 = x.reshape(-1, K)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    M = x2.shape[0]
    out = torch.empty(M, N // 2, dtype=BF16, device=x.device)
    gemm_raw(x2, weight_il, out, M, N, K, K, K, N // 2, bias=bias_il, act="geglu")
    return out.reshape(*x.shape[:-1], N // 2)


def geglu(h):
    if _wants_grad(h):
        return _grad().Geglu.apply(h)
    lib = _n.nn()
    D = h.shape[-1] // 2
    h = h.contiguous()
    out = torch.empty(*h.shape[:-1], D, dtype=BF16, device=h.device)
    _check(lib.gvd_geglu(h.data_ptr(), out.data_ptr(), h.numel() // (2 * D), int(D), _stream()), lib, "gvd_geglu")
    return out


def softmax_rows(scores, cols, ldy):
    """scores [rows, ld] (bf16 or fp32, contiguous) -> bf16 probabilities [rows, ldy] (columns >= cols zeroed)."""
    lib = _n.nn()
    rows = scores.numel() // scores.shape[-1]
    out = torch.empty(*scores.shape[:-1], ldy, dtype=BF16, device=scores.device)
    _check(lib.gvd_softmax_rows(scores.data_ptr(), int(scores.dtype == BF16), int(scores.shape[-1]), out.data_ptr(), int(ldy),
                                int(rows), int(cols), _stream()), lib, "gvd_softmax_rows")
    return out


IMPLICIT_CONV = os.environ.get("GVD_IMPLICIT_CONV", "1") != "0"  # A/B switch: 0 = every convolution through im2col


def _conv_implicit(kind, x, weight, geom, bias=None, bias2=None, residual=None, act="none"):
    """gvd_conv_bf16: implicit-GEMM convolution (no im2col buffer).  geom = (F, H, W) for kind 1, (B, T, S) for kind 2."""
    lib = _n.nn()
    a = _n.ConvArgs()
    a.kind = kind
    if kind == 1:
        a.F, a.H, a.W = (int(v) for v in geom)
        rows = a.F * a.H * a.W
    else:
        a.B, a.T, a.S = (int(v) for v in geom)
        rows = a.B * a.T * a.S
    a.Cin, a.Cout = int(x.shape[-1]), int(weight.shape[0])
    x = x if x.is_contiguous() else x.contiguous()
    y = torch.empty(rows, a.Cout, dtype=BF16, device=x.device)
    if residual is not None and not residual.is_contiguous():
        residual = residual.contiguous()
    a.x, a.weight, a.y = x.data_ptr(), weight.data_ptr(), y.data_ptr()
    a.bias, a.bias2, a.residual, a.act = _p(bias), _p(bias2), _p(residual), ACT[act]
    with _on_device(x.device):
        _check(lib.gvd_conv_bf16(C.byref(a), _stream()), lib, "gvd_conv_bf16")
    return y


def _implicit_ok(kind, H, W, Cin, Cout, x):
    return IMPLICIT_CONV and x.dtype == BF16 and bool(_n.nn().gvd_conv_bf16_supported(int(kind), int(H), int(W), int(Cin), int(Cout)))


def conv3x3(x, F, H, W, weight, bias=None, stride=1, upsample=False, bias2=None, residual=None, act="none"):
    """3x3 / pad 1 convolution of channels-last x[F, H*W, Cin] with weight [Cout, 9*Cin] (K order ky, kx, cin).
    Stride 1 without upsampling runs as an implicit GEMM (TMA fetches the shifted activation tiles); the rest builds
    the im2col matrix first."""
    Cin = x.shape[-1]
    Hin, Win = (2 * H, 2 * W) if upsample else (H, W)
    Ho, Wo = (Hin + 2 - 3) // stride + 1, (Win + 2 - 3) // stride + 1
    if _wants_grad(x, residual):
        return _grad().Conv3x3.apply(x, F, H, W, weight, bias, stride, bool(upsample), bias2, residual, act), Ho, Wo
    lib = _n.nn()
    if stride == 1 and not upsample and _implicit_ok(1, H, W, Cin, weight.shape[0], x):
        y = _conv_implicit(1, x, weight, (F, H, W), bias, bias2, residual, act)
        return y.view(F, Ho * Wo, -1), Ho, Wo
    if stride == 1 and upsample and _implicit_ok(1, 2 * H, 2 * W, Cin, weight.shape[0], x):
        # Upsample + conv: the 4x tensor is materialised (4 units written) and convolved implicitly, instead of a 36x im2col matrix
        y = _conv_implicit(1, upsample2x(x, F, H, W), weight, (F, 2 * H, 2 * W), bias, bias2, residual, act)
        return y.view(F, Ho * Wo, -1), Ho, Wo
    col = torch.empty(F * Ho * Wo, 9 * Cin, dtype=BF16, device=x.device)
    _check(lib.gvd_im2col3x3_cl(x.data_ptr(), col.data_ptr(), int(F), int(H), int(W), int(Cin), int(stride), int(upsample),
                                _stream()), lib, "gvd_im2col3x3_cl")
    y = linear(col, weight, bias=bias, bias2=bias2, residual=residual, act=act)
    return y.view(F, Ho * Wo, -1), Ho, Wo


def upsample2x(x, F, H, W):
    """Nearest-neighbour 2x upsampling of channels-last x[F, H*W, C] -> [F, 4*H*W, C] (gvd_upsample2x_cl)."""
    lib = _n.nn()
    Cc = x.shape[-1]
    x = x if x.is_contiguous() else x.contiguous()
    y = torch.empty(F, 4 * H * W, Cc, dtype=x.dtype, device=x.device)
    with _on_device(x.device):
        _check(lib.gvd_upsample2x_cl(x.data_ptr(), y.data_ptr(), int(F), int(H), int(W), int(Cc), _stream()), lib, "gvd_upsample2x_cl")
    return y


def upsample2x_bwd(dy, F, H, W):
    """Adjoint of `upsample2x`: dy[F, 4*H*W, C] -> dx[F, H*W, C], each the sum of its 2 x 2 block (gvd_upsample2x_bwd_cl)."""
    lib = _n.nn()
    Cc = dy.shape[-1]
    dy = dy if dy.is_contiguous() else dy.contiguous()
    dx = torch.empty(F, H * W, Cc, dtype=dy.dtype, device=dy.device)
    with _on_device(dy.device):
        _check(lib.gvd_upsample2x_bwd_cl(dy.data_ptr(), dx.data_ptr(), int(F), int(H), int(W), int(Cc), _stream()), lib,
               "gvd_upsample2x_bwd_cl")
    return dx


def conv3x3_down(x, F, H, W, weight, bias=None):
    """The VAE encoder's Downsample (ae_modules.py:93-106): pad right/bottom by one, 3x3 stride 2, no padding.
    x[F, H*W, Cin] -> (y[F, Ho*Wo, Cout], Ho, Wo).  Inference only (the encoder is never differentiated)."""
    if _wants_grad(x):
        raise RuntimeError("conv3x3_down: the VAE encoder is inference-only (no backward)")
    lib = _n.nn()
    Cin = x.shape[-1]
    Ho, Wo = (H - 2) // 2 + 1, (W - 2) // 2 + 1
    col = torch.empty(F * Ho * Wo, 9 * Cin, dtype=BF16, device=x.device)
    _check(lib.gvd_im2col3x3_down_cl(x.data_ptr(), col.data_ptr(), int(F), int(H), int(W), int(Cin), _stream()), lib,
           "gvd_im2col3x3_down_cl")
    return linear(col, weight, bias=bias).view(F, Ho * Wo, -1), Ho, Wo


def conv_t3(x, B, T, S, weight, bias=None, residual=None):
    """(3,1,1) / pad (1,0,0) temporal convolution of x[B*T, S, C] with weight [Cout, 3*Cin] (K order kt, cin)."""
    if _wants_grad(x, residual):
        return _grad().ConvT3.apply(x, B, T, S, weight, bias, residual)
    lib = _n.nn()
    Cin = x.shape[-1]
    if _implicit_ok(2, 0, 0, Cin, weight.shape[0], x):
        return _conv_implicit(2, x, weight, (B, T, S), bias, None, residual).view(B * T, S, weight.shape[0])
    col = torch.empty(B * T * S, 3 * Cin, dtype=BF16, device=x.device)
    _check(lib.gvd_im2col_t3_cl(x.data_ptr(), col.data_ptr(), int(B), int(T), int(S), int(Cin), _stream()), lib,
           "gvd_im2col_t3_cl")
    return linear(col, weight, bias=bias, residual=residual).view(B * T, S, weight.shape[0])


def temporal_attention(q, k, v, B, T, S, H, scale):
    if _wants_grad(q, k, v):
        return _grad().TemporalAttention.apply(q, k, v, B, T, S, H, scale)
    lib = _n.nn()
    out = torch.empty_like(q)
    _check(lib.gvd_temporal_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), int(B), int(T), int(S), int(H),
                                      float(scale), _stream()), lib, "gvd_temporal_attention")
    return out


def flash_attention(q, k, v, Bq, Nq, Nk, H, scale, shared_kv=False):
    """Fused softmax(q k^T * scale) v (gvd_flash_attention). Same arguments/layouts as `attention` below."""
    if _wants_grad(q, k, v):
        return _grad().FlashAttention.apply(q, k, v, Bq, Nq, Nk, H, scale, bool(shared_kv))
    lib = _n.nn()
    HD = H * 64
    out = torch.empty(Bq, Nq, HD, dtype=BF16, device=q.device)
    if shared_kv:
        B, nq, qs, ks = 1, Bq * Nq, Bq * Nq * HD, Nk * HD
    else:
        B, nq, qs, ks = Bq, Nq, Nq * HD, Nk * HD
    with _on_device(q.device):
        _check(lib.gvd_flash_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), int(B), int(nq), int(Nk), int(H),
                                       int(qs), int(ks), float(scale), _stream()), lib, "gvd_flash_attention")
    return out


# The guided sampler's attention adjoint: fused on chip (csrc/attn_bwd_tc.cu) unless GVD_FLASH_BWD=0, which restores the
# first backward (`attention_bwd`: scores, probabilities and their gradients materialised, five GEMM launches).
FUSED_FLASH_BWD = os.environ.get("GVD_FLASH_BWD", "1") != "0"


def _flash_geometry(Bq, Nq, Nk, H, shared_kv):
    HD = H * 64
    if shared_kv:
        return 1, Bq * Nq, Bq * Nq * HD, Nk * HD
    return Bq, Nq, Nq * HD, Nk * HD


def flash_attention_lse(q, k, v, Bq, Nq, Nk, H, scale, shared_kv=False):
    """`flash_attention` that also returns the per-row statistic of its backward (gvd_flash_attention_lse):
    (out, lse) with lse fp32 [B, H, rows rounded up to 128] = log2 sum_j exp2(s_ij * scale * log2 e)."""
    lib = _n.nn()
    B, nq, qs, ks = _flash_geometry(Bq, Nq, Nk, H, shared_kv)
    out = torch.empty(Bq, Nq, H * 64, dtype=BF16, device=q.device)
    lse = torch.empty(B, H, (nq + 127) // 128 * 128, dtype=torch.float32, device=q.device)
    with _on_device(q.device):
        _check(lib.gvd_flash_attention_lse(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(), int(B), int(nq),
                                           int(Nk), int(H), int(qs), int(ks), float(scale), _stream()), lib, "gvd_flash_attention_lse")
    return out, lse


def flash_attention_bwd(q, k, v, out, lse, dout, Bq, Nq, Nk, H, scale, shared_kv=False, need_kv=True):
    """(dq, dk, dv) of `flash_attention` from the forward's (out, lse), no score-sized intermediates
    (gvd_flash_attention_bwd); dk = dv = None when need_kv is False."""
    lib = _n.nn()
    if shared_kv and need_kv:
        raise NotImplementedError("flash_attention_bwd: shared keys/values are frozen-context projections (no dk/dv)")
    B, nq, qs, ks = _flash_geometry(Bq, Nq, Nk, H, shared_kv)
    q, k, v, out, dout = q.contiguous(), k.contiguous(), v.contiguous(), out.contiguous(), dout.contiguous()
    dq = torch.empty_like(q)
    dk, dv = (torch.empty_like(k), torch.empty_like(v)) if need_kv else (None, None)
    delta = torch.empty_like(lse)
    a = _n.FlashBwdArgs(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), delta.data_ptr(),
                        dq.data_ptr(), _p(dk), _p(dv), int(B), int(nq), int(Nk), int(H), int(qs), int(ks), float(scale))
    with _on_device(q.device):
        _check(lib.gvd_flash_attention_bwd(C.byref(a), _stream()), lib, "gvd_flash_attention_bwd")
    return dq, dk, dv


def attention(q, k, v, Bq, Nq, Nk, H, scale, shared_kv=False, max_score_bytes=6 << 30, head_dim=64):
    """softmax(q k^T * scale) v with head dim `head_dim` (64 in the U-Net; 512, one head, in the VAE's AttnBlock).
    q [Bq, Nq, H*64]; k, v [Bq, Nk, H*64] (or [1, Nk, H*64] when shared_kv: the same keys for every batch item, then
    the batch is folded into the query rows).  Scores are materialised in bf16 (rounded exactly where the reference
    rounds them) one chunk of batch items at a time, probabilities feed the PV product as bf16."""
    if _wants_grad(q, k, v):
        return _grad().MaterialisedAttention.apply(q, k, v, Bq, Nq, Nk, H, scale, bool(shared_kv), int(head_dim))
    D = int(head_dim)
    HD = H * D
    dev = q.device
    out = torch.empty(Bq, Nq, HD, dtype=BF16, device=dev)
    Nkp = (Nk + 7) // 8 * 8
    if shared_kv:
        # V^T per head, zero-padded along the key axis: [H, 64, Nkp]
        vt = torch.zeros(H, D, Nkp, dtype=BF16, device=dev)
        vt[:, :, :Nk] = v.view(Nk, H, D).permute(1, 2, 0)
        rows_total = Bq * Nq
        rows_chunk = max(128, min(rows_total, max_score_bytes // (H * Nkp * 2) // 128 * 128))
        q2 = q.view(rows_total, HD)
        o2 = out.view(rows_total, HD)
        for r0 in range(0, rows_total, rows_chunk):
            m = min(rows_chunk, rows_total - r0)
            sim = torch.empty(H, m, Nkp, dtype=BF16, device=dev)
            gemm_raw(q2[r0:], k, sim, m, Nk, D, HD, HD, Nkp, batch_h=H, a_strides=(D, 0), b_strides=(D, 0),
                     c_strides=(m * Nkp, 0), alpha=scale, act="round_scale")
            p = softmax_rows(sim, Nk, Nkp)
            gemm_raw(p, vt, o2[r0:], m, D, Nkp, Nkp, Nkp, HD, batch_h=H, a_strides=(m * Nkp, 0), b_strides=(D * Nkp, 0),
                     c_strides=(D, 0))
        return out
    bchunk = max(1, min(Bq, max_score_bytes // (H * Nq * Nkp * 2)))
    for b0 in range(0, Bq, bchunk):
        nb = min(bchunk, Bq - b0)
        vt = torch.zeros(nb, H, D, Nkp, dtype=BF16, device=dev)
        vt[..., :Nk] = v[b0:b0 + nb].view(nb, Nk, H, D).permute(0, 2, 3, 1)
        sim = torch.empty(nb, H, Nq, Nkp, dtype=BF16, device=dev)
        gemm_raw(q[b0:], k[b0:], sim, Nq, Nk, D, HD, HD, Nkp, batch_h=H, batch_b=nb, a_strides=(D, Nq * HD),
                 b_strides=(D, Nk * HD), c_strides=(Nq * Nkp, H * Nq * Nkp), alpha=scale, act="round_scale")
        p = softmax_rows(sim, Nk, Nkp)
        gemm_raw(p, vt, out[b0:], Nq, D, Nkp, Nkp, Nkp, HD, batch_h=H, batch_b=nb, a_strides=(Nq * Nkp, H * Nq * Nkp),
                 b_strides=(D * Nkp, H * D * Nkp), c_strides=(D, Nq * HD))
    return out


def ddim_step(x, e_cond, e_uncond, noise, coef):
    """Fused DDIM update; all tensors fp32, same shape. coef: dict from vc_b200.sampler. Returns (x_prev, pred_x0)."""
    lib = _n.nn()
    n = x.numel()
    x_prev, pred_x0 = torch.empty_like(x), torch.empty_like(x)
    scratch = torch.empty(32 + 4 * n, dtype=torch.uint8, device=x.device)
    a = _n.DdimArgs()
    a.n = n
    a.x, a.e_cond, a.e_uncond, a.noise = x.data_ptr(), e_cond.data_ptr(), _p(e_uncond), noise.data_ptr()
    a.x_prev, a.pred_x0, a.scratch = x_prev.data_ptr(), pred_x0.data_ptr(), scratch.data_ptr()
    for k in ("cfg_scale", "guidance_rescale", "sqrt_alphas_cumprod_t", "sqrt_one_minus_alphas_cumprod_t", "ddim_alpha_prev",
              "ddim_sigma", "temperature", "scale_t", "scale_prev"):
        setattr(a, k, float(coef[k]))
    a.use_dynamic_rescale = int(coef["use_dynamic_rescale"])
    with _on_device(x.device):
        _check(lib.gvd_ddim_step(C.byref(a), _stream()), lib, "gvd_ddim_step")
    return x_prev, pred_x0


# ---------------------------------------------------------------------------------------------------------------------
# Input-gradient operators (include/gvd_nn.h, csrc/nn_backward.cu): the adjoints vc_b200.grad strings together for the
# guided sampler (lvdm/models/samplers/ddim_guidance.py:259-337).  Parameters are frozen, only activations get gradients.
# ---------------------------------------------------------------------------------------------------------------------
def _transposed(weight):
    """weight [N, K] -> [K, Np] copy (Np = N rounded up to 8, zero padded): the B operand of dX = dY @ W.
    The copy lives ON the weight tensor object (layers hold their weights as attributes and pass the same object every
    call), so it dies with the layer and can never be served to another model whose weight lands at a recycled
    address -- a process-wide cache keyed on data_ptr did exactly that on the first hardware run: the second model
    built in a process got the first one's transposed weights.  `_version` catches in-place updates."""
    cached = getattr(weight, "_gvd_wt", None)
    if cached is not None and cached[0] == weight._version and cached[1].device == weight.device:
        return cached[1]
    N, K = weight.shape
    Np = (N + 7) // 8 * 8
    wt = torch.zeros(K, Np, dtype=weight.dtype, device=weight.device)
    wt[:, :N] = weight.t()
    try:
        weight._gvd_wt = (weight._version, wt)
    except AttributeError:  # objects that refuse attributes: no caching, still correct
        pass
    return wt


def linear_dx(dy, weight, alpha=1.0):
    """dX = alpha * dY @ W for y = alpha * x @ W^T: dy [..., N], weight [N, K] -> [..., K] (one tensor-core GEMM against
    the cached transposed weight)."""
    N, K = weight.shape
    wt = _transposed(weight)
    Np = wt.shape[1]
    d2 = dy.reshape(-1, N)
    if Np != N:
        d2 = torch.nn.functional.pad(d2, (0, Np - N))
    elif not d2.is_contiguous():
        d2 = d2.contiguous()
    M = d2.shape[0]
    out = torch.empty(M, K, dtype=dy.dtype, device=dy.device)
    gemm_raw(d2, wt, out, M, K, Np, Np, Np, K, alpha=alpha)
    return out.reshape(*dy.shape[:-1], K)


_gn_bwd_tmp = {}


def groupnorm_bwd(x, dy, gamma, beta, F, S, groups=32, eps=1e-5, silu=0, stats=None):
    """dx of `groupnorm` (same arguments); statistics are recomputed from x unless the forward's are handed in."""
    lib = _n.nn()
    Cc = x.shape[-1]
    x, dy = x.contiguous(), dy.contiguous()
    dx = torch.empty_like(x)
    nfl = int(lib.gvd_groupnorm_tmp_floats(int(F), int(S), int(groups)))
    nby = int(lib.gvd_groupnorm_bwd_tmp_bytes(int(F), int(S), int(groups)))
    key = (x.device, nfl, nby)
    tmp = _gn_bwd_tmp.get(key)
    if tmp is None:
        tmp = _gn_bwd_tmp[key] = (torch.empty(nfl, dtype=torch.float32, device=x.device),
                                  torch.empty(nby // 8 + 1, dtype=torch.float64, device=x.device))
    if stats is None:
        stats = torch.empty(F * groups * 2, dtype=torch.float32, device=x.device)
        _check(lib.gvd_groupnorm_cl_stats(x.data_ptr(), stats.data_ptr(), int(F), int(S), int(Cc), int(groups), tmp[0].data_ptr(), nfl,
                                          _stream()), lib, "gvd_groupnorm_cl_stats")
    _check(lib.gvd_groupnorm_cl_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), gamma.data_ptr(), beta.data_ptr(), stats.data_ptr(),
                                    int(F), int(S), int(Cc), int(groups), float(eps), int(silu), tmp[1].data_ptr(), nby, _stream()),
           lib, "gvd_groupnorm_cl_bwd")
    return dx


def groupnorm_sharded_bwd(x, dy, gamma, beta, F, S_local, S_total, part, groups=32, eps=1e-5, silu=0):
    """dx of `groupnorm_sharded`: local (sum x, sum x^2) -> all-reduce -> local (sum g, sum g xh) -> all-reduce -> local dx."""
    lib = _n.nn()
    Cc = x.shape[-1]
    x, dy = x.contiguous(), dy.contiguous()
    dx = torch.empty_like(x)
    Sl = int(max(S_local, 1))
    nfl = int(lib.gvd_groupnorm_tmp_floats(int(F), Sl, int(groups)))
    nby = int(lib.gvd_groupnorm_bwd_tmp_bytes(int(F), Sl, int(groups)))
    tmpf = torch.empty(nfl, dtype=torch.float32, device=x.device)
    tmpd = torch.empty(nby // 8 + 1, dtype=torch.float64, device=x.device)
    stats = torch.empty(F * groups * 2, dtype=torch.float32, device=x.device)
    _check(lib.gvd_groupnorm_cl_stats(x.data_ptr(), stats.data_ptr(), int(F), int(S_local), int(Cc), int(groups), tmpf.data_ptr(), nfl,
                                      _stream()), lib, "gvd_groupnorm_cl_stats")
    part.sum_stats(stats)
    sums = torch.empty(F * groups * 2, dtype=torch.float64, device=x.device)
    _check(lib.gvd_groupnorm_cl_bwd_sums(x.data_ptr(), dy.data_ptr(), gamma.data_ptr(), beta.data_ptr(), stats.data_ptr(), sums.data_ptr(),
                                         int(F), int(S_local), int(S_total), int(Cc), int(groups), float(eps), int(silu), tmpd.data_ptr(),
                                         nby, _stream()), lib, "gvd_groupnorm_cl_bwd_sums")
    part.sum_stats(sums)
    _check(lib.gvd_groupnorm_cl_bwd_apply(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), gamma.data_ptr(), beta.data_ptr(), stats.data_ptr(),
                                          sums.data_ptr(), int(F), int(S_local), int(S_total), int(Cc), int(groups), float(eps), int(silu),
                                          _stream()), lib, "gvd_groupnorm_cl_bwd_apply")
    return dx


def layernorm_bwd(x, dy, gamma, eps=1e-5):
    lib = _n.nn()
    x, dy = x.contiguous(), dy.contiguous()
    dx = torch.empty_like(x)
    Cc = x.shape[-1]
    _check(lib.gvd_layernorm_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), gamma.data_ptr(), x.numel() // Cc, int(Cc), float(eps),
                                 _stream()), lib, "gvd_layernorm_bwd")
    return dx


def geglu_bwd(h, dout):
    lib = _n.nn()
    D = h.shape[-1] // 2
    h, dout = h.contiguous(), dout.contiguous()
    dh = torch.empty_like(h)
    _check(lib.gvd_geglu_bwd(h.data_ptr(), dout.data_ptr(), dh.data_ptr(), h.numel() // (2 * D), int(D), _stream()), lib,
           "gvd_geglu_bwd")
    return dh


def softmax_bwd_rows(p, dp, cols):
    """ds = p o (dp - rowsum(p o dp)) over the first `cols` columns, written over dp (padding columns zeroed)."""
    lib = _n.nn()
    ld = p.shape[-1]
    _check(lib.gvd_softmax_bwd_rows(p.data_ptr(), dp.data_ptr(), dp.data_ptr(), int(ld), p.numel() // ld, int(cols), _stream()), lib,
           "gvd_softmax_bwd_rows")
    return dp


def _dgrad_weight(weight, taps):
    """weight [Cout, taps*Cin] (tap-major K) -> [Cin, taps*Cout] with the taps reversed: the weight of the convolution
    that maps dY to dX (dX[p] = sum_tap dY[p - off(tap)] W[tap] = sum_tap' dY[p + off(tap')] W[taps-1-tap']).  Cached on the
    weight tensor object like `_transposed`."""
    cached = getattr(weight, "_gvd_wd", None)
    if cached is not None and cached[0] == weight._version and cached[1].device == weight.device:
        return cached[1]
    Cout, K = weight.shape
    Cin = K // taps
    wd = weight.view(Cout, taps, Cin).flip(1).permute(2, 1, 0).contiguous().view(Cin, taps * Cout)
    try:
        weight._gvd_wd = (weight._version, wd)
    except AttributeError:
        pass
    return wd


def conv3x3_dx(dy, F, H, W, Cin, weight, stride=1, upsample=False):
    """dX of `conv3x3`.  dy [F, Ho*Wo, Cout] -> [F, H*W, Cin].  Stride 1 without upsampling: the same implicit-GEMM
    convolution over dY with the tap-reversed, transposed weight; otherwise dcol = dY @ W and the col2im gather."""
    lib = _n.nn()
    Cout = dy.shape[-1]
    if stride == 1 and not upsample and _implicit_ok(1, H, W, Cout, Cin, dy):
        return _conv_implicit(1, dy.reshape(F, H * W, Cout), _dgrad_weight(weight, 9), (F, H, W)).view(F, H * W, Cin)
    if stride == 1 and upsample and _implicit_ok(1, 2 * H, 2 * W, Cout, Cin, dy):
        dxu = _conv_implicit(1, dy.reshape(F, 4 * H * W, Cout), _dgrad_weight(weight, 9), (F, 2 * H, 2 * W))
        return upsample2x_bwd(dxu.view(F, 4 * H * W, Cin), F, H, W)
    dcol = linear_dx(dy.reshape(-1, dy.shape[-1]), weight)  # [F*Ho*Wo, 9*Cin]
    dx = torch.empty(F, H * W, Cin, dtype=dy.dtype, device=dy.device)
    _check(lib.gvd_col2im3x3_cl(dcol.data_ptr(), dx.data_ptr(), int(F), int(H), int(W), int(Cin), int(stride), int(upsample),
                                _stream()), lib, "gvd_col2im3x3_cl")
    return dx


def conv_t3_dx(dy, B, T, S, Cin, weight):
    lib = _n.nn()
    Cout = dy.shape[-1]
    if _implicit_ok(2, 0, 0, Cout, Cin, dy):
        return _conv_implicit(2, dy.reshape(B * T, S, Cout), _dgrad_weight(weight, 3), (B, T, S)).view(B * T, S, Cin)
    dcol = linear_dx(dy.reshape(-1, dy.shape[-1]), weight)  # [B*T*S, 3*Cin]
    dx = torch.empty(B * T, S, Cin, dtype=dy.dtype, device=dy.device)
    _check(lib.gvd_col2im_t3_cl(dcol.data_ptr(), dx.data_ptr(), int(B), int(T), int(S), int(Cin), _stream()), lib,
           "gvd_col2im_t3_cl")
    return dx


def temporal_attention_bwd(q, k, v, dout, B, T, S, H, scale):
    lib = _n.nn()
    q, k, v, dout = q.contiguous(), k.contiguous(), v.contiguous(), dout.contiguous()
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    _check(lib.gvd_temporal_attention_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), dout.data_ptr(), dq.data_ptr(), dk.data_ptr(),
                                          dv.data_ptr(), int(B), int(T), int(S), int(H), float(scale), _stream()), lib,
           "gvd_temporal_attention_bwd")
    return dq, dk, dv


def _heads_t(x, nb, N, H, Np, D=64):
    """x [nb, N, H*D] -> [nb, H, D, Np] (token axis last, zero padded to Np): the K-major B operand of a product that
    contracts over tokens."""
    if Np == N:
        return x.reshape(nb, N, H, D).permute(0, 2, 3, 1).contiguous()
    out = torch.zeros(nb, H, D, Np, dtype=x.dtype, device=x.device)
    out[..., :N] = x.reshape(nb, N, H, D).permute(0, 2, 3, 1)
    return out


def _mat_t(m, rows, cols, colsp):
    """m [..., rows, ld] (first `cols` columns valid) -> [..., cols, colsp'] transposed copy, rows padded to a multiple of 8."""
    rp = (rows + 7) // 8 * 8
    mt = m[..., :cols].transpose(-1, -2)
    if rp == rows:
        return mt.contiguous(), rp
    out = torch.zeros(*m.shape[:-2], cols, rp, dtype=m.dtype, device=m.device)
    out[..., :rows] = mt
    return out, rp


def attention_bwd(q, k, v, dout, Bq, Nq, Nk, H, scale, shared_kv=False, need_kv=True, max_score_bytes=4 << 30, head_dim=64):
    """Backward of `flash_attention` / `attention` (head dim 64): returns (dq, dk, dv); dk = dv = None when
    need_kv is False (keys/values that come from the frozen context).  The probabilities are recomputed and
    materialised one chunk of batch items at a time (bf16, the rounding points of `attention`); every product --
    S = QK^T, dP = dO V^T, dQ = dS K, dK = dS^T Q, dV = P^T dO -- is a launch of the tensor-core GEMM."""
    D = int(head_dim)
    HD = H * D
    dev = q.device
    q, k, v, dout = q.contiguous(), k.contiguous(), v.contiguous(), dout.contiguous()
    dq = torch.empty_like(q)
    Nkp = (Nk + 7) // 8 * 8
    if shared_kv:
        if need_kv:
            raise NotImplementedError("attention_bwd: shared keys/values are frozen-context projections (no dk/dv)")
        kt = _heads_t(k.reshape(1, Nk, HD), 1, Nk, H, Nkp, D)[0]  # [H, D, Nkp]
        rows_total = Bq * Nq
        rows_chunk = max(128, min(rows_total, max_score_bytes // (2 * H * Nkp * 2) // 128 * 128))
        q2, d2, dq2 = q.view(rows_total, HD), dout.view(rows_total, HD), dq.view(rows_total, HD)
        for r0 in range(0, rows_total, rows_chunk):
            m = min(rows_chunk, rows_total - r0)
            sim = torch.empty(H, m, Nkp, dtype=q.dtype, device=dev)
            gemm_raw(q2[r0:], k, sim, m, Nk, D, HD, HD, Nkp, batch_h=H, a_strides=(D, 0), b_strides=(D, 0),
                     c_strides=(m * Nkp, 0), alpha=scale, act="round_scale")
            p = softmax_rows(sim, Nk, Nkp)
            gemm_raw(d2[r0:], v, sim, m, Nk, D, HD, HD, Nkp, batch_h=H, a_strides=(D, 0), b_strides=(D, 0),
                     c_strides=(m * Nkp, 0))  # dP over the score buffer
            ds = softmax_bwd_rows(p, sim, Nk)
            gemm_raw(ds, kt, dq2[r0:], m, D, Nkp, Nkp, Nkp, HD, batch_h=H, a_strides=(m * Nkp, 0), b_strides=(D * Nkp, 0),
                     c_strides=(D, 0), alpha=scale)
        return dq, None, None
    dk, dv = (torch.empty_like(k), torch.empty_like(v)) if need_kv else (None, None)
    per_item = H * Nq * Nkp * 2
    bchunk = max(1, min(Bq, max_score_bytes // (4 * per_item)))  # scores/dS, P, and their two transposes
    for b0 in range(0, Bq, bchunk):
        nb = min(bchunk, Bq - b0)
        sim = torch.empty(nb, H, Nq, Nkp, dtype=q.dtype, device=dev)
        gemm_raw(q[b0:], k[b0:], sim, Nq, Nk, D, HD, HD, Nkp, batch_h=H, batch_b=nb, a_strides=(D, Nq * HD),
                 b_strides=(D, Nk * HD), c_strides=(Nq * Nkp, H * Nq * Nkp), alpha=scale, act="round_scale")
        p = softmax_rows(sim, Nk, Nkp)
        gemm_raw(dout[b0:], v[b0:], sim, Nq, Nk, D, HD, HD, Nkp, batch_h=H, batch_b=nb, a_strides=(D, Nq * HD),
                 b_strides=(D, Nk * HD), c_strides=(Nq * Nkp, H * Nq * Nkp))
        ds = softmax_bwd_rows(p, sim, Nk)
        kt = _heads_t(k[b0:b0 + nb], nb, Nk, H, Nkp, D)
        gemm_raw(ds, kt, dq[b0:], Nq, D, Nkp, Nkp, Nkp, HD, batch_h=H, batch_b=nb, a_strides=(Nq * Nkp, H * Nq * Nkp),
                 b_strides=(D * Nkp, H * D * Nkp), c_strides=(D, Nq * HD), alpha=scale)
        if need_kv:
            # dK^T = Q^T dS and dV^T = dO^T P, each [D, Nk] per (item, head): dS / P enter as the MN-major B operand exactly
            # as they lie in memory ([Nq, Nkp], keys contiguous).  Round 2's first hardware profile of the guided step showed
            # the transposed copies of these two score-sized matrices (dK = dS^T Q, dV = P^T dO as K-major GEMMs) costing
            # 73 ms of a 654 ms step; what is transposed now are the small [Nq, D] / [D, Nk] operands and results.
            Nqp = (Nq + 7) // 8 * 8
            qt = _heads_t(q[b0:b0 + nb], nb, Nq, H, Nqp, D)         # [nb, H, D, Nqp]
            dot = _heads_t(dout[b0:b0 + nb], nb, Nq, H, Nqp, D)
            dkt = torch.empty(nb, H, D, Nkp, dtype=q.dtype, device=dev)
            dvt = torch.empty(nb, H, D, Nkp, dtype=q.dtype, device=dev)
            # N = Nkp: the padding columns of dS and P are zeros, so are the result's (the epilogue wants N % 8 == 0)
            gemm_raw(qt, ds, dkt, D, Nkp, Nq, Nqp, Nkp, Nkp, batch_h=H, batch_b=nb, a_strides=(D * Nqp, H * D * Nqp),
                     b_strides=(Nq * Nkp, H * Nq * Nkp), c_strides=(D * Nkp, H * D * Nkp), alpha=scale, b_mn_major=True)
            gemm_raw(dot, p, dvt, D, Nkp, Nq, Nqp, Nkp, Nkp, batch_h=H, batch_b=nb, a_strides=(D * Nqp, H * D * Nqp),
                     b_strides=(Nq * Nkp, H * Nq * Nkp), c_strides=(D * Nkp, H * D * Nkp), b_mn_major=True)
            dk[b0:b0 + nb] = dkt[..., :Nk].permute(0, 3, 1, 2).reshape(nb, Nk, HD)
            dv[b0:b0 + nb] = dvt[..., :Nk].permute(0, 3, 1, 2).reshape(nb, Nk, HD)
    return dq, dk, dv


def ddim_pred_x0_vjp(e_cond, e_uncond, grad_pred_x0, coef):
    """(dx_direct, de_cond, de_uncond) for G = dL/dpred_x0 of one guided step (gvd_ddim_pred_x0_vjp); fp32 tensors of one
    batch item, `coef` from DdimSchedule.coefficients."""
    lib = _n.nn()
    n = e_cond.numel()
    e_cond, grad_pred_x0 = e_cond.contiguous(), grad_pred_x0.contiguous()
    dx, de_c = torch.empty_like(e_cond), torch.empty_like(e_cond)
    de_u = None
    if e_uncond is not None:
        e_uncond = e_uncond.contiguous()
        de_u = torch.empty_like(e_cond)
    scratch = torch.empty(8, dtype=torch.float64, device=e_cond.device)
    a = _n.DdimVjpArgs()
    a.n = n
    a.e_cond, a.e_uncond, a.grad_pred_x0 = e_cond.data_ptr(), _p(e_uncond), grad_pred_x0.data_ptr()
    a.dx, a.de_cond, a.de_uncond = dx.data_ptr(), de_c.data_ptr(), _p(de_u)
    a.scratch, a.scratch_bytes = scratch.data_ptr(), 64
    for key in ("cfg_scale", "guidance_rescale", "sqrt_alphas_cumprod_t", "sqrt_one_minus_alphas_cumprod_t", "scale_t", "scale_prev"):
        setattr(a, key, float(coef[key]))
    a.use_dynamic_rescale = int(coef["use_dynamic_rescale"])
    with _on_device(e_cond.device):
        _check(lib.gvd_ddim_pred_x0_vjp(C.byref(a), _stream()), lib, "gvd_ddim_pred_x0_vjp")
    return dx, de_c, de_u
