"""GPU numerics: the tcgen05 bf16 GEMM (include/gvd_nn.h::gvd_gemm_bf16) against a plain PyTorch fp32 reference of
the same op on the same bf16 inputs.  Tolerance: bf16 output rounding (2^-8 relative) plus fp32 accumulation noise."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(A, B, bias, res, alpha, act):
    y = alpha * (A.float() @ B.float().transpose(-1, -2))
    if bias is not None:
        y = y + bias
    if act == "silu":
        y = torch.nn.functional.silu(y)
    elif act == "gelu":
        y = torch.nn.functional.gelu(y)
    if res is not None:
        y = y + res.float()
    return y


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 384, 320), (1000, 320, 2880), (130, 72, 72), (77, 1280, 1024),
                                   (4096, 2560, 320), (300, 8, 2880), (64, 4, 320), (513, 640, 8)])
@pytest.mark.parametrize("opts", [dict(), dict(bias=True, act="silu"), dict(bias=True, res=True), dict(fp32=True, bias=True, act="gelu", alpha=0.125)])
def test_linear_shapes(M, N, K, opts):
    from vc_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) if opts.get("bias") else None
    out_dtype = torch.float32 if opts.get("fp32") else torch.bfloat16
    res = torch.randn(M, N, device="cuda", generator=g).to(out_dtype) if opts.get("res") else None
    alpha, act = opts.get("alpha", 1.0), opts.get("act", "none")
    y = ops.linear(A, B, bias=bias, act=act, residual=res, out_dtype=out_dtype, alpha=alpha)
    ref = _ref(A, B, bias, res, alpha, act)
    tol = 2e-5 if out_dtype == torch.float32 else 1.0 / 128
    err = (y.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= tol * scale + 1e-4, (err, scale)


def test_batched_strided_heads():
    """The q/k layout of CrossAttention: [b, n, h*64] viewed per head without a copy (attention.py:98-103)."""
    from vc_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(5)
    Bn, H, Nq, Nk, D = 3, 5, 200, 77, 64
    q = torch.randn(Bn, Nq, H * D, device="cuda", generator=g).bfloat16()
    k = torch.randn(Bn, Nk, H * D, device="cuda", generator=g).bfloat16()
    sim = torch.empty(Bn, H, Nq, Nk + 3, device="cuda", dtype=torch.float32)[..., :Nk]  # padded rows: ldc = Nk+3 -> scalar path
    sim = torch.empty(Bn, H, Nq, 80, device="cuda", dtype=torch.float32)
    ops.gemm_raw(q, k, sim, Nq, Nk, D, H * D, H * D, 80, batch_h=H, batch_b=Bn, a_strides=(D, Nq * H * D),
                 b_strides=(D, Nk * H * D), c_strides=(Nq * 80, H * Nq * 80), alpha=D ** -0.5)
    ref = torch.einsum("bihd,bjhd->bhij", q.float().view(Bn, Nq, H, D), k.float().view(Bn, Nk, H, D)) * D ** -0.5
    assert (sim[..., :Nk] - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("M,N,K,ldb", [(64, 256, 256, 256), (64, 200, 200, 200), (512, 640, 640, 640), (64, 77, 130, 80), (100, 2560, 1000, 2560)])
def test_mn_major_b_operand(M, N, K, ldb):
    """C = A * B with B given as B[k][n] (n contiguous): the tcgen05 MN-major operand used for dK^T = Q^T dS and
    dV^T = dO^T P in the attention backward (no transposed copy of the score-sized matrices).  Batched over two levels."""
    from vc_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    nb, H = 2, 3
    Kp = (K + 7) // 8 * 8
    A = torch.zeros(nb, H, M, Kp, device="cuda", dtype=torch.bfloat16)
    A[..., :K] = torch.randn(nb, H, M, K, device="cuda", generator=g).bfloat16()
    B = torch.zeros(nb, H, K, ldb, device="cuda", dtype=torch.bfloat16)
    B[..., :N] = (torch.randn(nb, H, K, N, device="cuda", generator=g) / K ** 0.5).bfloat16()
    Np = (N + 7) // 8 * 8
    Cout = torch.full((nb, H, M, Np), 7.0, device="cuda", dtype=torch.bfloat16)
    # the library wants N % 8 == 0 on this path: callers pass the padded width (B's padding columns are zeros)
    ops.gemm_raw(A, B, Cout, M, Np, K, Kp, ldb, Np, batch_h=H, batch_b=nb, a_strides=(M * Kp, H * M * Kp),
                 b_strides=(K * ldb, H * K * ldb), c_strides=(M * Np, H * M * Np), alpha=0.5, b_mn_major=True)
    assert float(Cout[..., N:].float().abs().max()) == 0.0 if Np > N else True
    ref = 0.5 * (A[..., :K].float() @ B[..., :N].float())
    err = (Cout[..., :N].float() - ref).abs().max().item()
    assert err <= ref.abs().max().item() / 128 + 1e-4, err


@pytest.mark.parametrize("M,D,K", [(300, 1280, 320), (4096, 2560, 640), (1000, 5120, 1280), (130, 48, 64)])
def test_fused_geglu_epilogue(M, D, K):
    """GVD_ACT_GEGLU: projection GEMM with the gate applied in the epilogue (interleaved weight rows) against the
    two-step route (GEMM, then gvd_geglu) on the same inputs.  Same accumulators and rounding points; the epilogue's GELU
    uses a 1.5e-7 erf approximation, so where the gate's GELU sat on a rounding boundary its bf16 value moves by one ulp and
    the product, rounded again, by up to two.  K >= 640 takes the CTA-pair kernel, K = 320 the one-CTA kernel."""
    from vc_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M + D)
    x = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(2 * D, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(2 * D, device="cuda", generator=g)
    il = ops.geglu_weight(w, b)
    assert il is not None
    fused = ops.linear_geglu(x, *il)
    two = ops.geglu(ops.linear(x, w, b))
    assert fused.shape == two.shape == (M, D)
    f, t = fused.float(), two.float()
    ulp = 2.0 * 2.0 ** -7 * t.abs() + 1e-6
    assert bool(((f - t).abs() <= ulp).all()), float(((f - t).abs() / ulp).max())
    assert (fused != two).float().mean().item() < 0.01
    ref = (x.float() @ w.float().T + b)
    ref = ref[:, :D] * torch.nn.functional.gelu(ref[:, D:])
    assert (f - ref).abs().max().item() <= 2.0 ** -6 * ref.abs().max().item() + 1e-3
