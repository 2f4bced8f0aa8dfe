"""Host-side parity of the denoiser's input-gradient path WITHOUT a GPU.

vc_b200 has no CPU path; here `gvd_native.nn()` is replaced by tests/fake_nn_lib.py::FakeNN, which implements every
entry point of include/gvd_nn.h over host pointers.  That pins, against the REFERENCE UNetModel run in fp32 on the CPU
(oracle/_ref/ViewCrafter) and torch.autograd over it:
  * the ctypes bindings (argument order, strides, padding) of every forward and input-gradient operator,
  * the composition of UNetB200's forward and of its backward (vc_b200.grad) -- d(output)/d(input latent),
  * the formulas the kernels of csrc/nn_backward.cu implement (FakeNN restates them in closed form, autograd is the judge).
The kernels themselves are checked on the GPU by tests/test_nn_bwd_gpu.py.
"""
import contextlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "guidedvd-3dgs_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import unet_ref  # noqa: E402
from fake_nn_lib import FakeNN  # noqa: E402

needs_ref = pytest.mark.skipif(not unet_ref.ref_available(), reason="oracle/_ref/ViewCrafter not installed (python oracle/build_ref.py vc)")


def install_fake(monkeypatch, act_dtype=torch.float32):
    import gvd_native
    from vc_b200 import ops, unet

    fake = FakeNN(act_dtype)
    monkeypatch.setattr(gvd_native, "nn", lambda: fake)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    monkeypatch.setattr(ops, "BF16", act_dtype)
    monkeypatch.setattr(unet, "BF16", act_dtype)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    for cache in (ops._gn_tmp, ops._gn_bwd_tmp):
        cache.clear()
    return fake


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.fixture(scope="module")
def tiny_ref():
    torch.manual_seed(0)
    ref, cfg = unet_ref.build_reference_unet(model_channels=64, device="cpu")
    x, cc, ctx, ctx_uc = unet_ref.synth_inputs(3, 8, 8, device="cpu")
    return ref, cfg, torch.cat([x, cc], 1), ctx, ctx_uc


@needs_ref
def test_forward_matches_reference_fp32(monkeypatch, tiny_ref):
    from vc_b200.unet import UNetB200

    fake = install_fake(monkeypatch)
    ref, cfg, xin, ctx, _ = tiny_ref
    ours = UNetB200(ref.state_dict(), device="cpu", **cfg)
    ts, fs = torch.tensor([481]), torch.tensor([10])
    with torch.no_grad():
        y_ref = ref(xin, ts, context=ctx, fs=fs)
    y = ours(xin, ts, ctx, fs=fs)
    assert y.shape == y_ref.shape
    assert _rel(y, y_ref) < 2e-5
    assert fake.calls["gemm"] > 100 and fake.calls["flash_attention"] > 10 and "groupnorm_bwd" not in fake.calls


@needs_ref
@pytest.mark.parametrize("seed,implicit,fused_attn", [(0, True, True), (1, True, True), (0, False, False)])
def test_input_gradient_matches_autograd_of_reference(monkeypatch, tiny_ref, seed, implicit, fused_attn):
    """d<y, g>/dx for a random cotangent g: ours (vc_b200.grad over the C-ABI stand-in) vs torch.autograd over the
    reference module -- the call `pred_x0.backward(gradient=..., inputs=x)` of ddim_guidance.py:309."""
    from vc_b200.unet import UNetB200

    fake = install_fake(monkeypatch)
    from vc_b200 import ops
    monkeypatch.setattr(ops, "IMPLICIT_CONV", implicit)  # False: every convolution through im2col / col2im
    monkeypatch.setattr(ops, "FUSED_FLASH_BWD", fused_attn)  # False: the attention adjoint re-materialises the scores
    ref, cfg, xin, ctx, _ = tiny_ref
    ours = UNetB200(ref.state_dict(), device="cpu", **cfg)
    ts, fs = torch.tensor([300 + 100 * seed]), torch.tensor([10])
    g = torch.randn(1, 4, *xin.shape[2:], generator=torch.Generator().manual_seed(seed))

    xr = xin.clone().requires_grad_(True)
    ref(xr, ts, context=ctx, fs=fs).backward(gradient=g, inputs=[xr])

    xo = xin.clone().requires_grad_(True)
    y = ours.forward_with_grad(xo, ts, ctx, fs=fs)
    assert y.requires_grad
    y.backward(gradient=g.to(y.dtype), inputs=[xo])
    assert xo.grad is not None and xo.grad.shape == xr.grad.shape
    err = _rel(xo.grad, xr.grad)
    print(f"input-gradient rel L2 vs autograd(reference): {err:.3e}")
    assert err < 1e-4
    # every adjoint took part
    convs = ("conv_implicit",) if implicit else ("col2im3x3", "col2im_t3")
    attn = ("flash_attention_lse", "flash_attention_bwd") if fused_attn else ("softmax_bwd",)
    for name in ("groupnorm_bwd", "layernorm_bwd", "geglu_bwd", "temporal_attention_bwd") + convs + attn:
        assert fake.calls.get(name, 0) > 0, name
    if fused_attn:
        assert "softmax_bwd" not in fake.calls
    if not implicit:
        assert "conv_implicit" not in fake.calls


@needs_ref
def test_inference_path_records_no_graph(monkeypatch, tiny_ref):
    """The plain sampler's call (`forward`, under no_grad) must never reach vc_b200.grad, even for a leaf that requires grad."""
    from vc_b200.unet import UNetB200

    fake = install_fake(monkeypatch)
    ref, cfg, xin, ctx, _ = tiny_ref
    ours = UNetB200(ref.state_dict(), device="cpu", **cfg)
    y = ours(xin.clone().requires_grad_(True), torch.tensor([481]), ctx, fs=torch.tensor([10]))
    assert not y.requires_grad and y.grad_fn is None
    assert not any(k.endswith("_bwd") or k.startswith("col2im") for k in fake.calls)


@needs_ref
def test_input_gradient_with_bf16_rounding_points(monkeypatch, tiny_ref):
    """Same comparison with bf16 activations in the stand-in (the kernels' storage type and rounding points): the
    gradient is as close to the fp32 one as the REFERENCE's own gradient under torch.autocast(bfloat16) is (both sit at
    ~5e-2 relative L2 on this small network -- the bf16 noise floor of a backward through ~150 layers)."""
    from vc_b200.unet import UNetB200

    install_fake(monkeypatch, torch.bfloat16)
    ref, cfg, xin, ctx, _ = tiny_ref
    ours = UNetB200(ref.state_dict(), device="cpu", **cfg)
    ts, fs = torch.tensor([481]), torch.tensor([10])
    g = torch.randn(1, 4, *xin.shape[2:], generator=torch.Generator().manual_seed(5))
    xr = xin.clone().requires_grad_(True)
    ref(xr, ts, context=ctx, fs=fs).backward(gradient=g, inputs=[xr])
    xo = xin.clone().requires_grad_(True)
    ours.forward_with_grad(xo, ts, ctx, fs=fs).backward(gradient=g.to(torch.bfloat16), inputs=[xo])
    xb = xin.clone().requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        yb = ref(xb, ts, context=ctx, fs=fs)
    yb.backward(gradient=g.to(yb.dtype), inputs=[xb])
    err, err_ref = _rel(xo.grad, xr.grad), _rel(xb.grad, xr.grad)
    print(f"bf16 stand-in: input-gradient rel L2 {err:.3e}; reference under bf16 autocast {err_ref:.3e}")
    assert err <= 1.25 * err_ref + 5e-3


def test_attention_bwd_all_paths_against_autograd(monkeypatch):
    """ops.attention_bwd (GEMM + softmax-backward composition, padded key/query counts, shared keys) vs autograd."""
    from vc_b200 import ops

    install_fake(monkeypatch)
    g = torch.Generator().manual_seed(3)
    H, D = 2, 64
    for (Bq, Nq, Nk, shared) in [(3, 20, 20, False), (2, 16, 13, False), (3, 12, 77, True), (2, 24, 256, True)]:
        q = torch.randn(Bq, Nq, H * D, generator=g).requires_grad_(True)
        k = torch.randn(1 if shared else Bq, Nk, H * D, generator=g).requires_grad_(not shared)
        v = torch.randn(1 if shared else Bq, Nk, H * D, generator=g).requires_grad_(not shared)
        do = torch.randn(Bq, Nq, H * D, generator=g)
        scale = D ** -0.5
        qh = q.view(Bq, Nq, H, D)
        kh = k.view(-1, Nk, H, D).expand(Bq, Nk, H, D)
        vh = v.view(-1, Nk, H, D).expand(Bq, Nk, H, D)
        o = torch.einsum("bhij,bjhd->bihd", torch.softmax(torch.einsum("bihd,bjhd->bhij", qh, kh) * scale, -1), vh).reshape(Bq, Nq, H * D)
        o.backward(do)
        dq, dk, dv = ops.attention_bwd(q.detach(), k.detach(), v.detach(), do, Bq, Nq, Nk, H, scale, shared_kv=shared, need_kv=not shared)
        assert _rel(dq, q.grad) < 1e-5
        if not shared:
            assert _rel(dk, k.grad) < 1e-5 and _rel(dv, v.grad) < 1e-5
        # chunked evaluation gives the same result
        dq2, _, _ = ops.attention_bwd(q.detach(), k.detach(), v.detach(), do, Bq, Nq, Nk, H, scale, shared_kv=shared, need_kv=not shared,
                                      max_score_bytes=1)
        assert _rel(dq2, q.grad) < 1e-5


@pytest.mark.parametrize("stride,up", [(1, False), (2, False), (1, True)])
def test_conv_adjoints_against_autograd(monkeypatch, stride, up):
    """conv3x3_dx / conv_t3_dx: the dcol GEMM + col2im index math vs autograd through F.conv2d / conv3d."""
    from vc_b200 import ops

    install_fake(monkeypatch)
    g = torch.Generator().manual_seed(11)
    F_, H, W, Cin, Cout = 2, 6, 5, 8, 16
    x = torch.randn(F_, H * W, Cin, generator=g).requires_grad_(True)
    w4 = torch.randn(Cout, Cin, 3, 3, generator=g)
    w = w4.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
    y, Ho, Wo = ops.conv3x3(x, F_, H, W, w, None, stride=stride, upsample=up)
    img = x.view(F_, H, W, Cin).permute(0, 3, 1, 2)
    if up:
        img = torch.nn.functional.interpolate(img, scale_factor=2, mode="nearest")
    y_ref = torch.nn.functional.conv2d(img, w4, stride=stride, padding=1)
    assert (Ho, Wo) == tuple(y_ref.shape[2:])
    assert _rel(y, y_ref.permute(0, 2, 3, 1).reshape(F_, Ho * Wo, Cout)) < 1e-5
    dy = torch.randn(F_, Ho * Wo, Cout, generator=g)
    y.backward(dy)
    gx_ours = x.grad.clone()
    x.grad = None
    y_ref.backward(dy.view(F_, Ho, Wo, Cout).permute(0, 3, 1, 2))
    assert _rel(gx_ours, x.grad) < 1e-5


def test_temporal_conv_adjoint_against_autograd(monkeypatch):
    from vc_b200 import ops

    install_fake(monkeypatch)
    g = torch.Generator().manual_seed(12)
    B, T, S, Cin, Cout = 1, 5, 7, 8, 8
    x = torch.randn(B * T, S, Cin, generator=g).requires_grad_(True)
    w5 = torch.randn(Cout, Cin, 3, 1, 1, generator=g)
    w = w5[:, :, :, 0, 0].permute(0, 2, 1).reshape(Cout, -1).contiguous()
    res = torch.randn(B * T, S, Cout, generator=g).requires_grad_(True)
    y = ops.conv_t3(x, B, T, S, w, None, residual=res)
    vol = x.view(B, T, S, 1, Cin).permute(0, 4, 1, 2, 3)
    y_ref = torch.nn.functional.conv3d(vol, w5, padding=(1, 0, 0)).permute(0, 2, 3, 4, 1).reshape(B * T, S, Cout) + res
    assert _rel(y, y_ref) < 1e-5
    dy = torch.randn(B * T, S, Cout, generator=g)
    y.backward(dy)
    gx, gr = x.grad.clone(), res.grad.clone()
    x.grad = res.grad = None
    y_ref.backward(dy)
    assert _rel(gx, x.grad) < 1e-5 and _rel(gr, res.grad) < 1e-6


def test_transposed_weight_copy_belongs_to_the_weight_object():
    """Regression for the first hardware run of the guided path: the transposed copies of dX = dY @ W were cached
    process-wide under data_ptr, so a second model whose weights landed at recycled addresses got the first model's.
    The copy now lives on the weight tensor object: another tensor at the same address must not see it, an in-place
    update of the weight invalidates it."""
    from vc_b200 import ops

    w = torch.arange(12.0).view(3, 4)
    t1 = ops._transposed(w)
    assert ops._transposed(w) is t1 and torch.equal(t1[:, :3], w.t()) and t1.shape == (4, 8)
    alias = w.view(3, 4)  # same storage and address, different Python object: as after free + reallocation
    w.mul_(2.0)
    assert torch.equal(ops._transposed(alias)[:, :3], w.t())
    t2 = ops._transposed(w)
    assert t2 is not t1 and torch.equal(t2[:, :3], w.t())
