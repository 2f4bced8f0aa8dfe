"""GPU parity of the guided sampler's input-gradient kernels (csrc/nn_backward.cu, include/gvd_nn.h) and of the
denoiser backward built from them (vc_b200.grad, vc_b200.guided), through the C ABI.

Each operator is compared with torch.autograd over a plain fp32 PyTorch statement of the same layer on the same
bf16-rounded inputs; the bar is the bf16 output rounding (relative L2 <= 1e-2 per operator, tolerance written at each
assert).  The whole network's d(output)/d(latent) is compared with autograd over the REFERENCE UNetModel in fp32 and
under torch.autocast(bfloat16), with the same rule as the forward test (tests/test_unet_gpu.py): at least as close to
fp32 as the reference's own bf16 run.

First hardware run (round 2, profiles/r02_first_hw_run.txt): all 23 operator-level cases green; the four model-level
cases failed because the transposed-weight copies of dX = dY @ W were cached process-wide under data_ptr and a second
model built in the same process was served the first one's (fixed in ops._transposed; regression test in
tests/test_unet_grad_cpu.py).  No xfail markers: a failure here fails the suite.
"""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu]

BF = torch.bfloat16


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _bf(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(BF).cuda()


@pytest.mark.parametrize("silu", [0, 1, 2])
@pytest.mark.parametrize("F,S,C", [(3, 64, 64), (2, 1000, 320), (1, 777, 2560)])
def test_groupnorm_bwd(F, S, C, silu):
    from vc_b200 import ops

    x, dy = _bf(F, S, C, seed=1, scale=2.0) + 0.5, _bf(F, S, C, seed=2)
    g = torch.Generator().manual_seed(3)
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).cuda()
    beta = (0.1 * torch.randn(C, generator=g)).cuda()
    xf = x.float().requires_grad_(True)
    z = torch.nn.functional.group_norm(xf.permute(0, 2, 1), 32, gamma, beta, 1e-5).permute(0, 2, 1)
    if silu == 1:
        z = z + (z.to(BF).float() - z).detach()  # SiLU sees the bf16-rounded value, gradient passes straight through
    y = torch.nn.functional.silu(z) if silu else z
    y.backward(dy.float())
    dx = ops.groupnorm_bwd(x, dy, gamma, beta, F, S, 32, 1e-5, silu)
    torch.cuda.synchronize()
    assert _rel(dx, xf.grad) < 1e-2


def test_layernorm_geglu_softmax_bwd():
    from vc_b200 import ops

    x, dy = _bf(300, 640, seed=4, scale=3.0), _bf(300, 640, seed=5)
    g = torch.Generator().manual_seed(6)
    gamma, beta = (1 + 0.1 * torch.randn(640, generator=g)).cuda(), (0.1 * torch.randn(640, generator=g)).cuda()
    xf = x.float().requires_grad_(True)
    torch.nn.functional.layer_norm(xf, (640,), gamma, beta, 1e-5).backward(dy.float())
    assert _rel(ops.layernorm_bwd(x, dy, gamma, 1e-5), xf.grad) < 1e-2

    h, do = _bf(257, 2 * 320, seed=7, scale=1.5), _bf(257, 320, seed=8)
    hf = h.float().requires_grad_(True)
    (hf[:, :320] * torch.nn.functional.gelu(hf[:, 320:])).backward(do.float())
    assert _rel(ops.geglu_bwd(h, do), hf.grad) < 1e-2

    rows, cols, ld = 100, 77, 80
    s = _bf(rows, ld, seed=9, scale=2.0)
    p = ops.softmax_rows(s, cols, ld)
    dp = _bf(rows, ld, seed=10)
    pf = p.float()[:, :cols]
    ref = pf * (dp.float()[:, :cols] - (pf * dp.float()[:, :cols]).sum(-1, keepdim=True))
    ds = ops.softmax_bwd_rows(p, dp.clone(), cols)
    assert _rel(ds[:, :cols], ref) < 1e-2 and float(ds[:, cols:].abs().max()) == 0.0


@pytest.mark.parametrize("stride,up", [(1, False), (2, False), (1, True)])
def test_conv3x3_dx(stride, up):
    from vc_b200 import ops

    F_, H, W, Cin, Cout = 2, 13, 10, 32, 64
    x = _bf(F_, H * W, Cin, seed=11)
    w4 = (torch.randn(Cout, Cin, 3, 3, generator=torch.Generator().manual_seed(12)) / (9 * Cin) ** 0.5).to(BF).cuda()
    w = w4.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
    xf = x.float().requires_grad_(True)
    img = xf.view(F_, H, W, Cin).permute(0, 3, 1, 2)
    if up:
        img = torch.nn.functional.interpolate(img, scale_factor=2, mode="nearest")
    y_ref = torch.nn.functional.conv2d(img, w4.float(), stride=stride, padding=1)
    Ho, Wo = y_ref.shape[2:]
    dy = _bf(F_, Ho * Wo, Cout, seed=13)
    y_ref.backward(dy.float().view(F_, Ho, Wo, Cout).permute(0, 3, 1, 2))
    dx = ops.conv3x3_dx(dy, F_, H, W, Cin, w, stride, up)
    # dcol is rounded to bf16 before the 9-tap (36-tap with upsampling) gather
    assert _rel(dx, xf.grad) < 1.5e-2


def test_conv_t3_dx_and_linear_dx_padding():
    from vc_b200 import ops

    B, T, S, Cin, Cout = 1, 7, 50, 64, 32
    w5 = (torch.randn(Cout, Cin, 3, 1, 1, generator=torch.Generator().manual_seed(14)) / (3 * Cin) ** 0.5).to(BF).cuda()
    w = w5[:, :, :, 0, 0].permute(0, 2, 1).reshape(Cout, -1).contiguous()
    x = _bf(B * T, S, Cin, seed=15)
    xf = x.float().requires_grad_(True)
    vol = xf.view(B, T, S, 1, Cin).permute(0, 4, 1, 2, 3)
    y_ref = torch.nn.functional.conv3d(vol, w5.float(), padding=(1, 0, 0))
    dy = _bf(B * T, S, Cout, seed=16)
    y_ref.backward(dy.float().view(B, T, S, 1, Cout).permute(0, 4, 1, 2, 3))
    assert _rel(ops.conv_t3_dx(dy, B, T, S, Cin, w), xf.grad) < 1.5e-2
    # output conv of the U-Net: N = 4 output channels -> the cotangent is padded to 8 for the GEMM's 16-byte rule
    wl = _bf(4, 320, seed=17, scale=0.05)
    d4 = _bf(1000, 4, seed=18)
    assert _rel(ops.linear_dx(d4, wl), d4.float() @ wl.float()) < 1e-2


@pytest.mark.parametrize("T", [25, 16, 3])
def test_temporal_attention_bwd(T):
    from vc_b200 import ops

    B, S, H = 1, 37, 5
    q, k, v = (_bf(B * T, S, H * 64, seed=20 + i) for i in range(3))
    do = _bf(B * T, S, H * 64, seed=24)
    scale = 64 ** -0.5
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    sp = lambda t: t.view(B, T, S, H, 64).permute(0, 3, 2, 1, 4)  # noqa: E731
    p = torch.softmax(torch.einsum("bhsid,bhsjd->bhsij", sp(qf), sp(kf)) * scale, -1)
    o = torch.einsum("bhsij,bhsjd->bhsid", p, sp(vf)).permute(0, 3, 2, 1, 4).reshape(B * T, S, H * 64)
    o.backward(do.float())
    assert _rel(ops.temporal_attention(q, k, v, B, T, S, H, scale), o) < 1e-2
    dq, dk, dv = ops.temporal_attention_bwd(q, k, v, do, B, T, S, H, scale)
    torch.cuda.synchronize()
    # the forward's logits are rounded to bf16 twice before the softmax (attention.py:103), the fp32 statement's are not
    assert _rel(dq, qf.grad) < 2e-2 and _rel(dk, kf.grad) < 2e-2 and _rel(dv, vf.grad) < 1e-2


@pytest.mark.parametrize("Bq,Nq,Nk,shared", [(3, 256, 256, False), (2, 200, 200, False), (5, 128, 77, True), (2, 384, 256, True)])
def test_attention_bwd(Bq, Nq, Nk, shared):
    from vc_b200 import ops

    H, D = 5, 64
    q = _bf(Bq, Nq, H * D, seed=30)
    k, v = _bf(1 if shared else Bq, Nk, H * D, seed=31), _bf(1 if shared else Bq, Nk, H * D, seed=32)
    do = _bf(Bq, Nq, H * D, seed=33)
    scale = D ** -0.5
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    kh, vh = (t.view(-1, Nk, H, D).expand(Bq, Nk, H, D) for t in (kf, vf))
    p = torch.softmax(torch.einsum("bihd,bjhd->bhij", qf.view(Bq, Nq, H, D), kh) * scale, -1)
    torch.einsum("bhij,bjhd->bihd", p, vh).reshape(Bq, Nq, H * D).backward(do.float())
    dq, dk, dv = ops.attention_bwd(q, k, v, do, Bq, Nq, Nk, H, scale, shared_kv=shared, need_kv=not shared)
    torch.cuda.synchronize()
    assert _rel(dq, qf.grad) < 2e-2
    if not shared:
        assert _rel(dk, kf.grad) < 2e-2 and _rel(dv, vf.grad) < 2e-2


@pytest.mark.parametrize("Bq,Nq,Nk,shared,spread", [(3, 256, 256, False, 1.0), (2, 200, 200, False, 3.0), (5, 128, 77, True, 1.0),
                                                    (2, 384, 256, True, 3.0), (1, 129, 513, False, 2.0), (2, 1000, 77, False, 1.0),
                                                    (2, 2560, 2560, False, 2.0), (1, 64, 1, False, 1.0)])
def test_flash_attention_bwd_fused(Bq, Nq, Nk, shared, spread):
    """The fused adjoint (csrc/attn_bwd_tc.cu, gvd_flash_attention_lse + gvd_flash_attention_bwd) against autograd over
    fp32 softmax attention on the same bf16 inputs -- ragged tiles, one key, cross-attention widths, shared keys, the C4
    self-attention length, and logits spread over +-20 (the lazily rescaled forward's statistic must still be exact).
    Bar: the bf16 rounding of P / dS and of the outputs, relative L2 <= 2e-2 (measured ~4e-3); the row statistic itself
    within 2e-3 absolute of torch.logsumexp in base 2."""
    from vc_b200 import ops

    H, D = 5, 64
    q = _bf(Bq, Nq, H * D, seed=40)
    k, v = _bf(1 if shared else Bq, Nk, H * D, seed=41, scale=spread), _bf(1 if shared else Bq, Nk, H * D, seed=42)
    do = _bf(Bq, Nq, H * D, seed=43)
    scale = D ** -0.5
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    kh, vh = (t.view(-1, Nk, H, D).expand(Bq, Nk, H, D) for t in (kf, vf))
    s = torch.einsum("bihd,bjhd->bhij", qf.view(Bq, Nq, H, D), kh) * scale
    o_ref = torch.einsum("bhij,bjhd->bihd", torch.softmax(s, -1), vh).reshape(Bq, Nq, H * D)
    o_ref.backward(do.float())
    out, lse = ops.flash_attention_lse(q, k, v, Bq, Nq, Nk, H, scale, shared_kv=shared)
    assert torch.equal(out, ops.flash_attention(q, k, v, Bq, Nq, Nk, H, scale, shared_kv=shared))
    lse_ref = torch.logsumexp(s.detach(), -1) * 1.4426950408889634  # [Bq, H, Nq]
    if shared:
        lse_ref = lse_ref.permute(1, 0, 2).reshape(1, H, Bq * Nq)
    assert (lse[:, :, :lse_ref.shape[-1]] - lse_ref).abs().max().item() < 2e-3
    dq, dk, dv = ops.flash_attention_bwd(q, k, v, out, lse, do, Bq, Nq, Nk, H, scale, shared_kv=shared, need_kv=not shared)
    torch.cuda.synchronize()
    assert torch.isfinite(dq.float()).all()
    if Nk == 1:  # softmax over one key is the constant 1: dQ = dK = 0 exactly in fp32, rounding dust here
        assert dq.float().abs().max().item() < 1e-5 and dk.float().abs().max().item() < 1e-4
        assert _rel(dv, vf.grad) < 1e-2
        return
    assert _rel(dq, qf.grad) < 2e-2, _rel(dq, qf.grad)
    if not shared:
        assert _rel(dk, kf.grad) < 2e-2 and _rel(dv, vf.grad) < 2e-2, (_rel(dk, kf.grad), _rel(dv, vf.grad))
        # and against the first backward (scores materialised): same rounding points, so closer than either is to fp32
        dq1, dk1, dv1 = ops.attention_bwd(q, k, v, do, Bq, Nq, Nk, H, scale, need_kv=True)
        assert _rel(dq, dq1) < 2e-2 and _rel(dk, dk1) < 2e-2 and _rel(dv, dv1) < 2e-2


def test_flash_attention_autograd_routes_to_fused_adjoint():
    """vc_b200.grad.FlashAttention: forward under autograd saves (out, lse) and backward is the fused kernel; the
    GVD_FLASH_BWD=0 route gives the same gradients within the bf16 bar."""
    from vc_b200 import ops

    H = 5
    q, k, v = (_bf(2, 300, H * 64, seed=50 + i).requires_grad_(True) for i in range(3))
    g = _bf(2, 300, H * 64, seed=53)
    grads = []
    for fused in (True, False):
        ops.FUSED_FLASH_BWD = fused
        try:
            for t in (q, k, v):
                t.grad = None
            ops.flash_attention(q, k, v, 2, 300, 300, H, 0.125).backward(g)
            grads.append([t.grad.clone() for t in (q, k, v)])
        finally:
            ops.FUSED_FLASH_BWD = True
    for a, b in zip(*grads):
        assert _rel(a, b) < 2e-2


def test_pred_x0_vjp():
    from vc_b200 import ops
    from vc_b200.schedule import DdimSchedule, ModelSchedule

    g = torch.Generator().manual_seed(4)
    shape = (1, 4, 25, 40, 64)
    coef = DdimSchedule(ModelSchedule(), 50, "uniform_trailing", 1.0).coefficients(30, 7.5, 0.7, 1.0)
    x = torch.randn(shape, generator=g).cuda().requires_grad_(True)
    e_c = torch.randn(shape, generator=g).cuda().requires_grad_(True)
    e_u = (e_c.detach() + 0.3 * torch.randn(shape, generator=g).cuda()).requires_grad_(True)
    G = torch.randn(shape, generator=g).cuda()
    mo = e_u + 7.5 * (e_c - e_u)
    v = 0.7 * (mo * (e_c.std() / mo.std())) + 0.3 * mo
    p0 = (coef["sqrt_alphas_cumprod_t"] * x - coef["sqrt_one_minus_alphas_cumprod_t"] * v) * (coef["scale_prev"] / coef["scale_t"])
    p0.backward(G)
    dx, de_c, de_u = ops.ddim_pred_x0_vjp(e_c.detach(), e_u.detach(), G, coef)
    # fp32 end to end: tolerance 1e-4 relative L2
    assert _rel(dx, x.grad) < 1e-6 and _rel(de_c, e_c.grad) < 1e-4 and _rel(de_u, e_u.grad) < 1e-4


@pytest.mark.parametrize("mc,t,h,w", [(64, 5, 16, 16), (64, 3, 16, 24)])
def test_unet_input_gradient_vs_reference(mc, t, h, w):
    import unet_ref
    from vc_b200.unet import UNetB200

    if not unet_ref.ref_available():
        pytest.skip("oracle/_ref/ViewCrafter not installed")
    ref, cfg = unet_ref.build_reference_unet(model_channels=mc)
    ours = UNetB200(ref.state_dict(), device="cuda", **cfg)
    x, cc, ctx, _ = unet_ref.synth_inputs(t, h, w)
    xin = torch.cat([x, cc], 1)
    ts, fs = torch.tensor([481], device="cuda"), torch.tensor([10], device="cuda")
    g = torch.randn(1, 4, t, h, w, generator=torch.Generator().manual_seed(5)).cuda()
    grads = {}
    for name in ("fp32", "bf16"):
        xr = xin.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=BF, enabled=(name == "bf16")):
            y = ref(xr, ts, context=ctx, fs=fs)
        y.backward(gradient=g.to(y.dtype), inputs=[xr])
        grads[name] = xr.grad
    xo = xin.clone().requires_grad_(True)
    y = ours.forward_with_grad(xo, ts, ctx, fs=fs)
    y.backward(gradient=g.to(y.dtype), inputs=[xo])
    torch.cuda.synchronize()
    e_ours, e_ref = _rel(xo.grad, grads["fp32"]), _rel(grads["bf16"], grads["fp32"])
    print(f"input-gradient rel L2: ours vs fp32 {e_ours:.3e}, ref-bf16 vs fp32 {e_ref:.3e}")
    assert torch.isfinite(xo.grad).all()
    assert e_ours <= 1.25 * e_ref + 5e-3


def test_dropin_routes_guided_calls_to_native_when_enabled(monkeypatch):
    import unet_ref
    from vc_b200.dropin import B200UNet

    if not unet_ref.ref_available():
        pytest.skip("oracle/_ref/ViewCrafter not installed")
    monkeypatch.setenv("GVD_GUIDED_NATIVE", "1")
    ref, _ = unet_ref.build_reference_unet(model_channels=64)
    ref.requires_grad_(True)  # what the guided sampler does to the wrapped module (ddim_guidance.py:256)
    mod = B200UNet(ref)
    x, cc, ctx, _ = unet_ref.synth_inputs(3, 16, 16)
    xg = torch.cat([x, cc], 1).requires_grad_(True)
    ts, fs = torch.tensor([300], device="cuda"), torch.tensor([10], device="cuda")
    y = mod(xg, ts, context=ctx, fs=fs)
    assert type(y.grad_fn).__name__ != "ConvolutionBackward0"
    y.float().sum().backward(inputs=[xg])
    xr = xg.detach().clone().requires_grad_(True)
    ref(xr, ts, context=ctx, fs=fs).sum().backward(inputs=[xr])
    assert _rel(xg.grad, xr.grad) < 0.15


def test_vae_decoder_forward_and_latent_gradient_vs_reference():
    """vc_b200.vae.DecoderB200 (full-width VAE, ch 128) on the GPU vs the reference Decoder in fp32 and under bf16 autocast."""
    import test_vae_cpu as tv
    from vc_b200.vae import DecoderB200

    if not tv.HAVE:
        pytest.skip("oracle/_ref/ViewCrafter/.../ae_modules.py not installed")
    vae = tv.RefFirstStage(ch=128).cuda().eval()
    ours = DecoderB200(vae.state_dict(), device="cuda", scale_factor=tv.SCALE)
    g = torch.Generator().manual_seed(3)
    z = (torch.randn(2, 4, 20, 32, generator=g) * tv.SCALE * 3).cuda()
    cot = torch.randn(2, 3, 160, 256, generator=g).cuda()
    res = {}
    for name in ("fp32", "bf16"):
        zr = z.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=BF, enabled=(name == "bf16")):
            y = vae(zr)
        y.float().backward(cot)
        res[name] = (y.detach().float(), zr.grad)
    y_inf = ours.decode(z)
    zo = z.clone().requires_grad_(True)
    yo = ours.differentiable_decode(zo)
    yo.backward(cot)
    torch.cuda.synchronize()
    assert torch.equal(y_inf, yo.detach())  # the tape does not change the forward
    e_y, e_y_ref = _rel(yo, res["fp32"][0]), _rel(res["bf16"][0], res["fp32"][0])
    e_g, e_g_ref = _rel(zo.grad, res["fp32"][1]), _rel(res["bf16"][1], res["fp32"][1])
    print(f"VAE decoder: image {e_y:.2e} (reference autocast {e_y_ref:.2e}); latent gradient {e_g:.2e} (reference autocast {e_g_ref:.2e})")
    assert e_y <= 1.25 * e_y_ref + 2e-3 and e_g <= 1.25 * e_g_ref + 5e-3


def test_guided_step_native_vs_reference_sampler_on_gpu(monkeypatch):
    """One guided DDIM step, everything native (U-Net fwd + dX, VAE decoder fwd + dX, fused DDIM update and its VJP),
    against the reference DDIMSamplerGuidance over the reference modules under bf16 autocast on the same GPU.  The step is
    x_prev = plain_step - rho * grad with rho normalising the gradient's RMS, so bf16 noise in the gradient enters
    x_prev scaled by 0.2 * sgw * cfg * rms(e_c - e_u): compared on x_prev with the bf16 noise bound of the forward test."""
    import test_guided_cpu as tg
    import test_vae_cpu as tv
    import unet_ref
    from vc_b200.guided import DDIMSamplerGuidance
    from vc_b200.schedule import ModelSchedule
    from vc_b200.unet import DiffusionModelB200, UNetB200
    from vc_b200.vae import DecoderB200

    if not (unet_ref.ref_available() and tv.HAVE):
        pytest.skip("oracle/_ref/ViewCrafter not installed")
    ref, cfg = unet_ref.build_reference_unet(model_channels=64)
    vae = tv.RefFirstStage().cuda().eval()
    T, h, w, index = 3, 16, 16, 30
    x, cc, ctx, ctx_uc = unet_ref.synth_inputs(T, h, w)
    cond, uc = {"c_concat": [cc], "c_crossattn": [ctx]}, {"c_concat": [cc], "c_crossattn": [ctx_uc]}
    fs = torch.tensor([10], device="cuda")
    g = torch.Generator().manual_seed(123)
    targets = [(torch.rand(3, 8 * h, 8 * w, generator=g) * 2 - 1).cuda() for _ in range(T)]
    masks = [(torch.rand(1, 8 * h, 8 * w, generator=g) > 0.3).float().cuda() for _ in range(T)]
    noises = [torch.randn(x.shape, generator=g).cuda() for _ in range(2)]

    class PerFrame(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.vae = vae

        def forward(self, z):
            return torch.stack([self.vae(z[:, :, f])[0] for f in range(z.shape[2])], dim=1).unsqueeze(0)

    sampler_ref, dg = tg._reference_sampler(ref, PerFrame())
    for name, val in list(vars(sampler_ref.model).items()):
        if isinstance(val, torch.Tensor):
            setattr(sampler_ref.model, name, val.cuda())
    sampler_ref.model.device = torch.device("cuda")
    sampler_ref.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0, verbose=False)
    ts = torch.full((1,), int(sampler_ref.ddim_timesteps[index]), dtype=torch.long, device="cuda")
    outs = {}
    for name in ("fp32", "bf16"):
        queue = list(noises)
        monkeypatch.setattr(dg, "noise_like", lambda shape, device, repeat=False: queue.pop(0))
        with torch.autocast("cuda", dtype=BF, enabled=(name == "bf16")):
            outs[name] = sampler_ref.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5,
                                                   unconditional_conditioning=uc, guidance_rescale=0.7, fs=fs,
                                                   loss_guidance_fn=tg.StubGuidance(targets, masks, 1))[0].float()
    model = DiffusionModelB200(UNetB200(ref.state_dict(), device="cuda", **cfg), ModelSchedule())
    dec = DecoderB200(vae.state_dict(), device="cuda", scale_factor=tv.SCALE)
    model.differentiable_decode_first_stage = dec.differentiable_decode
    model.guided_decode_frames = 3
    sampler = DDIMSamplerGuidance(model)
    sampler.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0)
    xp, _ = sampler.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5, unconditional_conditioning=uc,
                                  guidance_rescale=0.7, fs=fs, loss_guidance_fn=tg.StubGuidance(targets, masks, 1),
                                  noise=noises[0:1], recur_noise=noises[1:2])
    torch.cuda.synchronize()
    e_ours, e_ref = _rel(xp, outs["fp32"]), _rel(outs["bf16"], outs["fp32"])
    print(f"guided x_prev rel L2: ours vs fp32 {e_ours:.3e}, reference-bf16 vs fp32 {e_ref:.3e}")
    assert e_ours <= 1.25 * e_ref + 5e-3


def test_vae_encoder_moments_vs_reference():
    """vc_b200.vae.EncoderB200 (full width) on the GPU vs the reference Encoder + quant_conv in fp32 / bf16 autocast."""
    import test_vae_cpu as tv
    from vc_b200.vae import EncoderB200

    if not tv.HAVE:
        pytest.skip("oracle/_ref/ViewCrafter/.../ae_modules.py not installed")
    ref = tv.RefEncoderStage(ch=128).cuda().eval()
    ours = EncoderB200(ref.state_dict(), device="cuda")
    x = (torch.rand(3, 3, 160, 256, generator=torch.Generator().manual_seed(4)) * 2 - 1).cuda()
    with torch.no_grad():
        m32 = ref(x)
        with torch.autocast("cuda", dtype=BF):
            mbf = ref(x).float()
    m = ours.moments(x)
    torch.cuda.synchronize()
    e, e_ref = _rel(m, m32), _rel(mbf, m32)
    print(f"VAE encoder moments: ours vs fp32 {e:.2e}, reference autocast vs fp32 {e_ref:.2e}")
    assert m.shape == (3, 8, 20, 32) and e <= 1.25 * e_ref + 2e-3
