"""The guided DDIM step (vc_b200.guided.DDIMSamplerGuidance) against the REFERENCE sampler
(oracle/_ref/ViewCrafter/lvdm/models/samplers/ddim_guidance.py::DDIMSamplerGuidance.p_sample_ddim, :205-362) on the
CPU: same U-Net weights (reference UNetModel in fp32 vs UNetB200 over the C-ABI stand-in of tests/fake_nn_lib.py), same
stub VAE decoder, same LossGuidance stand-in, same noise draws.  Also pins the closed-form pred_x0 VJP
(gvd_ddim_pred_x0_vjp) against autograd over the reference's own rescale_noise_cfg / predict_start arithmetic.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "guidedvd-3dgs_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import unet_ref  # noqa: E402
from test_unet_grad_cpu import _rel, install_fake  # noqa: E402

needs_ref = pytest.mark.skipif(not unet_ref.ref_available(), reason="oracle/_ref/ViewCrafter not installed (python oracle/build_ref.py vc)")


class StubDecoder(torch.nn.Module):
    """Stands in for first_stage_model.decode: latent [1,4,1,h,w] -> image [1,3,1,2h,2w], smooth and non-linear."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(77)
        self.w1 = torch.nn.Parameter(torch.randn(8, 4, 3, 3, generator=g) * 0.3)
        self.w2 = torch.nn.Parameter(torch.randn(3, 8, 3, 3, generator=g) * 0.3)

    def forward(self, z):
        x = z[0].permute(1, 0, 2, 3)  # frames as batch
        x = torch.nn.functional.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        x = torch.tanh(torch.nn.functional.conv2d(x, self.w1, padding=1))
        return torch.tanh(torch.nn.functional.conv2d(x, self.w2, padding=1)).permute(1, 0, 2, 3).unsqueeze(0)


class StubGuidance:
    """The LossGuidance protocol (utils/viewcrafter_wrapper.py:100-160): masked L2 against target frames."""
    verbose = False
    scale_guidance_weight = True
    mean_loss = False
    current_train_iter = 1200

    def __init__(self, targets, masks, recur_steps):
        self.targets, self.masks, self.recur_steps = targets, masks, recur_steps
        self.saved = []

    def guidance_weight_fn(self, it):
        return 0.5 + it / 4000.0

    def __call__(self, d_x0, index, f0, f1):
        m = self.masks[f0]
        loss = (((d_x0[:, 0] - self.targets[f0]) ** 2) * m).sum()
        return {"recon": loss}, float(m.sum().item() * 3)

    def save_pred_x0(self, x, index):
        self.saved.append((index, x.detach().clone()))


def _reference_sampler(ref_unet, decoder):
    if unet_ref.REF_VC not in sys.path:
        sys.path.insert(0, unet_ref.REF_VC)
    import lvdm.models.samplers.ddim_guidance as dg
    from lvdm.models.utils_diffusion import make_beta_schedule, rescale_zero_terminal_snr

    class Wrapper(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.diffusion_model = m

    class Model:  # the slice of LatentDiffusion the sampler touches (ddpm3d.py:123-151,239-251,519-527,674-675)
        parameterization = "v"
        use_dynamic_rescale = True
        device = torch.device("cpu")
        num_timesteps = 1000

        def __init__(self):
            betas = rescale_zero_terminal_snr(make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012))
            ac = np.cumprod(1. - betas, axis=0)
            t32 = lambda a: torch.tensor(a, dtype=torch.float32)  # noqa: E731
            self.betas, self.alphas_cumprod = t32(betas), t32(ac)
            self.alphas_cumprod_prev = t32(np.append(1., ac[:-1]))
            self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod = t32(np.sqrt(ac)), t32(np.sqrt(1. - ac))
            self.scale_arr = t32(np.concatenate((np.linspace(1.0, 0.3, 400), np.full(1000, 0.3))))
            self.model, self.first_stage_model = Wrapper(ref_unet), decoder

        def apply_model(self, x, t, c, fs=None, **kw):
            return ref_unet(torch.cat([x] + c["c_concat"], 1), t, context=torch.cat(c["c_crossattn"], 1), fs=fs)

        def predict_start_from_z_and_v(self, x_t, t, v):
            return self.sqrt_alphas_cumprod[t].view(-1, 1, 1, 1, 1) * x_t - self.sqrt_one_minus_alphas_cumprod[t].view(-1, 1, 1, 1, 1) * v

        def predict_eps_from_z_and_v(self, x_t, t, v):
            return self.sqrt_alphas_cumprod[t].view(-1, 1, 1, 1, 1) * v + self.sqrt_one_minus_alphas_cumprod[t].view(-1, 1, 1, 1, 1) * x_t

        def differentiable_decode_first_stage(self, z, **kw):
            return decoder(z)

    dg.DDIMSamplerGuidance.register_buffer = lambda self, name, attr: setattr(self, name, attr)  # keep buffers on the CPU
    s = dg.DDIMSamplerGuidance(Model())
    s.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0, verbose=False)
    return s, dg


@needs_ref
@pytest.mark.parametrize("index,recur,decode_frames", [(40, 1, 1), (12, 2, 2), (25, 1, 3)])
def test_guided_step_matches_reference_sampler(monkeypatch, index, recur, decode_frames):
    from vc_b200.guided import DDIMSamplerGuidance
    from vc_b200.schedule import ModelSchedule
    from vc_b200.unet import DiffusionModelB200, UNetB200

    install_fake(monkeypatch)
    ref, cfg = unet_ref.build_reference_unet(model_channels=64, device="cpu")
    T, h, w = 3, 8, 8
    x, cc, ctx, ctx_uc = unet_ref.synth_inputs(T, h, w, device="cpu")
    cond = {"c_concat": [cc], "c_crossattn": [ctx]}
    uc = {"c_concat": [cc], "c_crossattn": [ctx_uc]}
    fs = torch.tensor([10])
    g = torch.Generator().manual_seed(99)
    targets = [torch.rand(3, 2 * h, 2 * w, generator=g) * 2 - 1 for _ in range(T)]
    masks = [(torch.rand(1, 2 * h, 2 * w, generator=g) > 0.3).float() for _ in range(T)]
    noises = [torch.randn(x.shape, generator=g) for _ in range(2 * recur)]  # reference order: step noise, recurrence noise, ...
    decoder = StubDecoder()

    # ---- reference
    sampler_ref, dg = _reference_sampler(ref, decoder)
    queue = list(noises)
    monkeypatch.setattr(dg, "noise_like", lambda shape, device, repeat=False: queue.pop(0))
    lg_ref = StubGuidance(targets, masks, recur)
    ts = torch.full((1,), int(sampler_ref.ddim_timesteps[index]), dtype=torch.long)
    xp_ref, p0_ref = sampler_ref.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5,
                                               unconditional_conditioning=uc, guidance_rescale=0.7, fs=fs,
                                               loss_guidance_fn=lg_ref)
    assert not queue

    # ---- ours
    model = DiffusionModelB200(UNetB200(ref.state_dict(), device="cpu", **cfg), ModelSchedule())
    model.differentiable_decode_first_stage = decoder
    model.guided_decode_frames = decode_frames  # the reference decodes one frame per call; any chunking must agree
    sampler = DDIMSamplerGuidance(model)
    sampler.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0)
    lg = StubGuidance(targets, masks, recur)
    xp, p0 = sampler.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5, unconditional_conditioning=uc,
                                   guidance_rescale=0.7, fs=fs, loss_guidance_fn=lg, noise=noises[0::2], recur_noise=noises[1::2])
    e_xp, e_p0 = _rel(xp, xp_ref), _rel(p0, p0_ref)
    print(f"index {index} recur {recur}: x_prev rel {e_xp:.2e}, pred_x0 rel {e_p0:.2e}")
    assert e_xp < 2e-4 and e_p0 < 2e-4
    # the guidance actually moved the sample (otherwise the comparison above would not see the gradient path)
    plain, _ = super(DDIMSamplerGuidance, sampler).p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5,
                                                                   unconditional_conditioning=uc, guidance_rescale=0.7, fs=fs,
                                                                   noise=noises[0])
    if recur == 1:
        assert _rel(xp, plain) > 1e-3
    assert len(lg.saved) == recur and len(lg_ref.saved) == recur
    assert _rel(lg.saved[-1][1], lg_ref.saved[-1][1]) < 2e-4


def test_pred_x0_vjp_against_autograd(monkeypatch):
    """gvd_ddim_pred_x0_vjp's closed form vs autograd over rescale_noise_cfg + predict_start (+ dynamic rescale)."""
    from vc_b200 import ops
    from vc_b200.schedule import DdimSchedule, ModelSchedule

    install_fake(monkeypatch)
    g = torch.Generator().manual_seed(4)
    shape = (1, 4, 3, 6, 5)
    coef = DdimSchedule(ModelSchedule(), 50, "uniform_trailing", 1.0).coefficients(30, 7.5, 0.7, 1.0)
    x = torch.randn(shape, generator=g, dtype=torch.float64).requires_grad_(True)
    e_c = torch.randn(shape, generator=g, dtype=torch.float64).requires_grad_(True)
    e_u = (e_c.detach() + 0.3 * torch.randn(shape, generator=g, dtype=torch.float64)).requires_grad_(True)
    G = torch.randn(shape, generator=g, dtype=torch.float64)
    mo = e_u + 7.5 * (e_c - e_u)
    std_c, std_m = e_c.std(dim=(1, 2, 3, 4), keepdim=True), mo.std(dim=(1, 2, 3, 4), keepdim=True)
    v = 0.7 * (mo * (std_c / std_m)) + 0.3 * mo   # utils_diffusion.py:147-158
    p0 = (coef["sqrt_alphas_cumprod_t"] * x - coef["sqrt_one_minus_alphas_cumprod_t"] * v) * (coef["scale_prev"] / coef["scale_t"])
    p0.backward(G)
    dx, de_c, de_u = ops.ddim_pred_x0_vjp(e_c.detach().float(), e_u.detach().float(), G.float(), coef)
    assert _rel(dx, x.grad) < 1e-6 and _rel(de_c, e_c.grad) < 1e-5 and _rel(de_u, e_u.grad) < 1e-5
    # no unconditional branch: v = e_c
    x.grad = e_c.grad = None
    p0 = (coef["sqrt_alphas_cumprod_t"] * x - coef["sqrt_one_minus_alphas_cumprod_t"] * e_c) * (coef["scale_prev"] / coef["scale_t"])
    p0.backward(G)
    dx, de_c, de_u = ops.ddim_pred_x0_vjp(e_c.detach().float(), None, G.float(), coef)
    assert de_u is None and _rel(dx, x.grad) < 1e-6 and _rel(de_c, e_c.grad) < 1e-6


@needs_ref
def test_guided_step_with_native_vae_decoder(monkeypatch):
    """The whole guided step native -- U-Net forward + input-gradient AND the VAE decoder forward + latent-gradient
    (vc_b200.vae.DecoderB200, three frames per decoder call) -- against the reference sampler driving the reference
    UNetModel and the reference Decoder one frame at a time (ddpm3d.py:646-667)."""
    import test_vae_cpu as tv
    if not tv.HAVE:
        pytest.skip("ae_modules.py not installed")
    from vc_b200.guided import DDIMSamplerGuidance
    from vc_b200.schedule import ModelSchedule
    from vc_b200.unet import DiffusionModelB200, UNetB200
    from vc_b200.vae import DecoderB200

    install_fake(monkeypatch)
    ref, cfg = unet_ref.build_reference_unet(model_channels=64, device="cpu")
    vae = tv.RefFirstStage().eval()
    T, h, w, index = 3, 8, 8, 30
    x, cc, ctx, ctx_uc = unet_ref.synth_inputs(T, h, w, device="cpu")
    cond, uc = {"c_concat": [cc], "c_crossattn": [ctx]}, {"c_concat": [cc], "c_crossattn": [ctx_uc]}
    fs = torch.tensor([10])
    g = torch.Generator().manual_seed(123)
    targets = [torch.rand(3, 8 * h, 8 * w, generator=g) * 2 - 1 for _ in range(T)]
    masks = [(torch.rand(1, 8 * h, 8 * w, generator=g) > 0.3).float() for _ in range(T)]
    noises = [torch.randn(x.shape, generator=g) for _ in range(2)]

    class PerFrame(torch.nn.Module):  # decode_core: frames one at a time through AutoencoderKL.decode
        def __init__(self):
            super().__init__()
            self.vae = vae

        def forward(self, z):
            return torch.stack([self.vae(z[:, :, f])[0] for f in range(z.shape[2])], dim=1).unsqueeze(0)

    sampler_ref, dg = _reference_sampler(ref, PerFrame())
    queue = list(noises)
    monkeypatch.setattr(dg, "noise_like", lambda shape, device, repeat=False: queue.pop(0))
    ts = torch.full((1,), int(sampler_ref.ddim_timesteps[index]), dtype=torch.long)
    xp_ref, p0_ref = sampler_ref.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5,
                                               unconditional_conditioning=uc, guidance_rescale=0.7, fs=fs,
                                               loss_guidance_fn=StubGuidance(targets, masks, 1))

    model = DiffusionModelB200(UNetB200(ref.state_dict(), device="cpu", **cfg), ModelSchedule())
    dec = DecoderB200(vae.state_dict(), device="cpu", scale_factor=tv.SCALE)
    model.differentiable_decode_first_stage = dec.differentiable_decode
    model.guided_decode_frames = 3
    sampler = DDIMSamplerGuidance(model)
    sampler.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0)
    lg = StubGuidance(targets, masks, 1)
    xp, p0 = sampler.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5, unconditional_conditioning=uc,
                                   guidance_rescale=0.7, fs=fs, loss_guidance_fn=lg, noise=noises[0:1], recur_noise=noises[1:2])
    print(f"fully native guided step: x_prev rel {_rel(xp, xp_ref):.2e}")
    assert _rel(xp, xp_ref) < 2e-4 and _rel(p0, p0_ref) < 2e-4
    assert lg.saved[0][1].shape == (1, 3, T, 8 * h, 8 * w)


@needs_ref
def test_reference_sampler_runs_unchanged_over_the_dropins(monkeypatch):
    """INTEGRATION.md 'Guided sampling': the REFERENCE DDIMSamplerGuidance, untouched, over a model whose U-Net is
    vc_b200.dropin.B200UNet and whose first_stage_model.decode was replaced, with GVD_GUIDED_NATIVE=1 -- same x_prev as
    the all-reference run (the sampler's own torch arithmetic and autograd glue around our tapes)."""
    import test_vae_cpu as tv
    if not tv.HAVE:
        pytest.skip("ae_modules.py not installed")
    from vc_b200.dropin import B200UNet, replace_first_stage_decoder

    fake = install_fake(monkeypatch)
    ref, _ = unet_ref.build_reference_unet(model_channels=64, device="cpu")
    vae = tv.RefFirstStage().eval()
    T, h, w, index = 2, 8, 8, 20
    x, cc, ctx, ctx_uc = unet_ref.synth_inputs(T, h, w, device="cpu")
    cond, uc = {"c_concat": [cc], "c_crossattn": [ctx]}, {"c_concat": [cc], "c_crossattn": [ctx_uc]}
    fs = torch.tensor([10])
    g = torch.Generator().manual_seed(321)
    targets = [torch.rand(3, 8 * h, 8 * w, generator=g) * 2 - 1 for _ in range(T)]
    masks = [(torch.rand(1, 8 * h, 8 * w, generator=g) > 0.3).float() for _ in range(T)]
    noises = [torch.randn(x.shape, generator=g) for _ in range(2)]

    class FirstStage(torch.nn.Module):  # AutoencoderKL.decode + decode_core's per-frame loop and 1/scale_factor
        def __init__(self):
            super().__init__()
            self.decoder, self.post_quant_conv = vae.decoder, vae.post_quant_conv

        def decode(self, z, **kwargs):
            return self.decoder(self.post_quant_conv(z))

        def forward(self, z):
            return torch.stack([self.decode(z[:, :, f] / tv.SCALE)[0] for f in range(z.shape[2])], dim=1).unsqueeze(0)

    results = []
    for native in (False, True):
        fs_model = FirstStage()
        sampler, dg = _reference_sampler(ref, fs_model)
        if native:
            monkeypatch.setenv("GVD_GUIDED_NATIVE", "1")
            unet = B200UNet(ref)
            sampler.model.model.diffusion_model = unet
            sampler.model.apply_model = lambda xx, t, c, fs=None, **kw: unet(torch.cat([xx] + c["c_concat"], 1), t,
                                                                             context=torch.cat(c["c_crossattn"], 1), fs=fs)
            replace_first_stage_decoder(sampler.model)
        queue = list(noises)
        monkeypatch.setattr(dg, "noise_like", lambda shape, device, repeat=False: queue.pop(0))
        ts = torch.full((1,), int(sampler.ddim_timesteps[index]), dtype=torch.long)
        results.append(sampler.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5,
                                             unconditional_conditioning=uc, guidance_rescale=0.7, fs=fs,
                                             loss_guidance_fn=StubGuidance(targets, masks, 1))[0])
    assert fake.calls.get("groupnorm_bwd", 0) > 0 and fake.calls.get("col2im3x3", 0) > 0
    assert _rel(results[1], results[0]) < 2e-4
