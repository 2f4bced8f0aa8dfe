"""Loss-guided DDIM step over the B200-native denoiser -- the call surface of the reference's
lvdm/models/samplers/ddim_guidance.py::DDIMSamplerGuidance.p_sample_ddim (:205-362, Algorithm 1 of the paper) for the
options the guidedvd pipeline uses (VC/utils_vc/diffusion_utils.py: v-prediction, CFG 7.5, guidance_rescale 0.7,
eta 1.0, dynamic rescale, `loss_guidance_fn` = utils/viewcrafter_wrapper.py::LossGuidance, recur_steps 1 or 2).

One guided step:
  1. x.requires_grad_(True); e_cond, e_uncond = U-Net(x) with the tape on (vc_b200.grad; both graphs are kept, like
     the reference keeps them -- rescale_noise_cfg couples the two outputs through their standard deviations);
  2. x_prev, pred_x0 = the fused DDIM update (gvd_ddim_step, same arithmetic as the plain sampler);
  3. per frame: decode pred_x0 through the VAE decoder with grad, evaluate the guidance loss against the 3DGS
     rendering, G_f = d loss / d pred_x0_f (/ numel unless mean_loss)                         (:282-302)
  4. dL/dx = gvd_ddim_pred_x0_vjp (explicit x term, cotangents of e_cond / e_uncond) + the U-Net backward of both
     forwards                                                                                  (:309-311)
  5. rho = rms(e_cond - e_uncond) * cfg / rms(dL/dx) * 0.2 * scale_guidance_weight; x_prev -= rho * dL/dx (:318-326)
  6. (recurrence, :334) x = sqrt(beta_t) x_prev + sqrt(1 - beta_t) N(0, 1), beta_t = a_t / a_prev.

`model` protocol beyond vc_b200.sampler.DDIMSampler's: `model.differentiable_decode_first_stage(z[1,4,n,h,w]) ->
[1,3,n,H,W]` (ddpm3d.py:674-675; vc_b200.vae.DecoderB200.differentiable_decode is the native one) and, optionally,
`model.guided_decode_frames` = how many frames to decode per call (default 1, the reference's loop; frames are a batch
dimension of the decoder, so any chunk size gives the same gradients).  `loss_guidance_fn` protocol: SURVEY.md 8b.
"""
import numpy as np
import torch

from . import ops
from .sampler import DDIMSampler, _randn


def _rms(t):
    return (t.float() * t.float()).mean().sqrt().item()


class DDIMSamplerGuidance(DDIMSampler):
    def p_sample_ddim(self, x, c, t, index, unconditional_guidance_scale=1., unconditional_conditioning=None,
                      guidance_rescale=0.0, temperature=1., noise=None, loss_guidance_fn=None, recur_noise=None, **kwargs):
        """noise / recur_noise: optional pre-drawn N(0,1) tensors (one per recurrence: lists or a single tensor) so a
        test can replay the reference's draws; drawn fresh otherwise, in the reference's order."""
        if loss_guidance_fn is None:
            return super().p_sample_ddim(x, c, t, index, unconditional_guidance_scale, unconditional_conditioning,
                                         guidance_rescale, temperature, noise, **kwargs)
        if x.shape[0] != 1:
            raise ValueError("guided sampling supports batch size 1 (ddim_guidance.py:246-247)")
        lg = loss_guidance_fn
        repeat = int(lg.recur_steps)
        if repeat not in (1, 2):
            raise ValueError("only support 1 or 2 recur steps (ddim_guidance.py:250-251)")
        sgw = 1.0
        if lg.scale_guidance_weight:
            sgw = lg.guidance_weight_fn(lg.current_train_iter)
        cfg = float(unconditional_guidance_scale)
        coef = self.ddim.coefficients(index, cfg, guidance_rescale, temperature)
        a_t, a_prev = np.float32(self.ddim.alphas[index]), np.float32(self.ddim.alphas_prev[index])
        beta_t = np.float32(a_t / a_prev)
        uc = unconditional_conditioning
        model = self.model
        n_frames = x.shape[2]
        pick = lambda src, j: None if src is None else (src[j] if isinstance(src, (list, tuple)) else src)  # noqa: E731
        x_prev = pred_x0 = None
        for j in range(repeat):
            x = x.detach().float().requires_grad_(True)
            with torch.enable_grad():
                e_c = model.apply_model(x, t, c, **kwargs)
                e_u = model.apply_model(x, t, uc, **kwargs) if uc is not None else None
            nz = pick(noise, j)
            if nz is None:
                nz = _randn(model, x.shape, x.device)
            e_cd = e_c.detach().float().contiguous()
            e_ud = None if e_u is None else e_u.detach().float().contiguous()
            x_prev, pred_x0 = ops.ddim_step(x.detach().contiguous(), e_cd, e_ud, nz.float().contiguous(), coef)
            grads, decoded = [], []
            chunk = max(1, int(getattr(model, "guided_decode_frames", 1)))
            for f0 in range(0, n_frames, chunk):
                f1 = min(n_frames, f0 + chunk)
                z = pred_x0[:, :, f0:f1].clone().requires_grad_(True)  # the decoder graph ends here (:285)
                with torch.enable_grad():
                    d_x0 = model.differentiable_decode_first_stage(z)
                    # frames are independent through the decoder, so one backward over sum_f loss_f / numel_f gives
                    # every frame the gradient the reference's one-frame-at-a-time loop gives it
                    total = None
                    for f in range(f0, f1):
                        loss_dict, numel = lg(d_x0[0][:, f - f0:f - f0 + 1], index, f, f + 1)
                        term = loss_dict["recon"] if lg.mean_loss else loss_dict["recon"] / numel
                        total = term if total is None else total + term
                    g = torch.autograd.grad(outputs=total, inputs=z)[0]
                grads.append(g.detach())
                decoded.append(d_x0.detach())
            lg.save_pred_x0(torch.cat(decoded, dim=2), index)
            G = torch.cat(grads, dim=2).float().contiguous()
            dx, de_c, de_u = ops.ddim_pred_x0_vjp(e_cd, e_ud, G, coef)
            if e_u is None:
                torch.autograd.backward([e_c], [de_c.to(e_c.dtype)], inputs=[x])
            else:
                torch.autograd.backward([e_c, e_u], [de_c.to(e_c.dtype), de_u.to(e_u.dtype)], inputs=[x])
            guided = x.grad.detach().float() + dx
            x.grad = None
            tmp_s = _rms(guided)
            rho = 0.0
            if tmp_s != 0:
                corr = _rms(e_cd - e_ud) if e_ud is not None else 0.0
                rho = corr * cfg / tmp_s * (0.2 * sgw)
            x_prev = x_prev - rho * guided
            rz = pick(recur_noise, j)
            if rz is None:
                rz = _randn(model, x.shape, x.device)
            x = float(np.sqrt(beta_t)) * x_prev + float(np.sqrt(np.float32(1) - beta_t)) * rz
        return x_prev.detach(), pred_x0.detach()
