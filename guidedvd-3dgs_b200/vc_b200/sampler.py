"""DDIM sampler over the B200-native denoiser -- same call surface as the reference's
lvdm/models/samplers/ddim.py::DDIMSampler for the options the guidedvd pipeline uses
(VC/utils_vc/diffusion_utils.py:180-206: S steps, eta, cfg scale, guidance_rescale, 'uniform_trailing', fs).

`model` protocol (what ddim.py uses of the LatentDiffusion object):
    model.apply_model(x[b,4,t,h,w] fp32, t[b] int64, cond, **kwargs) -> v-prediction, same shape
    model.schedule: vc_b200.schedule.ModelSchedule
The per-step arithmetic (CFG mix, rescale_noise_cfg, v->eps/x0, dynamic rescale, stochastic DDIM update) is ONE fused
launch pair in the sm_100a library (gvd_ddim_step); the schedule tables stay on the host in float64/float32 exactly as
the reference builds them.
"""
import numpy as np
import torch

from . import ops
from .schedule import DdimSchedule


def _randn(model, shape, device):
    """Fresh Gaussian noise; under a multi-GPU plan rank 0 draws it and broadcasts, so every rank steps the same latent."""
    z = torch.randn(shape, device=device)
    plan = getattr(model, "plan", None)
    if plan is not None and plan.world > 1:
        torch.distributed.broadcast(z, src=0)
    return z


class DDIMSampler:
    def __init__(self, model, schedule="linear", **kwargs):
        self.model = model
        self.ddpm_num_timesteps = model.schedule.num_timesteps
        self.schedule = schedule
        self.counter = 0

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        self.ddim = DdimSchedule(self.model.schedule, ddim_num_steps, ddim_discretize, ddim_eta)
        self.ddim_timesteps = self.ddim.timesteps

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, eta=0., x_T=None, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, verbose=False, timestep_spacing="uniform", guidance_rescale=0.0,
               temperature=1., noises=None, device="cuda", **kwargs):
        self.make_schedule(S, ddim_discretize=timestep_spacing, ddim_eta=eta, verbose=verbose)
        size = (batch_size, *shape)
        img = _randn(self.model, size, device) if x_T is None else x_T
        steps = self.ddim_timesteps.shape[0]
        inter = {"x_inter": [img], "pred_x0": [img]}
        for i, step in enumerate(np.flip(self.ddim_timesteps)):
            index = steps - i - 1
            ts = torch.full((batch_size,), int(step), device=device, dtype=torch.long)
            noise = noises[i] if noises is not None else None
            img, pred_x0 = self.p_sample_ddim(img, conditioning, ts, index=index,
                                              unconditional_guidance_scale=unconditional_guidance_scale,
                                              unconditional_conditioning=unconditional_conditioning,
                                              guidance_rescale=guidance_rescale, temperature=temperature, noise=noise,
                                              **kwargs)
        inter["x_inter"].append(img)
        inter["pred_x0"].append(pred_x0)
        return img, inter

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, unconditional_guidance_scale=1., unconditional_conditioning=None,
                      guidance_rescale=0.0, temperature=1., noise=None, **kwargs):
        if x.shape[0] != 1:
            # rescale_noise_cfg takes its std per batch item (utils_diffusion.py:152-153)
            outs = [self.p_sample_ddim(x[i:i + 1], _index_cond(c, i), t[i:i + 1], index, unconditional_guidance_scale,
                                       _index_cond(unconditional_conditioning, i), guidance_rescale, temperature,
                                       None if noise is None else noise[i:i + 1], **kwargs) for i in range(x.shape[0])]
            return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])
        e_u = None
        if unconditional_conditioning is not None and unconditional_guidance_scale != 1.:
            if hasattr(self.model, "apply_model_cfg"):  # lets a multi-GPU plan run the two forwards on different ranks
                e_c, e_u = self.model.apply_model_cfg(x, t, c, unconditional_conditioning, **kwargs)
            else:
                e_c = self.model.apply_model(x, t, c, **kwargs)
                e_u = self.model.apply_model(x, t, unconditional_conditioning, **kwargs)
        else:
            e_c = self.model.apply_model(x, t, c, **kwargs)
        if noise is None:
            noise = _randn(self.model, x.shape, x.device)
        coef = self.ddim.coefficients(index, unconditional_guidance_scale, guidance_rescale, temperature)
        return ops.ddim_step(x.float().contiguous(), e_c.float().contiguous(),
                             None if e_u is None else e_u.float().contiguous(), noise.float().contiguous(), coef)


def _index_cond(c, i):
    if c is None:
        return None
    if isinstance(c, dict):
        return {k: [v[i:i + 1] for v in vs] for k, vs in c.items()}
    return c[i:i + 1]
