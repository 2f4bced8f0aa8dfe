"""Generate tests/golden/ddim_schedule.npz and ddim_steps.npz by running the REFERENCE sampler
(third_party/ViewCrafter/lvdm/models/samplers/ddim.py::DDIMSampler) in this container on the CPU.

The reference's LatentDiffusion class needs pytorch_lightning (absent), so the sampler is driven with a stub model
whose schedule buffers are built by calling the reference's own helpers (make_beta_schedule, rescale_zero_terminal_snr)
exactly as DDPM.register_schedule does (lvdm/models/ddpm3d.py:123-151,519-527) and whose apply_model replays seeded
tensors.  Run:  python tests/make_golden_ddim.py   (needs /root/reference; not needed on the GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference/third_party/ViewCrafter")
from lvdm.models.samplers.ddim import DDIMSampler  # noqa: E402
from lvdm.models.utils_diffusion import make_beta_schedule, rescale_zero_terminal_snr  # noqa: E402
import lvdm.models.samplers.ddim as ddim_mod  # noqa: E402


class StubModel:
    """Just the attributes/methods ddim.py touches."""
    parameterization = "v"
    use_dynamic_rescale = True

    def __init__(self, outputs):
        betas = make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)
        betas = rescale_zero_terminal_snr(betas)
        alphas_cumprod = np.cumprod(1. - betas, axis=0)
        t32 = lambda a: torch.tensor(a, dtype=torch.float32)  # noqa: E731
        self.num_timesteps = 1000
        self.betas = t32(betas)
        self.alphas_cumprod = t32(alphas_cumprod)
        self.alphas_cumprod_prev = t32(np.append(1., alphas_cumprod[:-1]))
        self.sqrt_alphas_cumprod = t32(np.sqrt(alphas_cumprod))
        self.sqrt_one_minus_alphas_cumprod = t32(np.sqrt(1. - alphas_cumprod))
        self.scale_arr = t32(np.concatenate((np.linspace(1.0, 0.3, 400), np.full(1000, 0.3))))
        self.device = torch.device("cpu")
        self.outputs, self.calls = outputs, 0

    def apply_model(self, x, t, c, **kw):
        o = self.outputs[self.calls]
        self.calls += 1
        return o

    def predict_start_from_z_and_v(self, x_t, t, v):  # ddpm3d.py:239-245
        return self.sqrt_alphas_cumprod[t].view(-1, 1, 1, 1, 1) * x_t - self.sqrt_one_minus_alphas_cumprod[t].view(-1, 1, 1, 1, 1) * v

    def predict_eps_from_z_and_v(self, x_t, t, v):    # ddpm3d.py:247-251
        return self.sqrt_alphas_cumprod[t].view(-1, 1, 1, 1, 1) * v + self.sqrt_one_minus_alphas_cumprod[t].view(-1, 1, 1, 1, 1) * x_t


def main():
    out = os.path.join(ROOT, "tests", "golden")
    shape = (1, 4, 5, 8, 12)
    g = torch.Generator().manual_seed(20260003)
    n_steps = 50
    picks = [49, 30, 1, 0]          # DDIM indices exercised (first, middle, last two)
    outs = []
    for _ in picks:
        outs += [torch.randn(shape, generator=g), torch.randn(shape, generator=g)]
    model = StubModel(outs)
    # register_buffer in the reference moves tensors to cuda; keep them on the CPU here
    DDIMSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr)
    s = DDIMSampler(model)
    s.make_schedule(n_steps, ddim_discretize="uniform_trailing", ddim_eta=1.0, verbose=False)
    np.savez_compressed(os.path.join(out, "ddim_schedule.npz"), ddim_timesteps=np.asarray(s.ddim_timesteps),
                        ddim_alphas=np.asarray(s.ddim_alphas), ddim_alphas_prev=np.asarray(s.ddim_alphas_prev),
                        ddim_sigmas=np.asarray(s.ddim_sigmas), ddim_sqrt_one_minus_alphas=np.asarray(s.ddim_sqrt_one_minus_alphas),
                        ddim_scale_arr=s.ddim_scale_arr.numpy(), ddim_scale_arr_prev=s.ddim_scale_arr_prev.numpy(),
                        alphas_cumprod=model.alphas_cumprod.numpy(), sqrt_alphas_cumprod=model.sqrt_alphas_cumprod.numpy(),
                        sqrt_one_minus_alphas_cumprod=model.sqrt_one_minus_alphas_cumprod.numpy(), betas=model.betas.numpy())
    rec = dict(shape=np.asarray(shape), picks=np.asarray(picks), cfg=7.5, guidance_rescale=0.7)
    for k, idx in enumerate(picks):
        x = torch.randn(shape, generator=g)
        noise = torch.randn(shape, generator=g)
        ddim_mod.noise_like = lambda shp, dev, rep=False, _n=noise: _n  # inject the step's noise
        t = torch.full((1,), int(s.ddim_timesteps[idx]), dtype=torch.long)
        x_prev, pred_x0 = s.p_sample_ddim(x, torch.zeros(1), t, index=idx, unconditional_guidance_scale=7.5,
                                          unconditional_conditioning=torch.zeros(1), guidance_rescale=0.7)
        rec.update({f"x_{k}": x.numpy(), f"noise_{k}": noise.numpy(), f"e_c_{k}": outs[2 * k].numpy(),
                    f"e_u_{k}": outs[2 * k + 1].numpy(), f"x_prev_{k}": x_prev.numpy(), f"pred_x0_{k}": pred_x0.numpy()})
    np.savez_compressed(os.path.join(out, "ddim_steps.npz"), **rec)
    print("wrote ddim_schedule.npz, ddim_steps.npz")


if __name__ == "__main__":
    main()
