"""Noise schedule of the ViewCrafter latent-diffusion model and the DDIM sub-schedule (host side, float64 numpy).

Restates, for the configuration the reference runs (configs/inference_pvd_1024.yaml: linear betas 0.00085..0.012,
1000 steps, zero-terminal-SNR rescale, v-parameterisation, dynamic rescale with base_scale 0.3):
  lvdm/models/utils_diffusion.py:31-35   make_beta_schedule('linear')
  lvdm/models/utils_diffusion.py:112-144 rescale_zero_terminal_snr
  lvdm/models/ddpm3d.py:123-151          register_schedule (alphas_cumprod, sqrt tables, float32 buffers)
  lvdm/models/ddpm3d.py:519-527          scale_arr (dynamic rescale)
  lvdm/models/utils_diffusion.py:56-91   make_ddim_timesteps / make_ddim_sampling_parameters
  lvdm/models/samplers/ddim.py:24-59     DDIMSampler.make_schedule
"""
import numpy as np


def make_betas(n_timestep=1000, linear_start=0.00085, linear_end=0.012, zero_terminal_snr=True):
    betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2
    if zero_terminal_snr:
        abar_sqrt = np.sqrt(np.cumprod(1.0 - betas, axis=0))
        first, last = abar_sqrt[0].copy(), abar_sqrt[-1].copy()
        abar_sqrt -= last
        abar_sqrt *= first / (first - last)
        abar = abar_sqrt ** 2
        alphas = np.concatenate([abar[0:1], abar[1:] / abar[:-1]])
        betas = 1 - alphas
    return betas


class ModelSchedule:
    """The buffers DDPM.register_schedule keeps (float32, like the reference's to_torch)."""

    def __init__(self, n_timestep=1000, linear_start=0.00085, linear_end=0.012, zero_terminal_snr=True,
                 use_dynamic_rescale=True, base_scale=0.3, turning_step=400):
        betas = make_betas(n_timestep, linear_start, linear_end, zero_terminal_snr)
        alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
        self.num_timesteps = int(n_timestep)
        self.betas = betas.astype(np.float32)
        self.alphas_cumprod = alphas_cumprod.astype(np.float32)
        self.alphas_cumprod_prev = np.append(1.0, alphas_cumprod[:-1]).astype(np.float32)
        self.sqrt_alphas_cumprod = np.sqrt(alphas_cumprod).astype(np.float32)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - alphas_cumprod).astype(np.float32)
        self.use_dynamic_rescale = bool(use_dynamic_rescale)
        if use_dynamic_rescale:
            arr1 = np.linspace(1.0, base_scale, turning_step)
            arr2 = np.full(self.num_timesteps, base_scale)
            self.scale_arr = np.concatenate((arr1, arr2)).astype(np.float32)


def ddim_timesteps(method, num_ddim, num_ddpm=1000):
    if method == "uniform":
        c = num_ddpm // num_ddim
        return np.asarray(list(range(0, num_ddpm, c))) + 1
    if method == "uniform_trailing":
        c = num_ddpm / num_ddim
        return np.flip(np.round(np.arange(num_ddpm, 0, -c))).astype(np.int64) - 1
    if method == "quad":
        return ((np.linspace(0, np.sqrt(num_ddpm * .8), num_ddim)) ** 2).astype(int) + 1
    raise NotImplementedError(f'There is no ddim discretization method called "{method}"')


class DdimSchedule:
    """Per-index tables of DDIMSampler.make_schedule (index i <-> timestep ddim_timesteps[i])."""

    def __init__(self, model: ModelSchedule, num_steps=50, method="uniform_trailing", eta=1.0):
        self.timesteps = ddim_timesteps(method, num_steps, model.num_timesteps)
        ac = model.alphas_cumprod  # float32, as in the reference (alphas_cumprod.cpu())
        self.alphas = ac[self.timesteps]
        self.alphas_prev = np.asarray([ac[0]] + ac[self.timesteps[:-1]].tolist())
        # dtype flow of make_ddim_sampling_parameters when it is fed the float32 alphas_cumprod TENSOR (ddim.py:49-51):
        # `ndarray / tensor` dispatches to Tensor.__rtruediv__ = tensor.reciprocal() * ndarray, so 1/(1 - alphas) is
        # evaluated in float32; everything else is float64.  Reproduced so the sigma table is bit-identical.
        rcp = (np.float32(1) / (np.float32(1) - self.alphas.astype(np.float32))).astype(np.float64)
        self.sigmas = eta * np.sqrt((1 - self.alphas_prev) * rcp * (1 - self.alphas.astype(np.float64) / self.alphas_prev))
        self.sqrt_one_minus_alphas = np.sqrt(1.0 - self.alphas)
        if model.use_dynamic_rescale:
            self.scale_arr = model.scale_arr[self.timesteps]
            self.scale_arr_prev = np.concatenate([model.scale_arr[0:1], self.scale_arr[:-1]])
        self.model = model

    def coefficients(self, index, cfg_scale=1.0, guidance_rescale=0.0, temperature=1.0):
        """Scalars of one p_sample_ddim call (ddim.py:241-279) for gvd_ddim_step."""
        t = int(self.timesteps[index])
        m = self.model
        c = dict(cfg_scale=cfg_scale, guidance_rescale=guidance_rescale, temperature=temperature,
                 sqrt_alphas_cumprod_t=float(m.sqrt_alphas_cumprod[t]),
                 sqrt_one_minus_alphas_cumprod_t=float(m.sqrt_one_minus_alphas_cumprod[t]),
                 ddim_alpha_prev=float(np.float32(self.alphas_prev[index])), ddim_sigma=float(np.float32(self.sigmas[index])),
                 use_dynamic_rescale=int(m.use_dynamic_rescale), scale_t=1.0, scale_prev=1.0, timestep=t)
        if m.use_dynamic_rescale:
            c["scale_t"] = float(self.scale_arr[index])
            c["scale_prev"] = float(self.scale_arr_prev[index])
        return c
