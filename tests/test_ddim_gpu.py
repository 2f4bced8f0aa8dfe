"""GPU: the fused DDIM update (gvd_ddim_step through vc_b200.sampler.DDIMSampler.p_sample_ddim) against golden steps
recorded from the reference's DDIMSampler.p_sample_ddim (tests/make_golden_ddim.py).  fp32 throughout; tolerance 2e-6
relative to the tensor scale (fused vs. op-by-op rounding)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Replay:
    def __init__(self, outs):
        from vc_b200.schedule import ModelSchedule

        self.schedule = ModelSchedule()
        self.outs, self.calls = outs, 0

    def apply_model(self, x, t, c, **kw):
        o = self.outs[self.calls]
        self.calls += 1
        return o


def test_p_sample_ddim_matches_reference():
    from vc_b200.sampler import DDIMSampler

    g = np.load(os.path.join(ROOT, "tests", "golden", "ddim_steps.npz"))
    for k, idx in enumerate(g["picks"].tolist()):
        dev = "cuda"
        e_c, e_u = torch.from_numpy(g[f"e_c_{k}"]).to(dev), torch.from_numpy(g[f"e_u_{k}"]).to(dev)
        s = DDIMSampler(_Replay([e_c, e_u]))
        s.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0)
        x = torch.from_numpy(g[f"x_{k}"]).to(dev)
        t = torch.full((1,), int(s.ddim_timesteps[idx]), device=dev, dtype=torch.long)
        x_prev, pred_x0 = s.p_sample_ddim(x, "c", t, index=idx, unconditional_guidance_scale=float(g["cfg"]),
                                          unconditional_conditioning="uc", guidance_rescale=float(g["guidance_rescale"]),
                                          noise=torch.from_numpy(g[f"noise_{k}"]).to(dev))
        for name, got in (("x_prev", x_prev), ("pred_x0", pred_x0)):
            ref = torch.from_numpy(g[f"{name}_{k}"]).to(dev)
            err = (got - ref).abs().max().item() / ref.abs().max().item()
            assert err < 2e-6, (idx, name, err)


def test_no_guidance_and_full_loop_shapes():
    from vc_b200.sampler import DDIMSampler

    dev = "cuda"
    shape = (4, 3, 8, 8)
    gen = torch.Generator(device=dev).manual_seed(0)
    outs = [torch.randn(1, *shape, device=dev, generator=gen) for _ in range(10)]
    s = DDIMSampler(_Replay(outs))
    img, inter = s.sample(S=10, batch_size=1, shape=shape, conditioning="c", eta=0.0, timestep_spacing="uniform_trailing")
    assert img.shape == (1, *shape) and torch.isfinite(img).all() and s.model.calls == 10
