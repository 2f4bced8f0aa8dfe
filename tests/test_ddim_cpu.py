"""CPU: the DDIM schedule restatement (vc_b200/schedule.py) against golden tables produced by the reference's own
DDIMSampler.make_schedule (tests/make_golden_ddim.py) -- bit-exact -- and the host-side coefficient extraction."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def test_schedule_tables_bit_exact():
    from vc_b200.schedule import DdimSchedule, ModelSchedule

    g = np.load(os.path.join(G, "ddim_schedule.npz"))
    m = ModelSchedule()
    d = DdimSchedule(m, 50, "uniform_trailing", 1.0)
    assert d.timesteps.tolist() == list(range(19, 1000, 20))
    pairs = [(m.betas, "betas"), (m.alphas_cumprod, "alphas_cumprod"), (m.sqrt_alphas_cumprod, "sqrt_alphas_cumprod"),
             (m.sqrt_one_minus_alphas_cumprod, "sqrt_one_minus_alphas_cumprod"), (d.timesteps, "ddim_timesteps"),
             (d.alphas, "ddim_alphas"), (d.alphas_prev, "ddim_alphas_prev"), (d.sigmas, "ddim_sigmas"),
             (d.sqrt_one_minus_alphas, "ddim_sqrt_one_minus_alphas"), (d.scale_arr, "ddim_scale_arr"),
             (d.scale_arr_prev, "ddim_scale_arr_prev")]
    for a, name in pairs:
        assert np.array_equal(np.asarray(a), g[name]), name
    assert m.alphas_cumprod[-1] == 0.0  # zero terminal SNR


def test_other_discretisations_and_coefficients():
    from vc_b200.schedule import DdimSchedule, ModelSchedule, ddim_timesteps

    assert ddim_timesteps("uniform", 50).tolist() == [i + 1 for i in range(0, 1000, 20)]
    assert len(ddim_timesteps("quad", 50)) == 50
    d = DdimSchedule(ModelSchedule(), 50, "uniform_trailing", 0.0)
    assert float(np.abs(d.sigmas).max()) == 0.0
    c = DdimSchedule(ModelSchedule(), 50, "uniform_trailing", 1.0).coefficients(49, 7.5, 0.7)
    assert c["timestep"] == 999 and c["sqrt_alphas_cumprod_t"] == 0.0 and c["use_dynamic_rescale"] == 1
    assert abs(c["scale_t"] - 0.3) < 1e-7 and abs(c["scale_prev"] - 0.3) < 1e-7
