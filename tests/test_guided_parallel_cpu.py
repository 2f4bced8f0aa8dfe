"""CPU (gloo): the multi-GPU plan of the guided DDIM step (vc_b200.guided.GuidedPlan) -- cond / uncond U-Net forward +
backward on alternating ranks, decoder passes dealt out by frame, one exchange of the two outputs, one gather of
dL/dpred_x0 and of the decoded frames, one sum of the two dL/dx.  Every rank must end with the x_prev / pred_x0 of the
single-process step.  The library underneath is the pointer-level stand-in of tests/fake_nn_lib.py (host memory), so
everything that differs between 1 and N ranks is exercised here; world sizes 2 (pure CFG split), 3 (odd: no CFG split, frames
three ways) and 8 (cfg 2 x frames 4: the U-Net itself is frame-sharded, with the adjoint all-to-alls and the two-stage
sharded GroupNorm backward; four ranks own no decoder frame)."""
import contextlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _install_fake():
    for p in (os.path.join(ROOT, "guidedvd-3dgs_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gvd_native
    from fake_nn_lib import FakeNN
    from vc_b200 import ops, unet

    fake = FakeNN(torch.float32)
    gvd_native.nn = lambda: fake
    ops._stream = lambda: None
    ops.BF16 = unet.BF16 = torch.float32
    torch.cuda.device = lambda d: contextlib.nullcontext()
    return fake


def _worker(rank, world, port, q):
    try:
        _worker_body(rank, world, port, q)
    except Exception:  # report instead of leaving the peers (and the test) waiting for a collective
        import traceback
        q.put((rank, "error", traceback.format_exc()[-1500:]))


def _worker_body(rank, world, port, q):
    torch.set_num_threads(1 if world == 8 else 2)
    fake = _install_fake()
    import unet_ref
    from test_guided_cpu import StubDecoder, StubGuidance
    from vc_b200.guided import DDIMSamplerGuidance, GuidedPlan
    from vc_b200.schedule import ModelSchedule
    from vc_b200.unet import DiffusionModelB200, UNetB200

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref, cfg = unet_ref.build_reference_unet(model_channels=64, device="cpu")
    # frame-sharded worlds need >= frame_ways pixels at the coarsest level (16x16 -> 2x2); the pure CFG split runs the small
    # latent with two recurrences instead
    (h, w, recur) = (8, 8, 2) if world == 2 else (16, 16, 1)
    T, index = (4 if world == 8 else 3), 22   # world 8 = cfg 2 x frames 4: one frame per rank of a branch, four ranks without a decoder frame
    x, cc, ctx, ctx_uc = unet_ref.synth_inputs(T, h, w, device="cpu")
    cond, uc = {"c_concat": [cc], "c_crossattn": [ctx]}, {"c_concat": [cc], "c_crossattn": [ctx_uc]}
    fs = torch.tensor([10])
    g = torch.Generator().manual_seed(99)
    targets = [torch.rand(3, 2 * h, 2 * w, generator=g) * 2 - 1 for _ in range(T)]
    masks = [(torch.rand(1, 2 * h, 2 * w, generator=g) > 0.3).float() for _ in range(T)]
    noises = [torch.randn(x.shape, generator=g) for _ in range(2 * recur)]
    model = DiffusionModelB200(UNetB200(ref.state_dict(), device="cpu", **cfg), ModelSchedule())
    model.differentiable_decode_first_stage = StubDecoder()
    sampler = DDIMSamplerGuidance(model)
    sampler.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0)
    ts = torch.full((1,), int(sampler.ddim_timesteps[index]), dtype=torch.long)

    def step(lg):
        return sampler.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5, unconditional_conditioning=uc,
                                     guidance_rescale=0.7, fs=fs, loss_guidance_fn=lg, noise=noises[0::2], recur_noise=noises[1::2])

    lg1 = StubGuidance(targets, masks, recur)
    xp1, p01 = step(lg1)                                 # single-process answer (no plan)
    GuidedPlan(T, model)
    lgn = StubGuidance(targets, masks, recur)
    xpn, p0n = step(lgn)
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()  # noqa: E731
    # an un-guided step on the same (now sharded) model still returns the full clip, equal to the single-process one
    plain_n = sampler.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5, unconditional_conditioning=uc,
                                    guidance_rescale=0.7, fs=fs, noise=noises[0])[0]
    model.plan, part, model.unet.part = None, model.unet.part, None
    plain_1 = sampler.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5, unconditional_conditioning=uc,
                                    guidance_rescale=0.7, fs=fs, noise=noises[0])[0]
    model.unet.part = part
    assert plain_n.shape == plain_1.shape and rel(plain_n, plain_1) < 2e-5
    gp = model.guided_plan
    sharded_ok = (fake.calls.get("groupnorm_bwd_sums", 0) > 0) == gp.part.active   # the sharded GroupNorm backward ran iff frames are sharded
    q.put((rank, rel(xpn, xp1), rel(p0n, p01), rel(lgn.saved[-1][1], lg1.saved[-1][1]), gp.branch, (gp.f0, gp.f1), sharded_ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_guided_plan_matches_single_process(world):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ViewCrafter", "lvdm", "modules", "networks", "openaimodel3d.py")):
        pytest.skip("oracle/_ref/ViewCrafter not installed")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    for _ in range(world):
        item = q.get(timeout=900)
        if item[1] == "error":
            for p in procs:
                p.terminate()
            pytest.fail(f"rank {item[0]}: {item[2]}")
        res.append(item)
    res.sort()
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    frames = [r[5] for r in res]
    assert frames[0][0] == 0 and frames[-1][1] == (4 if world == 8 else 3) and all(a[1] == b[0] for a, b in zip(frames, frames[1:]))
    for rank, e_xp, e_p0, e_img, branch, _, sharded_ok in res:
        assert sharded_ok
        assert branch == ((rank // (world // 2)) if world % 2 == 0 else None)
        assert e_xp < 2e-5 and e_p0 < 2e-5 and e_img < 2e-5, (rank, e_xp, e_p0, e_img)  # fp32 sums in a different order
