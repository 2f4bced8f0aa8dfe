"""N-rank check + timing of the guided DDIM step on real GPUs (vc_b200.guided.GuidedPlan: cfg x frame-sharded U-Net
forward + backward, decoder passes dealt out by frame):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29534 \
        tools/run_guided_parallel.py [--t 25 --h 40 --w 64] [--mc 320] [--vae-ch 128] [--steps 2]
Every rank builds the same seeded U-Net and VAE decoder, runs (a) the single-GPU guided step and (b) the planned one with
the same noise, compares x_prev, then times (b) (CUDA events, max over ranks).  Prints one JSON line on rank 0.
First hardware numbers: profiles/r02_first_hw_run.txt (bench.py carries the measured blocks now)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "guidedvd-3dgs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--t", type=int, default=25)
    ap.add_argument("--h", type=int, default=40)
    ap.add_argument("--w", type=int, default=64)
    ap.add_argument("--mc", type=int, default=320)
    ap.add_argument("--vae-ch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--decode-frames", type=int, default=5)
    ap.add_argument("--skip-single", action="store_true", help="do not run the single-GPU step first (memory / time)")
    ap.add_argument("--dry-cpu", action="store_true", help="walk the script on the CPU (gloo, the tests' library stand-in): host-side check only")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if a.dry_cpu:
        import pytest
        import test_unet_grad_cpu
        test_unet_grad_cpu.install_fake(pytest.MonkeyPatch())
        dev = torch.device("cpu")
        if world > 1:
            dist.init_process_group("gloo")
    else:
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
        torch.cuda.set_device(dev)
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)
    import test_guided_cpu as tg
    import test_vae_cpu as tv
    import unet_ref
    from vc_b200.guided import DDIMSamplerGuidance, GuidedPlan
    from vc_b200.schedule import ModelSchedule
    from vc_b200.unet import DiffusionModelB200, UNetB200
    from vc_b200.vae import DecoderB200

    ref, cfg = unet_ref.build_reference_unet(model_channels=a.mc, device="cpu")
    unet = UNetB200(ref.state_dict(), device=dev, **cfg)
    del ref
    dec = DecoderB200(tv.RefFirstStage(ch=a.vae_ch).state_dict(), device=dev, scale_factor=tv.SCALE)
    T, h, w, index = a.t, a.h, a.w, 30
    x, cc, ctx, ctx_uc = unet_ref.synth_inputs(T, h, w, device=dev)
    cond, uc = {"c_concat": [cc], "c_crossattn": [ctx]}, {"c_concat": [cc], "c_crossattn": [ctx_uc]}
    fs = torch.tensor([10], device=dev)
    g = torch.Generator().manual_seed(123)
    targets = [(torch.rand(3, 8 * h, 8 * w, generator=g) * 2 - 1).to(dev) for _ in range(T)]
    masks = [(torch.rand(1, 8 * h, 8 * w, generator=g) > 0.3).float().to(dev) for _ in range(T)]
    noises = [torch.randn(x.shape, generator=g).to(dev) for _ in range(2)]
    model = DiffusionModelB200(unet, ModelSchedule())
    model.differentiable_decode_first_stage = dec.differentiable_decode
    model.guided_decode_frames = a.decode_frames
    sampler = DDIMSamplerGuidance(model)
    sampler.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0)
    ts = torch.full((1,), int(sampler.ddim_timesteps[index]), dtype=torch.long, device=dev)

    def step():
        return sampler.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5, unconditional_conditioning=uc,
                                     guidance_rescale=0.7, fs=fs, loss_guidance_fn=tg.StubGuidance(targets, masks, 1),
                                     noise=noises[0:1], recur_noise=noises[1:2])[0]

    def timed(n):
        if world > 1:
            dist.barrier()
        if a.dry_cpu:
            import time
            t0 = time.perf_counter()
            for _ in range(n):
                out = step()
            ms = torch.tensor([(time.perf_counter() - t0) / n * 1e3])
        else:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                out = step()
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), out

    res = {"metric": "guided DDIM steps/sec", "n_gpus": world, "config": {"frames": T, "latent": [h, w], "unet_model_channels": a.mc,
                                                                            "vae_ch": a.vae_ch}}
    xp_single = None
    if not a.skip_single:
        step()
        ms1, xp_single = timed(a.steps)
        res["single_gpu_ms_per_step"] = round(ms1, 1)
    if world > 1:
        gp = GuidedPlan(T, model)
        step()
        msn, xp_plan = timed(a.steps)
        res.update(value=round(1e3 / msn, 4), unit="steps/s", ms_per_step=round(msn, 1),
                   plan=f"cfg{gp.denoise.cfg_ways} x frames{gp.denoise.frame_ways}, decoder frames {gp.frames}")
        if xp_single is not None:
            res["x_prev_rel_l2_vs_single_gpu"] = float((xp_plan - xp_single).norm() / xp_single.norm())
            res["speedup"] = round(res["single_gpu_ms_per_step"] / msn, 3)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
