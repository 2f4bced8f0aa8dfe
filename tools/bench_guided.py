"""Guided DDIM step at the shape train_guidedvd.py runs (C4, 25 frames, latent 40x64 -> 320x512 images):
one `p_sample_ddim` with loss guidance = cond + uncond U-Net forward WITH the tape, 25 VAE decodes with the tape, the
guidance loss, both backward passes, the update.  Full-size U-Net (1.44 B parameters) and VAE decoder (ch 128), seeded
random weights; ours (vc_b200.guided, everything native) against the reference DDIMSamplerGuidance over the reference
modules under torch.autocast(bfloat16) on the same GPU.  Prints one JSON line per arm.

usage: python tools/bench_guided.py [--arm ours|reference|both] [--frames 25] [--latent 40 64] [--steps 2] [--decode-frames 5]
First hardware numbers: profiles/r02_first_hw_run.txt (bench.py carries the measured blocks now)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "guidedvd-3dgs_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch  # noqa: E402

import test_guided_cpu as tg  # noqa: E402  (reference-sampler harness, LossGuidance stand-in)
import test_vae_cpu as tv  # noqa: E402
import unet_ref  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", default="both", choices=["ours", "reference", "both"])
    ap.add_argument("--frames", type=int, default=25)
    ap.add_argument("--latent", type=int, nargs=2, default=[40, 64])
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--decode-frames", type=int, default=5)
    ap.add_argument("--mc", type=int, default=320, help="U-Net model_channels (320 = full size)")
    ap.add_argument("--vae-ch", type=int, default=128)
    ap.add_argument("--dry-cpu", action="store_true",
                    help="walk the script on the CPU over the tests' stand-in of the library (host-side check of this tool; no timing value)")
    a = ap.parse_args()
    dev = torch.device("cpu") if a.dry_cpu else torch.device("cuda", 0)
    if a.dry_cpu:
        import pytest
        import test_unet_grad_cpu
        test_unet_grad_cpu.install_fake(pytest.MonkeyPatch())
    T, (h, w) = a.frames, a.latent
    ref, cfg = unet_ref.build_reference_unet(model_channels=a.mc, device=dev)
    vae = tv.RefFirstStage(ch=a.vae_ch).to(dev).eval()
    x, cc, ctx, ctx_uc = unet_ref.synth_inputs(T, h, w, device=dev)
    cond, uc = {"c_concat": [cc], "c_crossattn": [ctx]}, {"c_concat": [cc], "c_crossattn": [ctx_uc]}
    fs = torch.tensor([10], device=dev)
    g = torch.Generator().manual_seed(123)
    targets = [(torch.rand(3, 8 * h, 8 * w, generator=g) * 2 - 1).to(dev) for _ in range(T)]
    masks = [(torch.rand(1, 8 * h, 8 * w, generator=g) > 0.3).float().to(dev) for _ in range(T)]
    index = 30
    # 2 U-Net forwards + their input-gradient (2x a forward) + 25 decoder forward + latent-gradient
    flops = 2 * 20.19e12 * 3 * (T * h * w) / (25 * 40 * 64) + T * 1.56e12 * 3 * (h * w) / (40 * 64)

    def timed(step):
        step()
        if a.dry_cpu:
            t0 = time.perf_counter()
            for _ in range(a.steps):
                step()
            return (time.perf_counter() - t0) / a.steps * 1e3
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.reset_peak_memory_stats()
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps

    def line(arm, ms):
        print(json.dumps({"metric": "guided DDIM steps/sec", "impl": arm, "value": round(1e3 / ms, 4), "unit": "steps/s",
                          "ms_per_step": round(ms, 1), "tflops_per_s": round(flops / ms / 1e9, 1), "dtype": "bf16",
                          "peak_mem_gb": None if a.dry_cpu else round(torch.cuda.max_memory_allocated() / 1e9, 1),
                          "config": {"workload": "C4 guided", "frames": T, "latent": [h, w], "cfg": 7.5, "recur_steps": 1,
                                     "unet_model_channels": a.mc, "vae_ch": a.vae_ch,
                                     "decode_frames_per_call": a.decode_frames if arm == "ours" else 1}}), flush=True)

    if a.arm in ("ours", "both"):
        from vc_b200.guided import DDIMSamplerGuidance
        from vc_b200.schedule import ModelSchedule
        from vc_b200.unet import DiffusionModelB200, UNetB200
        from vc_b200.vae import DecoderB200

        model = DiffusionModelB200(UNetB200(ref.state_dict(), device=dev, **cfg), ModelSchedule())
        dec = DecoderB200(vae.state_dict(), device=dev, scale_factor=tv.SCALE)
        model.differentiable_decode_first_stage = dec.differentiable_decode
        model.guided_decode_frames = a.decode_frames
        s = DDIMSamplerGuidance(model)
        s.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0)
        ts = torch.full((1,), int(s.ddim_timesteps[index]), dtype=torch.long, device=dev)
        lg = tg.StubGuidance(targets, masks, 1)
        t0 = time.perf_counter()
        ms = timed(lambda: s.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5,
                                           unconditional_conditioning=uc, guidance_rescale=0.7, fs=fs, loss_guidance_fn=lg))
        line("ours", ms)
        print(f"# ours: wall {time.perf_counter() - t0:.1f} s", file=sys.stderr)
        del model, dec, s
        if not a.dry_cpu:
            torch.cuda.empty_cache()
    if a.arm in ("reference", "both"):
        class PerFrame(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.vae = vae

            def forward(self, z):
                return torch.stack([self.vae(z[:, :, f])[0] for f in range(z.shape[2])], dim=1).unsqueeze(0)

        s, _ = tg._reference_sampler(ref, PerFrame())
        for name, val in list(vars(s.model).items()):
            if isinstance(val, torch.Tensor):
                setattr(s.model, name, val.to(dev))
        s.model.device = dev
        s.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0, verbose=False)
        ts = torch.full((1,), int(s.ddim_timesteps[index]), dtype=torch.long, device=dev)
        lg = tg.StubGuidance(targets, masks, 1)

        def step():
            with torch.autocast("cpu" if a.dry_cpu else "cuda", dtype=torch.bfloat16):
                s.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5, unconditional_conditioning=uc,
                                guidance_rescale=0.7, fs=fs, loss_guidance_fn=lg)
        try:
            line("reference", timed(step))
        except torch.OutOfMemoryError as ex:
            print(json.dumps({"impl": "reference", "unavailable": "out of memory: " + str(ex)[:120]}))


if __name__ == "__main__":
    main()
