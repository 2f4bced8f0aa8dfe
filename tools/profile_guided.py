"""Kernel-time breakdown of one native guided DDIM step (C4 guided shape) and of one plain U-Net forward (C3), from
torch.profiler (CUPTI): which kernels the time goes to.  usage: python tools/profile_guided.py [guided|unet]"""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "guidedvd-3dgs_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import unet_ref  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "guided"
dev = torch.device("cuda", 0)
ref, cfg = unet_ref.build_reference_unet(model_channels=320, device=dev)
from vc_b200.schedule import ModelSchedule  # noqa: E402
from vc_b200.unet import DiffusionModelB200, UNetB200  # noqa: E402

unet = UNetB200(ref.state_dict(), device=dev, **cfg)
del ref
torch.cuda.empty_cache()
fs = torch.tensor([10], device=dev)
if what == "guided":
    import test_guided_cpu as tg
    import test_vae_cpu as tv
    from vc_b200.guided import DDIMSamplerGuidance
    from vc_b200.vae import DecoderB200
    T, h, w = 25, 40, 64
    x, cc, ctx, ctx_uc = unet_ref.synth_inputs(T, h, w, device=dev)
    cond, uc = {"c_concat": [cc], "c_crossattn": [ctx]}, {"c_concat": [cc], "c_crossattn": [ctx_uc]}
    g = torch.Generator().manual_seed(123)
    targets = [(torch.rand(3, 8 * h, 8 * w, generator=g) * 2 - 1).to(dev) for _ in range(T)]
    masks = [(torch.rand(1, 8 * h, 8 * w, generator=g) > 0.3).float().to(dev) for _ in range(T)]
    model = DiffusionModelB200(unet, ModelSchedule())
    dec = DecoderB200(tv.RefFirstStage(ch=128).state_dict(), device=dev, scale_factor=tv.SCALE)
    model.differentiable_decode_first_stage = dec.differentiable_decode
    model.guided_decode_frames = 5
    s = DDIMSamplerGuidance(model)
    s.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0)
    ts = torch.full((1,), int(s.ddim_timesteps[30]), dtype=torch.long, device=dev)
    lg = tg.StubGuidance(targets, masks, 1)

    def step():
        s.p_sample_ddim(x, cond, ts, index=30, unconditional_guidance_scale=7.5, unconditional_conditioning=uc,
                        guidance_rescale=0.7, fs=fs, loss_guidance_fn=lg)
else:
    x, cc, ctx, _ = unet_ref.synth_inputs(25, 72, 128, device=dev)
    xin = torch.cat([x, cc], 1)
    ts = torch.tensor([481], device=dev)

    def step():
        unet(xin, ts, ctx, fs=fs)
step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
tot = 0.0
for e in prof.events():
    if e.device_type is not None and str(e.device_type).endswith("CUDA") and e.device_time_total > 0 and "Memcpy" not in e.name and "Memset" not in e.name:
        a = agg.setdefault(e.name[:100], [0, 0.0])
        a[0] += 1
        a[1] += e.device_time_total
        tot += e.device_time_total
print(f"# {what}: {tot / 1e3:.1f} ms of kernels, {sum(v[0] for v in agg.values())} launches")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t / 1e3:9.2f} ms x{c:5d} {100 * t / tot:5.1f}%  {n}")

if len(sys.argv) > 2 and sys.argv[2] == "host":
    # host side of the same step: wall time of issuing it (device left asynchronous) and a cProfile of the Python path
    import cProfile, pstats, time
    torch.cuda.synchronize()
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"# host: issue {1e3 * (t1 - t0):.1f} ms, drain {1e3 * (t2 - t1):.1f} ms")
    pr = cProfile.Profile(); pr.enable(); step(); pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(30)
