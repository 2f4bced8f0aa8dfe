"""Full-size ViewCrafter U-Net (1.44 B params, seeded random weights): parity vs the reference under bf16 autocast and
forward time, ours vs reference, at a given latent size.  usage: python tools/bench_unet.py T H W [check]"""
import sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, "guidedvd-3dgs_b200")
import torch, unet_ref
from vc_b200.unet import UNetB200
t, h, w = (int(a) for a in sys.argv[1:4])
check = len(sys.argv) > 4
ref, cfg = unet_ref.build_reference_unet(model_channels=320)
ours = UNetB200(ref.state_dict(), device="cuda", **cfg)
x, cc, ctx, _ = unet_ref.synth_inputs(t, h, w)
xin = torch.cat([x, cc], 1); ts = torch.tensor([481], device="cuda"); fs = torch.tensor([10], device="cuda")
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
def run_ref():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return ref(xin, ts, context=ctx, fs=fs)
y = ours(xin, ts, ctx, fs=fs); torch.cuda.synchronize()
print("ours ok", tuple(y.shape), "peak mem GB", torch.cuda.max_memory_allocated() / 1e9)
t_ours = timeit(lambda: ours(xin, ts, ctx, fs=fs))
print(f"ours forward {t_ours*1e3:.1f} ms")
try:
    y_ref = run_ref(); torch.cuda.synchronize()
    t_ref = timeit(run_ref)
    print(f"ref(bf16 autocast) forward {t_ref*1e3:.1f} ms; rel L2 ours vs ref {((y.float()-y_ref.float()).norm()/y_ref.float().norm()).item():.3e}")
    if check:
        with torch.no_grad():
            y32 = ref(xin, ts, context=ctx, fs=fs)
        print(f"rel L2 vs fp32: ours {((y.float()-y32).norm()/y32.norm()).item():.3e} ref-bf16 {((y_ref.float()-y32).norm()/y32.norm()).item():.3e}")
except torch.cuda.OutOfMemoryError as e:
    print("reference OOM:", str(e)[:100])
