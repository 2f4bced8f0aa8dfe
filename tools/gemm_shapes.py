"""Per-shape GEMM / attention time inside one U-Net forward (events around every call; serialised)."""
import sys, collections
sys.path.insert(0, "tests"); sys.path.insert(0, "guidedvd-3dgs_b200")
import torch, unet_ref
from vc_b200 import ops
from vc_b200.unet import UNetB200
t, h, w = (int(a) for a in sys.argv[1:4])
ref, cfg = unet_ref.build_reference_unet(model_channels=320)
ours = UNetB200(ref.state_dict(), device="cuda", **cfg); del ref
x, cc, ctx, _ = unet_ref.synth_inputs(t, h, w)
xin = torch.cat([x, cc], 1); ts = torch.tensor([481], device="cuda"); fs = torch.tensor([10], device="cuda")
ours(xin, ts, ctx, fs=fs); torch.cuda.synchronize()
rec = collections.defaultdict(lambda: [0, 0.0])
def wrap(name, fn, keyf):
    def g(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record(); torch.cuda.synchronize()
        key = (name,) + keyf(*a, **k); rec[key][0] += 1; rec[key][1] += e0.elapsed_time(e1); return r
    return g
ops.gemm_raw = wrap("gemm", ops.gemm_raw, lambda A, B, C, M, N, K, *a, **k: (M, N, K, k.get("batch_h", 1) * k.get("batch_b", 1), bool(k.get("residual") is not None)))
ops.flash_attention = wrap("flash", ops.flash_attention, lambda q, k, v, Bq, Nq, Nk, H, s, shared_kv=False: (Bq, Nq, Nk, H, shared_kv))
for n in ("groupnorm", "layernorm", "geglu", "temporal_attention"):
    setattr(ops, n, wrap(n, getattr(ops, n), lambda *a, **k: (tuple(a[0].shape),)))
ours(xin, ts, ctx, fs=fs)
tot = sum(v[1] for v in rec.values())
print(f"total {tot:.1f} ms")
for key, (c, ms) in sorted(rec.items(), key=lambda kv: -kv[1][1])[:40]:
    extra = ""
    if key[0] == "gemm":
        M, N, K, b = key[1:5]; extra = f"{2*M*N*K*b*c/ms/1e9:8.0f} TF/s"
    if key[0] == "flash":
        Bq, Nq, Nk, H = key[1:5]; extra = f"{4*Bq*Nq*Nk*H*64*c/ms/1e9:8.0f} TF/s"
    print(f"{ms:8.2f} ms x{c:3d} {extra} {key}")
