"""Diagnosis of the fat-Gaussian gradient discrepancy (tests/test_render_dropin_gpu.py::test_frame_that_outgrows...)."""
import os, sys, types
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gs_refload, synth
import test_render_dropin_gpu as t

ours, ref = gs_refload.load("ours"), gs_refload.load("reference")
P, seed = 60_000, 31
m_ref, _ = t._model(ref, P, seed)
m_ours, _ = t._model(ours, P, seed, like=m_ref)
with torch.no_grad():
    for m in (m_ref, m_ours):
        m._scaling.add_(1.2)
bg = torch.zeros(3, device="cuda")
W, H, fov = 1280, 960, 100.0
w2c, K = t._pose(seed + 2, W, H, fovx_deg=fov)
gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(2)).cuda()
res = {}
for tag, gs, m in (("ref", ref, m_ref), ("ref2", ref, m_ref), ("ours", ours, m_ours), ("ours2", ours, m_ours)):
    er = gs.EasyRenderer.__new__(gs.EasyRenderer)
    view = er.make_gs_view_format(w2c, K, H, W)
    for n in t.PARAMS:
        getattr(m, n).grad = None
    pkg = gs.render(view, m, t._pipe(), bg)
    (((pkg["render"] - gt) ** 2).mean() + 0.1 * pkg["depth"].mean()).backward()
    torch.cuda.synchronize()
    g = {n: getattr(m, n).grad.detach().clone() for n in t.PARAMS}
    g["viewspace"] = pkg["viewspace_points"].grad.detach().clone()
    res[tag] = (pkg, g)
for k in ("render", "depth", "alpha"):
    a, b = res["ours"][0][k], res["ref"][0][k]
    print(k, "max abs diff", float((a - b).abs().max()), "bits differ", int((a.view(torch.int32) != b.view(torch.int32)).sum()), "of", a.numel())
print("radii equal", bool(torch.equal(res["ours"][0]["radii"], res["ref"][0]["radii"])), "visible", int((res["ref"][0]["radii"] > 0).sum()))
for n in res["ref"][1]:
    go, gr, gr2, go2 = res["ours"][1][n], res["ref"][1][n], res["ref2"][1][n], res["ours2"][1][n]
    print(f"{n:16s} ours-vs-ref {t._rel(go, gr):.3e}  ref-jitter {t._rel(gr2, gr):.3e}  ours-jitter {t._rel(go2, go):.3e}")
# localise: which Gaussians carry the xyz error
d = (res["ours"][1]["_xyz"] - res["ref"][1]["_xyz"]).norm(dim=1)
top = torch.topk(d, 10).indices
rad = res["ref"][0]["radii"]
print("top-10 share of squared error", float((d[top] ** 2).sum() / (d ** 2).sum()))
for i in top.tolist():
    print(i, "err", float(d[i]), "|g|", float(res["ref"][1]["_xyz"][i].norm()), "radius", int(rad[i]),
          "scale", m_ref.get_scaling[i].tolist(), "opac", float(m_ref.get_opacity[i]))
