"""One WHOLE 3DGS training iteration at BASELINE configs[1] (C2: 500 k Gaussians, 640x480, SH degree 3), both arms on the
same GPU: activations -> rasterizer forward -> L1 + SSIM loss -> backward -> densification statistics -> Adam.
  ours:       diff_gaussian_rasterization (this repo) + train_ops.photometric_loss / add_densification_stats / FusedAdam
  folded:     the same with the activations folded into the rasterizer kernels (rasterize_gaussians_raw: no exp / sigmoid /
              normalize launches, no 96 MB SH concatenation, one backward node -- SURVEY 8 row f3)
  reference:  oracle/_ref rasterizer (the reference's CUDA code) + utils/loss_utils.py's l1_loss / ssim restated in torch
              (same conv2d calls) + the masked densification statements + torch.optim.Adam(eps=1e-15)
(train_baseline.py:73-120 without logging / densify-and-prune, which run every 100 iterations.)
Prints one JSON line per arm: iterations/s and ms per iteration (CUDA events, 5 warm-ups).

usage: python tools/bench_train_step.py [--arm ours|reference|both] [--iters 100] [--workload C2]
First hardware numbers: profiles/r02_first_hw_run.txt (bench.py carries the measured blocks now)."""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "guidedvd-3dgs_b200", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import synth  # noqa: E402

WORKLOADS = {"C2": (500_000, 640, 480, 20260002, 3), "small": (20_000, 160, 120, 20260001, 3)}


def ssim_reference(img1, img2):
    """utils/loss_utils.py:36-82 (create_window every call, five depthwise conv2d, elementwise map, mean)."""
    Cc = img1.size(-3)
    gw = torch.Tensor([math.exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    gw = (gw / gw.sum()).unsqueeze(1)
    win = gw.mm(gw.t()).float().unsqueeze(0).unsqueeze(0).expand(Cc, 1, 11, 11).contiguous().cuda(img1.get_device()).type_as(img1)
    F = torch.nn.functional
    mu1, mu2 = F.conv2d(img1, win, padding=5, groups=Cc), F.conv2d(img2, win, padding=5, groups=Cc)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, win, padding=5, groups=Cc) - mu1_sq
    s2 = F.conv2d(img2 * img2, win, padding=5, groups=Cc) - mu2_sq
    s12 = F.conv2d(img1 * img2, win, padding=5, groups=Cc) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()


def run(arm, P, W, H, seed, D, iters):
    dev = torch.device("cuda", 0)
    if arm in ("ours", "folded"):
        import diff_gaussian_rasterization as pkg
        import train_ops
    else:
        import refload
        pkg = refload.ref_dgr()
        if pkg is None:
            return {"impl": arm, "unavailable": "oracle/_ref not built"}
    sc = synth.synth_scene(P, seed, device=dev)
    cam = synth.synth_camera(seed + 1, W, H, device=dev)
    bg = torch.zeros(3, device=dev)
    gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(5)).to(dev)
    # raw (pre-activation) parameters, as GaussianModel holds them (scene/gaussian_model.py:46-60)
    M = sc["shs"].shape[1]
    raw = {"xyz": sc["means3D"].clone(), "f_dc": sc["shs"][:, :1].clone(), "f_rest": sc["shs"][:, 1:].clone(),
           "opacity": torch.logit(sc["opacities"].clamp(1e-4, 1 - 1e-4)), "scaling": torch.log(sc["scales"]),
           "rotation": sc["rotations"].clone()}
    params = {k: v.contiguous().requires_grad_(True) for k, v in raw.items()}
    lrs = {"xyz": 1.6e-4, "f_dc": 2.5e-3, "f_rest": 2.5e-3 / 20, "opacity": 5e-2, "scaling": 5e-3, "rotation": 1e-3}
    groups = [{"params": [params[k]], "lr": lrs[k], "name": k} for k in params]
    opt = (train_ops.FusedAdam if arm != "reference" else torch.optim.Adam)(groups, lr=0.0, eps=1e-15)
    accum, denom, maxr = torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev), torch.zeros(P, device=dev)
    lam = 0.2

    def iteration():
        means2D = torch.zeros_like(params["xyz"], requires_grad=True)
        settings = pkg.GaussianRasterizationSettings(
            image_height=cam["height"], image_width=cam["width"], tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=bg,
            scale_modifier=1.0, viewmatrix=cam["viewmatrix"], projmatrix=cam["projmatrix"], sh_degree=D, campos=cam["campos"],
            prefiltered=False, debug=False, confidence=sc["confidence"])
        rast = pkg.GaussianRasterizer(raster_settings=settings)
        if arm == "folded":
            image, radii, depth, alpha = pkg.rasterize_gaussians_raw(params["xyz"], means2D, params["f_dc"], params["f_rest"],
                                                                     params["opacity"].view(-1, 1), params["scaling"], params["rotation"],
                                                                     settings)
        else:
          # gaussian_renderer/__init__.py:60-87: activations + the SH concatenation
          image, radii, depth, alpha = rast(means3D=params["xyz"], means2D=means2D, opacities=torch.sigmoid(params["opacity"]),
                                          shs=torch.cat((params["f_dc"], params["f_rest"]), dim=1), scales=torch.exp(params["scaling"]),
                                          rotations=torch.nn.functional.normalize(params["rotation"]))
        if arm != "reference":
            loss = train_ops.photometric_loss(image, gt, lam)
        else:
            loss = (1.0 - lam) * torch.abs(image - gt).mean() + lam * (1.0 - ssim_reference(image, gt))
        loss.backward()
        with torch.no_grad():
            if arm != "reference":
                train_ops.add_densification_stats(means2D.grad, radii, accum.view(-1), denom.view(-1), maxr)
            else:
                vis = radii > 0
                maxr[vis] = torch.max(maxr[vis], radii[vis].float())
                accum[vis] += torch.norm(means2D.grad[vis, :2], dim=-1, keepdim=True)
                denom[vis] += 1
            opt.step()
            opt.zero_grad(set_to_none=True)
        return loss

    for _ in range(5):
        iteration()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        loss = iteration()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"metric": "3DGS training iterations/sec (render + L1/SSIM loss + backward + densification stats + Adam)", "impl": arm,
            "value": round(1e3 / ms, 2), "unit": "iterations/s", "ms_per_iteration": round(ms, 4), "iters": iters, "dtype": "f32",
            "final_loss": round(float(loss.detach()), 6), "config": {"P": P, "width": W, "height": H, "sh_degree": D, "lambda_dssim": lam}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", default="both", choices=["ours", "folded", "reference", "both"])
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--workload", default="C2", choices=list(WORKLOADS))
    a = ap.parse_args()
    P, W, H, seed, D = WORKLOADS[a.workload]
    for arm in (["ours", "folded", "reference"] if a.arm == "both" else [a.arm]):
        try:
            print(json.dumps(run(arm, P, W, H, seed, D, a.iters)), flush=True)
        except Exception as ex:  # one arm failing must not hide the other
            print(json.dumps({"impl": arm, "error": repr(ex)[:300]}), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
