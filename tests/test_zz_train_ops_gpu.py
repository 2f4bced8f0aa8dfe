"""GPU parity of the training-step kernels (row f3, include/gvd_train.h) through train_ops.py: fused L1 + SSIM loss
forward/backward against the reference's goldens (tests/golden/loss_*.npz) and, at the benchmark's image size
(3 x 480 x 640), against a plain PyTorch fp32 statement of utils/loss_utils.py:46-82 with autograd; FusedAdam against
torch.optim.Adam(eps=1e-15); densification statistics against the reference's masked statements.
Tolerances: 1e-5 relative on loss values, 2e-4 relative L2 on gradients (fp32, different summation order).

Green on B200 since round 2 (profiles/r02_first_hw_run.txt); the same source also runs on the host (tests/test_train_ops_cpu.py)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = [pytest.mark.gpu]


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _ssim_torch(img1, img2):
    """utils/loss_utils.py:36-82 restated (window 11, sigma 1.5, zero padding, global mean)."""
    Cc = img1.shape[0]
    gw = torch.tensor([np.exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)], dtype=torch.float32)
    gw = (gw / gw.sum()).unsqueeze(1)
    win = (gw @ gw.t()).expand(Cc, 1, 11, 11).contiguous().to(img1.device)
    conv = lambda t: torch.nn.functional.conv2d(t.unsqueeze(0), win, padding=5, groups=Cc)[0]  # noqa: E731
    mu1, mu2 = conv(img1), conv(img2)
    s1, s2, s12 = conv(img1 * img1) - mu1 * mu1, conv(img2 * img2) - mu2 * mu2, conv(img1 * img2) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))).mean()


@pytest.mark.parametrize("name", ["small", "tile_edges", "one_channel", "tiny"])
def test_loss_matches_reference_golden(name):
    import train_ops
    from make_golden_loss import CASES, LAMBDA, loss_case

    Cc, H, W, seed = CASES[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", f"loss_{name}.npz"))
    img, gt = (t.cuda() for t in loss_case(Cc, H, W, seed))
    x = img.clone().requires_grad_(True)
    loss = train_ops.photometric_loss(x, gt, LAMBDA)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert _rel(x.grad, g["grad"]) < 2e-4


def test_loss_at_benchmark_size_and_determinism():
    import train_ops
    from make_golden_loss import loss_case

    img, gt = (t.cuda() for t in loss_case(3, 480, 640, 11))
    xr = img.clone().requires_grad_(True)
    ref = 0.8 * (xr - gt).abs().mean() + 0.2 * (1.0 - _ssim_torch(xr, gt))
    ref.backward()
    x = img.clone().requires_grad_(True)
    loss = train_ops.photometric_loss(x, gt, 0.2)
    loss.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item()) and _rel(x.grad, xr.grad) < 2e-4
    assert train_ops.photometric_loss(img, gt, 0.2).item() == loss.item()


def test_fused_adam_and_densification_stats():
    import train_ops

    g = torch.Generator().manual_seed(3)
    rp = [torch.randn(s, generator=g).cuda().requires_grad_(True) for s in [(50001, 3), (50001, 45), (777,)]]
    op = [p.detach().clone().requires_grad_(True) for p in rp]
    ref = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(rp, (1.6e-4, 1.25e-4, 5e-2))], lr=0.0, eps=1e-15)
    ours = train_ops.FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(op, (1.6e-4, 1.25e-4, 5e-2))], lr=0.0, eps=1e-15)
    for it in range(10):
        for a, b in zip(rp, op):
            grad = (torch.randn(a.shape, generator=g) * 1e-3).cuda()
            grad[::3] = 0.0
            a.grad, b.grad = grad.clone(), grad.clone()
        ref.step()
        ours.step()
    for a, b in zip(rp, op):
        assert torch.allclose(b, a, rtol=2e-6, atol=1e-9)
        assert torch.allclose(ours.state[b]["exp_avg_sq"], ref.state[a]["exp_avg_sq"], rtol=1e-5, atol=1e-16)
    P = 200003
    grad = (torch.randn(P, 3, generator=g) * 1e-3).cuda()
    radii = torch.randint(-1, 40, (P,), generator=g, dtype=torch.int32).clamp_min(0).cuda()
    accum, denom, maxr = torch.rand(P, generator=g).cuda(), torch.zeros(P).cuda(), (torch.rand(P, generator=g) * 30).cuda()
    r_accum, r_denom, r_maxr = accum.clone().unsqueeze(1), denom.clone().unsqueeze(1), maxr.clone()
    vis = radii > 0
    r_accum[vis] += torch.norm(grad[vis, :2], dim=-1, keepdim=True)
    r_denom[vis] += 1
    r_maxr[vis] = torch.max(r_maxr[vis], radii[vis].float())
    train_ops.add_densification_stats(grad, radii, accum, denom, maxr)
    assert torch.allclose(accum, r_accum[:, 0], rtol=1e-6, atol=0) and torch.equal(denom, r_denom[:, 0]) and torch.equal(maxr, r_maxr)


def test_mask_morphology_matches_scipy_on_gpu():
    import train_ops
    from scipy import ndimage

    rng = np.random.default_rng(7)
    masks = (rng.random((25, 320, 512)) > 0.4).astype(np.float32)
    t = torch.from_numpy(masks).cuda()
    for k, dilate in ((3, False), (5, True), (5, False), (10, True)):
        fn = ndimage.binary_dilation if dilate else ndimage.binary_erosion
        want = np.stack([fn(m, structure=np.ones((k, k))).astype(np.float32) for m in masks[:4]])
        got = (train_ops.mask_dilation if dilate else train_ops.mask_erosion)(t, k)
        assert np.array_equal(got[:4].cpu().numpy(), want), (k, dilate)
