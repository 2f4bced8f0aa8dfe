"""Generate tests/golden/loss_*.npz by running the REFERENCE loss functions (utils/loss_utils.py: l1_loss, ssim) and
torch.autograd over them in this container on the CPU.  Inputs are seeded (loss_case below), outputs are stored.
Run: python tests/make_golden_loss.py   (needs /root/reference; not needed on the GPU box)."""
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {"small": (3, 37, 45, 0), "tile_edges": (3, 16, 33, 1), "one_channel": (1, 50, 20, 2), "tiny": (3, 5, 7, 3)}
LAMBDA = 0.2


def loss_case(C, H, W, seed):
    """A rendering-like pair: smooth image + noise vs its ground truth, both in [0, 1]."""
    g = torch.Generator().manual_seed(1000 + seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 3, H), torch.linspace(0, 4, W), indexing="ij")
    base = torch.stack([0.5 + 0.4 * torch.sin(xx * (c + 1) + yy) for c in range(C)])
    gt = (base + 0.05 * torch.randn(C, H, W, generator=g)).clamp(0, 1)
    img = (base + 0.15 * torch.randn(C, H, W, generator=g)).clamp(0, 1)
    img[:, : H // 3, : W // 4] = gt[:, : H // 3, : W // 4]   # a region of exact agreement (|x - y| = 0: sign(0) = 0)
    return img, gt


def main():
    spec = importlib.util.spec_from_file_location("ref_loss_utils", "/root/reference/utils/loss_utils.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for name, (C, H, W, seed) in CASES.items():
        img, gt = loss_case(C, H, W, seed)
        x = img.clone().requires_grad_(True)
        l1 = ref.l1_loss(x, gt)
        s = ref.ssim(x, gt)
        loss = (1.0 - LAMBDA) * l1 + LAMBDA * (1.0 - s)     # train_baseline.py:82-83
        loss.backward()
        xs = img.clone().requires_grad_(True)
        ref.ssim(xs, gt).backward()
        # the evaluation code's masked variants (train_guidedvd.py:691; loss_utils.py:24-28,50-52)
        mask = (torch.rand(1, H, W, generator=torch.Generator().manual_seed(77 + seed)) > 0.3).float()
        xm = img.clone().requires_grad_(True)
        sm = ref.ssim(xm, gt, mask)
        sm.backward()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"loss_{name}.npz"), l1=l1.item(), ssim=s.item(), loss=loss.item(),
                            grad=x.grad.numpy(), grad_ssim=xs.grad.numpy(), ssim_masked=sm.item(), grad_ssim_masked=xm.grad.numpy(),
                            l1_masked=ref.l1_loss_mask(img, gt, mask).item())
        print(name, "l1", l1.item(), "ssim", s.item())


if __name__ == "__main__":
    main()
