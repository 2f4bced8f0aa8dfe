"""Drop-ins for the per-iteration pieces of the reference's 3DGS training step that sit around the rasterizer
(SURVEY.md section 8f row f3), computed by the sm_100a kernels behind include/gvd_train.h:

  l1_loss, l1_loss_mask, ssim   same names / arguments as utils/loss_utils.py:18-28,46-82 (window 11, mean; optional mask)
  photometric_loss              (1 - lambda) * l1_loss + lambda * (1 - ssim) as train_baseline.py:82-83 combines them,
                                one fused forward and one fused backward launch
  add_densification_stats       scene/gaussian_model.py:524-527 + train_baseline.py:109 without boolean-mask indexing
                                (each `x[mask]` of the reference is a nonzero() and a host synchronisation)
  mask_erosion, mask_dilation,  utils/viewcrafter_wrapper.py:602-647 (scipy.ndimage on the CPU, one mask at a time) on device
  decide_unobserved_regions     tensors, any batch of masks in one launch
  FusedAdam                     torch.optim.Adam(eps=1e-15) semantics and state layout (`step`, `exp_avg`, `exp_avg_sq`), so
                                the reference's optimizer surgery at densification (gaussian_model.py:379-470) keeps working

There is no CPU path: a missing library or a CPU tensor raises.
"""
import ctypes as C

import torch

import gvd_native as _n


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch._C._cuda_getDevice()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed: " + (_n.train().gvd_train_last_error() or b"").decode())


def _chw(t):
    if t.dim() == 4 and t.shape[0] == 1:
        t = t[0]
    if t.dim() != 3:
        raise ValueError("expected an image [C, H, W] (or [1, C, H, W])")
    return t


class _PhotometricLoss(torch.autograd.Function):
    """-> (mean |img - gt|, mean SSIM); gradients with respect to img only (gt is data)."""

    @staticmethod
    def forward(ctx, img, gt):
        lib = _n.train()
        x, y = _chw(img).float().contiguous(), _chw(gt).float().contiguous()
        if x.shape != y.shape:
            raise ValueError("image and ground truth differ in shape")
        Cc, H, W = x.shape
        out = torch.empty(2, dtype=torch.float32, device=x.device)
        need = ctx.needs_input_grad[0]
        dmaps = torch.empty(3, Cc, H, W, dtype=torch.float32, device=x.device) if need else None
        nby = int(lib.gvd_photometric_loss_scratch_bytes(Cc, H, W))
        scratch = torch.empty(nby // 8 + 1, dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            _check(lib.gvd_photometric_loss_forward(x.data_ptr(), y.data_ptr(), Cc, H, W, out.data_ptr(),
                                                    None if dmaps is None else dmaps.data_ptr(), scratch.data_ptr(), nby, _stream()),
                   "gvd_photometric_loss_forward")
        if need:
            ctx.save_for_backward(x, y, dmaps)
        ctx.in_shape = img.shape
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_l1, g_ssim):
        lib = _n.train()
        x, y, dmaps = ctx.saved_tensors
        Cc, H, W = x.shape
        coef = torch.stack([g_l1.reshape(()), g_ssim.reshape(())]).float().contiguous()  # stays on the device
        dimg = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _check(lib.gvd_photometric_loss_backward(x.data_ptr(), y.data_ptr(), dmaps.data_ptr(), coef.data_ptr(), Cc, H, W,
                                                     dimg.data_ptr(), _stream()), "gvd_photometric_loss_backward")
        return dimg.view(ctx.in_shape), None


def l1_loss(network_output, gt, return_map=False):
    """utils/loss_utils.py:18-22."""
    if return_map:
        return torch.abs(network_output - gt)
    return _PhotometricLoss.apply(network_output, gt)[0]


def l1_loss_mask(network_output, gt, mask=None):
    """utils/loss_utils.py:24-28: without a mask the plain L1 mean (the training loops' call); with one, the masked mean
    of the evaluation code (three elementwise torch launches, not worth a kernel)."""
    if mask is None:
        return l1_loss(network_output, gt)
    return torch.abs((network_output - gt) * mask).sum() / mask.sum()


def ssim(img1, img2, mask=None, window_size=11, size_average=True):
    """utils/loss_utils.py:46-82 (11x11 window, global mean).  With a mask both images are blended towards 1 outside it
    first (:50-52), exactly as the reference does before its convolutions."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("train_ops.ssim: the reference's callers use window_size=11, size_average=True")
    if mask is not None:
        img1 = img1 * mask + (1 - mask)
        img2 = img2 * mask + (1 - mask)
    return _PhotometricLoss.apply(img1, img2)[1]


def photometric_loss(image, gt, lambda_dssim=0.2):
    """(1 - lambda) * L1 + lambda * (1 - SSIM)   (train_baseline.py:82-83, train_guidedvd.py's main loss)."""
    l1, s = _PhotometricLoss.apply(image, gt)
    return (1.0 - lambda_dssim) * l1 + lambda_dssim * (1.0 - s)


@torch.no_grad()
def add_densification_stats(means2D_grad, radii, xyz_gradient_accum, denom, max_radii2D):
    """In place, for every visible Gaussian (radii > 0): accum += ||grad[:, :2]||, denom += 1, max_radii2D = max(., radii)."""
    lib = _n.train()
    P = radii.shape[0]
    g = means2D_grad.float().contiguous()
    r = radii.to(torch.int32).contiguous()
    for t in (xyz_gradient_accum, denom, max_radii2D):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != P:
            raise ValueError("accumulators must be contiguous float32 tensors with one entry per Gaussian")
    with torch.cuda.device(g.device):
        _check(lib.gvd_densification_stats(g.data_ptr(), r.data_ptr(), P, xyz_gradient_accum.data_ptr(), denom.data_ptr(),
                                           max_radii2D.data_ptr(), _stream()), "gvd_densification_stats")


def _morph(mask, size, dilate):
    lib = _n.train()
    m = mask.float().contiguous()
    if m.dim() < 2:
        raise ValueError("mask must be [..., H, W]")
    H, W = m.shape[-2:]
    N = m.numel() // (H * W) if H * W else 0
    out = torch.empty_like(m)
    c = size // 2   # scipy centres a size-k structure at index k // 2; dilation uses the mirrored window
    lo, hi = (-c, size - 1 - c) if not dilate else (-(size - 1 - c), c)
    with torch.cuda.device(m.device):
        _check(lib.gvd_mask_morphology(m.data_ptr(), out.data_ptr(), N, H, W, lo, hi, lo, hi, int(dilate), _stream()), "gvd_mask_morphology")
    return out.to(mask.dtype)


def mask_erosion(mask, size=3):
    """ViewCrafterWrapper.mask_erosion (utils/viewcrafter_wrapper.py:617-619) on device masks [..., H, W], any batch."""
    return _morph(mask, size, False)


def mask_dilation(mask, size=5):
    """ViewCrafterWrapper.mask_dilation (:621-624)."""
    return _morph(mask, size, True)


def decide_unobserved_regions(gs_render_results):
    """utils/viewcrafter_wrapper.py:602-615: pixels no Gaussian covered (rendering exactly 0), eroded by 3, dilated by 5.
    [N, 3, H, W] in [0, 1] -> [N, 1, H, W] float masks, without the trip through numpy / scipy."""
    empty = (gs_render_results.sum(1) == 0.).to(torch.float32)
    return mask_dilation(mask_erosion(empty, 3), 5).unsqueeze(1)


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas, eps) without weight decay / amsgrad, one launch per parameter tensor."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _n.train()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise ValueError("FusedAdam: contiguous float32 parameters only")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                g = p.grad.float().contiguous()
                with torch.cuda.device(p.device):
                    _check(lib.gvd_adam_step(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(),
                                             float(group["lr"]), float(b1), float(b2), float(group["eps"]), int(st["step"].item()),
                                             _stream()), "gvd_adam_step")
        return loss
