"""Drop-in for the reference's pybind module `diff_gaussian_rasterization._C`
(DGR/ext.cpp:15-18; argument order DGR/rasterize_points.h:19-66), implemented over the
C ABI in include/gvd_raster.h.  Same three functions, same argument order, same returns.
Absent tensors are passed as empty tensors (numel()==0), exactly like the reference.
"""
import ctypes as C
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import gvd_native as _n  # noqa: E402


def _ptr(t):
    """Device pointer or None (NULL) for an absent (empty) tensor."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t, name):
    if t is None or t.numel() == 0:
        return t
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")
    return t.contiguous()


# Tests flip this to also get the sorted 64-bit keys (tile<<32 | depth bits) written next to point_list.
EXPORT_KEYS = False


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Allocs:
    """Caller-owned scratch, grown on demand by the library through C callbacks
    (the reference's resizeFunctional, DGR/rasterize_points.cu:27-33)."""

    def __init__(self, device):
        self.device = device
        self.geom = self.binning = self.img = None
        self.cb_geom = _n.ALLOC_FN(lambda u, n: self._alloc("geom", n))
        self.cb_binning = _n.ALLOC_FN(lambda u, n: self._alloc("binning", n))
        self.cb_img = _n.ALLOC_FN(lambda u, n: self._alloc("img", n))

    def _alloc(self, which, nbytes):
        t = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        setattr(self, which, t)
        return t.data_ptr()

    def take(self):
        """Hand the buffers over and drop the callback closures: they reference `self`, and a reference
        cycle would keep ~100 MB of scratch alive until the cyclic GC runs."""
        out = (self.geom, self.binning, self.img)
        self.geom = self.binning = self.img = None
        self.cb_geom = self.cb_binning = self.cb_img = None
        return out


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                        prefiltered, debug):
    """-> (num_rendered, out_color, out_depth, out_alpha, radii, geomBuffer, binningBuffer, imgBuffer)"""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    lib = _n.raster()
    dev = means3D.device
    P, H, W = means3D.size(0), int(image_height), int(image_width)
    f32 = dict(dtype=torch.float32, device=dev)
    u8 = dict(dtype=torch.uint8, device=dev)
    if P == 0:
        # the reference returns its zero-filled outputs untouched (rasterize_points.cu:68-85)
        return (0, torch.zeros(3, H, W, **f32), torch.zeros(1, H, W, **f32), torch.zeros(1, H, W, **f32),
                torch.zeros(0, dtype=torch.int32, device=dev), torch.empty(0, **u8), torch.empty(0, **u8),
                torch.empty(0, **u8))

    background = _f32c(background, "background")
    means3D = _f32c(means3D, "means3D")
    colors = _f32c(colors, "colors_precomp")
    opacity = _f32c(opacity, "opacities")
    scales = _f32c(scales, "scales")
    rotations = _f32c(rotations, "rotations")
    cov3D_precomp = _f32c(cov3D_precomp, "cov3D_precomp")
    viewmatrix = _f32c(viewmatrix, "viewmatrix")
    projmatrix = _f32c(projmatrix, "projmatrix")
    sh = _f32c(sh, "sh")
    campos = _f32c(campos, "campos")

    out_color = torch.empty(3, H, W, **f32)
    out_depth = torch.empty(1, H, W, **f32)
    out_alpha = torch.empty(1, H, W, **f32)
    radii = torch.empty(P, dtype=torch.int32, device=dev)

    allocs = _Allocs(dev)
    a = _n.RasterForwardArgs()
    a.P, a.D, a.M, a.width, a.height = P, int(degree), (sh.size(1) if sh is not None and sh.numel() else 0), W, H
    a.background = _ptr(background)
    a.means3D = _ptr(means3D)
    a.shs = _ptr(sh)
    a.colors_precomp = _ptr(colors)
    a.opacities = _ptr(opacity)
    a.scales = _ptr(scales)
    a.rotations = _ptr(rotations)
    a.cov3D_precomp = _ptr(cov3D_precomp)
    a.viewmatrix = _ptr(viewmatrix)
    a.projmatrix = _ptr(projmatrix)
    a.campos = _ptr(campos)
    a.scale_modifier, a.tan_fovx, a.tan_fovy = float(scale_modifier), float(tan_fovx), float(tan_fovy)
    a.prefiltered, a.debug = int(bool(prefiltered)), int(bool(debug))
    a.export_keys = int(bool(EXPORT_KEYS))
    a.out_color, a.out_depth, a.out_alpha, a.radii = (out_color.data_ptr(), out_depth.data_ptr(),
                                                        out_alpha.data_ptr(), radii.data_ptr())
    a.geom_alloc, a.binning_alloc, a.img_alloc = allocs.cb_geom, allocs.cb_binning, allocs.cb_img
    with torch.cuda.device(dev):
        rc = lib.gvd_raster_forward(C.byref(a), _stream())
    a.geom_alloc = a.binning_alloc = a.img_alloc = _n.ALLOC_FN(0)
    geom, binning, img = allocs.take()
    if rc != 0:
        raise RuntimeError("gvd_raster_forward failed: " + _n.last_error(lib))
    return (int(a.num_rendered), out_color, out_depth, out_alpha, radii, geom, binning, img)


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                 dL_dout_depth, dL_dout_alpha, sh, degree, campos, geomBuffer, R, binningBuffer,
                                 imageBuffer, alphas, debug, confidence=None):
    """-> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)

    Extra trailing `confidence` ([P,1] or [P]): when given, every gradient except dL_dmeans2D is
    multiplied by it inside the kernel (the reference does this in Python afterwards,
    diff_gaussian_rasterization/__init__.py:147-157)."""
    lib = _n.raster()
    dev = means3D.device
    P = means3D.size(0)
    H, W = dL_dout_color.size(1), dL_dout_color.size(2)
    M = sh.size(1) if sh is not None and sh.numel() else 0
    f32 = dict(dtype=torch.float32, device=dev)
    has_sh = M > 0
    has_scales = scales is not None and scales.numel() > 0
    has_cov = cov3D_precomp is not None and cov3D_precomp.numel() > 0
    has_colors = colors is not None and colors.numel() > 0

    if P == 0:
        z = lambda *s: torch.zeros(*s, **f32)  # noqa: E731
        return z(0, 3), z(0, 3), z(0, 1), z(0, 3), z(0, 6), z(0, M, 3), z(0, 3), z(0, 4)

    e = lambda *s: torch.empty(*s, **f32)  # noqa: E731
    dL_dmeans2D, dL_dmeans3D, dL_dopacity = e(P, 3), e(P, 3), e(P, 1)
    # gradients of absent inputs are never consumed; return zeros-shaped views like the reference
    dL_dcolors = e(P, 3) if has_colors else None
    dL_dcov3D = e(P, 6) if has_cov else None
    dL_dsh = e(P, M, 3) if has_sh else None
    dL_dscales = e(P, 3) if has_scales else None
    dL_drotations = e(P, 4) if has_scales else None
    scratch = torch.empty(int(lib.gvd_raster_backward_scratch_bytes(P)), dtype=torch.uint8, device=dev)

    keep = [_f32c(t, n) for t, n in ((background, "background"), (means3D, "means3D"), (colors, "colors_precomp"),
                                     (scales, "scales"), (rotations, "rotations"), (cov3D_precomp, "cov3D_precomp"),
                                     (viewmatrix, "viewmatrix"), (projmatrix, "projmatrix"), (sh, "sh"),
                                     (campos, "campos"), (dL_dout_color, "dL_dout_color"),
                                     (dL_dout_depth, "dL_dout_depth"), (dL_dout_alpha, "dL_dout_alpha"),
                                     (alphas, "alphas"), (confidence, "confidence"))]
    (background, means3D, colors, scales, rotations, cov3D_precomp, viewmatrix, projmatrix, sh, campos,
     dL_dout_color, dL_dout_depth, dL_dout_alpha, alphas, confidence) = keep
    radii = radii.contiguous()

    a = _n.RasterBackwardArgs()
    a.P, a.D, a.M, a.R, a.width, a.height = P, int(degree), M, int(R), W, H
    a.background, a.means3D, a.shs = _ptr(background), _ptr(means3D), _ptr(sh)
    a.colors_precomp, a.scales, a.rotations = _ptr(colors), _ptr(scales), _ptr(rotations)
    a.cov3D_precomp, a.viewmatrix, a.projmatrix, a.campos = (_ptr(cov3D_precomp), _ptr(viewmatrix),
                                                             _ptr(projmatrix), _ptr(campos))
    a.scale_modifier, a.tan_fovx, a.tan_fovy = float(scale_modifier), float(tan_fovx), float(tan_fovy)
    a.radii, a.alphas = _ptr(radii), _ptr(alphas)
    a.geom_buffer, a.binning_buffer, a.img_buffer = _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer)
    a.dL_dpix, a.dL_ddepth_pix, a.dL_dalpha_pix = _ptr(dL_dout_color), _ptr(dL_dout_depth), _ptr(dL_dout_alpha)
    a.confidence = _ptr(confidence)
    a.scratch = scratch.data_ptr()
    a.dL_dmeans2D, a.dL_dmeans3D, a.dL_dopacity = (dL_dmeans2D.data_ptr(), dL_dmeans3D.data_ptr(),
                                                   dL_dopacity.data_ptr())
    a.dL_dcolors, a.dL_dcov3D, a.dL_dsh = _ptr(dL_dcolors), _ptr(dL_dcov3D), _ptr(dL_dsh)
    a.dL_dscales, a.dL_drotations = _ptr(dL_dscales), _ptr(dL_drotations)
    a.debug = int(bool(debug))
    with torch.cuda.device(dev):
        rc = lib.gvd_raster_backward(C.byref(a), _stream())
    if rc != 0:
        raise RuntimeError("gvd_raster_backward failed: " + _n.last_error(lib))
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations


def mark_visible(means3D, viewmatrix, projmatrix):
    """-> bool[P] (DGR/rasterize_points.cu:210-229)"""
    lib = _n.raster()
    P = means3D.size(0)
    present = torch.zeros(P, dtype=torch.bool, device=means3D.device)
    if P != 0:
        means3D, viewmatrix, projmatrix = (_f32c(means3D, "means3D"), _f32c(viewmatrix, "viewmatrix"),
                                           _f32c(projmatrix, "projmatrix"))
        with torch.cuda.device(means3D.device):
            rc = lib.gvd_raster_mark_visible(P, means3D.data_ptr(), viewmatrix.data_ptr(), projmatrix.data_ptr(),
                                             present.data_ptr(), _stream())
        if rc != 0:
            raise RuntimeError("gvd_raster_mark_visible failed: " + _n.last_error(lib))
    return present
