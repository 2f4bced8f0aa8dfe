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


# Speculative sizing of the instance buffer: after the first frame on a device the shim allocates the buffer from the
# previous frame's R (+25 %) and lets the library queue every stage without the host round trip the reference makes
# (rasterizer_impl.cu:281-282); R arrives through pinned memory and is validated here, falling back to the exact,
# synchronous path when a frame outgrows the guess.
SPECULATE = os.environ.get("GVD_SPECULATE", "1") != "0"
_spec_state = {}


def _spec(dev):
    st = _spec_state.get(dev)
    if st is None:
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))  # forces creation of the underlying cudaEvent_t
        st = _spec_state[dev] = {"last_R": None, "pinned": torch.zeros(1, dtype=torch.int32).pin_memory(), "event": ev}
    return st


# Optional caller-owned gradient storage per device (set_gradient_buffer): the backward then writes its gradients there
# instead of into fresh memory, e.g. into the peer-mapped exchange buffer of view_parallel.GradientExchange so the
# cross-GPU sum needs no packing copy.  The gradients of two backward calls then alias: consume them in between.
_grad_buffer = {}


def gradient_buffer_floats(P, M=16):
    """Floats a gradient buffer must hold for P Gaussians with M SH coefficients (every optional gradient present)."""
    return sum((P * w + 3) // 4 * 4 for w in (3, 3 * M, 1, 3, 4, 3, 6, 3))


def set_gradient_buffer(buf):
    """buf: 1-D float32 CUDA tensor (>= 62 floats per Gaussian + padding) or None to restore fresh allocations."""
    if buf is None:
        _grad_buffer.clear()
    else:
        _grad_buffer[buf.device] = buf


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
    st = _spec(dev) if (SPECULATE and not debug and not EXPORT_KEYS) else None
    spec_buf = None
    with torch.cuda.device(dev):
        if st is not None and st["last_R"] is not None:
            nbytes = int(lib.gvd_raster_binning_bytes(int(st["last_R"] * 1.25) + 65536, 0))
            spec_buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            a.spec_binning_buffer, a.spec_binning_bytes = spec_buf.data_ptr(), nbytes
            a.num_rendered_pinned, a.r_ready_event = st["pinned"].data_ptr(), st["event"].cuda_event
        rc = lib.gvd_raster_forward(C.byref(a), _stream())
        if rc == 0 and spec_buf is not None:
            st["event"].synchronize()  # waits for the scan, not for the render kernels queued behind it
            R = int(st["pinned"][0])
            if int(lib.gvd_raster_binning_bytes(R, 0)) > a.spec_binning_bytes:
                # the guess was too small: redo this frame on the exact path
                a.spec_binning_buffer, a.spec_binning_bytes, spec_buf = None, 0, None
                rc = lib.gvd_raster_forward(C.byref(a), _stream())
            else:
                a.num_rendered = R
    a.geom_alloc = a.binning_alloc = a.img_alloc = _n.ALLOC_FN(0)
    geom, binning, img = allocs.take()
    if spec_buf is not None:
        binning = spec_buf
    if rc != 0:
        raise RuntimeError("gvd_raster_forward failed: " + _n.last_error(lib))
    if st is not None:
        st["last_R"] = int(a.num_rendered)
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

    # All gradients are views of ONE flat allocation (16-byte aligned slices), so a data-parallel caller can sum them
    # across ranks with a single collective on `grad.untyped_storage()` / `._base` without packing copies.
    widths = [("means3D", 3), ("sh", 3 * M if has_sh else 0), ("opacity", 1), ("scales", 3 if has_scales else 0),
              ("rot", 4 if has_scales else 0), ("colors", 3 if has_colors else 0), ("cov", 6 if has_cov else 0),
              ("means2D", 3)]
    offs, total = {}, 0
    for name, w in widths:
        offs[name] = total
        total += (P * w + 3) // 4 * 4
    ext = _grad_buffer.get(dev)
    flat = ext[:total] if (ext is not None and ext.numel() >= total) else torch.empty(total, **f32)

    def view(name, *shape):
        n = 1
        for d in shape:
            n *= d
        return flat[offs[name]:offs[name] + n].view(*shape)

    dL_dmeans2D, dL_dmeans3D, dL_dopacity = view("means2D", P, 3), view("means3D", P, 3), view("opacity", P, 1)
    # gradients of absent inputs are never consumed: None instead of the reference's unused zero tensors
    dL_dcolors = view("colors", P, 3) if has_colors else None
    dL_dcov3D = view("cov", P, 6) if has_cov else None
    dL_dsh = view("sh", P, M, 3) if has_sh else None
    dL_dscales = view("scales", P, 3) if has_scales else None
    dL_drotations = view("rot", P, 4) if has_scales else None
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
