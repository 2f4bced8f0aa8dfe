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


_F32 = torch.float32


def _f32c(t, name):
    if t is None or (t.dtype is _F32 and t.is_cuda and t.is_contiguous()):  # the common case first: one pass, no numel()
        return t
    if t.numel() == 0:
        return t
    if t.dtype is not _F32 or not t.is_cuda:
        raise RuntimeError(f"{name} must be a float32 CUDA tensor")
    return t.contiguous()


# Tests flip this to also get the sorted 64-bit keys (tile<<32 | depth bits) written next to point_list.
EXPORT_KEYS = False


# How the instance buffer (4 bytes x R) and the chunk histogram (sized by V) get their sizes.  R and V are produced by
# the second kernel of the forward, ~40 us into the frame (include/gvd_raster.h), and stored straight into pinned
# host memory.  GVD_SPECULATE selects:
#   sync (default): from the second frame on a device the buffers are sized from the largest R / V seen so far (x2 +
#     slack), every stage is queued in one go, and the call waits for the two counts -- not for the kernels behind them
#     -- BEFORE it returns: a frame that outgrew the guess is redone on the exact path and nobody ever sees its
#     clamped outputs.  Always valid, never raises later, and the GPU has ~0.25 ms of queued work when the caller gets
#     its tensors back.
#   exact: the library itself waits for the counts (the depth sort is already queued behind them) and calls back for
#     exactly sized buffers: the reference's contract (rasterizer_impl.cu:281-286) without its pipeline bubble; the
#     first frame on a device, debug mode and overflowed frames always take this path.
#   defer (opt-in; bench.py names it in `config` when used): like sync, but a frame that will get a backward is not
#     validated until that backward starts (or, for a frame that never gets one, at the next forward), so the host
#     can run a whole step ahead of the GPU.  A frame that outgrew its buffers has already been consumed by then; the
#     backward repairs what it can -- it re-renders the frame exactly, warns, and returns the gradients of the exact
#     frame -- so an unmodified trainer keeps running.
_MODE = os.environ.get("GVD_SPECULATE", "sync")
if _MODE == "1":
    _MODE = "defer"
elif _MODE in ("0",):
    _MODE = "exact"
DEFER = _MODE == "defer"
SPECULATE = _MODE in ("sync", "defer")
_spec_state = {}


class SpeculationOverflow(RuntimeError):
    """A deferred frame produced more instances / visible Gaussians than its speculative buffers held."""


class Counts(int):
    """R (the reference's `num_rendered`) as an int that also carries V, the number of visible Gaussians."""

    def __new__(cls, R, V):
        o = int.__new__(cls, R)
        o.visible = int(V)
        return o


def _new_slot(dev):
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))  # forces creation of the underlying cudaEvent_t
    pinned = torch.zeros(2, dtype=torch.int32).pin_memory()
    return {"pinned": pinned, "host": pinned.numpy(), "ptr": pinned.data_ptr(), "event": ev, "cuda_event": ev.cuda_event}


def _spec(dev):
    st = _spec_state.get(dev)
    if st is None:
        st = _spec_state[dev] = {"max_R": None, "max_V": None, "free": [_new_slot(dev)], "open": []}
    return st


def _capacity(max_R):
    return 2 * int(max_R) + (1 << 20)


def _visible_capacity(max_V, P):
    return min(int(P), 2 * int(max_V) + (1 << 14))


def _note(st, R, V):
    if st["max_R"] is None or R > st["max_R"]:
        st["max_R"] = R
    if st["max_V"] is None or V > st["max_V"]:
        st["max_V"] = V


class PendingR:
    """R and V of a forward whose validation was deferred (see DEFER above).  int(obj) / obj.resolve() waits for the
    second kernel of that frame (not for its render kernels), validates the speculative buffers and returns a Counts."""
    __slots__ = ("slot", "cap", "vcap", "st", "value", "error", "__weakref__")

    def __init__(self, slot, cap, vcap, st):
        self.slot, self.cap, self.vcap, self.st, self.value, self.error = slot, cap, vcap, st, None, None

    def done(self):
        return self.slot is None or self.slot["event"].query()

    def resolve(self):
        if self.value is None:
            slot, self.slot = self.slot, None
            slot["event"].synchronize()
            R, V = int(slot["host"][0]), int(slot["host"][1])
            self.value = Counts(R, V)
            self.st["free"].append(slot)
            _note(self.st, R, V)
            if R + 128 > self.cap or V > self.vcap:
                self.error = (f"diff_gaussian_rasterization: this frame produced R={R} instances from V={V} visible Gaussians, "
                              f"more than its speculative buffers (R <= {self.cap - 128}, V <= {self.vcap}) sized from earlier "
                              "frames; the image it returned is incomplete.  GVD_SPECULATE=exact (the default) never does this.")
        if self.error is not None:
            raise SpeculationOverflow(self.error)
        return self.value

    def __int__(self):
        return int(self.resolve())

    __index__ = __int__

    def __del__(self):
        if self.value is None and self.slot is not None:
            try:
                self.resolve()
            except Exception as ex:  # never raise from a finaliser
                import warnings
                warnings.warn(str(ex))


def _settle_open(st):
    """Deferred frames that never reached a backward (renders under grad mode that were only looked at): validate the
    ones whose counts have arrived, so an overflow is reported at the next render instead of at garbage collection."""
    import warnings

    keep = []
    for ref in st["open"]:
        p = ref()
        if p is None or p.value is not None or p.error is not None:
            continue
        if p.done():
            try:
                p.resolve()
            except SpeculationOverflow as ex:
                warnings.warn(str(ex))
        else:
            keep.append(ref)
    st["open"] = keep


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
        _last_views.clear()
        _view_cache.clear()
        _grad_pending.clear()
    else:
        _grad_buffer[buf.device] = buf


_last_views = {}
_grad_pending = {}   # device -> True while the caller-owned buffer holds a backward's gradients nobody has asked for yet
_warned_alias = []


def gradient_views(device):
    """With a caller-owned gradient buffer (set_gradient_buffer): the views of it that the most recent backward on
    `device` filled, keyed by the rasterizer input they belong to.  autograd's AccumulateGrad does not adopt gradient
    tensors that are views of a larger buffer (it copies them into `leaf.grad`), so a data-parallel caller sums the
    BUFFER across ranks and then points `leaf.grad` at these views (view_parallel.allreduce_gradients).
    Asking for the views marks the buffer as consumed: the next backward may write into it again.  A backward that
    arrives BEFORE that (two renders feeding one loss.backward(), as train_guidedvd.py does with a train view and a
    pseudo view) would overwrite gradients autograd still holds views of, so it gets fresh memory instead."""
    dev = torch.device(device)
    _grad_pending[dev] = False
    return _last_views.get(dev)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(dev=None):
    """The current stream of `dev` as a cudaStream_t.  torch.cuda.current_stream() builds a Stream object and resolves the
    device on every call (26 us per call on the bench box, twice per step -- 10 % of the host time of a 0.5 ms step);
    the raw-handle query behind it costs ~1 us."""
    if _raw_stream is not None:
        idx = dev.index if dev is not None and dev.index is not None else torch.cuda.current_device()
        return C.c_void_p(_raw_stream(idx))
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _on_device:
    """`with torch.cuda.device(dev)` only when dev is not already current (the guard costs ~10 us per use)."""
    __slots__ = ("guard",)

    def __init__(self, dev):
        self.guard = None if dev.index is None or torch.cuda.current_device() == dev.index else torch.cuda.device(dev)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()

    def __exit__(self, *exc):
        if self.guard is not None:
            self.guard.__exit__(*exc)
        return False


class _Alloc:
    """Caller-owned scratch handed to the library through a C callback once the size is known (the reference's
    resizeFunctional, DGR/rasterize_points.cu:27-33).  Every buffer it hands out stays referenced in `bufs`."""

    def __init__(self, device):
        self.device, self.bufs = device, []
        self.cb = _n.ALLOC_FN(self._alloc)

    def _alloc(self, user, nbytes):
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        self.bufs.append(buf)
        return buf.data_ptr()

    def take(self):
        """Hand the buffers over and drop the callback closure: it references `self`, and a reference cycle would
        keep the scratch alive until the cyclic GC runs."""
        out, self.bufs, self.cb = self.bufs, [], None
        return out


_scratch_sizes = {}


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                        prefiltered, debug, defer=False, force_defer=False, sh_rest=None):
    """-> (num_rendered, out_color, out_depth, out_alpha, radii, geomBuffer, binningBuffer, imgBuffer)

    defer=True (used by the autograd Function): under GVD_SPECULATE=defer num_rendered may come back as a PendingR, see
    DEFER above.  force_defer=True: always (when a history exists) -- for callers that validate a whole batch of frames
    themselves before anybody sees them (rasterize_views).
    sh_rest (B200 extension): when given, `scales`, `rotations`, `opacity` are the RAW GaussianModel
    parameters and `sh` / `sh_rest` are `_features_dc` / `_features_rest`; the kernels apply the activations."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    lib = _n.raster()
    dev = means3D.device
    P, H, W = means3D.size(0), int(image_height), int(image_width)
    if P == 0:
        # the reference returns its zero-filled outputs untouched (rasterize_points.cu:68-85)
        f32 = dict(dtype=torch.float32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
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
    sh_rest = _f32c(sh_rest, "sh_rest")
    campos = _f32c(campos, "campos")

    sizes = _scratch_sizes.get((P, W, H))
    if sizes is None:
        sizes = _scratch_sizes[(P, W, H)] = (int(lib.gvd_raster_geom_bytes(P, W, H)), int(lib.gvd_raster_img_bytes(W, H)),
                                             int(lib.gvd_raster_sort_bytes(P)), int(lib.gvd_raster_hist_bytes(0, W, H)),
                                             4 * ((W + 15) // 16) * ((H + 15) // 16))
    plain = debug or EXPORT_KEYS          # debug / parity runs: the library's own synchronous route, no pinned words
    st = None if plain else _spec(dev)
    pending = None
    with _on_device(dev):
        out_color = torch.empty((3, H, W), dtype=_F32, device=dev)
        out_depth = torch.empty((1, H, W), dtype=_F32, device=dev)
        out_alpha = torch.empty((1, H, W), dtype=_F32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        geom = torch.empty((sizes[0],), dtype=torch.uint8, device=dev)
        img = torch.empty((sizes[1],), dtype=torch.uint8, device=dev)
        sort = torch.empty((sizes[2],), dtype=torch.uint8, device=dev)  # forward-only; released when this call returns

        a = _n.RasterForwardArgs()
        a.P, a.D, a.M, a.width, a.height = P, int(degree), (sh.size(1) if sh is not None and sh.numel() else 0), W, H
        if sh_rest is not None:  # raw parameters: M counts the coefficients of both tensors
            a.M = 1 + sh_rest.size(1)
            a.raw_params, a.shs_rest = 1, _ptr(sh_rest)
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
        a.geom_buffer, a.geom_bytes, a.img_buffer, a.img_bytes = geom.data_ptr(), sizes[0], img.data_ptr(), sizes[1]
        a.sort_buffer, a.sort_bytes = sort.data_ptr(), sizes[2]
        stream = _stream(dev)

        binning, rc, done = None, 0, False
        if st is not None and st["open"]:
            _settle_open(st)
        if st is not None and SPECULATE and st["max_R"] is not None:
            cap, vcap = _capacity(st["max_R"]), _visible_capacity(st["max_V"], P)
            nbytes = 4 * cap + 1024  # >= gvd_raster_binning_bytes(cap, 0)
            hbytes = sizes[3] + sizes[4] * ((vcap + 63) // 64) + 1024  # >= gvd_raster_hist_bytes(vcap, W, H)
            binning = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            hist = torch.empty((hbytes,), dtype=torch.uint8, device=dev)
            slot = st["free"].pop() if st["free"] else _new_slot(dev)
            a.spec_binning_buffer, a.spec_binning_bytes = binning.data_ptr(), nbytes
            a.spec_hist_buffer, a.spec_hist_bytes = hist.data_ptr(), hbytes
            a.num_rendered_pinned, a.r_ready_event = slot["ptr"], slot["cuda_event"]
            rc = lib.gvd_raster_forward(C.byref(a), stream)
            if rc != 0:
                st["free"].append(slot)
            else:
                pending = PendingR(slot, cap, vcap, st)
                if force_defer:
                    done = True
                elif DEFER and defer:
                    import weakref
                    st["open"].append(weakref.ref(pending))
                    done = True
                else:
                    try:
                        counts = pending.resolve()  # waits for the second kernel of the frame, not for the ones behind it
                        a.num_rendered, a.num_visible = int(counts), counts.visible
                        done = True
                    except SpeculationOverflow:
                        # the guess was too small: redo this frame on the exact path below (its outputs were clamped,
                        # in bounds, and are simply overwritten)
                        a.spec_binning_buffer, a.spec_binning_bytes, a.spec_hist_buffer, a.spec_hist_bytes = None, 0, None, 0
                        binning = None
                    pending = None
        if rc == 0 and not done:
            alloc = _Alloc(dev)
            a.binning_alloc = a.temp_alloc = alloc.cb
            slot = None
            if st is not None:
                slot = st["free"].pop() if st["free"] else _new_slot(dev)
                a.num_rendered_pinned, a.r_ready_event = slot["ptr"], slot["cuda_event"]
            rc = lib.gvd_raster_forward(C.byref(a), stream)
            a.binning_alloc = a.temp_alloc = _n.ALLOC_FN(0)
            bufs = alloc.take()  # [instance list, chunk histogram]; the histogram is forward-only
            binning = bufs[0] if bufs else None
            if slot is not None:
                st["free"].append(slot)
            if rc == 0 and st is not None:
                _note(st, int(a.num_rendered), int(a.num_visible))
    if rc != 0:
        raise RuntimeError("gvd_raster_forward failed: " + _n.last_error(lib))
    if binning is None:
        binning = torch.empty(0, dtype=torch.uint8, device=dev)
    return (pending if pending is not None else Counts(a.num_rendered, a.num_visible), out_color, out_depth, out_alpha, radii,
            geom, binning, img)


_view_cache = {}


def _grad_views(flat, P, M, has_sh, has_scales, has_colors, has_cov):
    """All gradients are views of ONE flat allocation (16-byte aligned slices), so a data-parallel caller can sum them
    across ranks with a single collective on `grad._base` without packing copies."""
    widths = (("means3D", 3), ("sh", 3 * M if has_sh else 0), ("opacity", 1), ("scales", 3 if has_scales else 0),
              ("rot", 4 if has_scales else 0), ("colors", 3 if has_colors else 0), ("cov", 6 if has_cov else 0),
              ("means2D", 3))
    sizes = [(P * w + 3) // 4 * 4 for _, w in widths]
    parts = flat.split_with_sizes(sizes)
    v = {name: part for (name, _), part in zip(widths, parts)}
    return (v["means2D"][:P * 3].view(P, 3), v["means3D"][:P * 3].view(P, 3), v["opacity"][:P].view(P, 1),
            v["colors"][:P * 3].view(P, 3) if has_colors else None,
            v["cov"][:P * 6].view(P, 6) if has_cov else None,
            v["sh"][:P * 3 * M].view(P, M, 3) if has_sh else None,
            v["scales"][:P * 3].view(P, 3) if has_scales else None,
            v["rot"][:P * 4].view(P, 4) if has_scales else None)


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                 dL_dout_depth, dL_dout_alpha, sh, degree, campos, geomBuffer, R, binningBuffer,
                                 imageBuffer, alphas, debug, confidence=None, sh_rest=None, opacity_raw=None):
    """-> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)

    Extra trailing `confidence` ([P,1] or [P]): when given, every gradient except dL_dmeans2D is
    multiplied by it inside the kernel (the reference does this in Python afterwards,
    diff_gaussian_rasterization/__init__.py:147-157)."""
    lib = _n.raster()
    dev = means3D.device
    P = means3D.size(0)
    H, W = dL_dout_color.size(1), dL_dout_color.size(2)
    M = sh.size(1) if sh is not None and sh.numel() else 0
    raw = sh_rest is not None
    if raw:
        M = 1 + sh_rest.size(1)
    has_sh = M > 0
    has_scales = scales is not None and scales.numel() > 0
    has_cov = cov3D_precomp is not None and cov3D_precomp.numel() > 0
    has_colors = colors is not None and colors.numel() > 0

    if P == 0:
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)  # noqa: E731
        return z(0, 3), z(0, 3), z(0, 1), z(0, 3), z(0, 6), z(0, M, 3), z(0, 3), z(0, 4)
    if isinstance(R, PendingR):
        R = R.resolve()  # a deferred R is validated here (SpeculationOverflow if the frame outgrew its speculative buffers)
    V, R = getattr(R, "visible", -1), int(R)

    total = sum((P * w + 3) // 4 * 4 for w in (3, 3 * M if has_sh else 0, 1, 3 if has_scales else 0,
                                                4 if has_scales else 0, 3 if has_colors else 0, 6 if has_cov else 0, 3))
    with _on_device(dev):
        ext = _grad_buffer.get(dev)
        if ext is not None and _grad_pending.get(dev):
            # a second backward before the first one's gradients were taken: never alias them
            if not _warned_alias:
                import warnings
                _warned_alias.append(1)
                warnings.warn("diff_gaussian_rasterization: a second backward reached the caller-owned gradient buffer before "
                              "gradient_views() took the first one's gradients; it gets fresh memory (no in-place exchange for it)")
            ext = None
        if ext is not None and ext.numel() >= total:
            # caller-owned storage: the same memory every step, so the eight views are built once
            key = (ext.data_ptr(), P, M, has_sh, has_scales, has_colors, has_cov)
            views = _view_cache.get(key)
            if views is None:
                _view_cache.clear()
                views = _view_cache[key] = _grad_views(ext[:total], P, M, has_sh, has_scales, has_colors, has_cov)
        else:
            ext = None
            flat = torch.empty((total,), dtype=_F32, device=dev)
            views = _grad_views(flat, P, M, has_sh, has_scales, has_colors, has_cov)
        # gradients of absent inputs are never consumed: None instead of the reference's unused zero tensors
        dL_dmeans2D, dL_dmeans3D, dL_dopacity, dL_dcolors, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations = views
        zero_ptr = ext.data_ptr() if ext is not None else flat.data_ptr()
        if ext is not None:
            _grad_pending[dev] = True
            _last_views[dev] = {"means2D": dL_dmeans2D, "means3D": dL_dmeans3D, "opacities": dL_dopacity,
                                "colors_precomp": dL_dcolors, "cov3D_precomp": dL_dcov3D, "shs": dL_dsh,
                                "scales": dL_dscales, "rotations": dL_drotations, "_floats": total}
        scratch = torch.empty((int(lib.gvd_raster_backward_scratch_bytes(P)),), dtype=torch.uint8, device=dev)

        background = _f32c(background, "background")
        means3D = _f32c(means3D, "means3D")
        colors = _f32c(colors, "colors_precomp")
        scales = _f32c(scales, "scales")
        rotations = _f32c(rotations, "rotations")
        cov3D_precomp = _f32c(cov3D_precomp, "cov3D_precomp")
        viewmatrix = _f32c(viewmatrix, "viewmatrix")
        projmatrix = _f32c(projmatrix, "projmatrix")
        sh = _f32c(sh, "sh")
        sh_rest = _f32c(sh_rest, "sh_rest")
        opacity_raw = _f32c(opacity_raw, "opacity_raw")
        campos = _f32c(campos, "campos")
        dL_dout_color = _f32c(dL_dout_color, "dL_dout_color")
        dL_dout_depth = _f32c(dL_dout_depth, "dL_dout_depth")
        dL_dout_alpha = _f32c(dL_dout_alpha, "dL_dout_alpha")
        alphas = _f32c(alphas, "alphas")
        confidence = _f32c(confidence, "confidence")
        radii = radii.contiguous()

        a = _n.RasterBackwardArgs()
        a.P, a.D, a.M, a.R, a.width, a.height = P, int(degree), M, R, W, H
        a.num_visible = int(V)
        a.zero_region, a.zero_region_bytes = zero_ptr, 4 * total  # all eight gradients are views of this one region
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
        if raw:
            # the [P, M, 3] gradient region holds d/d_features_dc [P,1,3] followed by d/d_features_rest [P,M-1,3]
            flat_sh = dL_dsh.view(-1)
            dL_dsh, dL_dsh_rest = flat_sh[:P * 3].view(P, 1, 3), flat_sh[P * 3:].view(P, M - 1, 3)
            a.raw_params, a.shs_rest, a.opacities = 1, _ptr(sh_rest), _ptr(opacity_raw)
            a.dL_dsh, a.dL_dsh_rest = dL_dsh.data_ptr(), dL_dsh_rest.data_ptr()
        rc = lib.gvd_raster_backward(C.byref(a), _stream(dev))
    if rc != 0:
        raise RuntimeError("gvd_raster_backward failed: " + _n.last_error(lib))
    if raw:
        return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, (dL_dsh, dL_dsh_rest), dL_dscales, dL_drotations
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
