"""Drives the HOST build of the rasterizer's CUDA sources (tests/cuda_emu/build_emu.py raster: raster_api.cu,
raster_forward.cu, raster_backward.cu compiled as they are, CUDA threads as OS threads) through the C ABI of
include/gvd_raster.h with numpy buffers -- the exact (synchronous, callback-allocating) path of the library.
Test infrastructure.  Returns the same dictionary as oracle/raster_oracle.py::run."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "guidedvd-3dgs_b200"), os.path.join(ROOT, "tests", "cuda_emu")):
    if p not in sys.path:
        sys.path.insert(0, p)

_lib = None


def lib():
    global _lib
    if _lib is None:
        import build_emu
        import gvd_native as n

        L = C.CDLL(build_emu.build("raster", build_emu.RASTER_SOURCES))
        L.gvd_last_error.restype = C.c_char_p
        for name in ("gvd_raster_geom_bytes", "gvd_raster_binning_bytes", "gvd_raster_img_bytes", "gvd_raster_backward_scratch_bytes",
                     "gvd_raster_sort_bytes", "gvd_raster_hist_bytes"):
            getattr(L, name).restype = C.c_size_t
        L.gvd_raster_sort_bytes.argtypes = [C.c_int]
        L.gvd_raster_hist_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
        L.gvd_raster_geom_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
        L.gvd_raster_binning_bytes.argtypes = [C.c_int, C.c_int]
        L.gvd_raster_img_bytes.argtypes = [C.c_int, C.c_int]
        L.gvd_raster_backward_scratch_bytes.argtypes = [C.c_int]
        L.gvd_raster_forward.argtypes = [C.POINTER(n.RasterForwardArgs), C.c_void_p]
        L.gvd_raster_backward.argtypes = [C.POINTER(n.RasterBackwardArgs), C.c_void_p]
        L.gvd_raster_layout.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(n.RasterLayout)]
        _lib = L
    return _lib


def _aligned(nbytes, align=128):
    raw = np.zeros(nbytes + align + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw, raw.ctypes.data + off


def _p(a):
    return None if a is None else a.ctypes.data


def run(sc, cam, bg, D, cot=None, use_conf=False, precomp=None, spec_capacity=None, spec_visible=None, pinned=False, flat_grads=False,
        know_visible=True, raw=None):
    """spec_capacity: run the SPECULATIVE forward (no host round trip for R) with an instance buffer of that many entries
    (spec_visible: and a chunk-histogram buffer for that many visible Gaussians, default P).
    pinned: hand the exact path a host word pair for {R, V} (its event-wait route instead of the memcpy fallback).
    flat_grads: the gradient outputs are views of one allocation and the backward is told so (zero_region).
    know_visible=False: the backward is not told V (num_visible = -1).
    raw: dict(scaling, rotation, opacity, features_dc, features_rest) of UN-activated GaussianModel parameters: the
    raw_params mode of the ABI (activations folded into the kernels); gradients come back for the raw tensors, the SH
    gradient as grads["features_dc"] / grads["features_rest"]."""
    import gvd_native as n

    L = lib()
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)  # noqa: E731
    means3D, opac = f32(sc["means3D"]), f32(sc["opacities"]).reshape(-1)
    shs, scales, rots = f32(sc["shs"]), f32(sc["scales"]), f32(sc["rotations"])
    colors_pre = f32(precomp["colors_precomp"]) if precomp else None
    cov_pre = f32(precomp["cov3D_precomp"]) if precomp else None
    if precomp:
        shs = scales = rots = None
    sh_rest = None
    if raw is not None:
        scales, rots, opac = f32(raw["scaling"]), f32(raw["rotation"]), f32(raw["opacity"]).reshape(-1)
        shs, sh_rest = f32(raw["features_dc"]), f32(raw["features_rest"])
    P, W, H = means3D.shape[0], int(cam["width"]), int(cam["height"])
    M = 16 if shs is None else shs.shape[1]
    if raw is not None:
        M = 1 + sh_rest.shape[1]
    view, proj, campos = f32(np.asarray(cam["viewmatrix"]).reshape(-1)), f32(np.asarray(cam["projmatrix"]).reshape(-1)), f32(cam["campos"])
    bgf = f32(bg)
    color, depth, alpha = (np.zeros((c, H, W), np.float32) for c in (3, 1, 1))
    radii = np.zeros(P, np.int32)
    keep = {}

    def make_alloc(tag):
        def fn(user, nbytes):
            raw, ptr = _aligned(int(nbytes))
            keep[tag] = (raw, ptr, int(nbytes))
            return ptr
        return n.ALLOC_FN(fn)

    cbs = [make_alloc(t) for t in ("geom", "binning", "img")]
    temps = []

    def temp_fn(user, nbytes):
        raw, ptr = _aligned(int(nbytes))
        temps.append(raw)
        return ptr
    temp_cb = n.ALLOC_FN(temp_fn)
    a = n.RasterForwardArgs()
    a.P, a.D, a.M, a.width, a.height = P, D, M, W, H
    a.background, a.means3D, a.shs, a.colors_precomp = _p(bgf), _p(means3D), _p(shs), _p(colors_pre)
    a.opacities, a.scales, a.rotations, a.cov3D_precomp = _p(opac), _p(scales), _p(rots), _p(cov_pre)
    a.viewmatrix, a.projmatrix, a.campos = _p(view), _p(proj), _p(campos)
    a.scale_modifier, a.tan_fovx, a.tan_fovy = 1.0, float(cam["tanfovx"]), float(cam["tanfovy"])
    a.prefiltered, a.debug, a.export_keys = 0, 0, 1
    if raw is not None:
        a.raw_params, a.shs_rest = 1, _p(sh_rest)
    a.out_color, a.out_depth, a.out_alpha, a.radii = _p(color), _p(depth), _p(alpha), _p(radii)
    a.geom_alloc, a.binning_alloc, a.img_alloc = cbs
    a.temp_alloc = temp_cb
    r_word = np.full(2, -7, np.int32)
    if pinned:
        a.num_rendered_pinned = r_word.ctypes.data
    if spec_capacity is not None:
        nb = int(L.gvd_raster_binning_bytes(int(spec_capacity), 1)) + 512
        raw, ptr = _aligned(nb)
        keep["binning"] = (raw, ptr, nb)
        a.spec_binning_buffer, a.spec_binning_bytes = ptr, nb
        hb = int(L.gvd_raster_hist_bytes(int(P if spec_visible is None else spec_visible), W, H)) + 512
        hraw, hptr = _aligned(hb)
        keep["hist"] = (hraw, hptr, hb)
        a.spec_hist_buffer, a.spec_hist_bytes = hptr, hb
        a.num_rendered_pinned = r_word.ctypes.data
    rc = L.gvd_raster_forward(C.byref(a), None)
    if rc != 0:
        raise RuntimeError("gvd_raster_forward (host build): " + (L.gvd_last_error() or b"").decode())
    R, V = int(a.num_rendered), int(a.num_visible)
    if pinned:
        assert (R, V) == (int(r_word[0]), int(r_word[1]))
    if spec_capacity is not None:
        assert R == -1 and V == -1        # the library did not learn R on this path; it arrives through the pinned words
        R, V = int(r_word[0]), int(r_word[1])
        if R > spec_capacity or (spec_visible is not None and V > spec_visible):  # the caller's validation: this frame is invalid
            return dict(num_rendered=R, num_visible=V, overflow=True, color=color, radii=radii)
    lay = n.RasterLayout()
    L.gvd_raster_layout(P, R, W, H, C.byref(lay))
    T = ((W + 15) // 16) * ((H + 15) // 16)

    def view_of(tag, off, dtype, count):
        raw, ptr, _ = keep[tag]
        start = ptr - raw.ctypes.data + off
        return raw[start:start + count * np.dtype(dtype).itemsize].view(dtype).copy()

    out = dict(color=color, depth=depth, alpha=alpha, radii=radii, num_rendered=R, num_visible=V,
               visible_ids=view_of("geom", lay.geom_visible_ids, np.uint32, P)[:V],
               counts=view_of("geom", lay.geom_counts, np.uint32, 2),
               tiles_touched=view_of("geom", lay.geom_tiles_touched, np.uint32, P),
               ranges=view_of("img", lay.img_ranges, np.uint32, 2 * T).reshape(T, 2),
               n_contrib=view_of("img", lay.img_n_contrib, np.uint32, H * W).reshape(H, W))
    if R > 0:
        out["point_list"] = view_of("binning", lay.bin_point_list, np.uint32, R)
        out["point_list_keys"] = view_of("binning", lay.bin_point_list_keys, np.uint64, R)
    splat = view_of("geom", lay.geom_splat, np.float32, 16 * P).reshape(P, 16)
    out["g_xy"], out["g_depth"] = splat[:, 0:2], splat[:, 9]
    if cot is None:
        return out
    shapes = dict(means2D=(P, 3), means3D=(P, 3), opacities=(P,), colors_precomp=(P, 3), cov3D_precomp=(P, 6), shs=(P, M, 3),
                  scales=(P, 3), rotations=(P, 4))
    # every output starts as garbage: the backward has to produce the zeros of the invisible Gaussians itself
    if flat_grads:
        sizes = {k: (int(np.prod(s)) + 3) // 4 * 4 for k, s in shapes.items()}
        flat_raw = np.full(sum(sizes.values()) + 8, 7.5, np.float32)
        flat_off = (-flat_raw.ctypes.data // 4) % 4  # 16-byte aligned start
        g, pos = {}, flat_off
        for k, s_ in shapes.items():
            g[k] = flat_raw[pos:pos + int(np.prod(s_))].reshape(s_)
            pos += sizes[k]
    else:
        g = {k: np.full(s_, 7.5, np.float32) for k, s_ in shapes.items()}
    conf = f32(sc["confidence"]).reshape(-1) if use_conf else None
    sraw, sptr = _aligned(int(L.gvd_raster_backward_scratch_bytes(P)))
    b = n.RasterBackwardArgs()
    b.P, b.D, b.M, b.R, b.width, b.height = P, D, M, R, W, H
    b.num_visible = V if know_visible else -1
    if flat_grads:
        b.zero_region, b.zero_region_bytes = flat_raw.ctypes.data + 4 * flat_off, 4 * sum(sizes.values())
    b.background, b.means3D, b.shs, b.colors_precomp = _p(bgf), _p(means3D), _p(shs), _p(colors_pre)
    b.scales, b.rotations, b.cov3D_precomp = _p(scales), _p(rots), _p(cov_pre)
    b.viewmatrix, b.projmatrix, b.campos = _p(view), _p(proj), _p(campos)
    b.scale_modifier, b.tan_fovx, b.tan_fovy = 1.0, float(cam["tanfovx"]), float(cam["tanfovy"])
    b.radii, b.alphas = _p(radii), _p(alpha)
    b.geom_buffer, b.binning_buffer, b.img_buffer = keep["geom"][1], keep["binning"][1] if "binning" in keep else None, keep["img"][1]
    dpix, ddep, dalp = f32(cot["color"]), f32(cot["depth"]), f32(cot["alpha"])
    b.dL_dpix, b.dL_ddepth_pix, b.dL_dalpha_pix = _p(dpix), _p(ddep), _p(dalp)
    b.confidence, b.scratch = _p(conf), sptr
    b.dL_dmeans2D, b.dL_dmeans3D, b.dL_dopacity = _p(g["means2D"]), _p(g["means3D"]), _p(g["opacities"])
    b.dL_dcolors = _p(g["colors_precomp"]) if precomp else None
    b.dL_dcov3D = _p(g["cov3D_precomp"]) if precomp else None
    b.dL_dsh = None if precomp else _p(g["shs"])
    b.dL_dscales = None if precomp else _p(g["scales"])
    b.dL_drotations = None if precomp else _p(g["rotations"])
    b.debug = 0
    if raw is not None:
        g["features_dc"], g["features_rest"] = np.full((P, 1, 3), 7.5, np.float32), np.full((P, M - 1, 3), 7.5, np.float32)
        b.raw_params, b.shs_rest, b.opacities = 1, _p(sh_rest), _p(opac)
        b.dL_dsh, b.dL_dsh_rest = _p(g["features_dc"]), _p(g["features_rest"])
    rc = L.gvd_raster_backward(C.byref(b), None)
    if rc != 0:
        raise RuntimeError("gvd_raster_backward (host build): " + (L.gvd_last_error() or b"").decode())
    if precomp:
        for k in ("shs", "scales", "rotations"):
            g.pop(k)
    else:
        for k in ("colors_precomp", "cov3D_precomp"):
            g.pop(k)
    if raw is not None:
        g.pop("shs")
    out["grads"] = g
    return out
