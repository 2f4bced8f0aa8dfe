"""ctypes binding of the C-ABI libraries in ./lib (built from ./csrc for sm_100a).

This is the only place the product loads native code.  There is NO fallback: if the
library is missing the import raises, and every op raises RuntimeError with
gvd_last_error() when a call fails.  Signatures mirror include/gvd_raster.h.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_DIR = os.path.join(_HERE, "lib")

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)


class RasterForwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("width", C.c_int), ("height", C.c_int),
        ("background", C.c_void_p), ("means3D", C.c_void_p), ("shs", C.c_void_p),
        ("colors_precomp", C.c_void_p), ("opacities", C.c_void_p), ("scales", C.c_void_p),
        ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p), ("viewmatrix", C.c_void_p),
        ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
        ("scale_modifier", C.c_float), ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
        ("prefiltered", C.c_int), ("debug", C.c_int), ("export_keys", C.c_int),
        ("out_color", C.c_void_p), ("out_depth", C.c_void_p), ("out_alpha", C.c_void_p), ("radii", C.c_void_p),
        ("geom_alloc", ALLOC_FN), ("binning_alloc", ALLOC_FN), ("img_alloc", ALLOC_FN), ("temp_alloc", ALLOC_FN),
        ("alloc_user", C.c_void_p),
        ("geom_buffer", C.c_void_p), ("geom_bytes", C.c_size_t), ("img_buffer", C.c_void_p), ("img_bytes", C.c_size_t),
        ("sort_buffer", C.c_void_p), ("sort_bytes", C.c_size_t),
        ("spec_binning_buffer", C.c_void_p), ("spec_binning_bytes", C.c_size_t),
        ("spec_hist_buffer", C.c_void_p), ("spec_hist_bytes", C.c_size_t),
        ("num_rendered_pinned", C.c_void_p), ("r_ready_event", C.c_void_p),
        ("num_rendered", C.c_int), ("num_visible", C.c_int),
        ("raw_params", C.c_int), ("shs_rest", C.c_void_p),
    ]


class RasterBackwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("R", C.c_int), ("num_visible", C.c_int), ("width", C.c_int), ("height", C.c_int),
        ("background", C.c_void_p), ("means3D", C.c_void_p), ("shs", C.c_void_p),
        ("colors_precomp", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p),
        ("cov3D_precomp", C.c_void_p), ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p),
        ("campos", C.c_void_p),
        ("scale_modifier", C.c_float), ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
        ("radii", C.c_void_p), ("alphas", C.c_void_p),
        ("geom_buffer", C.c_void_p), ("binning_buffer", C.c_void_p), ("img_buffer", C.c_void_p),
        ("dL_dpix", C.c_void_p), ("dL_ddepth_pix", C.c_void_p), ("dL_dalpha_pix", C.c_void_p),
        ("confidence", C.c_void_p),
        ("scratch", C.c_void_p),
        ("zero_region", C.c_void_p), ("zero_region_bytes", C.c_size_t),
        ("dL_dmeans2D", C.c_void_p), ("dL_dmeans3D", C.c_void_p), ("dL_dopacity", C.c_void_p),
        ("dL_dcolors", C.c_void_p), ("dL_dcov3D", C.c_void_p), ("dL_dsh", C.c_void_p),
        ("dL_dscales", C.c_void_p), ("dL_drotations", C.c_void_p),
        ("debug", C.c_int),
        ("raw_params", C.c_int), ("shs_rest", C.c_void_p), ("opacities", C.c_void_p), ("dL_dsh_rest", C.c_void_p),
    ]


class RasterLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "geom_splat", "geom_clamped", "geom_tiles_touched", "geom_visible_ids", "geom_counts",
        "bin_point_list", "bin_point_list_keys",
        "img_ranges", "img_n_contrib")]


STAGE_NAMES = ("preprocess", "bin_count", "bin_fill", "depth_sort", "export_keys", "render_fwd", "render_bwd", "gaussian_bwd")


class RasterStageTimes(C.Structure):
    _fields_ = [("ms", C.c_double * len(STAGE_NAMES)), ("calls", C.c_int * len(STAGE_NAMES))]


RASTER_SYMBOLS = (
    "gvd_raster_profile_enable", "gvd_raster_profile_read",
    "gvd_raster_abi_version", "gvd_last_error", "gvd_raster_geom_bytes", "gvd_raster_binning_bytes",
    "gvd_raster_img_bytes", "gvd_raster_sort_bytes", "gvd_raster_hist_bytes", "gvd_raster_backward_scratch_bytes", "gvd_raster_layout",
    "gvd_raster_forward", "gvd_raster_backward", "gvd_raster_mark_visible",
    "gvd_exchange_alloc", "gvd_exchange_free", "gvd_exchange_open", "gvd_exchange_close", "gvd_exchange_allreduce_sum", "gvd_exchange_status",
)

EXCHANGE_MAX_RANKS = 8
EXCHANGE_HANDLE_BYTES = 64
EXCHANGE_FLAG_BYTES = 256


class ExchangeArgs(C.Structure):
    """include/gvd_exchange.h::GvdExchangeArgs"""
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("bufs", C.c_void_p * EXCHANGE_MAX_RANKS),
                ("payload_bytes", C.c_size_t), ("n_floats", C.c_size_t), ("epoch", C.c_uint32), ("multicast", C.c_void_p)]

_raster = None
ABI_VERSION = 9


def lib_path(name="libgvd_raster.so"):
    return os.path.join(_LIB_DIR, name)


def raster():
    """Load libgvd_raster.so once; raise loudly when it is absent (no CPU/torch fallback exists)."""
    global _raster
    if _raster is not None:
        return _raster
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C guidedvd-3dgs_b200/csrc`). The rasterizer has no fallback path.")
    lib = C.CDLL(path)
    lib.gvd_last_error.restype = C.c_char_p
    lib.gvd_raster_abi_version.restype = C.c_int
    lib.gvd_raster_geom_bytes.restype = C.c_size_t
    lib.gvd_raster_geom_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.gvd_raster_binning_bytes.restype = C.c_size_t
    lib.gvd_raster_binning_bytes.argtypes = [C.c_int, C.c_int]
    lib.gvd_raster_backward_scratch_bytes.restype = C.c_size_t
    lib.gvd_raster_backward_scratch_bytes.argtypes = [C.c_int]
    lib.gvd_raster_img_bytes.restype = C.c_size_t
    lib.gvd_raster_img_bytes.argtypes = [C.c_int, C.c_int]
    lib.gvd_raster_sort_bytes.restype = C.c_size_t
    lib.gvd_raster_sort_bytes.argtypes = [C.c_int]
    lib.gvd_raster_hist_bytes.restype = C.c_size_t
    lib.gvd_raster_hist_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.gvd_raster_layout.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(RasterLayout)]
    lib.gvd_raster_forward.argtypes = [C.POINTER(RasterForwardArgs), C.c_void_p]
    lib.gvd_raster_backward.argtypes = [C.POINTER(RasterBackwardArgs), C.c_void_p]
    lib.gvd_raster_mark_visible.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    for n in ("gvd_raster_layout", "gvd_raster_forward", "gvd_raster_backward", "gvd_raster_mark_visible"):
        getattr(lib, n).restype = C.c_int
    lib.gvd_exchange_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]
    lib.gvd_exchange_free.argtypes = [C.c_void_p]
    lib.gvd_exchange_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.gvd_exchange_close.argtypes = [C.c_void_p]
    lib.gvd_exchange_allreduce_sum.argtypes = [C.POINTER(ExchangeArgs), C.c_void_p]
    lib.gvd_exchange_status.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)]
    for n in ("gvd_exchange_alloc", "gvd_exchange_free", "gvd_exchange_open", "gvd_exchange_close", "gvd_exchange_allreduce_sum", "gvd_exchange_status"):
        getattr(lib, n).restype = C.c_int
    lib.gvd_raster_profile_enable.argtypes = [C.c_int]
    lib.gvd_raster_profile_read.argtypes = [C.POINTER(RasterStageTimes)]
    if lib.gvd_raster_abi_version() != ABI_VERSION:
        raise RuntimeError("libgvd_raster.so ABI version mismatch; rebuild")
    _raster = lib
    return lib


def last_error(lib):
    return (lib.gvd_last_error() or b"").decode("utf-8", "replace")


# ---- libgvd_knn.so (include/gvd_knn.h) -----------------------------------------------------------------
KNN_SYMBOLS = ("gvd_knn3_tmp_bytes", "gvd_knn3", "gvd_knn_last_error")
_knn = None


def knn():
    """Load libgvd_knn.so once; raise loudly when it is absent (there is no fallback)."""
    global _knn
    if _knn is not None:
        return _knn
    path = lib_path("libgvd_knn.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with `make -C guidedvd-3dgs_b200/csrc`. No fallback path exists.")
    lib = C.CDLL(path)
    lib.gvd_knn_last_error.restype = C.c_char_p
    lib.gvd_knn3_tmp_bytes.restype = C.c_size_t
    lib.gvd_knn3_tmp_bytes.argtypes = [C.c_int]
    lib.gvd_knn3.restype = C.c_int
    lib.gvd_knn3.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    _knn = lib
    return lib


# ---- libgvd_points.so (include/gvd_points.h) ------------------------------------------------------------
POINTS_SYMBOLS = ("gvd_point_project_scratch_bytes", "gvd_point_project", "gvd_points_last_error")
_points = None


def points():
    """Load libgvd_points.so once; raise loudly when it is absent (there is no fallback)."""
    global _points
    if _points is not None:
        return _points
    path = lib_path("libgvd_points.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with `make -C guidedvd-3dgs_b200/csrc`. No fallback path exists.")
    lib = C.CDLL(path)
    lib.gvd_points_last_error.restype = C.c_char_p
    lib.gvd_point_project_scratch_bytes.restype = C.c_size_t
    lib.gvd_point_project_scratch_bytes.argtypes = [C.c_int, C.c_int]
    lib.gvd_point_project.restype = C.c_int
    lib.gvd_point_project.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double,
                                      C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    _points = lib
    return lib


# ---- libgvd_train.so (include/gvd_train.h) --------------------------------------------------------------
TRAIN_SYMBOLS = ("gvd_photometric_loss_scratch_bytes", "gvd_photometric_loss_forward", "gvd_photometric_loss_backward",
                 "gvd_densification_stats", "gvd_adam_step", "gvd_mask_morphology", "gvd_train_last_error")
_train = None


def bind_train(lib):
    """argtypes / restypes of include/gvd_train.h (shared with the host build the CPU tests execute)."""
    vp, ll, i32, f32 = C.c_void_p, C.c_longlong, C.c_int, C.c_float
    lib.gvd_train_last_error.restype = C.c_char_p
    lib.gvd_photometric_loss_scratch_bytes.restype = C.c_size_t
    lib.gvd_photometric_loss_scratch_bytes.argtypes = [i32, i32, i32]
    lib.gvd_photometric_loss_forward.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, C.c_size_t, vp]
    lib.gvd_photometric_loss_backward.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp]
    lib.gvd_densification_stats.argtypes = [vp, vp, ll, vp, vp, vp, vp]
    lib.gvd_adam_step.argtypes = [vp, vp, vp, vp, ll, C.c_double, C.c_double, C.c_double, C.c_double, i32, vp]
    lib.gvd_mask_morphology.argtypes = [vp, vp, ll, i32, i32, i32, i32, i32, i32, i32, vp]
    for n in ("gvd_photometric_loss_forward", "gvd_photometric_loss_backward", "gvd_densification_stats", "gvd_adam_step", "gvd_mask_morphology"):
        getattr(lib, n).restype = C.c_int
    return lib


def train():
    """Load libgvd_train.so once; raise loudly when it is absent (there is no fallback)."""
    global _train
    if _train is not None:
        return _train
    path = lib_path("libgvd_train.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with `make -C guidedvd-3dgs_b200/csrc`. No fallback path exists.")
    _train = bind_train(C.CDLL(path))
    return _train


# ---- libgvd_nn.so (include/gvd_nn.h) --------------------------------------------------------------------
class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("batch_h", C.c_int), ("batch_b", C.c_int),
        ("A", C.c_void_p), ("lda", C.c_longlong), ("a_stride_h", C.c_longlong), ("a_stride_b", C.c_longlong),
        ("B", C.c_void_p), ("ldb", C.c_longlong), ("b_stride_h", C.c_longlong), ("b_stride_b", C.c_longlong),
        ("C", C.c_void_p), ("ldc", C.c_longlong), ("c_stride_h", C.c_longlong), ("c_stride_b", C.c_longlong),
        ("bias", C.c_void_p), ("bias2", C.c_void_p), ("residual", C.c_void_p), ("alpha", C.c_float), ("act", C.c_int),
        ("out_fp32", C.c_int), ("b_mn_major", C.c_int),
    ]


class ConvArgs(C.Structure):
    """include/gvd_nn.h::GvdConvArgs"""
    _fields_ = [
        ("kind", C.c_int), ("F", C.c_int), ("H", C.c_int), ("W", C.c_int), ("B", C.c_int), ("T", C.c_int), ("S", C.c_longlong),
        ("Cin", C.c_int), ("Cout", C.c_int), ("x", C.c_void_p), ("weight", C.c_void_p), ("y", C.c_void_p),
        ("bias", C.c_void_p), ("bias2", C.c_void_p), ("residual", C.c_void_p), ("act", C.c_int),
    ]


class FlashBwdArgs(C.Structure):
    """include/gvd_nn.h::GvdFlashBwdArgs"""
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p), ("dout", C.c_void_p),
        ("lse", C.c_void_p), ("delta", C.c_void_p), ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
        ("B", C.c_int), ("Nq", C.c_int), ("Nk", C.c_int), ("H", C.c_int),
        ("q_batch_stride", C.c_longlong), ("kv_batch_stride", C.c_longlong), ("scale", C.c_float),
    ]


class DdimArgs(C.Structure):
    _fields_ = [
        ("n", C.c_longlong), ("x", C.c_void_p), ("e_cond", C.c_void_p), ("e_uncond", C.c_void_p), ("noise", C.c_void_p),
        ("x_prev", C.c_void_p), ("pred_x0", C.c_void_p), ("scratch", C.c_void_p),
        ("cfg_scale", C.c_float), ("guidance_rescale", C.c_float),
        ("sqrt_alphas_cumprod_t", C.c_float), ("sqrt_one_minus_alphas_cumprod_t", C.c_float),
        ("ddim_alpha_prev", C.c_float), ("ddim_sigma", C.c_float), ("temperature", C.c_float),
        ("scale_t", C.c_float), ("scale_prev", C.c_float), ("use_dynamic_rescale", C.c_int),
    ]


class DdimVjpArgs(C.Structure):
    _fields_ = [
        ("n", C.c_longlong), ("e_cond", C.c_void_p), ("e_uncond", C.c_void_p), ("grad_pred_x0", C.c_void_p),
        ("dx", C.c_void_p), ("de_cond", C.c_void_p), ("de_uncond", C.c_void_p), ("scratch", C.c_void_p),
        ("scratch_bytes", C.c_size_t),
        ("cfg_scale", C.c_float), ("guidance_rescale", C.c_float),
        ("sqrt_alphas_cumprod_t", C.c_float), ("sqrt_one_minus_alphas_cumprod_t", C.c_float),
        ("scale_t", C.c_float), ("scale_prev", C.c_float), ("use_dynamic_rescale", C.c_int),
    ]


NN_SYMBOLS = ("gvd_gemm_bf16", "gvd_nn_last_error", "gvd_groupnorm_tmp_floats", "gvd_groupnorm_cl", "gvd_layernorm",
              "gvd_geglu", "gvd_softmax_rows", "gvd_im2col3x3_cl", "gvd_im2col_t3_cl", "gvd_temporal_attention",
              "gvd_ddim_step", "gvd_flash_attention", "gvd_groupnorm_cl_stats", "gvd_groupnorm_cl_apply",
              # input-gradient operators of the guided sampler (csrc/nn_backward.cu)
              "gvd_groupnorm_bwd_tmp_bytes", "gvd_groupnorm_cl_bwd", "gvd_groupnorm_cl_bwd_sums", "gvd_groupnorm_cl_bwd_apply", "gvd_layernorm_bwd", "gvd_geglu_bwd", "gvd_softmax_bwd_rows",
              "gvd_col2im3x3_cl", "gvd_col2im_t3_cl", "gvd_temporal_attention_bwd", "gvd_ddim_pred_x0_vjp",
              "gvd_im2col3x3_down_cl", "gvd_nn_set_fast", "gvd_conv_bf16", "gvd_conv_bf16_supported",
              "gvd_flash_attention_lse", "gvd_flash_attention_bwd", "gvd_groupnorm_cl_keep_stats",
              "gvd_upsample2x_cl", "gvd_upsample2x_bwd_cl")
_nn = None


def _nn_signatures():
    vp, ll, i32, f32, sz = C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_size_t
    I, S = C.c_int, C.c_size_t  # return types
    return {  # name: (restype, argtypes) -- include/gvd_nn.h
        "gvd_nn_last_error": (C.c_char_p, []),
        "gvd_nn_set_fast": (I, [i32]),
        "gvd_gemm_bf16": (I, [C.POINTER(GemmArgs), vp]),
        "gvd_conv_bf16": (I, [C.POINTER(ConvArgs), vp]),
        "gvd_conv_bf16_supported": (I, [i32, i32, i32, i32, i32]),
        "gvd_groupnorm_tmp_floats": (S, [i32, ll, i32]),
        "gvd_groupnorm_cl": (I, [vp, vp, vp, vp, i32, ll, i32, i32, f32, i32, vp, sz, vp]),
        "gvd_groupnorm_cl_keep_stats": (I, [vp, vp, vp, vp, vp, i32, ll, i32, i32, f32, i32, vp, sz, vp]),
        "gvd_groupnorm_cl_stats": (I, [vp, vp, i32, ll, i32, i32, vp, sz, vp]),
        "gvd_groupnorm_cl_apply": (I, [vp, vp, vp, vp, vp, i32, ll, ll, i32, i32, f32, i32, vp]),
        "gvd_layernorm": (I, [vp, vp, vp, vp, ll, i32, f32, vp]),
        "gvd_geglu": (I, [vp, vp, ll, i32, vp]),
        "gvd_softmax_rows": (I, [vp, i32, ll, vp, ll, ll, i32, vp]),
        "gvd_im2col3x3_cl": (I, [vp, vp, i32, i32, i32, i32, i32, i32, vp]),
        "gvd_im2col3x3_down_cl": (I, [vp, vp, i32, i32, i32, i32, vp]),
        "gvd_im2col_t3_cl": (I, [vp, vp, i32, i32, ll, i32, vp]),
        "gvd_upsample2x_cl": (I, [vp, vp, i32, i32, i32, i32, vp]),
        "gvd_upsample2x_bwd_cl": (I, [vp, vp, i32, i32, i32, i32, vp]),
        "gvd_temporal_attention": (I, [vp, vp, vp, vp, i32, i32, ll, i32, f32, vp]),
        "gvd_flash_attention": (I, [vp, vp, vp, vp, i32, i32, i32, i32, ll, ll, f32, vp]),
        "gvd_flash_attention_lse": (I, [vp, vp, vp, vp, vp, i32, i32, i32, i32, ll, ll, f32, vp]),
        "gvd_flash_attention_bwd": (I, [C.POINTER(FlashBwdArgs), vp]),
        "gvd_ddim_step": (I, [C.POINTER(DdimArgs), vp]),
        "gvd_groupnorm_bwd_tmp_bytes": (S, [i32, ll, i32]),
        "gvd_groupnorm_cl_bwd": (I, [vp, vp, vp, vp, vp, vp, i32, ll, i32, i32, f32, i32, vp, sz, vp]),
        "gvd_groupnorm_cl_bwd_sums": (I, [vp, vp, vp, vp, vp, vp, i32, ll, ll, i32, i32, f32, i32, vp, sz, vp]),
        "gvd_groupnorm_cl_bwd_apply": (I, [vp, vp, vp, vp, vp, vp, vp, i32, ll, ll, i32, i32, f32, i32, vp]),
        "gvd_layernorm_bwd": (I, [vp, vp, vp, vp, ll, i32, f32, vp]),
        "gvd_geglu_bwd": (I, [vp, vp, vp, ll, i32, vp]),
        "gvd_softmax_bwd_rows": (I, [vp, vp, vp, ll, ll, i32, vp]),
        "gvd_col2im3x3_cl": (I, [vp, vp, i32, i32, i32, i32, i32, i32, vp]),
        "gvd_col2im_t3_cl": (I, [vp, vp, i32, i32, ll, i32, vp]),
        "gvd_temporal_attention_bwd": (I, [vp, vp, vp, vp, vp, vp, vp, i32, i32, ll, i32, f32, vp]),
        "gvd_ddim_pred_x0_vjp": (I, [C.POINTER(DdimVjpArgs), vp]),
    }


def bind_nn(lib, partial=False):
    """restype / argtypes of include/gvd_nn.h on a loaded library.  partial=True: bind whatever subset the library exports
    (the host builds of single source files that the CPU tests execute)."""
    for name, (res, args) in _nn_signatures().items():
        if partial and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def nn():
    """Load libgvd_nn.so once; raise loudly when it is absent (there is no fallback)."""
    global _nn
    if _nn is not None:
        return _nn
    path = lib_path("libgvd_nn.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with `make -C guidedvd-3dgs_b200/csrc`. No fallback path exists.")
    _nn = bind_nn(C.CDLL(path))
    return _nn
