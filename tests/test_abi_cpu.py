"""CPU: the C-ABI library loads and exports every symbol include/gvd_raster.h declares (no compute)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"GVD_API\s+[\w\s\*]+?\b(gvd_\w+)\s*\(", txt)))


def test_header_symbols_exported():
    import gvd_native

    lib = gvd_native.raster()
    declared = _declared_symbols("gvd_raster.h") + _declared_symbols("gvd_exchange.h")
    assert len(declared) >= 17
    for s in declared:
        assert hasattr(lib, s), f"libgvd_raster.so does not export {s}"
    assert set(declared) == set(gvd_native.RASTER_SYMBOLS)
    assert lib.gvd_raster_abi_version() == gvd_native.ABI_VERSION


def test_sizes_and_layout_monotone():
    import gvd_native

    lib = gvd_native.raster()
    assert lib.gvd_raster_geom_bytes(0, 64, 64) >= 128
    g1, g2 = lib.gvd_raster_geom_bytes(1000, 640, 480), lib.gvd_raster_geom_bytes(2000, 640, 480)
    assert g2 > g1 >= 1000 * 64
    b1 = lib.gvd_raster_binning_bytes(10000, 0)
    assert b1 >= 10000 * 4 and lib.gvd_raster_binning_bytes(10000, 1) >= b1 + 10000 * 8
    assert lib.gvd_raster_img_bytes(640, 480) >= 640 * 480 * 4 + 1200 * 8
    assert lib.gvd_raster_backward_scratch_bytes(1000) >= 1000 * 48
    L = gvd_native.RasterLayout()
    assert lib.gvd_raster_layout(1000, 5000, 640, 480, C.byref(L)) == 0
    offs = [getattr(L, n) for n, _ in L._fields_]
    assert all(o % 128 == 0 for o in offs)


def test_struct_sizes_match_header():
    """ctypes mirrors must have the C struct sizes (checked by compiling a sizeof probe)."""
    import subprocess
    import tempfile

    import gvd_native

    src = '#include "gvd_raster.h"\n#include "gvd_exchange.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu\\n", ' \
          'sizeof(GvdRasterForwardArgs), sizeof(GvdRasterBackwardArgs), sizeof(GvdRasterLayout), sizeof(GvdRasterStageTimes), ' \
          'sizeof(GvdExchangeArgs));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(gvd_native.RasterForwardArgs), C.sizeof(gvd_native.RasterBackwardArgs),
                     C.sizeof(gvd_native.RasterLayout), C.sizeof(gvd_native.RasterStageTimes), C.sizeof(gvd_native.ExchangeArgs)]


def test_exchange_rejects_bad_arguments():
    """No GPU needed: argument validation of the peer-memory gradient sum happens before any CUDA call."""
    import gvd_native

    lib = gvd_native.raster()
    assert lib.gvd_exchange_allreduce_sum(None, None) != 0
    a = gvd_native.ExchangeArgs()
    a.world, a.rank = 1, 0
    assert lib.gvd_exchange_allreduce_sum(C.byref(a), None) != 0 and b"world" in lib.gvd_last_error()
    a.world, a.rank, a.payload_bytes, a.n_floats = 2, 0, 64, 6
    assert lib.gvd_exchange_allreduce_sum(C.byref(a), None) != 0 and b"multiple of 4" in lib.gvd_last_error()
    a.n_floats = 16
    assert lib.gvd_exchange_allreduce_sum(C.byref(a), None) != 0 and b"null buffer" in lib.gvd_last_error()


def test_null_args_fail_cleanly():
    import gvd_native

    lib = gvd_native.raster()
    assert lib.gvd_raster_forward(None, None) != 0
    assert b"null" in lib.gvd_last_error()
    a = gvd_native.RasterForwardArgs()
    a.P = 0
    assert lib.gvd_raster_forward(C.byref(a), None) == 0  # empty scene is a no-op, like the reference
    a.P, a.width, a.height = 10, 64, 64
    assert lib.gvd_raster_forward(C.byref(a), None) != 0  # neither SHs nor colours
    assert b"SHs or precomputed colors" in lib.gvd_last_error()


def test_shim_argument_contract():
    """Same exceptions as the reference front-end (DGR/diff_gaussian_rasterization/__init__.py:196-200)."""
    import torch

    import diff_gaussian_rasterization as d

    s = d.GaussianRasterizationSettings(8, 8, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                        torch.zeros(3), False, False, torch.ones(4, 1))
    assert s._fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix",
                         "projmatrix", "sh_degree", "campos", "prefiltered", "debug", "confidence")
    r = d.GaussianRasterizer(s)
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(m, m, torch.ones(4, 1), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(m, m, torch.ones(4, 1), shs=torch.zeros(4, 16, 3), colors_precomp=torch.zeros(4, 3), scales=m, rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.ones(4, 1), shs=torch.zeros(4, 16, 3), scales=m)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.ones(4, 1), shs=torch.zeros(4, 16, 3), scales=m, rotations=torch.ones(4, 4), cov3D_precomp=torch.zeros(4, 6))


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "guidedvd-3dgs_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "raster_oracle" not in txt and "oracle/" not in txt.replace("oracle/_ref", ""), os.path.join(dp, f)


def test_nn_header_symbols_exported():
    import gvd_native

    lib = gvd_native.nn()
    declared = _declared_symbols_api("gvd_nn.h", "GVD_NN_API")
    assert set(declared) == set(gvd_native.NN_SYMBOLS), (sorted(declared), sorted(gvd_native.NN_SYMBOLS))
    for s in declared:
        assert hasattr(lib, s), s
    # argument validation happens before any CUDA call
    assert lib.gvd_gemm_bf16(None, None) != 0 and b"null" in lib.gvd_nn_last_error()
    assert lib.gvd_ddim_step(None, None) != 0


def test_nn_struct_sizes_and_late_entry_points_validate():
    """ctypes mirrors of the include/gvd_nn.h argument structs have the C sizes, and the entry points added with the fused
    attention adjoint / GroupNorm / upsample route reject bad arguments before any CUDA call."""
    import subprocess
    import tempfile

    import gvd_native

    src = '#include "gvd_nn.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(GvdGemmArgs), sizeof(GvdConvArgs), ' \
          'sizeof(GvdFlashBwdArgs), sizeof(GvdDdimArgs));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(gvd_native.GemmArgs), C.sizeof(gvd_native.ConvArgs), C.sizeof(gvd_native.FlashBwdArgs),
                     C.sizeof(gvd_native.DdimArgs)]
    lib = gvd_native.nn()
    assert lib.gvd_flash_attention_bwd(None, None) == 2 and b"null" in lib.gvd_nn_last_error()
    a = gvd_native.FlashBwdArgs()
    assert lib.gvd_flash_attention_bwd(C.byref(a), None) == 2            # null tensors
    buf = (C.c_char * 4096)()
    ptr = C.addressof(buf) // 16 * 16 + 16
    for f in ("q", "k", "v", "out", "dout", "lse", "delta", "dq", "dk"):
        setattr(a, f, ptr)
    a.B, a.Nq, a.Nk, a.H, a.q_batch_stride, a.kv_batch_stride = 1, 8, 8, 1, 512, 512
    assert lib.gvd_flash_attention_bwd(C.byref(a), None) == 2 and b"go together" in lib.gvd_nn_last_error()   # dk without dv
    a.dv, a.Nk = ptr, 0
    assert lib.gvd_flash_attention_bwd(C.byref(a), None) == 2 and b"Nk" in lib.gvd_nn_last_error()
    a.Nk, a.q_batch_stride = 8, 513
    assert lib.gvd_flash_attention_bwd(C.byref(a), None) == 2 and b"strides" in lib.gvd_nn_last_error()
    a.q_batch_stride, a.B = 512, 0
    assert lib.gvd_flash_attention_bwd(C.byref(a), None) == 0            # empty batch: nothing to do
    assert lib.gvd_flash_attention_lse(ptr, ptr, ptr, ptr, None, 1, 8, 8, 1, 512, 512, 0.125, None) == 2
    assert lib.gvd_upsample2x_cl(ptr, ptr, 1, 2, 2, 12, None) == 2 and lib.gvd_upsample2x_cl(None, ptr, 1, 2, 2, 8, None) == 2
    assert lib.gvd_upsample2x_bwd_cl(ptr, ptr, 1, 2, 2, 12, None) == 2 and lib.gvd_upsample2x_cl(ptr, ptr, 0, 2, 2, 8, None) == 0
    assert lib.gvd_groupnorm_cl_keep_stats(ptr, ptr, ptr, ptr, None, 1, 8, 64, 32, 1e-5, 0, ptr, 1024, None) == 2


def _declared_symbols_api(header, macro):
    txt = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(macro + r"\s+[\w\s\*]+?\b(gvd_\w+)\s*\(", txt)))


def test_nn_fast_default_level():
    """csrc/nn_fast.cu: level 1 (GEGLU + im2col variants, measured faster on B200) is the library's default, the
    temporal-attention variant (level 2, measured slower) is not.  (Query only: no CUDA call.)"""
    import gvd_native

    if os.environ.get("GVD_NN_FAST") not in (None, "1"):
        return
    lib = gvd_native.nn()
    assert lib.gvd_nn_set_fast(-1) == 1
    assert lib.gvd_nn_set_fast(2) == 1 and lib.gvd_nn_set_fast(0) == 2 and lib.gvd_nn_set_fast(1) == 0 and lib.gvd_nn_set_fast(7) == 1
