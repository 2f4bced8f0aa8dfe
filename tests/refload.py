"""Load the compiled UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.py) under alias
module names so it can live beside the drop-in packages of the same name. Test infrastructure."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _load_pkg(alias, pkgdir):
    if alias in sys.modules:
        return sys.modules[alias]
    init = os.path.join(pkgdir, "__init__.py")
    if not os.path.exists(init) or not os.path.exists(os.path.join(pkgdir, "_C.so")):
        return None
    spec = importlib.util.spec_from_file_location(alias, init, submodule_search_locations=[pkgdir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        del sys.modules[alias]
        raise
    return mod


def ref_dgr():
    """The reference `diff_gaussian_rasterization` package (or None if it was not built)."""
    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    return _load_pkg("gvdref_dgr", os.path.join(REF, "diff_gaussian_rasterization"))


def ref_knn():
    import torch  # noqa: F401

    pkg = _load_pkg("gvdref_knn", os.path.join(REF, "simple_knn"))
    if pkg is None:
        return None
    import importlib

    return importlib.import_module("gvdref_knn._C")


def _align(x, a=128):
    return (x + a - 1) // a * a


def ref_binning_views(binning, R):
    """Views into the reference's binningBuffer (rasterizer_impl.cu:180-195 layout)."""
    import torch

    off = 0
    point_list = binning[off:off + 4 * R].view(torch.int32)
    off = _align(off + 4 * R)
    off = _align(off + 4 * R)  # point_list_unsorted
    keys = binning[off:off + 8 * R].view(torch.int64)
    return dict(point_list=point_list, point_list_keys=keys)


def ref_geom_views(geom, P):
    """Views into the reference's geomBuffer (rasterizer_impl.cu:155-170 layout)."""
    import torch

    out = {}
    off = 0
    out["depths"] = geom[off:off + 4 * P].view(torch.float32)
    off = _align(off + 4 * P)
    out["clamped"] = geom[off:off + 3 * P].view(torch.bool).view(P, 3)
    off = _align(off + 3 * P)
    off = _align(off + 4 * P)  # internal_radii
    out["means2D"] = geom[off:off + 8 * P].view(torch.float32).view(P, 2)
    off = _align(off + 8 * P)
    out["cov3D"] = geom[off:off + 24 * P].view(torch.float32).view(P, 6)
    off = _align(off + 24 * P)
    out["conic_opacity"] = geom[off:off + 16 * P].view(torch.float32).view(P, 4)
    off = _align(off + 16 * P)
    out["rgb"] = geom[off:off + 12 * P].view(torch.float32).view(P, 3)
    off = _align(off + 12 * P)
    out["tiles_touched"] = geom[off:off + 4 * P].view(torch.int32)
    return out


def ref_img_views(img, W, H):
    """rasterizer_impl.cu:172-178: n_contrib[N] then ranges[N] (N = W*H)."""
    import torch

    N = W * H
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    n_contrib = img[0:4 * N].view(torch.int32)
    off = _align(4 * N)
    ranges = img[off:off + 8 * tiles].view(torch.int32).view(tiles, 2)
    return dict(n_contrib=n_contrib, ranges=ranges)
