"""Load the reference's OWN Python above the rasterizer boundary -- `gaussian_renderer.render`
(gaussian_renderer/__init__.py:19-132), `scene.gaussian_model.GaussianModel`, `scene.cameras.PseudoCamera`,
`utils.easy_renderer.EasyRenderer`, `arguments.PipelineParams`, `utils.loss_utils` -- from oracle/_ref/gs (installed
unmodified by oracle/build_ref.py gs) with `diff_gaussian_rasterization` / `simple_knn` resolving to a chosen backend:

    gs = gs_refload.load("ours")        # the drop-in packages of guidedvd-3dgs_b200/
    gs = gs_refload.load("reference")   # the compiled reference extensions of oracle/_ref

so a test can run the same reference code over both and compare.  Stand-ins are provided only for what the image lacks
and the path never executes: `matplotlib` (colour maps for debug dumps), `plyfile` (checkpoint IO), and the `scene`
package's __init__ (dataset readers).  Test infrastructure; nothing in the product imports this.
"""
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "guidedvd-3dgs_b200")
GS = os.path.join(ROOT, "oracle", "_ref", "gs")
_TOP = ("scene", "utils", "gaussian_renderer", "arguments", "diff_gaussian_rasterization", "simple_knn", "matplotlib", "plyfile")
_loaded = {}


def available():
    return os.path.exists(os.path.join(GS, "gaussian_renderer", "__init__.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load(backend):
    """-> namespace with render, GaussianModel, PseudoCamera, EasyRenderer, PipelineParams, ModelParams, loss_utils,
    rasterizer (the diff_gaussian_rasterization package in use), distCUDA2."""
    if backend in _loaded:
        return _loaded[backend]
    assert backend in ("ours", "reference")
    if not available():
        raise RuntimeError("oracle/_ref/gs not installed (python oracle/build_ref.py gs)")
    import refload

    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in _TOP}
    saved_path = list(sys.path)
    try:
        mpl = _stub("matplotlib")
        mpl.__path__ = []
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.cm = _stub("matplotlib.cm")
        _stub("plyfile", PlyData=object, PlyElement=object)
        scene = _stub("scene", Scene=None)
        scene.__path__ = [os.path.join(GS, "scene")]
        if backend == "ours":
            for k, v in saved.items():  # reuse the already-loaded drop-in modules (one ctypes library per process)
                if k.split(".")[0] in ("diff_gaussian_rasterization", "simple_knn") and PKG in (getattr(v, "__file__", "") or ""):
                    sys.modules[k] = v
            sys.path.insert(0, PKG)
            dgr = importlib.import_module("diff_gaussian_rasterization")
            knn_c = importlib.import_module("simple_knn._C")
        else:
            dgr, knn_c = refload.ref_dgr(), refload.ref_knn()
            if dgr is None or knn_c is None:
                raise RuntimeError("oracle/_ref extensions not built")
            sys.modules["diff_gaussian_rasterization"] = dgr
            sys.modules["simple_knn"] = sys.modules["gvdref_knn"]
            sys.modules["simple_knn._C"] = knn_c
        sys.path.insert(0, GS)
        gr = importlib.import_module("gaussian_renderer")
        er = importlib.import_module("utils.easy_renderer")
        cams = importlib.import_module("scene.cameras")
        args = importlib.import_module("arguments")
        lu = importlib.import_module("utils.loss_utils")
        gu = importlib.import_module("utils.graphics_utils")
        assert gr.GaussianRasterizer is dgr.GaussianRasterizer
        ns = types.SimpleNamespace(backend=backend, render=gr.render, GaussianModel=gr.GaussianModel, PseudoCamera=cams.PseudoCamera,
                                   EasyRenderer=er.EasyRenderer, PipelineParams=args.PipelineParams, ModelParams=args.ModelParams,
                                   loss_utils=lu, BasicPointCloud=gu.BasicPointCloud, rasterizer=dgr, distCUDA2=knn_c.distCUDA2,
                                   gaussian_renderer=gr)
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in _TOP]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path[:] = saved_path
    _loaded[backend] = ns
    return ns
