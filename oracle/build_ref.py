#!/usr/bin/env python
"""Build the UNMODIFIED reference native extensions from /root/reference into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
(guidedvd-3dgs_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
baseline legs use what this script builds.

What it builds (sources compiled where they lie; no reference source is copied
into the git history -- oracle/_ref/ is git-ignored but travels to the GPU box):

  oracle/_ref/diff_gaussian_rasterization/_C.so   <- submodules/diff-gaussian-rasterization-confidence
        {ext.cpp, rasterize_points.cu, cuda_rasterizer/{rasterizer_impl,forward,backward}.cu}
        flags: -I third_party/glm  -include cstdint   (rasterizer_impl.h needs <cstdint>)
  oracle/_ref/simple_knn/_C.so                    <- submodules/simple-knn/{ext.cpp,spatial.cu,simple_knn.cu}
        flags: -include cfloat  (simple_knn.cu uses FLT_MAX)

The python front-end file of the rasterizer package (the autograd.Function) is
*installed* next to the .so (like `pip install --target`) so that the reference
can be driven through its own public API on the GPU box, where /root/reference
does not exist.

  oracle/_ref/ViewCrafter/...                     <- pure-python denoiser / sampler / VAE modules (build_vc)
  oracle/_ref/gs/...                              <- the reference's own Python ABOVE the rasterizer boundary
        (gaussian_renderer/__init__.py::render, utils/easy_renderer.py::EasyRenderer, scene/gaussian_model.py,
        scene/cameras.py, arguments/, utils/{sh,graphics,general,system,loss}_utils.py), installed unmodified so the
        GPU tests can drive `render()` over the drop-in packages exactly as train_*.py does (tests/gs_refload.py)

Usage:  python oracle/build_ref.py [dgr] [knn] [vc] [gs]
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GVD_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def _load(name, sources, extra_cuda, build_dir, extra_include=()):
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load

    os.makedirs(build_dir, exist_ok=True)
    load(
        name=name,
        sources=sources,
        extra_cuda_cflags=list(extra_cuda),
        extra_include_paths=list(extra_include),
        build_directory=build_dir,
        is_python_module=False,  # just build; do not import (no GPU here)
        verbose=True,
    )
    return os.path.join(build_dir, name + ".so")


def build_dgr():
    src = os.path.join(REF, "submodules", "diff-gaussian-rasterization-confidence")
    if not os.path.isdir(src):
        print("reference not present, skip dgr")
        return False
    pkg = os.path.join(OUT, "diff_gaussian_rasterization")
    os.makedirs(pkg, exist_ok=True)
    so = _load(
        "_C",
        [
            os.path.join(src, "ext.cpp"),
            os.path.join(src, "rasterize_points.cu"),
            os.path.join(src, "cuda_rasterizer", "rasterizer_impl.cu"),
            os.path.join(src, "cuda_rasterizer", "forward.cu"),
            os.path.join(src, "cuda_rasterizer", "backward.cu"),
        ],
        ["-I" + os.path.join(src, "third_party", "glm"), "-include", "cstdint", "-lineinfo"],
        os.path.join(OUT, "build_dgr"),
    )
    shutil.copy2(so, os.path.join(pkg, "_C.so"))
    # install the package front-end (unmodified) beside the extension
    dst = os.path.join(pkg, "__init__.py")
    if os.path.exists(dst):
        os.chmod(dst, 0o644)
    shutil.copyfile(os.path.join(src, "diff_gaussian_rasterization", "__init__.py"), dst)
    print("built", os.path.join(pkg, "_C.so"))
    return True


def build_knn():
    src = os.path.join(REF, "submodules", "simple-knn")
    if not os.path.isdir(src):
        print("reference not present, skip knn")
        return False
    pkg = os.path.join(OUT, "simple_knn")
    os.makedirs(pkg, exist_ok=True)
    so = _load(
        "_C",
        [os.path.join(src, "ext.cpp"), os.path.join(src, "spatial.cu"), os.path.join(src, "simple_knn.cu")],
        ["-include", "cfloat", "-lineinfo"],
        os.path.join(OUT, "build_knn"),
    )
    shutil.copy2(so, os.path.join(pkg, "_C.so"))
    open(os.path.join(pkg, "__init__.py"), "w").close()
    print("built", os.path.join(pkg, "_C.so"))
    return True


def build_vc():
    """Install (copy) the pure-python ViewCrafter denoiser/sampler modules the parity tests import on the GPU box:
    lvdm/{basics,common}.py, lvdm/modules/{attention.py,networks/openaimodel3d.py,networks/ae_modules.py},
    lvdm/models/{utils_diffusion.py,samplers/*.py}, utils_vc/diffusion_utils.py.  They need torch, einops, cv2, tqdm."""
    src = os.path.join(REF, "third_party", "ViewCrafter")
    if not os.path.isdir(src):
        print("reference not present, skip vc")
        return False
    dst = os.path.join(OUT, "ViewCrafter")
    # utils_vc/diffusion_utils.py is pulled in by lvdm/basics.py (instantiate_from_config) and itself imports the
    # three sampler modules
    for rel in ("lvdm/basics.py", "lvdm/common.py", "lvdm/modules/attention.py", "lvdm/modules/networks/openaimodel3d.py",
                "lvdm/modules/networks/ae_modules.py", "lvdm/models/utils_diffusion.py", "lvdm/models/samplers/ddim.py", "lvdm/models/samplers/ddim_guidance.py",
                "lvdm/models/samplers/ddim_multiplecond.py", "utils_vc/diffusion_utils.py"):
        d = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.exists(d):
            os.chmod(d, 0o644)
        shutil.copyfile(os.path.join(src, rel), d)
        os.chmod(d, 0o644)
    print("installed", dst)
    return True


def build_gs():
    """Install (copy, unmodified) the reference's Python above the rasterizer boundary; see the module docstring."""
    if not os.path.isdir(os.path.join(REF, "gaussian_renderer")):
        print("reference not present, skip gs")
        return False
    dst = os.path.join(OUT, "gs")
    for rel in ("gaussian_renderer/__init__.py", "utils/easy_renderer.py", "utils/sh_utils.py", "utils/graphics_utils.py",
                "utils/general_utils.py", "utils/system_utils.py", "utils/loss_utils.py", "scene/gaussian_model.py",
                "scene/cameras.py", "arguments/__init__.py"):
        d = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.exists(d):
            os.chmod(d, 0o644)
        shutil.copyfile(os.path.join(REF, rel), d)
        os.chmod(d, 0o644)
    print("installed", dst)
    return True


if __name__ == "__main__":
    which = sys.argv[1:] or ["dgr", "knn", "vc", "gs"]
    if len(which) > 1:
        # one process per extension: torch's JIT loader renames a second "_C" built in the
        # same process to "_C_v1", which the packages' `from . import _C` would not find.
        import subprocess

        for w in which:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), w])
    elif which[0] == "dgr":
        build_dgr()
    elif which[0] == "knn":
        build_knn()
    elif which[0] == "vc":
        build_vc()
    elif which[0] == "gs":
        build_gs()
