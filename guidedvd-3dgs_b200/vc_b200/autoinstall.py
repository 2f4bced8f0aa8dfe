"""Zero-edit activation of the B200-native diffusion path inside the UNMODIFIED reference tree.

The reference reaches the denoiser through `utils/viewcrafter_wrapper.py:27,225-228` ->
`third_party.ViewCrafter.viewcrafter.ViewCrafter(opts, setup_diffusion=..., device=...)`, whose `setup_diffusion`
(third_party/ViewCrafter/viewcrafter.py:315-335) instantiates the LatentDiffusion model, loads the checkpoint and stores
it in `self.diffusion`.  `install()` puts a post-import hook on `sys.meta_path`: when a module whose last name component
is `viewcrafter` has been executed and exposes a class `ViewCrafter` with a `setup_diffusion` method, that method is
wrapped so that `vc_b200.dropin.replace_unet / replace_first_stage_decoder / replace_first_stage_encoder` run on
`self.diffusion` right after the reference's own code.  Nothing under the reference tree is edited:

    PYTHONPATH=/path/to/guidedvd-3dgs_b200 python train_guidedvd.py ...

`guidedvd-3dgs_b200/sitecustomize.py` (imported by the interpreter at start-up from that PYTHONPATH entry) calls
`install()`; the same directory provides `diff_gaussian_rasterization` and `simple_knn`, so the rasterizer half drops
in through the same variable.  GVD_AUTOINSTALL=0 turns the hook off.

This module must stay import-light (no torch): it runs in every interpreter started with that PYTHONPATH.
"""
import importlib.abc
import importlib.machinery
import os
import sys

_TARGET_LEAF = "viewcrafter"
_state = {"installed": False, "patched": []}


def _apply(view_crafter):
    """What runs after the reference's setup_diffusion: swap in the native U-Net / VAE decoder / VAE encoder."""
    from . import dropin

    model = view_crafter.diffusion
    dropin.replace_unet(model)
    dropin.replace_first_stage_decoder(model)
    try:
        dropin.replace_first_stage_encoder(model)
    except ImportError:  # the encoder hook needs the reference's lvdm.distributions on sys.path
        pass


def patch_class(cls):
    """Wrap cls.setup_diffusion once.  Returns True when the class was patched by this call."""
    orig = getattr(cls, "setup_diffusion", None)
    if orig is None or getattr(orig, "_gvd_wrapped", False):
        return False

    def setup_diffusion(self, *args, **kwargs):
        out = orig(self, *args, **kwargs)
        if getattr(self, "diffusion", None) is not None:
            _apply(self)
        return out

    setup_diffusion._gvd_wrapped = True
    setup_diffusion.__wrapped__ = orig
    setup_diffusion.__doc__ = orig.__doc__
    cls.setup_diffusion = setup_diffusion
    _state["patched"].append(f"{cls.__module__}.{cls.__qualname__}")
    return True


def _patch_module(module):
    cls = getattr(module, "ViewCrafter", None)
    if isinstance(cls, type):
        patch_class(cls)


class _Loader(importlib.abc.Loader):
    def __init__(self, inner):
        self.inner = inner

    def create_module(self, spec):
        return self.inner.create_module(spec)

    def exec_module(self, module):
        self.inner.exec_module(module)
        _patch_module(module)

    def __getattr__(self, name):  # get_code, get_source, is_package ... of the real loader
        return getattr(self.inner, name)


class _Finder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.rpartition(".")[2] != _TARGET_LEAF:
            return None
        for finder in sys.meta_path:
            if finder is self or not hasattr(finder, "find_spec"):
                continue
            spec = finder.find_spec(fullname, path, target)
            if spec is not None and spec.loader is not None and hasattr(spec.loader, "exec_module"):
                spec.loader = _Loader(spec.loader)
                return spec
            if spec is not None:
                return spec
        return None


def install():
    """Idempotent.  Also patches a matching module that was imported before the hook existed."""
    if os.environ.get("GVD_AUTOINSTALL", "1") == "0" or _state["installed"]:
        return _state["installed"]
    sys.meta_path.insert(0, _Finder())
    _state["installed"] = True
    for name, module in list(sys.modules.items()):
        if module is not None and name.rpartition(".")[2] == _TARGET_LEAF:
            _patch_module(module)
    return True


def installed():
    return _state["installed"]


def patched():
    """Qualified names of the classes whose setup_diffusion has been wrapped so far."""
    return list(_state["patched"])
