"""Start-up hook: with `PYTHONPATH=/path/to/guidedvd-3dgs_b200`, the interpreter imports this file before the user's
script (site.py -> `import sitecustomize`).  It arms vc_b200.autoinstall, which swaps the B200-native U-Net / VAE into
`ViewCrafter.setup_diffusion` (third_party/ViewCrafter/viewcrafter.py:315-335) when -- and only when -- the reference's
module is imported.  No torch import, no CUDA work here.  GVD_AUTOINSTALL=0 disables it.

If another `sitecustomize` sits later on sys.path (a virtualenv's, a distribution's), it is run afterwards so this
file does not mask it."""
import os
import sys


def _chain_next():
    here = os.path.dirname(os.path.abspath(__file__))
    import importlib.machinery
    import importlib.util

    rest = [p for p in sys.path if p and os.path.abspath(p) != here]
    spec = importlib.machinery.PathFinder.find_spec("sitecustomize", rest)
    if spec is None or spec.loader is None or (spec.origin and os.path.abspath(spec.origin) == os.path.abspath(__file__)):
        return
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)


if os.environ.get("GVD_AUTOINSTALL", "1") != "0":
    try:
        from vc_b200 import autoinstall as _gvd_autoinstall

        _gvd_autoinstall.install()
    except Exception as _ex:  # a broken hook must never take the interpreter down
        sys.stderr.write(f"guidedvd-3dgs_b200/sitecustomize.py: autoinstall failed: {_ex!r}\n")
try:
    _chain_next()
except Exception as _ex:
    sys.stderr.write(f"guidedvd-3dgs_b200/sitecustomize.py: chained sitecustomize failed: {_ex!r}\n")
