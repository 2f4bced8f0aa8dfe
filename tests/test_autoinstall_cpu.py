"""Zero-edit activation (guidedvd-3dgs_b200/sitecustomize.py + vc_b200/autoinstall.py): with the package directory on
PYTHONPATH, an unmodified `third_party.ViewCrafter.viewcrafter.ViewCrafter.setup_diffusion`
(/root/reference/third_party/ViewCrafter/viewcrafter.py:315-335) gets the native U-Net / VAE swapped in right after
it has built `self.diffusion`.  Here the reference module is a stand-in with the same shape (class name, method name,
attribute name, import path as utils/viewcrafter_wrapper.py:27 uses it); every check runs in a fresh interpreter."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "guidedvd-3dgs_b200")

STUB = '''
class ViewCrafter:
    def __init__(self, opts=None, gradio=False, setup_diffusion=True, device="cuda:0"):
        self.opts, self.device = opts, device
        if setup_diffusion:
            self.setup_diffusion()

    def setup_diffusion(self):
        """reference docstring"""
        self.diffusion = "latent-diffusion-model"
        self.noise_shape = [1, 4, 25, 40, 64]
'''


def _tree(tmp_path):
    pkg = tmp_path / "third_party" / "ViewCrafter"
    pkg.mkdir(parents=True)
    (tmp_path / "third_party" / "__init__.py").write_text("")
    (pkg / "__init__.py").write_text("")
    (pkg / "viewcrafter.py").write_text(STUB)


def _run(tmp_path, script, **env):
    e = dict(os.environ)
    e["PYTHONPATH"] = os.pathsep.join([PKG, str(tmp_path)])
    e.update(env)
    return subprocess.run([sys.executable, "-c", textwrap.dedent(script)], env=e, capture_output=True, text=True, timeout=120)


def test_hook_wraps_setup_diffusion_without_touching_the_reference(tmp_path):
    _tree(tmp_path)
    r = _run(tmp_path, """
        import sys
        assert "torch" not in sys.modules, "the start-up hook must stay import-light"
        import vc_b200.autoinstall as ai
        assert ai.installed()                      # armed by sitecustomize before this script ran
        calls = []
        ai._apply = lambda vc: calls.append((vc.diffusion, list(vc.noise_shape)))
        from third_party.ViewCrafter.viewcrafter import ViewCrafter   # utils/viewcrafter_wrapper.py:27
        vc = ViewCrafter(None, setup_diffusion=True, device="cpu")    # utils/viewcrafter_wrapper.py:225-228
        assert calls == [("latent-diffusion-model", [1, 4, 25, 40, 64])], calls
        assert ViewCrafter.setup_diffusion.__doc__ == "reference docstring"
        ViewCrafter(None, setup_diffusion=False)                      # the other constructor path: nothing to swap yet
        assert len(calls) == 1
        assert ai.patched() == ["third_party.ViewCrafter.viewcrafter.ViewCrafter"], ai.patched()
        import importlib, third_party.ViewCrafter.viewcrafter as m
        importlib.reload(m)                                           # a reload is wrapped again, once
        m.ViewCrafter(None)
        assert len(calls) == 2
        print("ok")
    """)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr + r.stdout


def test_hook_can_be_disabled_and_patches_already_imported_modules(tmp_path):
    _tree(tmp_path)
    r = _run(tmp_path, """
        import os
        import vc_b200.autoinstall as ai
        assert not ai.installed()
        from third_party.ViewCrafter.viewcrafter import ViewCrafter
        assert not hasattr(ViewCrafter.setup_diffusion, "_gvd_wrapped")
        os.environ["GVD_AUTOINSTALL"] = "1"
        calls = []
        ai._apply = lambda vc: calls.append(vc.diffusion)
        assert ai.install() and ai.install()                          # idempotent; catches the module imported above
        ViewCrafter(None)
        assert calls == ["latent-diffusion-model"]
        print("ok")
    """, GVD_AUTOINSTALL="0")
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr + r.stdout


def test_a_later_sitecustomize_still_runs(tmp_path):
    _tree(tmp_path)
    (tmp_path / "sitecustomize.py").write_text("import os\nos.environ['OTHER_SITECUSTOMIZE_RAN'] = 'yes'\n")
    r = _run(tmp_path, """
        import os
        import vc_b200.autoinstall as ai
        assert ai.installed() and os.environ.get("OTHER_SITECUSTOMIZE_RAN") == "yes"
        print("ok")
    """)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr + r.stdout
