"""GPU parity tests: the sm_100a rasterizer (through the drop-in `diff_gaussian_rasterization` API and the
C ABI beneath it) against (1) the compiled UNMODIFIED reference in oracle/_ref on the same seeded inputs,
(2) the committed golden vectors the reference produced, (3) the CPU oracle, and (4) size-independent
properties at the BASELINE.json sizes.

Bars (BASELINE.json north_star): bit-exact on radii / sorted keys / point list / tile ranges / n_contrib;
<= 1e-4 relative on colour, depth, alpha (we in fact require bit-identical images against the compiled
reference); gradients <= 1e-4 relative to the tensor scale, and never worse than 4x the reference's own
run-to-run jitter (its backward uses float atomics in arbitrary order, so it does not reproduce itself
more tightly than that)."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "raster_*.npz")))


def _pkgs():
    import diff_gaussian_rasterization as ours
    import refload

    return ours, refload.ref_dgr()


def _grad_err(a, b):
    """max |a-b| relative to the tensor scale (rms), and relative L2."""
    a, b = a.double().flatten(), b.double().flatten()
    scale = b.pow(2).mean().sqrt().clamp_min(1e-30)
    return ((a - b).abs().max() / scale).item(), ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("cfg,D,use_conf,precomp", [
    ("tiny", 3, True, False), ("tiny", 0, False, False), ("tiny", 1, True, False), ("tiny", 2, True, False),
    ("small", 3, True, False), ("small", 0, True, True), ("C2", 3, True, False), ("C2", 0, False, False),
])
def test_against_compiled_reference(cfg, D, use_conf, precomp):
    import parity_raster as pr
    import synth

    ours, ref = _pkgs()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    P, W, H, seed = synth.CONFIGS[cfg]
    res = pr.compare(P, W, H, seed, sh_degree=D, use_conf=use_conf, precomp=precomp, verbose=False)
    assert res["R_ours"] == res["R_ref"]
    for k in ("radii_mismatch", "point_list_mismatch", "keys_mismatch", "tiles_touched_mismatch", "ranges_mismatch",
              "n_contrib_mismatch", "means2D_bits_mismatch", "depth_bits_mismatch", "conic_bits_mismatch"):
        assert res[k] == 0, (k, res[k])
    for k in ("color", "depth", "alpha"):
        assert res[f"{k}_bits_mismatch"] == 0, k          # stronger than the 1e-4 bar
        assert res[f"{k}_relerr"][0] <= 1e-4
    for k in [k for k in res if k.startswith("grad_") and k.endswith("_scale_err")]:
        name = k[len("grad_"):-len("_scale_err")]
        mx, l2 = res[k]                                   # max |a-b| / rms(ref), ||a-b|| / ||ref||
        jmx, jl2 = res[f"grad_{name}_ref_scale_jitter"]   # the reference against a second run of itself
        assert l2 <= max(1e-5, 4 * jl2), (name, l2, jl2)
        assert mx <= max(1e-3 if cfg == "tiny" else 1e-4, 4 * jmx), (name, mx, jmx)  # tiny: a few hundred terms per sum


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_against_golden(path):
    import parity_raster as pr

    ours, _ = _pkgs()
    g = np.load(path)
    P, W, H, seed, D = int(g["P"]), int(g["W"]), int(g["H"]), int(g["seed"]), int(g["sh_degree"])
    sc, cam, cot, bg, _ = pr.make_inputs(P, W, H, seed, D)
    o = pr.run(ours, sc, cam, cot, bg, D, bool(g["use_conf"]), bool(g["precomp"]))
    v = pr.ours_views(o, P, W, H)
    assert o["num_rendered"] == int(g["num_rendered"])
    assert np.array_equal(o["radii"].cpu().numpy(), g["radii"])
    assert np.array_equal(v["point_list"].cpu().numpy(), g["point_list"])
    assert np.array_equal(v["point_list_keys"].cpu().numpy(), g["point_list_keys"])
    assert np.array_equal(v["ranges"].cpu().numpy(), g["ranges"])
    assert np.array_equal(v["n_contrib"].cpu().numpy(), g["n_contrib"])
    for k in ("color", "depth", "alpha"):
        assert np.array_equal(o[k].cpu().numpy(), g[k]), k
    for k, go in o["grads"].items():
        if go is None:
            continue
        mx, l2 = _grad_err(go, torch.from_numpy(g["grad_" + k]).to(go.device).reshape(go.shape))
        assert l2 < 5e-5 and mx < 2e-3, (k, mx, l2)  # float-atomic order jitter on an 800-Gaussian scene


def test_against_cpu_oracle():
    import parity_raster as pr
    import raster_oracle as ro
    import synth

    ours, _ = _pkgs()
    P, W, H, seed = synth.CONFIGS["tiny"]
    sc, cam, cot, bg, D = pr.make_inputs(P, W, H, seed, 3)
    o = pr.run(ours, sc, cam, cot, bg, D, True)
    r = ro.run(ro.to_numpy_scene(sc), ro.to_numpy_scene(cam), bg.cpu().numpy(), D,
               cot={k: v.cpu().numpy() for k, v in cot.items()}, use_conf=True)
    v = pr.ours_views(o, P, W, H)
    assert np.array_equal(o["radii"].cpu().numpy(), r["radii"])
    assert np.array_equal(v["point_list"].cpu().numpy().astype(np.uint32), r["point_list"])
    assert np.array_equal(v["point_list_keys"].cpu().numpy().astype(np.uint64), r["point_list_keys"])
    for k in ("color", "depth", "alpha"):
        np.testing.assert_allclose(o[k].cpu().numpy(), r[k], rtol=1e-4, atol=2e-6)
    for k, gr in r["grads"].items():
        mx, l2 = _grad_err(o["grads"][k], torch.from_numpy(gr).to("cuda").reshape(o["grads"][k].shape))
        assert l2 < 5e-3, (k, l2)


def _settings(ours, cam, bg, D, conf, debug=False):
    return ours.GaussianRasterizationSettings(
        image_height=cam["height"], image_width=cam["width"], tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=bg,
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"], projmatrix=cam["projmatrix"], sh_degree=D,
        campos=cam["campos"], prefiltered=False, debug=debug, confidence=conf)


def test_edge_cases():
    """Empty scene, everything culled, ragged image size (not a multiple of 16), debug mode, markVisible."""
    import synth

    ours, ref = _pkgs()
    dev = "cuda"
    cam = synth.synth_camera(5, 70, 37, device=dev)  # ragged: 5x3 tiles, partial right/bottom tiles
    bg = torch.tensor([0.2, 0.4, 0.6], device=dev)
    # (1) P == 0
    r = ours.GaussianRasterizer(_settings(ours, cam, bg, 0, torch.ones(0, 1, device=dev)))
    e = torch.zeros(0, 3, device=dev)
    color, radii, depth, alpha = r(e, e, torch.zeros(0, 1, device=dev), shs=torch.zeros(0, 16, 3, device=dev),
                                   scales=e, rotations=torch.zeros(0, 4, device=dev))
    assert color.shape == (3, 37, 70) and radii.numel() == 0 and float(color.abs().sum()) == 0.0
    # (2) all Gaussians behind the camera: background everywhere, R == 0, zero grads
    sc = synth.synth_scene(500, 3, device=dev)
    fwd = cam["viewmatrix"][:3, 2]
    sc["means3D"] = (cam["campos"] - 3.0 * fwd)[None, :] + 0.01 * sc["means3D"]
    leaf = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    r = ours.GaussianRasterizer(_settings(ours, cam, bg, 3, sc["confidence"], debug=True))
    color, radii, depth, alpha = r(leaf["means3D"], torch.zeros_like(leaf["means3D"]), leaf["opacities"], shs=leaf["shs"],
                                   scales=leaf["scales"], rotations=leaf["rotations"])
    assert int((radii != 0).sum()) == 0
    assert torch.equal(color, bg[:, None, None].expand_as(color)) and float(alpha.abs().sum()) == 0
    (color.sum() + depth.sum() + alpha.sum()).backward()
    assert all(float(v.grad.abs().sum()) == 0 for v in leaf.values())
    # (3) ragged image against the reference, if present
    sc = synth.synth_scene(3000, 7, device=dev)
    if ref is not None:
        import parity_raster as pr

        g = torch.Generator().manual_seed(11)
        cot = dict(color=torch.randn(3, 37, 70, generator=g).to(dev), depth=torch.randn(1, 37, 70, generator=g).to(dev),
                   alpha=torch.randn(1, 37, 70, generator=g).to(dev))
        a = pr.run(ours, sc, cam, cot, bg, 3)
        b = pr.run(ref, sc, cam, cot, bg, 3)
        assert torch.equal(a["radii"], b["radii"]) and torch.equal(a["color"], b["color"])
        assert torch.equal(a["depth"], b["depth"]) and torch.equal(a["alpha"], b["alpha"])
        for k in a["grads"]:
            assert _grad_err(a["grads"][k], b["grads"][k])[1] < 1e-5, k
    # (4) markVisible == (view-space z > 0.2)
    r = ours.GaussianRasterizer(_settings(ours, cam, bg, 3, sc["confidence"]))
    vis = r.markVisible(sc["means3D"])
    z = (torch.cat([sc["means3D"], torch.ones(3000, 1, device=dev)], 1) @ cam["viewmatrix"])[:, 2]
    assert vis.dtype == torch.bool and int((vis != (z > 0.2)).sum()) <= 1


@pytest.mark.parametrize("cfg", ["C2", "C4"])
def test_properties_at_full_size(cfg):
    """Size-independent properties at BASELINE.json sizes: keys sorted and stable, ranges partition the list,
    alpha in [0, 1), colour linear in the SH DC term's cotangent (gradient linearity), determinism of forward."""
    import parity_raster as pr
    import synth

    ours, _ = _pkgs()
    P, W, H, seed = synth.CONFIGS[cfg]
    sc, cam, cot, bg, D = pr.make_inputs(P, W, H, seed, 3)
    a = pr.run(ours, sc, cam, cot, bg, D)
    v = pr.ours_views(a, P, W, H)
    keys, pl, ranges = v["point_list_keys"], v["point_list"], v["ranges"].long()
    R = a["num_rendered"]
    assert R == int(v["tiles_touched"].long().sum())
    assert bool((keys[1:] >= keys[:-1]).all())                      # sortedness
    same = keys[1:] == keys[:-1]
    assert bool((pl[1:][same] > pl[:-1][same]).all())               # stability: ties keep Gaussian-index order
    lens = ranges[:, 1] - ranges[:, 0]
    assert int(lens.sum()) == R and bool((lens >= 0).all())          # ranges partition the instance list
    tile_of = (keys >> 32)
    nz = lens > 0
    assert bool((tile_of[ranges[nz, 0]] == torch.nonzero(nz).flatten()).all())
    assert float(a["alpha"].min()) >= 0 and float(a["alpha"].max()) < 1.0 + 1e-5
    assert bool((v["n_contrib"].view(H, W)[::16, ::16].flatten().long() <= lens.view(-1, )[
        (torch.arange(0, H, 16, device="cuda")[:, None] // 16 * ((W + 15) // 16) + torch.arange(0, W, 16, device="cuda")[None, :] // 16).flatten()]).all())
    b = pr.run(ours, sc, cam, cot, bg, D, backward=False)
    assert torch.equal(a["color"], b["color"]) and torch.equal(a["depth"], b["depth"])  # forward is deterministic
    # gradient linearity in the cotangent: grads(2*cot) == 2*grads(cot) up to atomic-order rounding
    cot2 = {k: 2 * t for k, t in cot.items()}
    c = pr.run(ours, sc, cam, cot2, bg, D)
    for k in a["grads"]:
        assert _grad_err(c["grads"][k], 2 * a["grads"][k])[1] < 1e-5, k


def test_default_path_validates_before_returning(monkeypatch):
    """The production path (no key export), GVD_SPECULATE=sync: buffers sized from history, everything queued in one go,
    and the {R, V} the second kernel stored in pinned memory are checked BEFORE the forward returns -- a frame that outgrew
    the guess is redone exactly and its clamped outputs are never seen.  Results must equal the parity (synchronous)
    path's in every case; nothing is deferred, nothing warns, nothing raises."""
    import warnings

    import parity_raster as pr
    import synth

    ours, _ = _pkgs()
    assert not ours._C.DEFER and ours._C.SPECULATE
    P, W, H, seed = synth.CONFIGS["small"]
    sc, cam, cot, bg, D = pr.make_inputs(P, W, H, seed, 3)
    exact = pr.run(ours, sc, cam, cot, bg, D)
    ours._C._spec_state.clear()
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        runs = [pr.run(ours, sc, cam, cot, bg, D, export_keys=False) for _ in range(3)]  # 1st: no history -> exact path
        monkeypatch.setattr(ours._C, "_capacity", lambda max_R: 4096)        # every later guess is far too small ...
        runs.append(pr.run(ours, sc, cam, cot, bg, D, export_keys=False))
        monkeypatch.undo()
        monkeypatch.setattr(ours._C, "_visible_capacity", lambda max_V, P: 64)   # ... for the chunk histogram too
        runs.append(pr.run(ours, sc, cam, cot, bg, D, export_keys=False))
        monkeypatch.undo()
        monkeypatch.setattr(ours._C, "SPECULATE", False)                     # GVD_SPECULATE=exact
        runs.append(pr.run(ours, sc, cam, cot, bg, D, export_keys=False))
        monkeypatch.undo()
    st = ours._C._spec_state[sc["means3D"].device]
    assert st["max_R"] == exact["num_rendered"] and st["max_V"] == int((exact["radii"] > 0).sum()) and len(st["free"]) >= 1
    for r in runs:
        assert r["num_rendered"] == exact["num_rendered"] and isinstance(r["num_rendered"], int)
        for k in ("color", "depth", "alpha", "radii"):
            assert torch.equal(r[k], exact[k]), k
        for k in exact["grads"]:
            assert _grad_err(r["grads"][k], exact["grads"][k])[1] < 1e-5, k
    with torch.no_grad():
        nograd = pr.run(ours, sc, cam, cot, bg, D, export_keys=False, backward=False)
    assert torch.equal(nograd["color"], exact["color"])


def test_deferred_mode_matches_exact_and_repairs_an_overflow(monkeypatch):
    """GVD_SPECULATE=defer (opt-in): no host wait between forward and backward.  Same results as the exact path while the
    guesses fit; a frame that outgrows them is re-rendered exactly by its backward, with a warning -- never an
    exception inside an unmodified trainer -- and the grown history makes the next frame fit."""
    import parity_raster as pr
    import synth

    ours, _ = _pkgs()
    P, W, H, seed = synth.CONFIGS["small"]
    sc, cam, cot, bg, D = pr.make_inputs(P, W, H, seed, 3)
    exact = pr.run(ours, sc, cam, cot, bg, D)
    dev = sc["means3D"].device
    monkeypatch.setattr(ours._C, "DEFER", True)
    ours._C._spec_state.clear()
    runs = [pr.run(ours, sc, cam, cot, bg, D, export_keys=False) for _ in range(3)]  # 1st: no history -> exact path
    for r in runs:
        assert r["num_rendered"] == exact["num_rendered"]
        for k in ("color", "depth", "alpha", "radii"):
            assert torch.equal(r[k], exact[k]), k
        for k in exact["grads"]:
            assert _grad_err(r["grads"][k], exact["grads"][k])[1] < 1e-5, k
    with torch.no_grad():  # no backward to come: never deferred
        nograd = pr.run(ours, sc, cam, cot, bg, D, export_keys=False, backward=False)
    assert torch.equal(nograd["color"], exact["color"])
    real_capacity = ours._C._capacity
    monkeypatch.setattr(ours._C, "_capacity", lambda max_R: 4096)  # every later guess is far too small
    with pytest.warns(UserWarning, match="Re-rendered exactly"):
        broken = pr.run(ours, sc, cam, cot, bg, D, export_keys=False)
    torch.cuda.synchronize()
    assert torch.equal(broken["radii"], exact["radii"])
    for k in exact["grads"]:  # gradients of the frame as it should have been
        assert _grad_err(broken["grads"][k], exact["grads"][k])[1] < 1e-5, k
    monkeypatch.setattr(ours._C, "_capacity", real_capacity)
    again = pr.run(ours, sc, cam, cot, bg, D, export_keys=False)
    assert torch.equal(again["color"], exact["color"]) and ours._C._spec_state[dev]["max_R"] == exact["num_rendered"]
    # a deferred frame that never gets a backward is validated at the next render on that device, not at GC
    # (or, as here where nothing keeps the graph alive, when its autograd node is dropped)
    monkeypatch.setattr(ours._C, "_capacity", lambda max_R: 4096)
    with pytest.warns(UserWarning, match="speculative buffers"):
        lost = pr.run(ours, sc, cam, cot, bg, D, export_keys=False, backward=False)
        torch.cuda.synchronize()
        monkeypatch.setattr(ours._C, "_capacity", real_capacity)
        pr.run(ours, sc, cam, cot, bg, D, export_keys=False)
    del lost


def test_two_renders_into_one_backward_with_a_gradient_buffer():
    """set_gradient_buffer (the in-place cross-GPU exchange) hands every backward the same memory.  Two renders that feed
    ONE loss.backward() -- train_guidedvd.py's train view + pseudo view -- must still accumulate both gradients: the second
    backward arrives before anybody asked for the first one's views and gets fresh memory (with a warning)."""
    import warnings

    import synth

    ours, _ = _pkgs()
    dev = "cuda"
    sc = synth.synth_scene(4000, 11, device=dev)
    cams = [synth.synth_camera(21 + i, 96, 64, device=dev) for i in range(2)]
    bg = torch.zeros(3, device=dev)

    def grads(use_buffer):
        leaf = {k: sc[k].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
        if use_buffer:
            ours.set_gradient_buffer(torch.empty(ours.gradient_buffer_floats(4000), device=dev))
        try:
            total = 0.0
            for cam in cams:
                r = ours.GaussianRasterizer(_settings(ours, cam, bg, 3, sc["confidence"]))
                color, radii, depth, alpha = r(leaf["means3D"], torch.zeros_like(leaf["means3D"]), leaf["opacities"], shs=leaf["shs"],
                                               scales=leaf["scales"], rotations=leaf["rotations"])
                total = total + (color ** 2).sum() + depth.sum()
            with warnings.catch_warnings(record=True) as w:
                warnings.simplefilter("always")
                total.backward()
            torch.cuda.synchronize()
            return {k: v.grad.clone() for k, v in leaf.items()}, [str(x.message) for x in w]
        finally:
            ours.set_gradient_buffer(None)

    plain, _ = grads(False)
    buffered, msgs = grads(True)
    for k in plain:
        rel = ((buffered[k] - plain[k]).double().norm() / plain[k].double().norm().clamp_min(1e-30)).item()
        assert rel < 1e-5, (k, rel)
