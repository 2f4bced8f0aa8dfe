"""The rasterizer's CUDA sources executed on the HOST (tests/cuda_emu: raster_api.cu, raster_forward.cu,
raster_backward.cu compiled as they are; CUDA threads are OS threads, warp collectives / barriers / atomics keep their
meaning, the TMA bulk copy + mbarrier pair and the programmatic-dependent-launch intrinsics are replaced by host
equivalents under GVD_HOST_EMU in raster_common.cuh) through the C ABI, against the golden vectors of the compiled
REFERENCE (tests/golden/raster_*.npz, produced on a B200 by tests/make_golden.py).

What this pins without a GPU: the whole kernel chain's logic -- preprocess with its per-CTA counts, the compaction of
the visible Gaussians (V, R), the hand-written 4-pass radix depth sort, the rect-aware counting sort (bin_count /
bin_prefix / bin_ranges / the mask-ranked bin_fill), both render kernels with their sub-tile culling and batched id
staging, the transposing-butterfly reduction of the backward with its fused zero fill, the per-visible-Gaussian
backward.  Integer buffers are
compared exactly; floats with the bounds of tests/test_oracle_cpu.py (x86 expf / no FMA contraction differ from the
GPU's in the last ulp and flip isolated alpha < 1/255 decisions on a 2 000-Gaussian scene)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "guidedvd-3dgs_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import test_oracle_cpu as toc  # noqa: E402


@pytest.mark.parametrize("path", toc.GOLDEN, ids=[os.path.basename(p) for p in toc.GOLDEN])
def test_kernel_chain_matches_reference_golden(path):
    import raster_emu

    g = np.load(path)
    _, sc, cam, cot, bg, D, precomp = toc._inputs(g)
    backward = "d0" not in os.path.basename(path)      # two of the three fixtures run the backward as well (time)
    flat = "precomp" in os.path.basename(path)          # one fixture hands the backward one flat gradient region
    o = raster_emu.run(sc, cam, bg, D, cot=cot if backward else None, use_conf=bool(g["use_conf"]), precomp=precomp,
                       pinned="d3" in os.path.basename(path), flat_grads=flat, know_visible=not flat)
    P = int(g["P"])
    # the compacted id list is exactly the Gaussians with a radius, ascending; V and R are the device counts
    assert np.array_equal(o["visible_ids"], np.nonzero(o["radii"] > 0)[0].astype(np.uint32))
    assert o["num_visible"] == int((o["radii"] > 0).sum()) == int(o["counts"][0]) and int(o["counts"][1]) == o["num_rendered"]
    assert o["num_rendered"] == int(o["tiles_touched"].astype(np.int64).sum())
    assert (o["radii"] != g["radii"]).sum() <= max(1, P // 2000)
    assert (o["tiles_touched"].astype(np.int64) != g["tiles_touched"].astype(np.int64)).sum() <= max(1, P // 2000)
    assert abs(o["num_rendered"] - int(g["num_rendered"])) <= 64
    if o["num_rendered"] == int(g["num_rendered"]):
        assert (o["point_list"].astype(np.int64) != g["point_list"].astype(np.int64)).mean() < 2e-3
        # keys = tile << 32 | depth bits: the tile half exactly; the depth half to 2 ulp (this host build does not contract
        # the view-space z into FMAs as nvcc does -- the ORDER, i.e. point_list above, is what has to agree)
        ke, kg = o["point_list_keys"].astype(np.uint64), g["point_list_keys"].astype(np.uint64)
        assert np.array_equal(ke >> np.uint64(32), kg >> np.uint64(32))
        lo = np.uint64(0xffffffff)
        assert np.abs((ke & lo).astype(np.int64) - (kg & lo).astype(np.int64)).max() <= 2
        assert (o["ranges"] != g["ranges"]).sum() == 0
    assert (o["n_contrib"].reshape(-1).astype(np.int64) != g["n_contrib"].reshape(-1).astype(np.int64)).mean() < 5e-3
    for k in ("color", "depth", "alpha"):
        a, b = o[k].astype(np.float64), g[k].astype(np.float64).reshape(o[k].shape)
        bad = np.abs(a - b) > 1e-4 * np.abs(b) + 1e-5
        assert bad.mean() < 2e-3, (k, bad.mean())
    if backward:
        for k, go in o["grads"].items():
            gr = g["grad_" + k].astype(np.float64).reshape(go.shape)
            rel = np.sqrt(((go - gr) ** 2).sum()) / max(np.sqrt((gr ** 2).sum()), 1e-30)
            assert rel < 5e-3, (k, rel)
        # untouched Gaussians get exact zeros (the outputs start as garbage: the backward's own zero fill produced them)
        inv = g["radii"] == 0
        assert all(not np.any(v[inv]) for v in o["grads"].values())


def test_everything_behind_the_camera_and_tiny_inputs():
    import raster_emu
    import raster_oracle as ro
    import synth

    sc = ro.to_numpy_scene(synth.synth_scene(64, 1))
    cam = ro.to_numpy_scene(synth.synth_camera(2, 64, 48))
    sc["means3D"] = (sc["means3D"] * 0 - np.array(cam["viewmatrix"])[2, :3] * 5.0 + np.asarray(cam["campos"])).astype(np.float32)
    bg = np.array([0.3, 0.6, 0.9], np.float32)
    cot = dict(color=np.ones((3, 48, 64), np.float32), depth=np.ones((1, 48, 64), np.float32), alpha=np.ones((1, 48, 64), np.float32))
    o = raster_emu.run(sc, cam, bg, 3, cot=cot)
    assert o["num_rendered"] == 0 and (o["radii"] == 0).all()
    assert np.allclose(o["color"], bg[:, None, None]) and (o["alpha"] == 0).all() and (o["depth"] == 0).all()
    assert all((v == 0).all() for v in o["grads"].values())
    # a ragged image (not a multiple of the 16x16 tile) with a handful of Gaussians: same answer as the C oracle
    sc = ro.to_numpy_scene(synth.synth_scene(40, 3))
    cam = ro.to_numpy_scene(synth.synth_camera(4, 37, 29))
    cot = {k: np.random.default_rng(0).normal(size=s).astype(np.float32) for k, s in dict(color=(3, 29, 37), depth=(1, 29, 37), alpha=(1, 29, 37)).items()}
    o = raster_emu.run(sc, cam, bg, 2, cot=cot, use_conf=True)
    r = ro.run(sc, cam, bg, 2, cot=cot, use_conf=True)
    assert o["num_rendered"] == r["num_rendered"] and np.array_equal(o["radii"], r["radii"])
    if o["num_rendered"]:
        assert np.array_equal(o["point_list"], r["point_list"]) and np.array_equal(o["ranges"], r["ranges"])
    for k in ("color", "depth", "alpha"):
        assert np.abs(o[k] - r[k].reshape(o[k].shape)).max() < 1e-4


def test_speculative_forward_equals_exact_and_survives_overflow():
    """The no-host-round-trip forward (`GVD_SPECULATE`): with an instance buffer sized from a guess it
    writes R through the caller's pinned word and produces exactly the exact path's buffers; with a buffer that is too
    small every write and read is clamped (no out-of-bounds access -- the buffer ends where the allocation ends) and the
    caller sees R > capacity."""
    import raster_emu

    g = np.load(toc.GOLDEN[0])
    _, sc, cam, cot, bg, D, precomp = toc._inputs(g)
    exact = raster_emu.run(sc, cam, bg, D, use_conf=bool(g["use_conf"]), precomp=precomp)
    R = exact["num_rendered"]
    spec = raster_emu.run(sc, cam, bg, D, use_conf=bool(g["use_conf"]), precomp=precomp, spec_capacity=2 * R + 1000)
    assert spec["num_rendered"] == R
    for k in ("radii", "point_list", "ranges", "n_contrib", "color", "depth", "alpha"):
        assert np.array_equal(spec[k], exact[k]), k
    small = raster_emu.run(sc, cam, bg, D, use_conf=bool(g["use_conf"]), precomp=precomp, spec_capacity=R // 3)
    assert small.get("overflow") and small["num_rendered"] == R and np.array_equal(small["radii"], exact["radii"])
    # the chunk-histogram buffer can be the one that is too small (V outgrew the guess): clamped as well, reported through V
    V = exact["num_visible"]
    few = raster_emu.run(sc, cam, bg, D, use_conf=bool(g["use_conf"]), precomp=precomp, spec_capacity=2 * R + 1000, spec_visible=V // 3)
    assert few.get("overflow") and few["num_visible"] == V and few["num_rendered"] == R
    fit = raster_emu.run(sc, cam, bg, D, use_conf=bool(g["use_conf"]), precomp=precomp, spec_capacity=R, spec_visible=V)
    for k in ("radii", "point_list", "ranges", "n_contrib", "color", "depth", "alpha"):
        assert np.array_equal(fit[k], exact[k]), k


def test_depth_sort_and_fill_at_sizes_that_cross_tile_boundaries():
    """More visible Gaussians than one sort tile (1024) and one super-tile of the offset lookup would need 32 768, which
    the host emulation cannot afford; 5 000 visible ones cross four sort tiles and exercise the per-tile histograms the
    passes hand to each other.  Big rects (scale x 6) push chunks over the row-walk threshold of bin_fill.  Checked
    against the C oracle (oracle/raster_oracle.c), integer buffers exactly."""
    import raster_emu
    import raster_oracle as ro
    import synth

    sc = ro.to_numpy_scene(synth.synth_scene(9000, 11))
    sc["scales"] = (sc["scales"] * 6.0).astype(np.float32)
    cam = ro.to_numpy_scene(synth.synth_camera(12, 208, 160))
    bg = np.zeros(3, np.float32)
    o = raster_emu.run(sc, cam, bg, 1)
    r = ro.run(sc, cam, bg, 1)
    assert o["num_visible"] > 4 * 1024 and o["num_rendered"] > 64 * 4096 // 8
    assert o["num_rendered"] == r["num_rendered"] and np.array_equal(o["radii"], r["radii"])
    assert np.array_equal(o["ranges"], r["ranges"])
    assert np.array_equal(o["point_list"], r["point_list"])


def test_raw_parameter_mode_folds_the_activations():
    """raw_params = 1 (exp / normalize / sigmoid and the dc | rest SH split inside the kernels) against the
    standard mode fed with the activations computed the way gaussian_renderer.render() computes them
    (scene/gaussian_model.py:36-43,106-130): same images, and raw-parameter gradients equal to torch.autograd's chain
    rule over the standard mode's gradients."""
    import torch

    import raster_emu
    import synth

    sc = synth.synth_scene(600, 5)
    cam = synth.synth_camera(6, 96, 64)
    bg = np.array([0.1, 0.2, 0.3], np.float32)
    g = np.random.default_rng(3)
    f = lambda k: np.asarray(sc[k].cpu() if hasattr(sc[k], "cpu") else sc[k], dtype=np.float32)  # noqa: E731
    shs = f("shs")
    raw_t = dict(scaling=torch.tensor(np.log(f("scales"))), rotation=torch.tensor(f("rotations") * g.uniform(0.5, 2.0, (600, 1)).astype(np.float32)),
                 opacity=torch.tensor(np.log(f("opacities").reshape(-1) / (1 - f("opacities").reshape(-1)))),
                 features_dc=torch.tensor(shs[:, :1].copy()), features_rest=torch.tensor(shs[:, 1:].copy()))
    for v in raw_t.values():
        v.requires_grad_(True)
    act = dict(scales=torch.exp(raw_t["scaling"]), rotations=torch.nn.functional.normalize(raw_t["rotation"]),
               opacities=torch.sigmoid(raw_t["opacity"]), shs=torch.cat((raw_t["features_dc"], raw_t["features_rest"]), dim=1))
    sc_act = dict(sc)
    for k, v in act.items():
        sc_act[k] = v.detach().numpy()
    cot = dict(color=g.standard_normal((3, 64, 96)).astype(np.float32), depth=g.standard_normal((1, 64, 96)).astype(np.float32),
               alpha=g.standard_normal((1, 64, 96)).astype(np.float32))
    std = raster_emu.run(sc_act, cam, bg, 3, cot=cot)
    raw = raster_emu.run(sc_act, cam, bg, 3, cot=cot, raw={k: v.detach().numpy() for k, v in raw_t.items()})
    assert raw["num_rendered"] > 1000
    assert (raw["radii"] != std["radii"]).mean() < 0.01 and abs(raw["num_rendered"] - std["num_rendered"]) <= 0.01 * std["num_rendered"]
    for k in ("color", "depth", "alpha"):
        assert np.abs(raw[k] - std[k]).max() <= 2e-5 * max(1.0, np.abs(std[k]).max()), k
    # chain rule through the activations, by autograd, on the standard mode's gradients
    gs = std["grads"]
    torch.autograd.backward([act["scales"], act["rotations"], act["opacities"], act["shs"]],
                            [torch.tensor(gs["scales"]), torch.tensor(gs["rotations"]), torch.tensor(gs["opacities"]).reshape(-1),
                             torch.tensor(gs["shs"])])
    want = {"scales": raw_t["scaling"].grad, "rotations": raw_t["rotation"].grad, "opacities": raw_t["opacity"].grad,
            "features_dc": raw_t["features_dc"].grad, "features_rest": raw_t["features_rest"].grad}
    for k, w in want.items():
        got, w = raw["grads"][k].reshape(-1), w.numpy().reshape(-1)
        rel = np.linalg.norm(got - w) / max(np.linalg.norm(w), 1e-12)
        assert rel < 1e-3, (k, rel)  # the host build reaches 1-3e-3 against the goldens itself (libm expf, no FMA contraction)
    for k in ("means3D", "means2D"):
        rel = np.linalg.norm(raw["grads"][k] - gs[k]) / np.linalg.norm(gs[k])
        assert rel < 1e-3, (k, rel)  # the host build reaches 1-3e-3 against the goldens itself (libm expf, no FMA contraction)
