"""Multi-view forward rendering of a trained Gaussian set: what `utils/easy_renderer.py::EasyRenderer.render`
does per pose (easy_renderer.py:59-81 -> gaussian_renderer/__init__.py:19-132), for a whole trajectory at once.

train_guidedvd.py renders the 25 (or 75) poses of every diffusion round with one `easy_renderer.render(w2c, K, h, w)` call
each (train_guidedvd.py:157-165,521-527).  Each call re-runs the Gaussian activations (`exp`, `sigmoid`, `normalize`) and
the 96 MB `torch.cat` of the SH coefficients (gaussian_renderer/__init__.py:60-87, scene/gaussian_model.py get_features),
builds a camera module, and the rasterizer waits for its instance count before the host can prepare the next pose.  Here
the activations run once per batch, the camera matrices come from the same formulas without the nn.Module, and
`diff_gaussian_rasterization.rasterize_views` queues all poses as one uninterrupted GPU pipeline.

`render_trajectory` returns exactly what a loop over `EasyRenderer.render` returns (bit-identical; pinned by
tests/test_render_dropin_gpu.py against the reference's own EasyRenderer over both backends)."""
import math

import numpy as np
import torch

from diff_gaussian_rasterization import GaussianRasterizationSettings, rasterize_views


def _world_to_view(R, t):
    """utils/graphics_utils.py:38-49 getWorld2View2 with translate = 0, scale = 1 (what PseudoCamera passes)."""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    C2W = np.linalg.inv(Rt)           # the reference inverts twice (camera centre shift + rescale in between, both
    return np.float32(np.linalg.inv(C2W))  # identities here); kept so the matrices are bit-identical to PseudoCamera's


def _projection(fovX, fovY):
    """utils/graphics_utils.py:51-75 getProjectionMatrix -- the reference's non-standard P (P[2,2] = P[3,2] = 1: p_hom.w
    = p_hom.z = view-space z; znear / zfar do not enter)."""
    P = torch.zeros([4, 4])
    P[0, 0] = 1.0 / math.tan(fovX / 2)
    P[1, 1] = 1.0 / math.tan(fovY / 2)
    P[2, 2] = 1.0
    P[3, 2] = 1.0
    return P


def _focal2fov(focal, pixels):
    return 2 * math.atan(pixels / (2 * focal))   # utils/graphics_utils.py:80-81


def camera_settings(w2c, intrinsic, h, w, bg, sh_degree, confidence, device="cuda", scale_modifier=1.0, debug=False):
    """One pose -> GaussianRasterizationSettings, following EasyRenderer.make_gs_view_format (easy_renderer.py:68-81) and
    PseudoCamera (scene/cameras.py:67-93: znear 0.01, zfar 100) and render()'s settings (gaussian_renderer/__init__.py:35-56)."""
    fovx, fovy = _focal2fov(intrinsic[0, 0], w), _focal2fov(intrinsic[1, 1], h)
    R, T = np.transpose(w2c[:3, :3]), w2c[:3, 3]
    wvt = torch.tensor(_world_to_view(R, T)).transpose(0, 1).to(device)
    proj = _projection(fovx, fovy).transpose(0, 1).to(device)
    full = (wvt.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
    return GaussianRasterizationSettings(
        image_height=int(h), image_width=int(w), tanfovx=math.tan(fovx * 0.5), tanfovy=math.tan(fovy * 0.5), bg=bg,
        scale_modifier=scale_modifier, viewmatrix=wvt, projmatrix=full, sh_degree=sh_degree, campos=wvt.inverse()[3, :3],
        prefiltered=False, debug=debug, confidence=confidence)


def render_trajectory(gaussians, pipe, background, w2cs, intrinsics, h, w):
    """gaussians: the reference GaussianModel (any object with get_xyz / get_opacity / get_scaling / get_rotation /
    get_features / active_sh_degree / confidence / bg_color); pipe: PipelineParams; w2cs, intrinsics: sequences of [4,4] /
    [3,3] numpy arrays.  -> (renders [N,3,h,w], alphas [N,1,h,w], depths [N,1,h,w]) as EasyRenderer.render per pose."""
    with torch.no_grad():
        means3D, opacity = gaussians.get_xyz, gaussians.get_opacity
        scales = rotations = cov = None
        if pipe.compute_cov3D_python:
            cov = gaussians.get_covariance(1.0)
        else:
            scales, rotations = gaussians.get_scaling, gaussians.get_rotation
        shs = gaussians.get_features
        if getattr(pipe, "convert_SHs_python", False):
            raise NotImplementedError("render_trajectory: convert_SHs_python colours are view dependent; use render() per pose")
        train_bg = min(gaussians.bg_color.shape) != 0
        bg = torch.tensor([0., 0., 0.], device=means3D.device) if train_bg else background
        conf = gaussians.confidence if pipe.use_confidence else torch.ones_like(gaussians.confidence)
        settings = [camera_settings(np.asarray(p), np.asarray(k), h, w, bg, gaussians.active_sh_degree, conf, device=means3D.device,
                                    debug=pipe.debug) for p, k in zip(w2cs, intrinsics)]
        outs = rasterize_views(settings, means3D, opacity, shs=shs, scales=scales, rotations=rotations, cov3D_precomp=cov)
        color = torch.stack([o[0] for o in outs])
        depth = torch.stack([o[2] for o in outs])
        alpha = torch.stack([o[3] for o in outs])
        if train_bg:
            color = color + (1 - alpha) * torch.sigmoid(gaussians.bg_color)
    return color, alpha, depth
