"""`render()` with the GaussianModel activations folded into the rasterizer kernels.

Same signature and the same result dict as the reference's `gaussian_renderer.render` (gaussian_renderer/__init__.py:19-132):
    render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None, white_bg=False)
      -> {"render", "viewspace_points", "visibility_filter", "radii", "depth", "alpha"}
For the common configuration (SH colours, scale + rotation covariance: `pipe.convert_SHs_python` and
`pipe.compute_cov3D_python` off, no `override_color`) it hands the RAW parameters `pc._scaling`, `pc._rotation`,
`pc._opacity`, `pc._features_dc`, `pc._features_rest` to `diff_gaussian_rasterization.rasterize_gaussians_raw`: no
`torch.exp` / `sigmoid` / `normalize` launches, no 96 MB `torch.cat` per call, and one backward node instead of six.
Every other configuration goes through the standard `GaussianRasterizer` exactly like the reference's function.

    import gaussian_renderer_b200
    gaussian_renderer.render = gaussian_renderer_b200.render        # or import it in train_*.py
"""
import math

import torch

from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians_raw


def render(viewpoint_camera, pc, pipe, bg_color, scaling_modifier=1.0, override_color=None, white_bg=False):
    xyz = pc.get_xyz
    # the gradient carrier of the screen-space means (gaussian_renderer/__init__.py:28-32)
    screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device=xyz.device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    if min(pc.bg_color.shape) != 0:  # learnable background: composited below
        bg_color = torch.zeros(3, device=xyz.device)
    confidence = pc.confidence if pipe.use_confidence else torch.ones_like(pc.confidence)
    settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height), image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5), tanfovy=math.tan(viewpoint_camera.FoVy * 0.5), bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform, sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center, prefiltered=False, debug=pipe.debug, confidence=confidence)

    folded = override_color is None and not pipe.convert_SHs_python and not pipe.compute_cov3D_python and not pipe.debug
    if folded:
        image, radii, depth, alpha = rasterize_gaussians_raw(xyz, screenspace_points, pc._features_dc, pc._features_rest,
                                                             pc._opacity, pc._scaling, pc._rotation, settings)
    else:
        scales = rotations = cov3D = shs = colors = None
        if pipe.compute_cov3D_python:
            cov3D = pc.get_covariance(scaling_modifier)
        else:
            scales, rotations = pc.get_scaling, pc.get_rotation
        if override_color is not None:
            colors = override_color
        elif pipe.convert_SHs_python:
            from utils.sh_utils import eval_sh  # the reference's own helper (only this configuration needs it)
            feats = pc.get_features
            shs_view = feats.transpose(1, 2).view(-1, 3, (pc.max_sh_degree + 1) ** 2)
            dirs = xyz - viewpoint_camera.camera_center.repeat(feats.shape[0], 1)
            colors = torch.clamp_min(eval_sh(pc.active_sh_degree, shs_view, dirs / dirs.norm(dim=1, keepdim=True)) + 0.5, 0.0)
        else:
            shs = pc.get_features
        image, radii, depth, alpha = GaussianRasterizer(raster_settings=settings)(
            means3D=xyz, means2D=screenspace_points, shs=shs, colors_precomp=colors, opacities=pc.get_opacity, scales=scales,
            rotations=rotations, cov3D_precomp=cov3D)
    if min(pc.bg_color.shape) != 0:
        image = image + (1 - alpha) * torch.sigmoid(pc.bg_color)
    return {"render": image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii,
            "depth": depth, "alpha": alpha}
