"""B200-native drop-in for the reference package `diff_gaussian_rasterization`
(DGR/diff_gaussian_rasterization/__init__.py).  Same public names, signatures and semantics:

    GaussianRasterizationSettings(image_height, image_width, tanfovx, tanfovy, bg, scale_modifier,
                                  viewmatrix, projmatrix, sh_degree, campos, prefiltered, debug, confidence)
    GaussianRasterizer(raster_settings)(means3D, means2D, opacities, shs=None, colors_precomp=None,
                                        scales=None, rotations=None, cov3D_precomp=None)
        -> (color[3,H,W], radii[P] int32, depth[1,H,W], alpha[1,H,W])
    GaussianRasterizer.markVisible(positions) -> bool[P]
    rasterize_gaussians(...)  (functional form)

so `gaussian_renderer.render()` (gaussian_renderer/__init__.py:14,42-101) runs unchanged when this
directory's parent is put on sys.path ahead of the reference submodule.  The compute is the sm_100a
library behind include/gvd_raster.h; there is no fallback path.
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _C
from ._C import gradient_buffer_floats, gradient_views, set_gradient_buffer  # noqa: F401  (B200 extension: caller-owned gradient storage)


def cpu_deep_copy_tuple(input_tuple):
    """Host snapshot of a call's arguments, written to snapshot_{fw,bw}.dump when debug=True fails."""
    return tuple(x.detach().cpu().clone() if torch.is_tensor(x) else x for x in input_tuple)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    # Will a backward come?  (Inside Function.forward grad mode is always off and ctx.needs_input_grad ignores
    # torch.no_grad(), so this is decided here.)  Only matters under the opt-in GVD_SPECULATE=defer mode, where the
    # validation of the speculative buffers then moves to the backward (see _C.DEFER); the default sizes every
    # buffer exactly inside the forward.
    defer = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad
        for t in (means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp))
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings, defer)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, defer=False):
        rs = raster_settings
        args = (rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh,
                rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
        # defer: num_rendered may be a _C.PendingR until the backward (or int()) resolves it.
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args)  # copy before they can be corrupted
            try:
                out = _C.rasterize_gaussians(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            out = _C.rasterize_gaussians(*args, defer=defer)
        num_rendered, color, depth, alpha, radii, geomBuffer, binningBuffer, imgBuffer = out

        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        # a deferred frame may have to be re-rendered by its backward (see there); opacities is the one forward input
        # the backward does not otherwise keep
        ctx.opacities = opacities if isinstance(num_rendered, _C.PendingR) else None
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer, alpha)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
         imgBuffer, alpha) = ctx.saved_tensors
        conf = rs.confidence
        if isinstance(ctx.num_rendered, _C.PendingR):
            try:
                ctx.num_rendered = ctx.num_rendered.resolve()
            except _C.SpeculationOverflow as ex:
                # GVD_SPECULATE=defer only.  The image this frame returned was incomplete and has been consumed; what
                # can still be done is to keep the trainer alive and hand it the gradients of the frame as it should
                # have been: re-render exactly (same inputs, exact buffers) and continue with those buffers.
                import warnings
                warnings.warn(str(ex) + "  Re-rendered exactly for the backward; this step's loss saw the incomplete image.")
                redo = _C.rasterize_gaussians(rs.bg, means3D, colors_precomp, ctx.opacities, scales, rotations, rs.scale_modifier,
                                              cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy,
                                              rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered,
                                              rs.debug, defer=False)
                ctx.num_rendered, _, _, alpha, radii, geomBuffer, binningBuffer, imgBuffer = redo
        if conf is not None and conf.numel() != means3D.size(0):
            raise RuntimeError("confidence must hold one value per Gaussian")
        args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_color, grad_depth, grad_alpha, sh,
                rs.sh_degree, rs.campos, geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer, alpha, rs.debug)
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                out = _C.rasterize_gaussians_backward(*args, confidence=conf)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            out = _C.rasterize_gaussians_backward(*args, confidence=conf)
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh,
         grad_scales, grad_rotations) = out
        # confidence is already applied in-kernel to everything except grad_means2D
        # (reference: diff_gaussian_rasterization/__init__.py:147-157)
        return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales,
                grad_rotations, grad_cov3Ds_precomp, None, None)


class _RasterizeGaussiansRaw(torch.autograd.Function):
    """B200 extension: the rasterizer over the RAW GaussianModel parameters.  gaussian_renderer.render()
    activates them with four torch launches and concatenates the SH tensors into a fresh [P,16,3] (96 MB at 500 k
    Gaussians) on every call (gaussian_renderer/__init__.py:60-87, scene/gaussian_model.py:106-130), and autograd
    replays all of it backwards; here `preprocess` applies exp / normalize / sigmoid and reads `_features_dc` and
    `_features_rest` where they lie, and `gaussian_backward` returns the gradients with respect to the raw parameters
    (two SH gradient tensors, no split)."""

    @staticmethod
    def forward(ctx, means3D, means2D, features_dc, features_rest, opacity_raw, scaling_raw, rotation_raw, raster_settings):
        rs = raster_settings
        absent = torch.Tensor([])
        out = _C.rasterize_gaussians(rs.bg, means3D, absent, opacity_raw, scaling_raw, rotation_raw, rs.scale_modifier, absent,
                                     rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width,
                                     features_dc, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug, sh_rest=features_rest)
        num_rendered, color, depth, alpha, radii, geomBuffer, binningBuffer, imgBuffer = out
        ctx.raster_settings, ctx.num_rendered = rs, num_rendered
        ctx.save_for_backward(means3D, features_dc, features_rest, opacity_raw, scaling_raw, rotation_raw, radii, geomBuffer,
                              binningBuffer, imgBuffer, alpha)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth, alpha

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_alpha):
        rs = ctx.raster_settings
        (means3D, features_dc, features_rest, opacity_raw, scaling_raw, rotation_raw, radii, geomBuffer, binningBuffer,
         imgBuffer, alpha) = ctx.saved_tensors
        conf = rs.confidence
        if conf is not None and conf.numel() != means3D.size(0):
            raise RuntimeError("confidence must hold one value per Gaussian")
        absent = torch.Tensor([])
        out = _C.rasterize_gaussians_backward(rs.bg, means3D, radii, absent, scaling_raw, rotation_raw, rs.scale_modifier, absent,
                                              rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_color, grad_depth,
                                              grad_alpha, features_dc, rs.sh_degree, rs.campos, geomBuffer, ctx.num_rendered,
                                              binningBuffer, imgBuffer, alpha, rs.debug, confidence=conf,
                                              sh_rest=features_rest, opacity_raw=opacity_raw)
        grad_means2D, _, grad_opacity, grad_means3D, _, (grad_dc, grad_rest), grad_scales, grad_rotations = out
        return (grad_means3D, grad_means2D, grad_dc, grad_rest, grad_opacity.view_as(opacity_raw), grad_scales, grad_rotations,
                None)


def rasterize_gaussians_raw(means3D, means2D, features_dc, features_rest, opacity_raw, scaling_raw, rotation_raw,
                            raster_settings):
    """-> (color[3,H,W], radii[P], depth[1,H,W], alpha[1,H,W]) from the un-activated parameters: equal to
    GaussianRasterizer(raster_settings)(means3D, means2D, sigmoid(opacity_raw), shs=cat(features_dc, features_rest),
    scales=exp(scaling_raw), rotations=normalize(rotation_raw)) up to the rounding of the normalisation."""
    return _RasterizeGaussiansRaw.apply(means3D, means2D, features_dc, features_rest, opacity_raw, scaling_raw, rotation_raw,
                                        raster_settings)


def rasterize_views(settings_list, means3D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                    cov3D_precomp=None):
    """B200 extension: forward-only render of MANY views of one Gaussian set -- the 25-75 trajectory
    poses train_guidedvd.py renders one `easy_renderer.render` call at a time per diffusion round
    (train_guidedvd.py:157-165,521-527; utils/easy_renderer.py:59-66).

    settings_list: one GaussianRasterizationSettings per view (same Gaussians, any cameras / image sizes).
    -> list of (color[3,H,W], radii[P], depth[1,H,W], alpha[1,H,W]), identical to calling GaussianRasterizer per view.

    All views are queued back to back without a host wait in between (instance buffers sized from history); the counts
    every frame stored in pinned memory are validated once at the end and a frame that outgrew its buffers is rendered
    again on the exact path BEFORE anything is returned -- so the results are always valid, and the GPU runs the batch as
    one uninterrupted pipeline instead of idling while the host prepares each next view."""
    if (shs is None) == (colors_precomp is None):
        raise Exception('Please provide excatly one of either SHs or precomputed colors!')
    pair_given = scales is not None or rotations is not None
    pair_complete = scales is not None and rotations is not None
    if (cov3D_precomp is None and not pair_complete) or (cov3D_precomp is not None and pair_given):
        raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
    absent = torch.Tensor([])
    sh, col, sc, rot, cov = [absent if t is None else t.detach() for t in (shs, colors_precomp, scales, rotations, cov3D_precomp)]
    means3D, opacities = means3D.detach(), opacities.detach()

    def args_of(rs):
        return (rs.bg, means3D, col, opacities, sc, rot, rs.scale_modifier, cov, rs.viewmatrix, rs.projmatrix, rs.tanfovx,
                rs.tanfovy, rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)

    with torch.no_grad():
        outs = [_C.rasterize_gaussians(*args_of(rs), force_defer=not rs.debug) for rs in settings_list]
        results = []
        for rs, out in zip(settings_list, outs):
            if isinstance(out[0], _C.PendingR):
                try:
                    out[0].resolve()
                except _C.SpeculationOverflow:
                    out = _C.rasterize_gaussians(*args_of(rs))  # this frame outgrew its buffers: render it again, exactly
            results.append((out[1], out[4], out[2], out[3]))
    return results


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    confidence: torch.Tensor


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        # Mark visible points (based on frustum culling for camera) with a boolean
        with torch.no_grad():
            rs = self.raster_settings
            visible = _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        # argument contract of the reference module (DGR/diff_gaussian_rasterization/__init__.py:192-212):
        # exactly one colour source and exactly one covariance source; messages kept verbatim because
        # callers/tests match on them.
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        pair_given = scales is not None or rotations is not None
        pair_complete = scales is not None and rotations is not None
        if (cov3D_precomp is None and not pair_complete) or (cov3D_precomp is not None and pair_given):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        absent = torch.Tensor([])  # empty CPU tensor == "not provided" (NULL at the C ABI)
        opt = [absent if t is None else t for t in (shs, colors_precomp, scales, rotations, cov3D_precomp)]
        return rasterize_gaussians(means3D, means2D, opt[0], opt[1], opacities, opt[2], opt[3], opt[4],
                                   self.raster_settings)
