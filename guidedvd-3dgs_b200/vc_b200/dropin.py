"""Drop-in boundary for the diffusion half: swap the reference's U-Net module for the B200-native forward.

The reference pipeline (utils/viewcrafter_wrapper.py:550-573 -> VC/viewcrafter.py:92-112 ->
VC/utils_vc/diffusion_utils.py:118-223 -> DDIMSampler.p_sample_ddim -> model.apply_model ->
DiffusionWrapper.forward, lvdm/models/ddpm3d.py:1426-1443) reaches the denoiser as
`model.model.diffusion_model(xc, t, context=cc, **kwargs)`.  `B200UNet` is an nn.Module with that call signature whose
forward runs vc_b200.unet.UNetB200; `replace_unet(latent_diffusion)` installs it, so `ViewCrafterWrapper`,
`ViewCrafter.run_diffusion`, `image_guided_synthesis` and the reference samplers run unchanged on top of it.

Scope: the plain DDIM path (`no_guidance=True`, and every U-Net call made under torch.no_grad()) always runs native.
The guided sampler (lvdm/models/samplers/ddim_guidance.py:259-337) differentiates through the U-Net with respect to
the latent: such a call runs `UNetB200.forward_with_grad` (vc_b200.grad: the same forward kernels with the tape on,
input-gradient kernels in the backward).  GVD_GUIDED_NATIVE=0 hands calls that need a graph back to the reference
module the wrapper keeps (A/B switch; the native path is the default since its hardware parity run in round 2,
profiles/r02_*guided*).  Calls that want PARAMETER gradients (nobody on this path: the guided sampler sets
requires_grad on the modules but only ever asks for `inputs=x`) go to the reference module unless the latent itself
requires grad.

Zero-edit activation: vc_b200.autoinstall (loaded by guidedvd-3dgs_b200/sitecustomize.py when that directory is on
PYTHONPATH) wraps `ViewCrafter.setup_diffusion` (third_party/ViewCrafter/viewcrafter.py:315-335) so the three
replace_* calls below run right after the reference has built and loaded its model.
"""
import os

import torch


def guided_native():
    """Native input-gradient path on (default) or off (GVD_GUIDED_NATIVE=0)."""
    return os.environ.get("GVD_GUIDED_NATIVE", "1") != "0"


import torch.nn as nn

from .unet import UNetB200


class B200UNet(nn.Module):
    def __init__(self, reference_unet, device=None):
        super().__init__()
        self.reference = reference_unet  # kept for the autograd (guided) path and for state_dict round trips
        dev = device or next(reference_unet.parameters()).device
        cfg = dict(in_channels=reference_unet.in_channels, model_channels=reference_unet.model_channels,
                   out_channels=reference_unet.out_channels, num_res_blocks=reference_unet.num_res_blocks,
                   attention_resolutions=tuple(reference_unet.attention_resolutions),
                   channel_mult=tuple(reference_unet.channel_mult), temporal_conv=True,
                   addition_attention=bool(reference_unet.addition_attention),
                   fs_condition=bool(reference_unet.fs_condition), default_fs=int(reference_unet.default_fs))
        self.native = UNetB200(reference_unet.state_dict(), device=dev, **cfg)

    def forward(self, x, timesteps, context=None, features_adapter=None, fs=None, **kwargs):
        if features_adapter is not None:
            return self.reference(x, timesteps, context=context, features_adapter=features_adapter, fs=fs, **kwargs)
        if torch.is_grad_enabled() and x.requires_grad:
            if guided_native():
                return self.native.forward_with_grad(x.float(), timesteps, context.float(), fs=fs).to(x.dtype)
            return self.reference(x, timesteps, context=context, fs=fs, **kwargs)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.reference.parameters()):
            return self.reference(x, timesteps, context=context, fs=fs, **kwargs)
        return self.native(x.float(), timesteps, context.float(), fs=fs).to(x.dtype)


def replace_unet(latent_diffusion):
    """latent_diffusion.model.diffusion_model <- B200UNet(original).  Returns the wrapper."""
    wrapper = latent_diffusion.model
    unet = wrapper.diffusion_model
    if isinstance(unet, B200UNet):
        return unet
    new = B200UNet(unet)
    wrapper.diffusion_model = new
    return new


def replace_first_stage_decoder(latent_diffusion):
    """latent_diffusion.first_stage_model.decode <- the B200-native VAE decoder (vc_b200.vae.DecoderB200).

    `decode_core` (lvdm/models/ddpm3d.py:646-667) keeps its per-frame loop and its 1/scale_factor; each
    `first_stage_model.decode(frame_z)` (autoencoder.py:104-107: post_quant_conv + Decoder) lands here.  Same routing
    rule as B200UNet: no graph wanted -> native forward; graph with respect to the latent wanted (the guided sampler,
    ddim_guidance.py:288) -> native forward with the tape on (GVD_GUIDED_NATIVE=0: the reference module).
    Returns the DecoderB200."""
    from .vae import DecoderB200

    fs = latent_diffusion.first_stage_model
    if getattr(fs, "_b200_decoder", None) is not None:
        return fs._b200_decoder
    dev = next(fs.parameters()).device
    native = DecoderB200(fs.state_dict(), device=dev, scale_factor=1.0)
    reference_decode = fs.decode

    def decode(z, **kwargs):
        if torch.is_grad_enabled() and z.requires_grad:
            if guided_native():
                return native.differentiable_decode(z.float()).to(z.dtype)
            return reference_decode(z, **kwargs)
        if torch.is_grad_enabled() and any(p.requires_grad for p in fs.parameters()):
            return reference_decode(z, **kwargs)
        return native.decode(z.float()).to(z.dtype)

    fs.decode = decode
    fs._b200_decoder = native
    return native


def replace_first_stage_encoder(latent_diffusion):
    """latent_diffusion.first_stage_model.encode <- the B200-native VAE encoder (vc_b200.vae.EncoderB200).

    `encode_first_stage` (lvdm/models/ddpm3d.py:620-644, reached from `get_latent_z`, VC/utils_vc/diffusion_utils.py:111-116,
    once per diffusion round for the 25 conditioning frames) keeps its own posterior object: the hook computes the
    moments natively and hands them to the reference's `DiagonalGaussianDistribution` (lvdm/distributions.py), so
    sampling, scale_factor and the per-frame / batched switch stay the reference's.  Returns the EncoderB200."""
    from lvdm.distributions import DiagonalGaussianDistribution  # the reference package this drop-in lives next to

    from .vae import EncoderB200

    fs = latent_diffusion.first_stage_model
    if getattr(fs, "_b200_encoder", None) is not None:
        return fs._b200_encoder
    dev = next(fs.parameters()).device
    native = EncoderB200(fs.state_dict(), device=dev)
    reference_encode = fs.encode

    def encode(x, **kwargs):
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in fs.parameters())):
            return reference_encode(x, **kwargs)  # nobody on this path differentiates the encoder
        return DiagonalGaussianDistribution(native.moments(x.float()).to(x.dtype))

    fs.encode = encode
    fs._b200_encoder = native
    return native
