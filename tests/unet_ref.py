"""Helpers to run the REFERENCE denoiser (oracle/_ref/ViewCrafter, installed by oracle/build_ref.py vc) beside ours.
Test infrastructure."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_VC = os.path.join(ROOT, "oracle", "_ref", "ViewCrafter")

FULL_CFG = dict(in_channels=8, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
                channel_mult=[1, 2, 4, 4], dropout=0.1, num_head_channels=64, transformer_depth=1, context_dim=1024,
                use_linear=True, use_checkpoint=False, temporal_conv=True, temporal_attention=True,
                temporal_selfatt_only=True, use_relative_position=False, use_causal_attention=False, temporal_length=16,
                addition_attention=True, image_cross_attention=True, default_fs=10, fs_condition=True)


def ref_available():
    return os.path.exists(os.path.join(REF_VC, "lvdm", "modules", "networks", "openaimodel3d.py"))


def build_reference_unet(model_channels=320, seed=1234, device="cuda"):
    """Reference UNetModel (configs/inference_pvd_1024.yaml:33-64) with EVERY parameter re-drawn from a seeded generator:
    a freshly constructed model outputs exactly 0 (zero_module on all output projections)."""
    if REF_VC not in sys.path:
        sys.path.insert(0, REF_VC)
    from lvdm.modules.networks.openaimodel3d import UNetModel

    cfg = dict(FULL_CFG, model_channels=model_channels)
    m = UNetModel(**cfg)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if p.dim() >= 2:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(fan_in))
            elif name.endswith("weight"):
                p.copy_(1.0 + 0.02 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
    return m.to(device).eval(), cfg


def synth_inputs(t, h, w, seed=20260003, device="cuda"):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, 4, t, h, w, generator=g)
    c_concat = 0.18215 * torch.randn(1, 4, t, h, w, generator=g)
    ctx = torch.randn(1, 333, 1024, generator=g)
    ctx_uc = torch.randn(1, 333, 1024, generator=g)
    return x.to(device), c_concat.to(device), ctx.to(device), ctx_uc.to(device)
