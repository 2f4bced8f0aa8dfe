"""B200-native VAE decoder of the ViewCrafter latent-diffusion model, forward and input-gradient.

Same network as third_party/ViewCrafter/lvdm/modules/networks/ae_modules.py::Decoder (:466-579) behind
AutoencoderKL.decode (lvdm/models/autoencoder.py:104-107: post_quant_conv, then the decoder) and LatentDiffusion.
decode_core (lvdm/models/ddpm3d.py:646-667: 1/scale_factor, one frame at a time) for the configuration the reference
runs (configs/inference_pvd_1024.yaml:66-87: ch 128, ch_mult 1-2-4-4, 2 res blocks per level, no attention besides the
single-head 512-channel AttnBlock of the middle, z_channels 4).  The guided sampler decodes every frame of pred_x0 WITH
a graph and differentiates the guidance loss back to the latent (ddim_guidance.py:282-302): 25 decoder forward +
backward passes per guided DDIM step, ~40 % of its arithmetic.

Parameters come from a reference state_dict (`first_stage_model.state_dict()`: keys `post_quant_conv.*`, `decoder.*`),
repacked once like the U-Net's (vc_b200.unet._P).  Activations are channels-last bf16 [frames, pixels, channels];
every operator is a launch of the sm_100a library (vc_b200.ops) and, with the tape on, records its input-gradient
(vc_b200.grad).  Rounding points follow torch.autocast(bfloat16) over the reference modules: nn.GroupNorm and the swish
run in fp32 with one rounding at the next convolution's input (groupnorm mode 2), the AttnBlock's logits are rounded
to bf16 before and after the scale, the softmax runs in fp32.

Frames are a batch dimension for every layer (GroupNorm statistics are per frame), so any number of frames can be
decoded in one call; the reference loops over single frames to bound its memory.
"""
import torch

from . import ops
from .unet import _P

EPS = 1e-6  # ae_modules.py:15


def _pad_cols(w, k_to):
    """[N, K] -> [N, k_to] zero padded (a 4-channel latent is carried as 8 channels: 16-byte rows for TMA / im2col)."""
    out = torch.zeros(w.shape[0], k_to, dtype=w.dtype, device=w.device)
    out[:, :w.shape[1]] = w
    return out


def _pad_rows(w, b, n_to):
    wo = torch.zeros(n_to, w.shape[1], dtype=w.dtype, device=w.device)
    wo[:w.shape[0]] = w
    bo = torch.zeros(n_to, dtype=b.dtype, device=b.device)
    bo[:b.shape[0]] = b
    return wo, bo


class _ResnetBlock:
    """ae_modules.py:150-212 with temb_channels = 0, dropout 0."""

    def __init__(self, p, pre):
        self.n1, self.c1 = p.norm(pre + ".norm1"), p.conv3x3(pre + ".conv1")
        self.n2, self.c2 = p.norm(pre + ".norm2"), p.conv3x3(pre + ".conv2")
        self.skip = p.lin(pre + ".nin_shortcut") if p.has(pre + ".nin_shortcut.weight") else None
        if p.has(pre + ".conv_shortcut.weight"):
            raise NotImplementedError("conv_shortcut ResnetBlocks are not part of the ViewCrafter VAE")

    def __call__(self, x, F, H, W):
        h = ops.groupnorm(x, *self.n1, F, H * W, eps=EPS, silu=2)
        h, _, _ = ops.conv3x3(h, F, H, W, *self.c1)
        h = ops.groupnorm(h, *self.n2, F, H * W, eps=EPS, silu=2)
        skip = x if self.skip is None else ops.linear(x, *self.skip)
        h, _, _ = ops.conv3x3(h, F, H, W, *self.c2, residual=skip)
        return h


class _AttnBlock:
    """ae_modules.py:25-78: one head over all channels, softmax(q k^T c^-1/2) v, 1x1 projections."""

    def __init__(self, p, pre):
        self.norm = p.norm(pre + ".norm")
        self.q, self.k, self.v, self.proj = (p.lin(pre + n) for n in (".q", ".k", ".v", ".proj_out"))

    def __call__(self, x, F, S):
        Cc = x.shape[-1]
        n = ops.groupnorm(x, *self.norm, F, S, eps=EPS, silu=0)
        q, k, v = ops.linear(n, *self.q), ops.linear(n, *self.k), ops.linear(n, *self.v)
        o = ops.attention(q, k, v, F, S, S, 1, int(Cc) ** (-0.5), head_dim=Cc)
        return ops.linear(o, *self.proj, residual=x)


class DecoderB200:
    def __init__(self, state_dict, device="cuda", scale_factor=0.18215, ch_mult=(1, 2, 4, 4), num_res_blocks=2):
        p = _P(state_dict, device)
        self.dev, self.scale_factor = device, float(scale_factor)
        w, b = p.lin("post_quant_conv")                       # [4, 4] -> [8, 8]: the latent travels as 8 channels
        self.zc = w.shape[1]
        self.post_quant = _pad_rows(_pad_cols(w, 8), b, 8)
        w = state_dict["decoder.conv_in.weight"].detach().to(device)   # [C, 4, 3, 3] -> [C, (ky, kx, 8)]
        w8 = torch.zeros(w.shape[0], 8, 3, 3, dtype=w.dtype, device=device)
        w8[:, :w.shape[1]] = w
        self.conv_in = (w8.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(ops.BF16).contiguous(),
                        p.f32("decoder.conv_in.bias"))
        self.mid = (_ResnetBlock(p, "decoder.mid.block_1"), _AttnBlock(p, "decoder.mid.attn_1"), _ResnetBlock(p, "decoder.mid.block_2"))
        self.up = []
        for level in range(len(ch_mult)):
            blocks = [_ResnetBlock(p, f"decoder.up.{level}.block.{i}") for i in range(num_res_blocks + 1)]
            if p.has(f"decoder.up.{level}.attn.0.norm.weight"):
                raise NotImplementedError("attn_resolutions is empty in the ViewCrafter VAE")
            upsample = p.conv3x3(f"decoder.up.{level}.upsample.conv") if level != 0 else None
            self.up.append((blocks, upsample))
        self.norm_out = p.norm("decoder.norm_out")
        w, b = p.conv3x3("decoder.conv_out")                  # 3 output channels -> 8 rows: 16-byte output rows
        self.out_ch = w.shape[0]
        self.conv_out = _pad_rows(w, b, 8)

    def _decode(self, z):
        """z [F, 4, h, w] (latent frames, fp32) -> [F, 3, 8h, 8w] fp32."""
        F, zc, hh, ww = z.shape
        x = (z.float() * (1.0 / self.scale_factor)).permute(0, 2, 3, 1).reshape(F, hh * ww, zc)
        x = torch.nn.functional.pad(x, (0, 8 - zc)).to(ops.BF16).contiguous()
        x = ops.linear(x, *self.post_quant)
        H, W = hh, ww
        h, _, _ = ops.conv3x3(x, F, H, W, *self.conv_in)
        h = self.mid[0](h, F, H, W)
        h = self.mid[1](h, F, H * W)
        h = self.mid[2](h, F, H, W)
        for level in reversed(range(len(self.up))):
            blocks, upsample = self.up[level]
            for blk in blocks:
                h = blk(h, F, H, W)
            if upsample is not None:
                h, H, W = ops.conv3x3(h, F, H, W, *upsample, upsample=True)
        h = ops.groupnorm(h, *self.norm_out, F, H * W, eps=EPS, silu=2)
        y, _, _ = ops.conv3x3(h, F, H, W, *self.conv_out)
        return y.view(F, H, W, 8)[..., :self.out_ch].permute(0, 3, 1, 2).float()

    def decode(self, z):
        """Inference (ddpm3d.py:669-671 `decode_first_stage`): no graph."""
        with torch.no_grad():
            return self._video(z)

    def differentiable_decode(self, z):
        """ddpm3d.py:673-675 `differentiable_decode_first_stage`: the tape is on; z.requires_grad -> d(image)/dz."""
        with torch.enable_grad():
            return self._video(z)

    def _video(self, z):
        if z.dim() == 5:  # [b, c, t, h, w] -> frames as batch and back (decode_core's rearrange)
            b, c, t, hh, ww = z.shape
            y = self._decode(z.permute(0, 2, 1, 3, 4).reshape(b * t, c, hh, ww))
            return y.view(b, t, *y.shape[1:]).permute(0, 2, 1, 3, 4)
        return self._decode(z)

    __call__ = decode


class EncoderB200:
    """VAE encoder (ae_modules.py:363-464) + quant_conv (autoencoder.py:97-102): frames [F, 3, H, W] in [-1, 1] -> the
    posterior's moments [F, 2 * z_channels, H/8, W/8] (mean, log-variance) -- what `get_latent_z`
    (VC/utils_vc/diffusion_utils.py:111-116) runs over the 25 conditioning frames before every diffusion round.
    Inference only.  The 3 image channels travel as 8 (16-byte rows); the Downsample layers use the right/bottom-padded
    stride-2 im2col (gvd_im2col3x3_down_cl)."""

    def __init__(self, state_dict, device="cuda", ch_mult=(1, 2, 4, 4), num_res_blocks=2):
        p = _P(state_dict, device)
        self.dev = device
        w = state_dict["encoder.conv_in.weight"].detach().to(device)   # [C, 3, 3, 3] -> [C, (ky, kx, 8)]
        self.in_ch = w.shape[1]
        w8 = torch.zeros(w.shape[0], 8, 3, 3, dtype=w.dtype, device=device)
        w8[:, :w.shape[1]] = w
        self.conv_in = (w8.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(ops.BF16).contiguous(), p.f32("encoder.conv_in.bias"))
        self.down = []
        for level in range(len(ch_mult)):
            blocks = [_ResnetBlock(p, f"encoder.down.{level}.block.{i}") for i in range(num_res_blocks)]
            if p.has(f"encoder.down.{level}.attn.0.norm.weight"):
                raise NotImplementedError("attn_resolutions is empty in the ViewCrafter VAE")
            ds = p.conv3x3(f"encoder.down.{level}.downsample.conv") if level != len(ch_mult) - 1 else None
            self.down.append((blocks, ds))
        self.mid = (_ResnetBlock(p, "encoder.mid.block_1"), _AttnBlock(p, "encoder.mid.attn_1"), _ResnetBlock(p, "encoder.mid.block_2"))
        self.norm_out = p.norm("encoder.norm_out")
        self.conv_out = p.conv3x3("encoder.conv_out")          # 2 * z_channels = 8 outputs
        self.quant = p.lin("quant_conv")                       # 1x1, 8 -> 8

    @torch.no_grad()
    def moments(self, x):
        F, c, H, W = x.shape
        h = x.float().permute(0, 2, 3, 1).reshape(F, H * W, c)
        h = torch.nn.functional.pad(h, (0, 8 - c)).to(ops.BF16).contiguous()
        h, _, _ = ops.conv3x3(h, F, H, W, *self.conv_in)
        for blocks, ds in self.down:
            for blk in blocks:
                h = blk(h, F, H, W)
            if ds is not None:
                h, H, W = ops.conv3x3_down(h, F, H, W, *ds)
        h = self.mid[0](h, F, H, W)
        h = self.mid[1](h, F, H * W)
        h = self.mid[2](h, F, H, W)
        h = ops.groupnorm(h, *self.norm_out, F, H * W, eps=EPS, silu=2)
        h, _, _ = ops.conv3x3(h, F, H, W, *self.conv_out)
        m = ops.linear(h, *self.quant)
        return m.view(F, H, W, -1).permute(0, 3, 1, 2).float()

    def encode(self, x, scale_factor=0.18215, noise=None, sample=True):
        """`encode_first_stage` (ddpm3d.py:620-644 + :611-618): scale_factor * posterior.sample() (or the mode).
        x [F, 3, H, W] or a video [b, 3, t, H, W]; noise: optional N(0, 1) of the latent's shape."""
        video = x.dim() == 5
        if video:
            b, c, t, H, W = x.shape
            x = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, H, W)
        m = self.moments(x)
        zc = m.shape[1] // 2
        mean, logvar = m[:, :zc], m[:, zc:].clamp(-30.0, 20.0)   # lvdm/distributions.py: DiagonalGaussianDistribution
        z = mean
        if sample:
            if noise is None:
                noise = torch.randn(mean.shape, device=mean.device)
            z = mean + torch.exp(0.5 * logvar) * noise
        z = scale_factor * z
        return z.view(b, t, *z.shape[1:]).permute(0, 2, 1, 3, 4) if video else z
