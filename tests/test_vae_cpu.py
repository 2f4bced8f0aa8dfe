"""Host-side parity of the B200-native VAE decoder (vc_b200.vae.DecoderB200), forward and d(image)/d(latent), against
the REFERENCE Decoder (oracle/_ref/ViewCrafter/lvdm/modules/networks/ae_modules.py:466-579) + post_quant_conv run in fp32
on the CPU and torch.autograd over it.  vc_b200 reaches "the library" through the pointer-level stand-in of
tests/fake_nn_lib.py (see tests/test_unet_grad_cpu.py); the kernels underneath are the ones the U-Net tests cover."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "guidedvd-3dgs_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import unet_ref  # noqa: E402
from test_unet_grad_cpu import _rel, install_fake  # noqa: E402

HAVE = os.path.exists(os.path.join(unet_ref.REF_VC, "lvdm", "modules", "networks", "ae_modules.py"))
pytestmark = pytest.mark.skipif(not HAVE, reason="oracle/_ref/ViewCrafter/.../ae_modules.py not installed (python oracle/build_ref.py vc)")

SCALE = 0.18215


class RefFirstStage(torch.nn.Module):
    """AutoencoderKL.decode (autoencoder.py:104-107) + decode_core's scaling (ddpm3d.py:655) around the reference Decoder,
    with every parameter re-drawn (a fresh GroupNorm-heavy net with default init is too tame a test)."""

    def __init__(self, ch=32, seed=5):
        super().__init__()
        if unet_ref.REF_VC not in sys.path:
            sys.path.insert(0, unet_ref.REF_VC)
        from lvdm.modules.networks.ae_modules import Decoder

        self.decoder = Decoder(ch=ch, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, attn_resolutions=[], dropout=0.0,
                               in_channels=3, resolution=256, z_channels=4, double_z=True)
        self.post_quant_conv = torch.nn.Conv2d(4, 4, 1)
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for name, p in self.named_parameters():
                if p.dim() >= 2:
                    p.copy_(torch.randn(p.shape, generator=g) / p[0].numel() ** 0.5)
                elif name.endswith("weight"):
                    p.copy_(1.0 + 0.05 * torch.randn(p.shape, generator=g))
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=g))

    def forward(self, z):
        return self.decoder(self.post_quant_conv(z / SCALE))


@pytest.fixture(scope="module")
def ref_vae():
    return RefFirstStage().eval()


def test_decoder_forward_and_latent_gradient(monkeypatch, ref_vae):
    from vc_b200.vae import DecoderB200

    fake = install_fake(monkeypatch)
    ours = DecoderB200(ref_vae.state_dict(), device="cpu", scale_factor=SCALE)
    g = torch.Generator().manual_seed(1)
    z = torch.randn(2, 4, 6, 5, generator=g) * SCALE * 3
    with torch.no_grad():
        y_ref = ref_vae(z)
    y = ours.decode(z)
    assert y.shape == y_ref.shape == (2, 3, 48, 40) and not y.requires_grad
    assert _rel(y, y_ref) < 2e-5

    cot = torch.randn(y.shape, generator=g)
    zr = z.clone().requires_grad_(True)
    ref_vae(zr).backward(cot)
    zo = z.clone().requires_grad_(True)
    yo = ours.differentiable_decode(zo)
    assert yo.requires_grad
    yo.backward(cot)
    err = _rel(zo.grad, zr.grad)
    print(f"decoder latent-gradient rel L2 vs autograd(reference): {err:.3e}")
    assert err < 1e-4
    for name in ("groupnorm_bwd", "softmax_bwd", "col2im3x3"):
        assert fake.calls.get(name, 0) > 0, name


def test_decoder_video_layout_matches_per_frame_loop(monkeypatch, ref_vae):
    """decode_core (ddpm3d.py:646-667) decodes '(b t) c h w' one frame at a time; frames are a batch dimension here."""
    from vc_b200.vae import DecoderB200

    install_fake(monkeypatch)
    ours = DecoderB200(ref_vae.state_dict(), device="cpu", scale_factor=SCALE)
    z = torch.randn(1, 4, 3, 4, 4, generator=torch.Generator().manual_seed(2)) * SCALE
    y = ours.decode(z)
    assert y.shape == (1, 3, 3, 32, 32)
    with torch.no_grad():
        per_frame = torch.stack([ref_vae(z[:, :, f])[0] for f in range(3)], dim=1)
    assert _rel(y[0], per_frame) < 2e-5


def test_decoder_bf16_rounding_points(monkeypatch, ref_vae):
    """bf16 storage in the stand-in: output and latent gradient as close to fp32 as the reference under bf16 autocast."""
    from vc_b200.vae import DecoderB200

    install_fake(monkeypatch, torch.bfloat16)
    ours = DecoderB200(ref_vae.state_dict(), device="cpu", scale_factor=SCALE)
    g = torch.Generator().manual_seed(3)
    z = torch.randn(1, 4, 6, 6, generator=g) * SCALE * 3
    cot = torch.randn(1, 3, 48, 48, generator=g)
    res = {}
    for name in ("fp32", "bf16"):
        zr = z.clone().requires_grad_(True)
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=(name == "bf16")):
            y = ref_vae(zr)
        y.float().backward(cot)
        res[name] = (y.detach().float(), zr.grad)
    zo = z.clone().requires_grad_(True)
    yo = ours.differentiable_decode(zo)
    yo.backward(cot)
    e_y, e_y_ref = _rel(yo, res["fp32"][0]), _rel(res["bf16"][0], res["fp32"][0])
    e_g, e_g_ref = _rel(zo.grad, res["fp32"][1]), _rel(res["bf16"][1], res["fp32"][1])
    print(f"bf16: image {e_y:.2e} (reference autocast {e_y_ref:.2e}); latent gradient {e_g:.2e} (reference autocast {e_g_ref:.2e})")
    assert e_y <= 1.25 * e_y_ref + 2e-3 and e_g <= 1.25 * e_g_ref + 5e-3


def test_first_stage_dropin_routing(monkeypatch, ref_vae):
    """vc_b200.dropin.replace_first_stage_decoder: `first_stage_model.decode` (autoencoder.py:104-107) runs native without
    a graph, forwards to the reference module when a graph is wanted, and runs native WITH the tape under
    GVD_GUIDED_NATIVE=1 -- the three routes give the same image / latent gradient."""
    from vc_b200.dropin import replace_first_stage_decoder

    install_fake(monkeypatch)

    class FirstStage(torch.nn.Module):  # AutoencoderKL's decode half
        def __init__(self, src):
            super().__init__()
            self.decoder, self.post_quant_conv = src.decoder, src.post_quant_conv

        def decode(self, z, **kwargs):
            return self.decoder(self.post_quant_conv(z))

    class LD:
        pass

    ld = LD()
    ld.first_stage_model = FirstStage(ref_vae)
    fs_ref = FirstStage(ref_vae)
    native = replace_first_stage_decoder(ld)
    assert replace_first_stage_decoder(ld) is native
    z = torch.randn(1, 4, 5, 4, generator=torch.Generator().manual_seed(8)) * 2
    with torch.no_grad():
        y_ref = fs_ref.decode(z)
        y = ld.first_stage_model.decode(z)
    assert _rel(y, y_ref) < 2e-5
    cot = torch.randn(y.shape, generator=torch.Generator().manual_seed(9))
    grads = []
    for env in ("0", "1"):
        monkeypatch.setenv("GVD_GUIDED_NATIVE", env)
        zg = z.clone().requires_grad_(True)
        yg = ld.first_stage_model.decode(zg)
        yg.backward(cot)
        grads.append(zg.grad)
        assert ("ConvolutionBackward" in type(yg.grad_fn).__name__) == (env == "0")
    assert _rel(grads[1], grads[0]) < 1e-4


class RefEncoderStage(torch.nn.Module):
    """AutoencoderKL.encode's arithmetic (autoencoder.py:97-102) around the reference Encoder, parameters re-drawn."""

    def __init__(self, ch=32, seed=6):
        super().__init__()
        if unet_ref.REF_VC not in sys.path:
            sys.path.insert(0, unet_ref.REF_VC)
        from lvdm.modules.networks.ae_modules import Encoder

        self.encoder = Encoder(ch=ch, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, attn_resolutions=[], dropout=0.0,
                               in_channels=3, resolution=256, z_channels=4, double_z=True)
        self.quant_conv = torch.nn.Conv2d(8, 8, 1)
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for name, p in self.named_parameters():
                if p.dim() >= 2:
                    p.copy_(torch.randn(p.shape, generator=g) / p[0].numel() ** 0.5)
                elif name.endswith("weight"):
                    p.copy_(1.0 + 0.05 * torch.randn(p.shape, generator=g))
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=g))

    def forward(self, x):
        return self.quant_conv(self.encoder(x))


@pytest.mark.parametrize("H,W", [(32, 48), (24, 40)])
def test_encoder_moments_and_latent(monkeypatch, H, W):
    """EncoderB200 (incl. the right/bottom-padded stride-2 Downsample) vs the reference Encoder + quant_conv in fp32, and the
    sampled latent vs DiagonalGaussianDistribution's arithmetic with the same noise."""
    from vc_b200.vae import EncoderB200

    fake = install_fake(monkeypatch)
    ref = RefEncoderStage().eval()
    ours = EncoderB200(ref.state_dict(), device="cpu")
    g = torch.Generator().manual_seed(4)
    x = torch.rand(2, 3, H, W, generator=g) * 2 - 1
    with torch.no_grad():
        m_ref = ref(x)
    m = ours.moments(x)
    assert m.shape == m_ref.shape == (2, 8, H // 8, W // 8)
    assert _rel(m, m_ref) < 2e-5 and fake.calls["im2col3x3_down"] == 3
    noise = torch.randn(2, 4, H // 8, W // 8, generator=g)
    z = ours.encode(x, scale_factor=SCALE, noise=noise)
    mean, logvar = m_ref[:, :4], m_ref[:, 4:].clamp(-30.0, 20.0)
    assert _rel(z, SCALE * (mean + torch.exp(0.5 * logvar) * noise)) < 2e-5
    zv = ours.encode(x.view(1, 2, 3, H, W).permute(0, 2, 1, 3, 4), scale_factor=SCALE, sample=False)
    assert zv.shape == (1, 4, 2, H // 8, W // 8) and _rel(zv[0].permute(1, 0, 2, 3), SCALE * mean) < 2e-5
