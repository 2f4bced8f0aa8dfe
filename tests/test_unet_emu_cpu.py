"""The whole tiny denoiser (reference architecture, model_channels 64, one frame, 8 x 8 latent) with every forward operator
executed by the product's CUDA SOURCES on the host: tensor-core GEMMs / im2col convolutions (gemm_tc.cu), the tcgen05
flash attention (attn_tc.cu), GroupNorm / LayerNorm / GEGLU / temporal attention / im2col (nn_kernels.cu) -- through
tests/cuda_emu and tc_emu.h, bound under the same ctypes layer the GPU library sits behind.  Compared with the same network
over tests/fake_nn_lib.py's closed forms (bf16 storage in both) and with the fp32 reference module."""
import ctypes as C
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "guidedvd-3dgs_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "cuda_emu")):
    if p not in sys.path:
        sys.path.insert(0, p)

import unet_ref  # noqa: E402
from test_unet_grad_cpu import _rel, install_fake, needs_ref  # noqa: E402

BF = torch.bfloat16


@pytest.fixture(scope="module")
def emu_libs():
    import build_emu
    import gvd_native

    nn = gvd_native.bind_nn(C.CDLL(build_emu.build("nn_kernels", ["nn_kernels.cu", "nn_fast.cu"])), partial=True)
    gemm = gvd_native.bind_nn(C.CDLL(build_emu.build("gemm_tc")), partial=True)
    attn = gvd_native.bind_nn(C.CDLL(build_emu.build("attn_tc")), partial=True)
    return nn, gemm, attn


@needs_ref
def test_tiny_unet_forward_through_the_emulated_kernels(monkeypatch, emu_libs):
    from vc_b200.unet import UNetB200

    nn, gemm, attn = emu_libs
    torch.manual_seed(0)
    ref, cfg = unet_ref.build_reference_unet(model_channels=64, device="cpu")
    x, cc, ctx, _ = unet_ref.synth_inputs(1, 8, 8, device="cpu")
    xin = torch.cat([x, cc], 1)
    ts, fs = torch.tensor([481]), torch.tensor([10])
    with torch.no_grad():
        y_ref = ref(xin, ts, context=ctx, fs=fs)

    fake = install_fake(monkeypatch, BF)
    taken = {}
    for lib, names in ((nn, ("gvd_groupnorm_tmp_floats", "gvd_groupnorm_cl", "gvd_groupnorm_cl_stats", "gvd_groupnorm_cl_apply",
                             "gvd_groupnorm_cl_keep_stats", "gvd_layernorm", "gvd_geglu", "gvd_softmax_rows", "gvd_im2col3x3_cl",
                             "gvd_im2col_t3_cl", "gvd_temporal_attention", "gvd_upsample2x_cl")),
                       (gemm, ("gvd_gemm_bf16", "gvd_conv_bf16", "gvd_conv_bf16_supported")),
                       (attn, ("gvd_flash_attention",))):
        for name in names:
            fn = getattr(lib, name)

            def counted(*a, _fn=fn, _name=name):
                taken[_name] = taken.get(_name, 0) + 1
                return _fn(*a)
            setattr(fake, name, counted)
    y_emu = UNetB200(ref.state_dict(), device="cpu", **cfg)(xin, ts, ctx, fs=fs)
    assert taken.get("gvd_gemm_bf16", 0) > 100 and taken.get("gvd_flash_attention", 0) > 10 and taken.get("gvd_groupnorm_cl", 0) > 20

    install_fake(monkeypatch, BF)
    y_fake = UNetB200(ref.state_dict(), device="cpu", **cfg)(xin, ts, ctx, fs=fs)
    print(f"tiny U-Net, emulated kernels vs closed forms {_rel(y_emu, y_fake):.3e}; vs fp32 reference {_rel(y_emu, y_ref):.3e} "
          f"(closed forms vs reference {_rel(y_fake, y_ref):.3e})")
    assert y_emu.shape == y_ref.shape and torch.isfinite(y_emu.float()).all()
    assert _rel(y_emu, y_fake) < 3e-2          # both carry bf16 activations; they differ by bf16 roundings along different fp32 sums
    assert _rel(y_emu, y_ref) < max(4e-2, 1.5 * _rel(y_fake, y_ref))


@needs_ref
@pytest.mark.skipif(os.environ.get("GVD_EMU_FULL") != "1", reason="~100 s of host emulation: GVD_EMU_FULL=1 (tools/emu_memcheck.sh sets it)")
def test_tiny_unet_input_gradient_through_the_emulated_kernels(monkeypatch, emu_libs):
    """d<y, g>/dx -- the call `pred_x0.backward(gradient=..., inputs=x)` of ddim_guidance.py:309 -- with the tape's forward AND
    backward operators executed by the CUDA sources on the host: the fused tcgen05 attention adjoint behind its forward's
    row statistic, the GroupNorm / LayerNorm / GEGLU / temporal-attention adjoints, data-gradient GEMMs and col2im."""
    import build_emu
    import gvd_native
    from vc_b200.unet import UNetB200

    nn, gemm, attn = emu_libs
    bwd = gvd_native.bind_nn(C.CDLL(build_emu.build("nn_backward")), partial=True)
    abwd = gvd_native.bind_nn(C.CDLL(build_emu.build("attn_bwd_tc")), partial=True)
    torch.manual_seed(0)
    ref, cfg = unet_ref.build_reference_unet(model_channels=64, device="cpu")
    x, cc, ctx, _ = unet_ref.synth_inputs(1, 8, 8, device="cpu")  # one frame: T > 1 is the business of the operator-level emulation tests
    xin = torch.cat([x, cc], 1)
    ts, fs = torch.tensor([300]), torch.tensor([10])
    g = torch.randn(1, 4, *xin.shape[2:], generator=torch.Generator().manual_seed(1))

    def grad_of(install_emu):
        fake = install_fake(monkeypatch, BF)
        taken = {}
        if install_emu:
            table = ((nn, ("gvd_groupnorm_tmp_floats", "gvd_groupnorm_cl", "gvd_groupnorm_cl_stats", "gvd_groupnorm_cl_apply",
                           "gvd_groupnorm_cl_keep_stats", "gvd_layernorm", "gvd_geglu", "gvd_softmax_rows", "gvd_im2col3x3_cl",
                           "gvd_im2col_t3_cl", "gvd_temporal_attention", "gvd_upsample2x_cl")),
                     (gemm, ("gvd_gemm_bf16", "gvd_conv_bf16", "gvd_conv_bf16_supported")),
                     (attn, ("gvd_flash_attention", "gvd_flash_attention_lse")),
                     (abwd, ("gvd_flash_attention_bwd",)),
                     (bwd, ("gvd_groupnorm_bwd_tmp_bytes", "gvd_groupnorm_cl_bwd", "gvd_layernorm_bwd", "gvd_geglu_bwd", "gvd_softmax_bwd_rows",
                            "gvd_col2im3x3_cl", "gvd_col2im_t3_cl", "gvd_temporal_attention_bwd", "gvd_upsample2x_bwd_cl")))
            for lib, names in table:
                for name in names:
                    def counted(*a, _fn=getattr(lib, name), _name=name):
                        taken[_name] = taken.get(_name, 0) + 1
                        return _fn(*a)
                    setattr(fake, name, counted)
        xo = xin.clone().requires_grad_(True)
        y = UNetB200(ref.state_dict(), device="cpu", **cfg).forward_with_grad(xo, ts, ctx, fs=fs)
        y.backward(gradient=g.to(y.dtype), inputs=[xo])
        return xo.grad.detach().clone(), taken

    g_emu, taken = grad_of(True)
    for name in ("gvd_flash_attention_lse", "gvd_flash_attention_bwd", "gvd_groupnorm_cl_bwd", "gvd_layernorm_bwd", "gvd_geglu_bwd",
                 "gvd_temporal_attention_bwd", "gvd_gemm_bf16"):
        assert taken.get(name, 0) > 0, name
    g_fake, _ = grad_of(False)
    xr = xin.clone().requires_grad_(True)
    ref(xr, ts, context=ctx, fs=fs).backward(gradient=g, inputs=[xr])
    print(f"tiny U-Net input gradient, emulated kernels vs closed forms {_rel(g_emu, g_fake):.3e}; vs autograd(reference, fp32) "
          f"{_rel(g_emu, xr.grad):.3e} (closed forms vs reference {_rel(g_fake, xr.grad):.3e})")
    assert torch.isfinite(g_emu.float()).all()
    assert _rel(g_emu, g_fake) < 6e-2
    assert _rel(g_emu, xr.grad) < max(8e-2, 1.5 * _rel(g_fake, xr.grad))
