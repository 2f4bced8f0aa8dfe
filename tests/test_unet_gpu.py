"""GPU parity of the B200-native denoiser against the reference UNetModel run on the same GPU under
torch.autocast(bfloat16) with identical seeded weights and inputs.

Bar: <= 1e-3 relative on the denoiser output is the north-star; in bf16 two implementations that round at the same
points but accumulate in different orders (tcgen05 vs cuBLAS/cuDNN) differ by a few bf16 ulps per layer, so the test
states what is measured: relative L2 error of the output <= 1e-2, and it must not exceed 2x the reference's own
sensitivity to accumulation order (reference with TF32-free fp32 matmuls vs the autocast run)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.mark.parametrize("mc,t,h,w", [(64, 5, 16, 16), (64, 3, 16, 24)])
def test_unet_small_vs_reference(mc, t, h, w):
    import unet_ref
    from vc_b200.unet import UNetB200

    if not unet_ref.ref_available():
        pytest.skip("oracle/_ref/ViewCrafter not installed")
    ref, cfg = unet_ref.build_reference_unet(model_channels=mc)
    ours = UNetB200(ref.state_dict(), device="cuda", **cfg)
    x, cc, ctx, _ = unet_ref.synth_inputs(t, h, w)
    xin = torch.cat([x, cc], 1)
    ts = torch.tensor([481], device="cuda")
    fs = torch.tensor([10], device="cuda")
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y_ref = ref(xin, ts, context=ctx, fs=fs)
    with torch.no_grad():
        y_fp32 = ref(xin, ts, context=ctx, fs=fs)
    y = ours(xin, ts, ctx, fs=fs)
    torch.cuda.synchronize()
    assert y.shape == y_ref.shape and torch.isfinite(y.float()).all()
    e_ours, e_ref = _rel(y, y_fp32), _rel(y_ref, y_fp32)
    e_pair = _rel(y, y_ref)
    print(f"rel L2: ours vs fp32 {e_ours:.3e}, ref-bf16 vs fp32 {e_ref:.3e}, ours vs ref-bf16 {e_pair:.3e}")
    assert e_pair <= 1e-2
    assert e_ours <= 2.0 * e_ref + 1e-3
