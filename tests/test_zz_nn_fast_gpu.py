"""GPU: the re-indexed GEGLU / im2col variants (csrc/nn_fast.cu, GVD_NN_FAST / gvd_nn_set_fast) against the kernels they
replace -- the same bits at the denoiser's real shapes.  Green on B200 (round 2, profiles/r02_first_hw_run.txt)."""
import pytest
import torch

pytestmark = [pytest.mark.gpu]
BF = torch.bfloat16


def test_fast_variants_bit_identical_on_gpu(monkeypatch):
    import gvd_native
    from vc_b200 import ops

    monkeypatch.setattr(ops, "IMPLICIT_CONV", False)  # the im2col kernels are what is compared here

    lib = gvd_native.nn()
    g = torch.Generator().manual_seed(0)

    def both(fn):
        was = lib.gvd_nn_set_fast(0)
        a = fn()
        lib.gvd_nn_set_fast(2)
        b = fn()
        lib.gvd_nn_set_fast(was)
        torch.cuda.synchronize()
        return a, b

    h = (torch.randn(25 * 640, 2 * 2560, generator=g) * 1.5).to(BF).cuda()
    a, b = both(lambda: ops.geglu(h))
    assert torch.equal(a, b)
    w = torch.zeros(8, 9 * 320, dtype=BF, device="cuda")
    x = torch.randn(5, 72 * 128, 320, generator=g).to(BF).cuda()
    for stride, up in ((1, False), (2, False), (1, True)):
        xs = x[:2] if up else x
        a, b = both(lambda: ops.conv3x3(xs, xs.shape[0], 72, 128, w, None, stride=stride, upsample=up)[0])
        assert torch.equal(a, b)
    wt = torch.zeros(8, 3 * 320, dtype=BF, device="cuda")
    a, b = both(lambda: ops.conv_t3(x, 1, 5, 72 * 128, wt, None))
    assert torch.equal(a, b)
    q, k, v = (torch.randn(25, 2000, 5 * 64, generator=g).to(BF).cuda() for _ in range(3))
    a, b = both(lambda: ops.temporal_attention(q, k, v, 1, 25, 2000, 5, 0.125))
    assert torch.equal(a, b)
    # level 1 (default): the mma.sync kernel (csrc/tattn_mma.cu) -- same rounding points, another fp32 summation order
    was = lib.gvd_nn_set_fast(1)
    c = ops.temporal_attention(q, k, v, 1, 25, 2000, 5, 0.125)
    lib.gvd_nn_set_fast(was)
    err = (c.float() - a.float()).abs().max().item()
    assert err <= 2.0 ** -7 * a.float().abs().max().item(), err
    assert (c != a).float().mean().item() < 0.05  # a bf16 ulp here and there, where a score sat on a rounding boundary
