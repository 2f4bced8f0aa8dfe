"""GPU numerics of the memory-bound U-Net layers (include/gvd_nn.h) against plain PyTorch fp32 references of the same
ops on the same bf16 inputs (tolerance: one bf16 rounding of the output, 2^-8 relative, plus fp32 reduction noise)."""
import pytest
import torch
import torch.nn.functional as Fn

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _close(a, b, tol=1.0 / 128):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    assert err <= tol * b.abs().max().item() + 1e-3, err


@pytest.mark.parametrize("F,S,C,silu", [(5, 256, 320, 1), (2, 1000, 64, 0), (1, 5 * 64, 2560, 2), (3, 77, 1920, 1), (25, 144, 1280, 2)])
def test_groupnorm(F, S, C, silu):
    from vc_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(C + S)
    x = (torch.randn(F, S, C, device="cuda", generator=g) * 2 + 0.5).to(BF)
    gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    y = ops.groupnorm(x, gamma, beta, F, S, groups=32, eps=1e-5, silu=silu)
    ref = Fn.group_norm(x.float().permute(0, 2, 1), 32, gamma, beta, 1e-5).permute(0, 2, 1)
    if silu == 1:
        ref = Fn.silu(ref.to(BF).float())
    elif silu == 2:
        ref = Fn.silu(ref)
    _close(y, ref)


@pytest.mark.parametrize("S,C,cut,silu", [(25 * 144, 1280, 25 * 36, 2), (1000, 320, 333, 0)])
def test_groupnorm_sharded_rows(S, C, cut, silu):
    """Statistics added up across two row shards (what two ranks exchange) reproduce the unsharded norm."""
    from vc_b200 import ops

    class TwoShards:  # stands in for FramePartition.sum_stats: adds the other shard's (sum, sumsq)
        def __init__(self):
            self.seen, self.other = [], None

        def sum_stats(self, st):
            if self.other is None:
                self.seen.append(st.clone())
            else:
                st += self.other
            return st

    g = torch.Generator(device="cuda").manual_seed(C + S)
    x = (torch.randn(1, S, C, device="cuda", generator=g) * 2 + 0.5).to(BF)
    gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    full = ops.groupnorm(x, gamma, beta, 1, S, groups=32, eps=1e-6, silu=silu)
    halves = [x[:, :cut].contiguous(), x[:, cut:].contiguous()]
    probe = TwoShards()
    for hx in halves:  # first pass only records each shard's local statistics
        ops.groupnorm_sharded(hx, gamma, beta, 1, hx.shape[1], S, probe, groups=32, eps=1e-6, silu=silu)
    outs = []
    for i, hx in enumerate(halves):
        part = TwoShards()
        part.other = probe.seen[1 - i]
        outs.append(ops.groupnorm_sharded(hx, gamma, beta, 1, hx.shape[1], S, part, groups=32, eps=1e-6, silu=silu))
    got = torch.cat(outs, dim=1)
    # identical up to the fp32 summation order of the statistics: at most one bf16 ulp on isolated elements
    d = (got.float() - full.float()).abs()
    assert float(d.max()) <= 2 ** -6 * float(full.float().abs().max()) and float((d > 0).float().mean()) < 0.01


def test_layernorm_geglu_softmax():
    from vc_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(1000, 640, device="cuda", generator=g).to(BF)
    gamma, beta = 1 + 0.1 * torch.randn(640, device="cuda", generator=g), 0.1 * torch.randn(640, device="cuda", generator=g)
    _close(ops.layernorm(x, gamma, beta), Fn.layer_norm(x.float(), (640,), gamma, beta))
    h = torch.randn(777, 2 * 1280, device="cuda", generator=g).to(BF)
    a, gate = h.float().chunk(2, -1)
    _close(ops.geglu(h), a * Fn.gelu(gate).to(BF).float())
    for dt in (BF, torch.float32):
        sc = (4 * torch.randn(300, 80, device="cuda", generator=g)).to(dt)
        p = ops.softmax_rows(sc, 77, 80)
        _close(p[:, :77], torch.softmax(sc[:, :77].float(), -1))
        assert float(p[:, 77:].abs().max()) == 0.0


@pytest.mark.parametrize("F,H,W,C,stride,up", [(3, 16, 24, 64, 1, False), (2, 16, 16, 320, 2, False), (2, 8, 12, 128, 1, True), (1, 9, 7, 8, 1, False)])
def test_conv3x3(F, H, W, C, stride, up):
    from vc_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(H * W + C)
    Cout = 96
    x = torch.randn(F, C, H, W, device="cuda", generator=g).to(BF)
    w = (torch.randn(Cout, C, 3, 3, device="cuda", generator=g) / (3 * C ** 0.5)).to(BF)
    b = torch.randn(Cout, device="cuda", generator=g)
    xin = Fn.interpolate(x.float(), scale_factor=2, mode="nearest") if up else x.float()
    ref = Fn.conv2d(xin, w.float(), b, stride=stride, padding=1)
    x_cl = x.permute(0, 2, 3, 1).reshape(F, H * W, C).contiguous()
    w_cl = w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
    y, Ho, Wo = ops.conv3x3(x_cl, F, H, W, w_cl, b, stride=stride, upsample=up)
    assert (Ho, Wo) == tuple(ref.shape[2:])
    _close(y.view(F, Ho, Wo, Cout).permute(0, 3, 1, 2), ref)


@pytest.mark.parametrize("F,H,W,C,Cout", [(3, 8, 16, 64, 72), (2, 18, 32, 128, 64), (2, 5, 128, 64, 320), (1, 3, 256, 192, 136),
                                          (2, 36, 64, 320, 320), (1, 16, 8, 64, 8)])
def test_conv3x3_implicit_gemm(F, H, W, C, Cout, monkeypatch):
    """gvd_conv_bf16 kind 1 (TMA-shifted activation tiles, zero padding by out-of-bounds fill): bit-identical to the
    im2col + GEMM route (same K order, same accumulation), right against fp32 torch, with every epilogue input; its
    tap-reversed form is the data gradient."""
    from vc_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(H * W + C)
    x = torch.randn(F, C, H, W, device="cuda", generator=g).to(BF)
    w = (torch.randn(Cout, C, 3, 3, device="cuda", generator=g) / (3 * C ** 0.5)).to(BF)
    b, b2 = torch.randn(Cout, device="cuda", generator=g), torch.randn(Cout, device="cuda", generator=g)
    res = torch.randn(F, H * W, Cout, device="cuda", generator=g).to(BF)
    x_cl = x.permute(0, 2, 3, 1).reshape(F, H * W, C).contiguous()
    w_cl = w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
    assert ops._implicit_ok(1, H, W, C, Cout, x_cl)
    assert not ops._implicit_ok(1, 9, 16, C, Cout, x_cl)  # 144 pixels per frame: two tiles 56 % full -> im2col route
    y, Ho, Wo = ops.conv3x3(x_cl, F, H, W, w_cl, b, bias2=b2, residual=res)
    y0, _, _ = ops.conv3x3(x_cl, F, H, W, w_cl, b)
    monkeypatch.setattr(ops, "IMPLICIT_CONV", False)
    y_col, _, _ = ops.conv3x3(x_cl, F, H, W, w_cl, b, bias2=b2, residual=res)
    assert torch.equal(y, y_col)
    ref = Fn.conv2d(x.float(), w.float(), b, padding=1)
    _close(y0.view(F, H, W, Cout).permute(0, 3, 1, 2), ref)
    # data gradient: implicit (tap-reversed weight) vs the col2im route vs autograd
    dy = torch.randn(F, H * W, Cout, device="cuda", generator=g).to(BF)
    dx_col = ops.conv3x3_dx(dy, F, H, W, C, w_cl)
    monkeypatch.setattr(ops, "IMPLICIT_CONV", True)
    if ops._implicit_ok(1, H, W, Cout, C, dy):
        dx = ops.conv3x3_dx(dy, F, H, W, C, w_cl)
        xr = x.float().requires_grad_(True)
        Fn.conv2d(xr, w.float(), None, padding=1).backward(dy.float().view(F, H, W, Cout).permute(0, 3, 1, 2))
        want = xr.grad.permute(0, 2, 3, 1).reshape(F, H * W, C)
        _close(dx, want)
        _close(dx_col, want, tol=1.0 / 64)  # the col2im route rounds the nine partial products to bf16 first


@pytest.mark.parametrize("F,H,W,C,Cout", [(2, 8, 16, 64, 64), (3, 20, 32, 128, 320), (1, 5, 64, 256, 128)])
def test_upsample_conv_on_the_implicit_route(F, H, W, C, Cout, monkeypatch):
    """Upsample + 3x3 convolution (openaimodel3d.py / ae_modules.py Upsample.forward): the materialised 4x tensor through
    gvd_conv_bf16 is bit-identical to the fused-upsampling im2col + GEMM route; its data gradient (implicit dgrad over the
    4x grid, then the 2 x 2 block sum) matches autograd of interpolate + conv2d within the bf16 bar."""
    from vc_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(H * W + C)
    x = torch.randn(F, C, H, W, device="cuda", generator=g).to(BF)
    w = (torch.randn(Cout, C, 3, 3, device="cuda", generator=g) / (3 * C ** 0.5)).to(BF)
    b = torch.randn(Cout, device="cuda", generator=g)
    x_cl = x.permute(0, 2, 3, 1).reshape(F, H * W, C).contiguous()
    w_cl = w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
    assert ops._implicit_ok(1, 2 * H, 2 * W, C, Cout, x_cl)
    xu = ops.upsample2x(x_cl, F, H, W).view(F, 2 * H, 2 * W, C)
    assert torch.equal(xu, x_cl.view(F, H, W, C).repeat_interleave(2, dim=1).repeat_interleave(2, dim=2))
    y, Ho, Wo = ops.conv3x3(x_cl, F, H, W, w_cl, b, upsample=True)
    assert (Ho, Wo) == (2 * H, 2 * W)
    dy = torch.randn(F, Ho * Wo, Cout, device="cuda", generator=g).to(BF)
    dx = ops.conv3x3_dx(dy, F, H, W, C, w_cl, 1, True)
    monkeypatch.setattr(ops, "IMPLICIT_CONV", False)
    y_col, _, _ = ops.conv3x3(x_cl, F, H, W, w_cl, b, upsample=True)
    dx_col = ops.conv3x3_dx(dy, F, H, W, C, w_cl, 1, True)
    assert torch.equal(y, y_col)
    xr = x.float().requires_grad_(True)
    ref = Fn.conv2d(Fn.interpolate(xr, scale_factor=2, mode="nearest"), w.float(), b, padding=1)
    _close(y.view(F, Ho, Wo, Cout).permute(0, 3, 1, 2), ref)
    ref.backward(dy.float().view(F, Ho, Wo, Cout).permute(0, 3, 1, 2))
    want = xr.grad.permute(0, 2, 3, 1).reshape(F, H * W, C)
    _close(dx, want, tol=1.0 / 64)      # two bf16 roundings: per upsampled pixel, then the block sum (as autograd under autocast)
    _close(dx_col, want, tol=1.0 / 64)


@pytest.mark.parametrize("B,T,S,C,Cout", [(1, 7, 50, 64, 128), (2, 1, 129, 128, 64), (1, 25, 300, 320, 320)])
def test_conv_t3_implicit_gemm(B, T, S, C, Cout, monkeypatch):
    from vc_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(T * S + C)
    x = torch.randn(B * T, S, C, device="cuda", generator=g).to(BF)
    w = (torch.randn(Cout, 3 * C, device="cuda", generator=g) / (3 * C) ** 0.5).to(BF)
    b = torch.randn(Cout, device="cuda", generator=g)
    res = torch.randn(B * T, S, Cout, device="cuda", generator=g).to(BF)
    y = ops.conv_t3(x, B, T, S, w, b, residual=res)
    dy = torch.randn(B * T, S, Cout, device="cuda", generator=g).to(BF)
    dx = ops.conv_t3_dx(dy, B, T, S, C, w)
    monkeypatch.setattr(ops, "IMPLICIT_CONV", False)
    assert torch.equal(y, ops.conv_t3(x, B, T, S, w, b, residual=res))
    xr = x.float().view(B, T, S, C).permute(0, 3, 1, 2).unsqueeze(-1).requires_grad_(True)       # [B, C, T, S, 1]
    w5 = w.float().view(Cout, 3, C).permute(0, 2, 1)[..., None, None]
    Fn.conv3d(xr, w5, None, padding=(1, 0, 0)).backward(dy.float().view(B, T, S, Cout).permute(0, 3, 1, 2).unsqueeze(-1))
    _close(dx, xr.grad[..., 0].permute(0, 2, 3, 1).reshape(B * T, S, C))
    _close(ops.conv_t3_dx(dy, B, T, S, C, w), xr.grad[..., 0].permute(0, 2, 3, 1).reshape(B * T, S, C), tol=1.0 / 64)


def test_conv_t3_and_temporal_attention():
    from vc_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(9)
    B, T, S, C, Cout = 1, 7, 50, 64, 128
    x = torch.randn(B, C, T, S, 1, device="cuda", generator=g).to(BF)
    w = (torch.randn(Cout, C, 3, 1, 1, device="cuda", generator=g) / (3 * C) ** 0.5).to(BF)
    b = torch.randn(Cout, device="cuda", generator=g)
    ref = Fn.conv3d(x.float(), w.float(), b, padding=(1, 0, 0))[..., 0]           # [B, Cout, T, S]
    x_cl = x[..., 0].permute(0, 2, 3, 1).reshape(B * T, S, C).contiguous()
    w_cl = w[:, :, :, 0, 0].permute(0, 2, 1).reshape(Cout, -1).contiguous()
    y = ops.conv_t3(x_cl, B, T, S, w_cl, b)
    _close(y.view(B, T, S, Cout).permute(0, 3, 1, 2), ref)
    for T in (1, 5, 25, 32):
        H, S = 5, 37
        q, k, v = (torch.randn(2, T, S, H * 64, device="cuda", generator=g).to(BF) for _ in range(3))
        o = ops.temporal_attention(q, k, v, 2, T, S, H, 0.125)
        qh, kh, vh = (t.float().view(2, T, S, H, 64).permute(0, 2, 3, 1, 4) for t in (q, k, v))  # [B,S,H,T,64]
        sim = ((qh @ kh.transpose(-1, -2)).to(BF).float() * 0.125).to(BF).float()
        ref = (torch.softmax(sim, -1).to(BF).float() @ vh).permute(0, 3, 1, 2, 4).reshape(2, T, S, H * 64)
        _close(o, ref)


@pytest.mark.parametrize("fn", ["attention", "flash_attention"])
def test_attention_paths(fn):
    from vc_b200 import ops

    attn = getattr(ops, fn)

    g = torch.Generator(device="cuda").manual_seed(11)
    Bq, Nq, H = 3, 200, 5
    q = torch.randn(Bq, Nq, H * 64, device="cuda", generator=g).to(BF)

    def ref(q, k, v):
        qh = q.float().view(q.shape[0], -1, H, 64).transpose(1, 2)
        kh = k.float().view(k.shape[0], -1, H, 64).transpose(1, 2)
        vh = v.float().view(v.shape[0], -1, H, 64).transpose(1, 2)
        sim = ((qh @ kh.transpose(-1, -2)).to(BF).float() * 0.125).to(BF).float()
        return (torch.softmax(sim, -1).to(BF).float() @ vh).transpose(1, 2).reshape(q.shape[0], -1, H * 64)

    k, v = (torch.randn(Bq, Nq, H * 64, device="cuda", generator=g).to(BF) for _ in range(2))
    _close(attn(q, k, v, Bq, Nq, Nq, H, 0.125), ref(q, k, v), tol=1.0 / 64)
    ks, vs = (torch.randn(1, 77, H * 64, device="cuda", generator=g).to(BF) for _ in range(2))
    _close(attn(q, ks, vs, Bq, Nq, 77, H, 0.125, shared_kv=True), ref(q, ks.expand(Bq, -1, -1), vs.expand(Bq, -1, -1)), tol=1.0 / 64)
    # longer sequences: several key blocks, ragged tails on both axes
    q2 = torch.randn(2, 700, H * 64, device="cuda", generator=g).to(BF)
    k2, v2 = (torch.randn(2, 700, H * 64, device="cuda", generator=g).to(BF) for _ in range(2))
    _close(attn(q2, k2, v2, 2, 700, 700, H, 0.125), ref(q2, k2, v2), tol=1.0 / 64)
    if fn != "flash_attention":
        return
    # Fused kernel only (it keeps fp32 logits; the bf16-rounded logits of the materialised path are too coarse at this
    # magnitude): logits that climb by ~0.05 per key (35 over the row) -- the running max overtakes the lazy-rescale
    # threshold several times, so the in-TMEM rescale of O is exercised -- and logits that fall (it never rescales).
    def ref32(q, k, v):
        qh, kh, vh = (t.float().view(t.shape[0], -1, H, 64).transpose(1, 2) for t in (q, k, v))
        return (torch.softmax(qh @ kh.transpose(-1, -2) * 0.125, -1) @ vh).transpose(1, 2).reshape(q.shape[0], -1, H * 64)

    for sign in (1.0, -1.0):
        q3, k3 = q2.clone(), k2.clone()
        q3.view(2, 700, H, 64)[..., 0] = 4.0
        k3.view(2, 700, H, 64)[..., 0] = (sign * 0.1 * torch.arange(700, device="cuda"))[None, :, None].to(BF)
        _close(attn(q3, k3, v2, 2, 700, 700, H, 0.125), ref32(q3, k3, v2), tol=1.0 / 64)


@pytest.mark.parametrize("variant", ["v1", "v2", "v5", "v6", "v8"])
def test_flash_attention_variants(variant):
    """The A/B variants of gvd_flash_attention (GVD_FLASH, read once per process): first generation, two-tile ping-pong,
    one-pass chunk-pipelined softmax -- each in its own process against fp32 attention, ragged sizes included."""
    import os
    import subprocess
    import sys

    code = r'''
import sys, torch
sys.path.insert(0, "guidedvd-3dgs_b200")
from vc_b200 import ops
g = torch.Generator(device="cuda").manual_seed(5)
for (B, Nq, Nk, H) in ((3, 300, 300, 2), (2, 1000, 77, 5), (1, 129, 513, 1), (2, 2304, 2304, 3)):
    q, k, v = (torch.randn(B, n, H * 64, device="cuda", generator=g).bfloat16() for n in (Nq, Nk, Nk))
    k = k * 3.0  # logits spread over +-20: exercises the running-maximum updates
    o = ops.flash_attention(q, k, v, B, Nq, Nk, H, 0.125).float().view(B, Nq, H, 64)
    qh, kh, vh = (t.float().view(B, -1, H, 64).permute(0, 2, 1, 3) for t in (q, k, v))
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * 0.125, -1) @ vh).permute(0, 2, 1, 3)
    err = (o - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1.5e-2, (B, Nq, Nk, H, err)
print("ok")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GVD_FLASH=variant)
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
