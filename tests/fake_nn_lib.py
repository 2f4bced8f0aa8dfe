"""CPU stand-in for libgvd_nn.so at the POINTER level (test infrastructure, never shipped).

The product has no CPU path: vc_b200.ops hands raw device pointers to the C ABI of include/gvd_nn.h.  To check the host
side without a GPU -- the ctypes argument order of every binding, the strides handed to the GEMM, the composition of
the U-Net forward and of its input-gradient (vc_b200.grad), the guided sampler -- this module implements every entry
point of the header over HOST pointers with plain torch arithmetic, restating what each kernel computes (formulas of
csrc/nn_kernels.cu / nn_backward.cu, index math of the im2col / col2im kernels).  tests/conftest-style fixtures install
it in place of `gvd_native.nn()`; activations are then fp32 (`act_dtype`) so the comparison against the reference
module in fp32 is tight, or bf16 to exercise the kernels' rounding points.
"""
import ctypes
import math

import torch


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, ctypes.c_void_p):
        return p.value or 0
    return int(p)


class FakeNN:
    def __init__(self, act_dtype=torch.float32):
        self.act = act_dtype
        self.calls = {}

    # ---- plumbing ----
    def _t(self, ptr, numel, dtype):
        a = _addr(ptr)
        if a == 0:
            return None
        nbytes = int(numel) * torch.empty((), dtype=dtype).element_size()
        buf = (ctypes.c_char * nbytes).from_address(a)
        return torch.frombuffer(buf, dtype=dtype)

    def _a(self, ptr, *shape):
        return self._t(ptr, math.prod(shape), self.act).view(*shape)

    def _f(self, ptr, *shape):
        return self._t(ptr, math.prod(shape), torch.float32).view(*shape)

    def _rnd(self, x):
        return x.to(self.act).float()

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    def gvd_nn_last_error(self):
        return b"fake"

    def gvd_nn_set_fast(self, on):
        return 0

    # ---- GEMM ----
    def gvd_gemm_bf16(self, args, stream):
        self._count("gemm")
        a = args._obj
        M, N, K, bh, bb = a.M, a.N, a.K, a.batch_h, a.batch_b
        if M <= 0 or N <= 0:
            return 0
        for v in (a.lda, a.ldb) + ((a.a_stride_h, a.b_stride_h) if bh > 1 else ()) + ((a.a_stride_b, a.b_stride_b) if bb > 1 else ()):
            assert v % 8 == 0, "gvd_gemm_bf16: operand strides must be multiples of 8 elements"

        def view(ptr, rows, cols, ld, sh, sb, dtype):
            ext = (bb - 1) * sb + (bh - 1) * sh + (rows - 1) * ld + cols
            return self._t(ptr, ext, dtype).as_strided((bb, bh, rows, cols), (sb, sh, ld, 1))

        A = view(a.A, M, K, a.lda, a.a_stride_h, a.a_stride_b, self.act).float()
        if getattr(a, "b_mn_major", 0):  # B[k][n], N contiguous
            B = view(a.B, K, N, a.ldb, a.b_stride_h, a.b_stride_b, self.act).float().transpose(-1, -2)
        else:
            B = view(a.B, N, K, a.ldb, a.b_stride_h, a.b_stride_b, self.act).float()
        cdt = torch.float32 if a.out_fp32 else self.act
        Cv = view(a.C, M, N, a.ldc, a.c_stride_h, a.c_stride_b, cdt)
        acc = A @ B.transpose(-1, -2)
        if a.act == 4:  # fused GEGLU: B's rows interleaved in blocks of 16 values | 16 gates; C has N / 2 columns
            assert N % 32 == 0 and not a.residual and not a.bias2  # (out_fp32 is set when the tests run with fp32 activations)
            y = acc * a.alpha
            if a.bias:
                y = y + self._f(a.bias, N)
            y = self._rnd(y).view(bb, bh, M, N // 32, 2, 16)
            out = self._rnd(y[..., 0, :] * self._rnd(torch.nn.functional.gelu(y[..., 1, :]))).reshape(bb, bh, M, N // 2)
            view(a.C, M, N // 2, a.ldc, a.c_stride_h, a.c_stride_b, self.act).copy_(out.to(self.act))
            return 0
        if a.act == 3:  # C = bf16(bf16(acc) * alpha)
            y = self._rnd(acc) * a.alpha
        else:
            y = acc * a.alpha
            if a.bias:
                y = y + self._f(a.bias, N)
            if a.act == 1:
                y = torch.nn.functional.silu(y)
            elif a.act == 2:
                y = torch.nn.functional.gelu(y)
        if not a.out_fp32:
            y = self._rnd(y)
        if a.bias2:
            y = y + self._f(a.bias2, N)
        if a.residual:
            y = y + view(a.residual, M, N, a.ldc, a.c_stride_h, a.c_stride_b, cdt).float()
        Cv.copy_(y.to(cdt))
        return 0

    # ---- implicit-GEMM convolution (gvd_conv_bf16): same arithmetic as im2col + GEMM, stated as a convolution ----
    def gvd_conv_bf16_supported(self, kind, H, W, Cin, Cout):
        if Cin <= 0 or Cout <= 0 or Cin % 64 or Cout % 8:
            return 0
        if kind == 2:
            return 1
        if not (kind == 1 and H > 0 and W > 0 and (W % 128 == 0 or (W <= 128 and 128 % W == 0))):
            return 0
        px = H * W
        return int((px + 127) // 128 * 128 * 8 <= px * 9)

    def gvd_conv_bf16(self, args, stream):
        self._count("conv_implicit")
        a = args._obj
        assert self.gvd_conv_bf16_supported(a.kind, a.H, a.W, a.Cin, a.Cout)
        Ci, Co = a.Cin, a.Cout
        if a.kind == 1:
            x = self._a(a.x, a.F, a.H, a.W, Ci).float().permute(0, 3, 1, 2)
            w = self._a(a.weight, Co, 3, 3, Ci).float().permute(0, 3, 1, 2)
            acc = torch.nn.functional.conv2d(x, w, padding=1).permute(0, 2, 3, 1).reshape(-1, Co)
        else:
            x = self._a(a.x, a.B, a.T, a.S, Ci).float().permute(0, 2, 3, 1).reshape(a.B * a.S, Ci, a.T)
            w = self._a(a.weight, Co, 3, Ci).float().permute(0, 2, 1)
            acc = torch.nn.functional.conv1d(x, w, padding=1).view(a.B, a.S, Co, a.T).permute(0, 3, 1, 2).reshape(-1, Co)
        rows = acc.shape[0]
        y = acc
        if a.bias:
            y = y + self._f(a.bias, Co)
        if a.act == 1:
            y = torch.nn.functional.silu(y)
        elif a.act == 2:
            y = torch.nn.functional.gelu(y)
        y = self._rnd(y)
        if a.bias2:
            y = y + self._f(a.bias2, Co)
        if a.residual:
            y = y + self._a(a.residual, rows, Co).float()
        self._a(a.y, rows, Co).copy_(y.to(self.act))
        return 0

    # ---- GroupNorm ----
    def gvd_groupnorm_tmp_floats(self, F, S, groups):
        return F * groups * 2 * 4

    def _gn(self, x, gamma, beta, mean, rstd, C, groups, silu):
        F, S = x.shape[0], x.shape[1]
        cpg = C // groups
        xh = (x.view(F, S, groups, cpg) - mean.view(F, 1, groups, 1)) * rstd.view(F, 1, groups, 1)
        z = xh.view(F, S, C) * gamma + beta
        if silu == 1:
            return torch.nn.functional.silu(self._rnd(z)), xh.view(F, S, C), z
        if silu == 2:
            return torch.nn.functional.silu(z), xh.view(F, S, C), z
        return z, xh.view(F, S, C), z

    def gvd_groupnorm_cl(self, x, y, gamma, beta, F, S, C, groups, eps, silu, tmp, tmp_floats, stream):
        self._count("groupnorm")
        assert tmp_floats >= 1 and _addr(tmp)
        xx = self._a(x, F, S, C).float()
        xg = xx.view(F, S, groups, C // groups)
        mean = xg.mean(dim=(1, 3))
        var = xg.var(dim=(1, 3), unbiased=False)
        out, _, _ = self._gn(xx, self._f(gamma, C), self._f(beta, C), mean, (var + eps).rsqrt(), C, groups, silu)
        self._a(y, F, S, C).copy_(out.to(self.act))
        return 0

    def gvd_groupnorm_cl_keep_stats(self, x, y, gamma, beta, stats, F, S, C, groups, eps, silu, tmp, tmp_floats, stream):
        assert tmp_floats >= 1 and _addr(tmp)
        self.gvd_groupnorm_cl_stats(x, stats, F, S, C, groups, tmp, tmp_floats, stream)
        return self.gvd_groupnorm_cl_apply(x, y, gamma, beta, stats, F, S, S, C, groups, eps, silu, stream)

    def gvd_groupnorm_cl_stats(self, x, stats, F, S, C, groups, tmp, tmp_floats, stream):
        self._count("groupnorm_stats")
        xg = self._a(x, F, S, groups, C // groups).double()
        st = self._f(stats, F, groups, 2)
        st[..., 0] = xg.sum(dim=(1, 3)).float()
        st[..., 1] = (xg * xg).sum(dim=(1, 3)).float()
        return 0

    def _stats_to_moments(self, stats, F, groups, n, eps):
        st = self._f(stats, F, groups, 2).double()
        mean = st[..., 0] / n
        var = (st[..., 1] / n - mean * mean).clamp_min(0.0)
        return mean.float(), (1.0 / (var + eps).sqrt()).float()

    def gvd_groupnorm_cl_apply(self, x, y, gamma, beta, stats, F, S, stat_rows, C, groups, eps, silu, stream):
        self._count("groupnorm_apply")
        mean, rstd = self._stats_to_moments(stats, F, groups, stat_rows * (C // groups), eps)
        out, _, _ = self._gn(self._a(x, F, S, C).float(), self._f(gamma, C), self._f(beta, C), mean, rstd, C, groups, silu)
        self._a(y, F, S, C).copy_(out.to(self.act))
        return 0

    def gvd_groupnorm_bwd_tmp_bytes(self, F, S, groups):
        return F * groups * 2 * 8 * 4

    def gvd_groupnorm_cl_bwd(self, x, dy, dx, gamma, beta, stats, F, S, C, groups, eps, silu, tmp, tmp_bytes, stream):
        """g = dy * act'(.) * gamma;  dx = rstd * (g - mean(g) - xh * mean(g * xh))   (csrc/nn_backward.cu)."""
        self._count("groupnorm_bwd")
        assert _addr(tmp) % 8 == 0 and tmp_bytes >= F * groups * 2 * 8
        cpg = C // groups
        mean, rstd = self._stats_to_moments(stats, F, groups, S * cpg, eps)
        gm = self._f(gamma, C)
        _, xh, z = self._gn(self._a(x, F, S, C).float(), gm, self._f(beta, C), mean, rstd, C, groups, 0)
        g = self._a(dy, F, S, C).float() * gm
        if silu:
            zz = self._rnd(z) if silu == 1 else z
            sg = torch.sigmoid(zz)
            g = g * (sg * (1 + zz * (1 - sg)))
        gg, xg = g.view(F, S, groups, cpg), xh.view(F, S, groups, cpg)
        m1 = gg.mean(dim=(1, 3), keepdim=True)
        m2 = (gg * xg).mean(dim=(1, 3), keepdim=True)
        out = rstd.view(F, 1, groups, 1) * (gg - m1 - xg * m2)
        self._a(dx, F, S, C).copy_(out.view(F, S, C).to(self.act))
        return 0

    def _gn_bwd_terms(self, x, dy, gamma, beta, stats, F, S, stat_rows, C, groups, eps, silu):
        cpg = C // groups
        mean, rstd = self._stats_to_moments(stats, F, groups, stat_rows * cpg, eps)
        gm = self._f(gamma, C)
        _, xh, z = self._gn(self._a(x, F, S, C).float(), gm, self._f(beta, C), mean, rstd, C, groups, 0)
        g = self._a(dy, F, S, C).float() * gm
        if silu:
            zz = self._rnd(z) if silu == 1 else z
            sg = torch.sigmoid(zz)
            g = g * (sg * (1 + zz * (1 - sg)))
        return g.view(F, S, groups, cpg), xh.view(F, S, groups, cpg), rstd

    def gvd_groupnorm_cl_bwd_sums(self, x, dy, gamma, beta, stats, sums, F, S, stat_rows, C, groups, eps, silu, tmp, tmp_bytes, stream):
        self._count("groupnorm_bwd_sums")
        out = self._t(sums, F * groups * 2, torch.float64).view(F, groups, 2)
        if S <= 0:
            out.zero_()
            return 0
        gg, xg, _ = self._gn_bwd_terms(x, dy, gamma, beta, stats, F, S, stat_rows, C, groups, eps, silu)
        out[..., 0] = gg.double().sum(dim=(1, 3))
        out[..., 1] = (gg * xg).double().sum(dim=(1, 3))
        return 0

    def gvd_groupnorm_cl_bwd_apply(self, x, dy, dx, gamma, beta, stats, sums, F, S, stat_rows, C, groups, eps, silu, stream):
        self._count("groupnorm_bwd_apply")
        if S <= 0:
            return 0
        gg, xg, rstd = self._gn_bwd_terms(x, dy, gamma, beta, stats, F, S, stat_rows, C, groups, eps, silu)
        sm = self._t(sums, F * groups * 2, torch.float64).view(F, groups, 2) / (stat_rows * (C // groups))
        m1, m2 = sm[..., 0].float().view(F, 1, groups, 1), sm[..., 1].float().view(F, 1, groups, 1)
        out = rstd.view(F, 1, groups, 1) * (gg - m1 - xg * m2)
        self._a(dx, F, S, C).copy_(out.view(F, S, C).to(self.act))
        return 0

    # ---- LayerNorm / GEGLU / softmax ----
    def gvd_layernorm(self, x, y, gamma, beta, rows, C, eps, stream):
        self._count("layernorm")
        out = torch.nn.functional.layer_norm(self._a(x, rows, C).float(), (C,), self._f(gamma, C), self._f(beta, C), eps)
        self._a(y, rows, C).copy_(out.to(self.act))
        return 0

    def gvd_layernorm_bwd(self, x, dy, dx, gamma, rows, C, eps, stream):
        self._count("layernorm_bwd")
        xx = self._a(x, rows, C).float()
        mean = xx.mean(-1, keepdim=True)
        rstd = (xx.var(-1, unbiased=False, keepdim=True) + eps).rsqrt()
        xh = (xx - mean) * rstd
        g = self._a(dy, rows, C).float() * self._f(gamma, C)
        out = rstd * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
        self._a(dx, rows, C).copy_(out.to(self.act))
        return 0

    def gvd_geglu(self, h, out, rows, D, stream):
        self._count("geglu")
        hh = self._a(h, rows, 2 * D).float()
        self._a(out, rows, D).copy_((hh[:, :D] * self._rnd(torch.nn.functional.gelu(hh[:, D:]))).to(self.act))
        return 0

    def gvd_geglu_bwd(self, h, dout, dh, rows, D, stream):
        self._count("geglu_bwd")
        hh = self._a(h, rows, 2 * D).float()
        a, g = hh[:, :D], hh[:, D:]
        d = self._a(dout, rows, D).float()
        dgelu = 0.5 * (1 + torch.erf(g * 0.7071067811865476)) + g * 0.3989422804014327 * torch.exp(-0.5 * g * g)
        o = self._a(dh, rows, 2 * D)
        o[:, :D] = (d * self._rnd(torch.nn.functional.gelu(g))).to(self.act)
        o[:, D:] = (d * a * dgelu).to(self.act)
        return 0

    def gvd_softmax_rows(self, x, x_is_bf16, ldx, y, ldy, rows, cols, stream):
        self._count("softmax")
        xx = (self._a(x, rows, ldx) if x_is_bf16 else self._f(x, rows, ldx)).float()
        yy = self._a(y, rows, ldy)
        yy[:, :cols] = torch.softmax(xx[:, :cols], dim=-1).to(self.act)
        yy[:, cols:] = 0
        return 0

    def gvd_softmax_bwd_rows(self, p, dp, ds, ld, rows, cols, stream):
        self._count("softmax_bwd")
        pp = self._a(p, rows, ld).float()[:, :cols]
        dd = self._a(dp, rows, ld).float()[:, :cols].clone()
        out = self._a(ds, rows, ld)
        out[:, :cols] = (pp * (dd - (pp * dd).sum(-1, keepdim=True))).to(self.act)
        out[:, cols:] = 0
        return 0

    # ---- im2col / col2im (index math of the kernels) ----
    @staticmethod
    def _conv_geom(H, W, stride, up):
        Hin, Win = (2 * H, 2 * W) if up else (H, W)
        return Hin, Win, (Hin + 2 - 3) // stride + 1, (Win + 2 - 3) // stride + 1

    def gvd_im2col3x3_cl(self, x, col, F, H, W, C, stride, up, stream):
        self._count("im2col3x3")
        Hin, Win, Ho, Wo = self._conv_geom(H, W, stride, up)
        xx = self._a(x, F, H, W, C)
        cc = self._a(col, F, Ho, Wo, 9, C)
        oy = torch.arange(Ho).view(Ho, 1)
        ox = torch.arange(Wo).view(1, Wo)
        for tap in range(9):
            iy, ix = oy * stride + tap // 3 - 1, ox * stride + tap % 3 - 1
            ok = ((iy >= 0) & (iy < Hin) & (ix >= 0) & (ix < Win)).expand(Ho, Wo)
            sy, sx = (iy >> 1, ix >> 1) if up else (iy, ix)
            sy, sx = sy.clamp(0, H - 1).expand(Ho, Wo), sx.clamp(0, W - 1).expand(Ho, Wo)
            cc[:, :, :, tap, :] = xx[:, sy, sx, :] * ok.view(1, Ho, Wo, 1).to(self.act)
        return 0

    def gvd_im2col3x3_down_cl(self, x, col, F, H, W, C, stream):
        self._count("im2col3x3_down")
        Ho, Wo = (H - 2) // 2 + 1, (W - 2) // 2 + 1
        xp = torch.nn.functional.pad(self._a(x, F, H, W, C), (0, 0, 0, 1, 0, 1))  # right / bottom only (ae_modules.py:101-102)
        cc = self._a(col, F, Ho, Wo, 9, C)
        for tap in range(9):
            ky, kx = tap // 3, tap % 3
            cc[:, :, :, tap, :] = xp[:, ky:ky + 2 * Ho:2, kx:kx + 2 * Wo:2, :][:, :Ho, :Wo]
        return 0

    def gvd_col2im3x3_cl(self, dcol, dx, F, H, W, C, stride, up, stream):
        self._count("col2im3x3")
        _, _, Ho, Wo = self._conv_geom(H, W, stride, up)
        dc = self._a(dcol, F, Ho, Wo, 9, C).float()
        acc = torch.zeros(F, H, W, C)
        sub = 2 if up else 1
        iy = torch.arange(H).view(H, 1)
        ix = torch.arange(W).view(1, W)
        for a in range(sub):
            for ky in range(3):
                ty = iy * sub + a + 1 - ky
                oky = (ty >= 0) & (ty % stride == 0) & (ty // stride < Ho)
                for b in range(sub):
                    for kx in range(3):
                        tx = ix * sub + b + 1 - kx
                        okx = (tx >= 0) & (tx % stride == 0) & (tx // stride < Wo)
                        ok = (oky & okx).expand(H, W)
                        oy = (ty // stride).clamp(0, Ho - 1).expand(H, W)
                        ox = (tx // stride).clamp(0, Wo - 1).expand(H, W)
                        acc += dc[:, oy, ox, ky * 3 + kx, :] * ok.view(1, H, W, 1).float()
        self._a(dx, F, H, W, C).copy_(acc.to(self.act))
        return 0

    def gvd_im2col_t3_cl(self, x, col, B, T, S, C, stream):
        self._count("im2col_t3")
        xx = self._a(x, B, T, S, C)
        cc = self._a(col, B, T, S, 3, C)
        cc.zero_()
        for tap in range(3):
            for tt in range(T):
                it = tt + tap - 1
                if 0 <= it < T:
                    cc[:, tt, :, tap, :] = xx[:, it]
        return 0

    def gvd_col2im_t3_cl(self, dcol, dx, B, T, S, C, stream):
        self._count("col2im_t3")
        dc = self._a(dcol, B, T, S, 3, C).float()
        acc = torch.zeros(B, T, S, C)
        for it in range(T):
            for tap in range(3):
                tt = it - tap + 1
                if 0 <= tt < T:
                    acc[:, it] += dc[:, tt, :, tap, :]
        self._a(dx, B, T, S, C).copy_(acc.to(self.act))
        return 0

    # ---- attention ----
    def _tattn_probs(self, q, k, scale):
        s = self._rnd(self._rnd(torch.einsum("bhsid,bhsjd->bhsij", q, k)) * scale)
        return torch.softmax(s, dim=-1)

    def gvd_temporal_attention(self, q, k, v, out, B, T, S, H, scale, stream):
        self._count("temporal_attention")
        sp = lambda p: self._a(p, B, T, S, H, 64).float().permute(0, 3, 2, 1, 4)  # noqa: E731  [B, H, S, T, 64]
        p = self._rnd(self._tattn_probs(sp(q), sp(k), scale))
        o = torch.einsum("bhsij,bhsjd->bhsid", p, sp(v))
        self._a(out, B, T, S, H, 64).copy_(o.permute(0, 3, 2, 1, 4).to(self.act))
        return 0

    def gvd_temporal_attention_bwd(self, q, k, v, dout, dq, dk, dv, B, T, S, H, scale, stream):
        self._count("temporal_attention_bwd")
        sp = lambda p: self._a(p, B, T, S, H, 64).float().permute(0, 3, 2, 1, 4)  # noqa: E731
        qq, kk, vv, do = sp(q), sp(k), sp(v), sp(dout)
        p = self._tattn_probs(qq, kk, scale)
        dp = torch.einsum("bhsid,bhsjd->bhsij", do, vv)
        ds = p * (dp - (p * dp).sum(-1, keepdim=True)) * scale
        back = lambda t: t.permute(0, 3, 2, 1, 4).to(self.act)  # noqa: E731
        self._a(dq, B, T, S, H, 64).copy_(back(torch.einsum("bhsij,bhsjd->bhsid", ds, kk)))
        self._a(dk, B, T, S, H, 64).copy_(back(torch.einsum("bhsij,bhsid->bhsjd", ds, qq)))
        self._a(dv, B, T, S, H, 64).copy_(back(torch.einsum("bhsij,bhsid->bhsjd", self._rnd(p), do)))
        return 0

    def gvd_flash_attention(self, q, k, v, out, B, Nq, Nk, H, qs, ks, scale, stream):
        self._count("flash_attention")
        HD = H * 64

        def view(ptr, n, bs):
            ext = (B - 1) * bs + n * HD
            return self._t(ptr, ext, self.act).as_strided((B, n, H, 64), (bs, HD, 64, 1))

        qq, kk, vv = view(q, Nq, qs).float(), view(k, Nk, ks).float(), view(v, Nk, ks).float()
        p = torch.softmax(torch.einsum("bihd,bjhd->bhij", qq, kk) * scale, dim=-1)
        o = torch.einsum("bhij,bjhd->bihd", self._rnd(p), vv)
        view(out, Nq, qs).copy_(o.to(self.act))
        return 0

    def gvd_upsample2x_cl(self, x, y, F, H, W, C, stream):
        self._count("upsample2x")
        xx = self._a(x, F, H, W, C)
        self._a(y, F, 2 * H, 2 * W, C).copy_(xx.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2))
        return 0

    def gvd_upsample2x_bwd_cl(self, dy, dx, F, H, W, C, stream):
        self._count("upsample2x_bwd")
        d = self._a(dy, F, H, 2, W, 2, C).float()
        self._a(dx, F, H, W, C).copy_(d.sum(dim=(2, 4)).to(self.act))
        return 0

    def gvd_flash_attention_lse(self, q, k, v, out, lse, B, Nq, Nk, H, qs, ks, scale, stream):
        self._count("flash_attention_lse")
        HD = H * 64

        def view(ptr, n, bs):
            ext = (B - 1) * bs + n * HD
            return self._t(ptr, ext, self.act).as_strided((B, n, H, 64), (bs, HD, 64, 1))

        qq, kk, vv = view(q, Nq, qs).float(), view(k, Nk, ks).float(), view(v, Nk, ks).float()
        s = torch.einsum("bihd,bjhd->bhij", qq, kk) * scale
        p = torch.softmax(s, dim=-1)
        view(out, Nq, qs).copy_(torch.einsum("bhij,bjhd->bihd", self._rnd(p), vv).to(self.act))
        ldl = (Nq + 127) // 128 * 128
        L = self._f(lse, B, H, ldl)
        L.zero_()
        L[:, :, :Nq] = torch.logsumexp(s, dim=-1) * 1.4426950408889634
        return 0

    def gvd_flash_attention_bwd(self, args, stream):
        self._count("flash_attention_bwd")
        a = args._obj
        B, Nq, Nk, H, scale = a.B, a.Nq, a.Nk, a.H, a.scale
        HD = H * 64

        def view(ptr, n, bs):
            ext = (B - 1) * bs + n * HD
            return self._t(ptr, ext, self.act).as_strided((B, n, H, 64), (bs, HD, 64, 1))

        qs, ks = a.q_batch_stride, a.kv_batch_stride
        qq, kk, vv = view(a.q, Nq, qs).float(), view(a.k, Nk, ks).float(), view(a.v, Nk, ks).float()
        oo, do = view(a.out, Nq, qs).float(), view(a.dout, Nq, qs).float()
        ldl = (Nq + 127) // 128 * 128
        L = self._f(a.lse, B, H, ldl)[:, :, :Nq]
        p = torch.exp2(torch.einsum("bihd,bjhd->bhij", qq, kk) * (scale * 1.4426950408889634) - L[..., None])
        D = (do * oo).sum(-1).permute(0, 2, 1)  # [B, H, Nq]
        dp = torch.einsum("bihd,bjhd->bhij", do, vv)
        ds = self._rnd(p * (dp - D[..., None]))
        view(a.dq, Nq, qs).copy_((torch.einsum("bhij,bjhd->bihd", ds, kk) * scale).to(self.act))
        if a.dk:
            view(a.dk, Nk, ks).copy_((torch.einsum("bhij,bihd->bjhd", ds, qq) * scale).to(self.act))
            view(a.dv, Nk, ks).copy_(torch.einsum("bhij,bihd->bjhd", self._rnd(p), do).to(self.act))
        return 0

    # ---- DDIM ----
    @staticmethod
    def _std_ratio(e_c, mo):
        return e_c.double().std() / mo.double().std()

    def gvd_ddim_step(self, args, stream):
        self._count("ddim_step")
        a = args._obj
        n = a.n
        x, e_c, noise = self._f(a.x, n), self._f(a.e_cond, n), self._f(a.noise, n)
        e_u = self._f(a.e_uncond, n) if a.e_uncond else None
        v = e_c
        if e_u is not None:
            mo = e_u + a.cfg_scale * (e_c - e_u)
            v = mo
            if a.guidance_rescale > 0:
                ratio = float(self._std_ratio(e_c, mo))
                v = a.guidance_rescale * (mo * ratio) + (1 - a.guidance_rescale) * mo
        sa, s1 = a.sqrt_alphas_cumprod_t, a.sqrt_one_minus_alphas_cumprod_t
        e_t = sa * v + s1 * x
        p0 = sa * x - s1 * v
        if a.use_dynamic_rescale:
            p0 = p0 * (a.scale_prev / a.scale_t)
        self._f(a.pred_x0, n).copy_(p0)
        dirc = math.sqrt(max(1.0 - a.ddim_alpha_prev - a.ddim_sigma ** 2, 0.0))
        self._f(a.x_prev, n).copy_(math.sqrt(a.ddim_alpha_prev) * p0 + dirc * e_t + a.ddim_sigma * a.temperature * noise)
        return 0

    def gvd_ddim_pred_x0_vjp(self, args, stream):
        """Closed form of csrc/nn_backward.cu::ddim_vjp_* (checked against autograd in tests/test_guided_cpu.py)."""
        self._count("ddim_vjp")
        a = args._obj
        n = a.n
        assert a.scratch and a.scratch_bytes >= 64
        e_c, G = self._f(a.e_cond, n).double(), self._f(a.grad_pred_x0, n).double()
        r = (a.scale_prev / a.scale_t) if a.use_dynamic_rescale else 1.0
        gv, gx = -r * a.sqrt_one_minus_alphas_cumprod_t, r * a.sqrt_alphas_cumprod_t
        self._f(a.dx, n).copy_((gx * G).float())
        if not a.e_uncond:
            self._f(a.de_cond, n).copy_((gv * G).float())
            return 0
        e_u = self._f(a.e_uncond, n).double()
        s, phi = a.cfg_scale, a.guidance_rescale
        mo = e_u + s * (e_c - e_u)
        factor, kc, km = 1.0, 0.0, 0.0
        if phi > 0:
            sd_c, sd_m = e_c.std(), mo.std()
            A = (gv * G * mo).sum()
            factor = phi * sd_c / sd_m + (1 - phi)
            kc = A * phi / ((n - 1) * sd_c * sd_m)
            km = -A * phi * sd_c / ((n - 1) * sd_m ** 3)
        dmo = gv * G * factor + km * (mo - mo.mean())
        self._f(a.de_cond, n).copy_((s * dmo + kc * (e_c - e_c.mean())).float())
        self._f(a.de_uncond, n).copy_(((1 - s) * dmo).float())
        return 0
