"""Backward of the B200-native denoiser with respect to its INPUT latent: what the guided sampler needs
(lvdm/models/samplers/ddim_guidance.py:259-337 -- `x.requires_grad_(True)`, two U-Net forwards, then
`pred_x0.backward(gradient=accum_grad_loss2x0, inputs=x)`).  The reference leaves this to autograd over the ATen /
cuDNN kernels of `UNetModel`; here every operator of vc_b200.ops is a `torch.autograd.Function` whose forward is the
same sm_100a launch the inference path makes and whose backward is built from the library's input-gradient operators
(csrc/nn_backward.cu + the tensor-core GEMM against transposed weights).  torch.autograd only records the tape and
adds gradients where the graph forks (residual branches, skip connections); no arithmetic of a layer runs in torch.

Parameters are frozen in the guidance loop, so no weight gradients exist here: a Function returns gradients for its
activation inputs only, and saves only what its input-gradient needs (norms: their input; GEGLU: its input; attention:
q, k, v; linear / conv layers: nothing but a reference to the weight).

`ops.<name>` dispatches here when autograd is on and an activation requires grad (ops._wants_grad); inside a
Function.forward grad mode is off, so the nested `ops.<name>` call takes the plain inference path.
"""
import torch
from torch.autograd import Function

from . import ops


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


class Linear(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act, residual, out_dtype, alpha, bias2):
        if act != "none":
            raise NotImplementedError("vc_b200.grad.Linear: fused activations sit on the embedding path, which carries no gradient")
        ctx.weight, ctx.alpha, ctx.has_res = weight, alpha, residual is not None
        ctx.x_dtype = x.dtype
        return ops.linear(x, weight, bias, act, residual, out_dtype, alpha, bias2)

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        dx = ops.linear_dx(dy.to(ctx.x_dtype), ctx.weight, ctx.alpha) if ctx.needs_input_grad[0] else None
        dres = dy if (ctx.has_res and ctx.needs_input_grad[4]) else None
        return dx, None, None, None, dres, None, None, None


class GroupNorm(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, F, S, groups, eps, silu):
        y, stats = ops.groupnorm_with_stats(x, gamma, beta, F, S, groups, eps, silu)  # the backward reuses the statistics
        ctx.save_for_backward(x, stats)
        ctx.args = (gamma, beta, F, S, groups, eps, silu)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats = ctx.saved_tensors
        gamma, beta, F, S, groups, eps, silu = ctx.args
        dx = ops.groupnorm_bwd(x, _c(dy), gamma, beta, F, S, groups, eps, silu, stats=stats).view_as(x)
        return dx, None, None, None, None, None, None, None


class GroupNormSharded(Function):
    """GroupNorm whose rows are spread over the ranks of a FramePartition (TemporalConvBlock / TemporalTransformer norms
    under the frame-sharded plan): the backward needs the mirror-image exchanges (two small all-reduces)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, F, S_local, S_total, part, groups, eps, silu):
        ctx.save_for_backward(x)
        ctx.args = (gamma, beta, F, S_local, S_total, part, groups, eps, silu)
        return ops.groupnorm_sharded(x, gamma, beta, F, S_local, S_total, part, groups, eps, silu)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dx = ops.groupnorm_sharded_bwd(x, _c(dy), *ctx.args).view_as(x)
        return (dx,) + (None,) * 9


class ToPixels(Function):
    """FramePartition.to_pixels with its adjoint: the all-to-all back (a re-sharding is a permutation)."""

    @staticmethod
    def forward(ctx, x, part):
        ctx.part, ctx.S = part, x.shape[1]
        return part.to_pixels(x)

    @staticmethod
    def backward(ctx, dy):
        return ctx.part.to_frames(_c(dy), ctx.S), None


class ToFrames(Function):
    @staticmethod
    def forward(ctx, y, part, S):
        ctx.part = part
        return part.to_frames(y, S)

    @staticmethod
    def backward(ctx, dx):
        return ctx.part.to_pixels(_c(dx)), None, None


class LayerNorm(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        ctx.save_for_backward(x)
        ctx.args = (gamma, eps)
        return ops.layernorm(x, gamma, beta, eps)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        gamma, eps = ctx.args
        return ops.layernorm_bwd(x, _c(dy), gamma, eps).view_as(x), None, None, None


class Geglu(Function):
    @staticmethod
    def forward(ctx, h):
        ctx.save_for_backward(h)
        return ops.geglu(h)

    @staticmethod
    def backward(ctx, dout):
        (h,) = ctx.saved_tensors
        return ops.geglu_bwd(h, _c(dout)).view_as(h)


class Conv3x3(Function):
    @staticmethod
    def forward(ctx, x, F, H, W, weight, bias, stride, upsample, bias2, residual, act):
        if act != "none":
            raise NotImplementedError("vc_b200.grad.Conv3x3: no fused activation on the differentiated path")
        ctx.weight, ctx.geom, ctx.has_res = weight, (F, H, W, x.shape[-1], stride, upsample), residual is not None
        y, _, _ = ops.conv3x3(x, F, H, W, weight, bias, stride, upsample, bias2, residual, act)
        return y

    @staticmethod
    def backward(ctx, dy):
        F, H, W, Cin, stride, upsample = ctx.geom
        dy = _c(dy)
        dx = ops.conv3x3_dx(dy, F, H, W, Cin, ctx.weight, stride, upsample) if ctx.needs_input_grad[0] else None
        dres = dy if (ctx.has_res and ctx.needs_input_grad[9]) else None
        return dx, None, None, None, None, None, None, None, None, dres, None


class ConvT3(Function):
    @staticmethod
    def forward(ctx, x, B, T, S, weight, bias, residual):
        ctx.weight, ctx.geom, ctx.has_res = weight, (B, T, S, x.shape[-1]), residual is not None
        return ops.conv_t3(x, B, T, S, weight, bias, residual)

    @staticmethod
    def backward(ctx, dy):
        B, T, S, Cin = ctx.geom
        dy = _c(dy)
        dx = ops.conv_t3_dx(dy, B, T, S, Cin, ctx.weight) if ctx.needs_input_grad[0] else None
        dres = dy if (ctx.has_res and ctx.needs_input_grad[6]) else None
        return dx, None, None, None, None, None, dres


class TemporalAttention(Function):
    @staticmethod
    def forward(ctx, q, k, v, B, T, S, H, scale):
        ctx.save_for_backward(q, k, v)
        ctx.args = (B, T, S, H, scale)
        return ops.temporal_attention(q, k, v, B, T, S, H, scale)

    @staticmethod
    def backward(ctx, dout):
        q, k, v = ctx.saved_tensors
        dq, dk, dv = ops.temporal_attention_bwd(q, k, v, _c(dout), *ctx.args)
        return dq.view_as(q), dk.view_as(k), dv.view_as(v), None, None, None, None, None


class FlashAttention(Function):
    """Fused attention with its fused adjoint (csrc/attn_bwd_tc.cu): the forward keeps the output and one fp32 per query
    row; GVD_FLASH_BWD=0 restores the first backward, which re-materialises the scores."""

    @staticmethod
    def forward(ctx, q, k, v, Bq, Nq, Nk, H, scale, shared_kv):
        ctx.args = (Bq, Nq, Nk, H, scale, shared_kv)
        ctx.fused = ops.FUSED_FLASH_BWD
        if ctx.fused:
            out, lse = ops.flash_attention_lse(q, k, v, Bq, Nq, Nk, H, scale, shared_kv)
            ctx.save_for_backward(q, k, v, out, lse)
            return out
        ctx.save_for_backward(q, k, v)
        return ops.flash_attention(q, k, v, Bq, Nq, Nk, H, scale, shared_kv)

    @staticmethod
    def backward(ctx, dout):
        Bq, Nq, Nk, H, scale, shared_kv = ctx.args
        need_kv = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        if ctx.fused:
            q, k, v, out, lse = ctx.saved_tensors
            dq, dk, dv = ops.flash_attention_bwd(q, k, v, out, lse, _c(dout), Bq, Nq, Nk, H, scale, shared_kv, need_kv)
        else:
            q, k, v = ctx.saved_tensors
            dq, dk, dv = ops.attention_bwd(q, k, v, _c(dout), Bq, Nq, Nk, H, scale, shared_kv, need_kv)
        return (dq.view_as(q), dk.view_as(k) if dk is not None else None, dv.view_as(v) if dv is not None else None,
                None, None, None, None, None, None)


class MaterialisedAttention(Function):
    """`ops.attention` (scores materialised, any head dim): the VAE decoder's single-head 512-channel AttnBlock."""

    @staticmethod
    def forward(ctx, q, k, v, Bq, Nq, Nk, H, scale, shared_kv, head_dim):
        ctx.save_for_backward(q, k, v)
        ctx.args = (Bq, Nq, Nk, H, scale, shared_kv, head_dim)
        return ops.attention(q, k, v, Bq, Nq, Nk, H, scale, shared_kv, head_dim=head_dim)

    @staticmethod
    def backward(ctx, dout):
        q, k, v = ctx.saved_tensors
        Bq, Nq, Nk, H, scale, shared_kv, head_dim = ctx.args
        need_kv = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dq, dk, dv = ops.attention_bwd(q, k, v, _c(dout), Bq, Nq, Nk, H, scale, shared_kv, need_kv, head_dim=head_dim)
        return (dq.view_as(q), dk.view_as(k) if dk is not None else None, dv.view_as(v) if dv is not None else None,
                None, None, None, None, None, None, None)
