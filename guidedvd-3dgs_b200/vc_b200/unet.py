"""B200-native forward of the ViewCrafter 3-D U-Net denoiser.

Same network as third_party/ViewCrafter/lvdm/modules/networks/openaimodel3d.py::UNetModel (+ attention.py) for the
configuration the reference runs (configs/inference_pvd_1024.yaml:33-64): 2-D ResBlocks with TemporalConvBlocks,
SpatialTransformers (self + text/image cross attention, GEGLU FF), TemporalTransformers (two self-attentions over the
frame axis), fps embedding, `addition_attention`.  Parameters are loaded from a reference state_dict (same key names,
so ViewCrafter checkpoints load unchanged) and repacked once: bf16 [N, K] matrices in the K order of the
channels-last im2col, fp32 biases / norm affine.

Activations are channels-last bf16 [frames, pixels, channels]; every arithmetic operator is a launch of the sm_100a
library (vc_b200.ops); bf16 rounding points follow `torch.autocast(bfloat16)` over the reference modules, which is
how the parity tests run the reference (the reference itself uses autocast fp16, VC/viewcrafter.py:102).
"""
import math

import torch

from . import ops

BF16 = torch.bfloat16


_freqs = {}


def timestep_embedding(timesteps, dim, max_period=10000):
    """lvdm/models/utils_diffusion.py:8-28 (host-side scalar plumbing: b x dim floats).  The frequency table is computed
    on the CPU exactly like the reference's (same libm exp) and kept per device: no host-to-device copy per call, which
    also keeps the forward capturable into a CUDA graph."""
    half = dim // 2
    freqs = _freqs.get((dim, max_period, timesteps.device))
    if freqs is None:
        freqs = _freqs[(dim, max_period, timesteps.device)] = torch.exp(
            -math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class _P:
    """Parameter access by reference key prefix, with repacking helpers."""

    def __init__(self, sd, device):
        self.sd, self.dev = sd, device

    def has(self, key):
        return key in self.sd

    def f32(self, key):
        return self.sd[key].detach().to(self.dev, torch.float32).contiguous()

    def lin(self, key):
        w = self.sd[key + ".weight"].detach().to(self.dev)
        w = w.reshape(w.shape[0], -1)  # Linear [N,K], 1x1 Conv2d/Conv1d [N,K,1(,1)]
        b = self.f32(key + ".bias") if (key + ".bias") in self.sd else None
        return w.to(BF16).contiguous(), b

    def conv3x3(self, key):
        w = self.sd[key + ".weight"].detach().to(self.dev)  # [Cout, Cin, 3, 3] -> [Cout, (ky, kx, cin)]
        w = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)
        return w.to(BF16).contiguous(), self.f32(key + ".bias")

    def conv_t3(self, key):
        w = self.sd[key + ".weight"].detach().to(self.dev)  # [Cout, Cin, 3, 1, 1] -> [Cout, (kt, cin)]
        w = w[:, :, :, 0, 0].permute(0, 2, 1).reshape(w.shape[0], -1)
        return w.to(BF16).contiguous(), self.f32(key + ".bias")

    def norm(self, key):
        return self.f32(key + ".weight"), self.f32(key + ".bias")


class ResBlock:
    """openaimodel3d.py:109-236 (+ TemporalConvBlock :239-279)."""

    def __init__(self, p, pre, temporal):
        self.gn1 = p.norm(pre + ".in_layers.0")
        self.conv1 = p.conv3x3(pre + ".in_layers.2")
        self.emb = p.lin(pre + ".emb_layers.1")
        self.gn2 = p.norm(pre + ".out_layers.0")
        self.conv2 = p.conv3x3(pre + ".out_layers.3")
        self.skip = p.lin(pre + ".skip_connection") if p.has(pre + ".skip_connection.weight") else None
        self.tconv = None
        if temporal and p.has(pre + ".temopral_conv.conv1.0.weight"):
            t = pre + ".temopral_conv"
            self.tconv = [(p.norm(f"{t}.conv{i}.0"), p.conv_t3(f"{t}.conv{i}.{2 if i == 1 else 3}")) for i in (1, 2, 3, 4)]

    def __call__(self, x, F, H, W, emb_silu, B, T=None, part=None):
        S = H * W
        h = ops.groupnorm(x, *self.gn1, F, S, eps=1e-5, silu=1)
        emb_out = ops.linear(emb_silu, *self.emb).float().reshape(-1).contiguous()  # [Cout], bf16-rounded values
        h, _, _ = ops.conv3x3(h, F, H, W, *self.conv1, bias2=emb_out)
        h = ops.groupnorm(h, *self.gn2, F, S, eps=1e-5, silu=1)
        skip = x if self.skip is None else ops.linear(x, *self.skip)
        h, _, _ = ops.conv3x3(h, F, H, W, *self.conv2, residual=skip)
        if self.tconv is not None:
            T = F // B if T is None else T
            sharded = part is not None and part.active
            if sharded:  # frames -> pixels: every pixel of this rank's slice sees all T frames
                h = part.to_pixels(h)
            Sl = h.shape[1]
            ident = h
            for i, (gn, conv) in enumerate(self.tconv):
                # plain nn.GroupNorm: fp32 out under autocast, SiLU in fp32, one rounding at the conv input
                if sharded:
                    g = ops.groupnorm_sharded(h, *gn, B, T * Sl, T * S, part, eps=1e-5, silu=2)
                else:
                    g = ops.groupnorm(h, *gn, B, T * S, eps=1e-5, silu=2)
                h = ops.conv_t3(g, B, T, Sl, *conv, residual=ident if i == 3 else None)
            if sharded:
                h = part.to_frames(h, S)
        return h


class _Attn:
    def __init__(self, p, pre, image_cross=False):
        self.q, _ = p.lin(pre + ".to_q")
        self.k, _ = p.lin(pre + ".to_k")
        self.v, _ = p.lin(pre + ".to_v")
        self.o = p.lin(pre + ".to_out.0")
        self.k_ip = self.v_ip = None
        if image_cross and p.has(pre + ".to_k_ip.weight"):
            self.k_ip, _ = p.lin(pre + ".to_k_ip")
            self.v_ip, _ = p.lin(pre + ".to_v_ip")
        self.heads = self.q.shape[0] // 64
        self.scale = 64 ** -0.5


class _TBlock:
    """BasicTransformerBlock parameters (attention.py:212-246)."""

    def __init__(self, p, pre, image_cross):
        self.attn1 = _Attn(p, pre + ".attn1")
        self.attn2 = _Attn(p, pre + ".attn2", image_cross)
        self.n1, self.n2, self.n3 = p.norm(pre + ".norm1"), p.norm(pre + ".norm2"), p.norm(pre + ".norm3")
        self.ff1 = p.lin(pre + ".ff.net.0.proj")
        self.ff1_il = ops.geglu_weight(*self.ff1)  # rows interleaved for the fused GEGLU epilogue (inference path)
        self.ff2 = p.lin(pre + ".ff.net.2")

    def ff(self, h):
        a = ops.layernorm(h, *self.n3)
        if self.ff1_il is not None and ops.FUSED_GEGLU and not ops._wants_grad(a):
            g = ops.linear_geglu(a, *self.ff1_il)
        else:
            g = ops.geglu(ops.linear(a, *self.ff1))
        return ops.linear(g, *self.ff2, residual=h)


class SpatialTransformer:
    """attention.py:249-310 with use_linear=True; CrossAttention.forward :81-144 (einsum path)."""

    def __init__(self, p, pre):
        self.norm = p.norm(pre + ".norm")
        self.proj_in = p.lin(pre + ".proj_in")
        self.blk = _TBlock(p, pre + ".transformer_blocks.0", True)
        self.proj_out = p.lin(pre + ".proj_out")

    def __call__(self, x, F, S, ctx_text, ctx_img):
        b = self.blk
        h = ops.linear(ops.groupnorm(x, *self.norm, F, S, eps=1e-6, silu=0), *self.proj_in)
        # self attention, per frame
        a = ops.layernorm(h, *b.n1)
        at = b.attn1
        o = ops.flash_attention(ops.linear(a, at.q), ops.linear(a, at.k), ops.linear(a, at.v), F, S, S, at.heads, at.scale)
        h = ops.linear(o, *at.o, residual=h)
        # cross attention: 77 text tokens + 256 image tokens, identical for every frame (openaimodel3d.py:555-562)
        a = ops.layernorm(h, *b.n2)
        at = b.attn2
        q = ops.linear(a, at.q)
        o = ops.flash_attention(q, ops.linear(ctx_text, at.k), ops.linear(ctx_text, at.v), F, S, ctx_text.shape[1], at.heads, at.scale,
                          shared_kv=True)
        if at.k_ip is not None:
            o_ip = ops.flash_attention(q, ops.linear(ctx_img, at.k_ip), ops.linear(ctx_img, at.v_ip), F, S, ctx_img.shape[1], at.heads,
                                 at.scale, shared_kv=True)
            o = o + o_ip  # image_cross_attention_scale = 1.0, not learnable (attention.py:141-142)
        h = ops.linear(o, *at.o, residual=h)
        h = b.ff(h)
        return ops.linear(h, *self.proj_out, residual=x)


class TemporalTransformer:
    """attention.py:313-412 with use_linear=True, only_self_att=True, no relative position, no causal mask."""

    def __init__(self, p, pre):
        self.norm = p.norm(pre + ".norm")
        self.proj_in = p.lin(pre + ".proj_in")
        self.blk = _TBlock(p, pre + ".transformer_blocks.0", False)
        self.proj_out = p.lin(pre + ".proj_out")

    def __call__(self, x, B, T, S, part=None):
        b = self.blk
        sharded = part is not None and part.active
        if sharded:  # the whole block runs pixel-sharded: attention over t needs every frame of a pixel
            x = part.to_pixels(x)
            Sl = x.shape[1]
            n = ops.groupnorm_sharded(x, *self.norm, B, T * Sl, T * S, part, eps=1e-6, silu=0)
        else:
            Sl = S
            n = ops.groupnorm(x, *self.norm, B, T * S, eps=1e-6, silu=0)
        h = ops.linear(n, *self.proj_in)
        for at, n in ((b.attn1, b.n1), (b.attn2, b.n2)):
            a = ops.layernorm(h, *n)
            o = ops.temporal_attention(ops.linear(a, at.q), ops.linear(a, at.k), ops.linear(a, at.v), B, T, Sl, at.heads, at.scale)
            h = ops.linear(o, *at.o, residual=h)
        h = b.ff(h)
        y = ops.linear(h, *self.proj_out, residual=x)
        return part.to_frames(y, S) if sharded else y


class UNetB200:
    def __init__(self, state_dict, device="cuda", in_channels=8, model_channels=320, out_channels=4, num_res_blocks=2,
                 attention_resolutions=(4, 2, 1), channel_mult=(1, 2, 4, 4), temporal_conv=True, addition_attention=True,
                 fs_condition=True, default_fs=10, **_ignored):
        p = _P(state_dict, device)
        self.dev = device
        self.model_channels, self.default_fs, self.fs_condition = model_channels, default_fs, fs_condition
        self.time_embed = (p.lin("time_embed.0"), p.lin("time_embed.2"))
        self.fps_embedding = (p.lin("fps_embedding.0"), p.lin("fps_embedding.2")) if fs_condition else None
        # ---- same construction walk as UNetModel.__init__ (openaimodel3d.py:388-543), by key prefix ----
        self.input_blocks = [[("conv", p.conv3x3("input_blocks.0.0"))]]
        self.init_attn = TemporalTransformer(p, "init_attn.0") if addition_attention else None
        idx, ds = 1, 1
        for level, _mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                pre = f"input_blocks.{idx}"
                layers = [("res", ResBlock(p, pre + ".0", temporal_conv))]
                if ds in attention_resolutions:
                    layers += [("st", SpatialTransformer(p, pre + ".1")), ("tt", TemporalTransformer(p, pre + ".2"))]
                self.input_blocks.append(layers)
                idx += 1
            if level != len(channel_mult) - 1:
                self.input_blocks.append([("down", p.conv3x3(f"input_blocks.{idx}.0.op"))])
                idx += 1
                ds *= 2
        self.middle = [("res", ResBlock(p, "middle_block.0", temporal_conv)), ("st", SpatialTransformer(p, "middle_block.1")),
                       ("tt", TemporalTransformer(p, "middle_block.2")), ("res", ResBlock(p, "middle_block.3", temporal_conv))]
        self.output_blocks = []
        idx = 0
        for level, _mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                pre = f"output_blocks.{idx}"
                layers = [("res", ResBlock(p, pre + ".0", temporal_conv))]
                j = 1
                if ds in attention_resolutions:
                    layers += [("st", SpatialTransformer(p, pre + ".1")), ("tt", TemporalTransformer(p, pre + ".2"))]
                    j = 3
                if level and i == num_res_blocks:
                    layers.append(("up", p.conv3x3(f"{pre}.{j}.conv")))
                    ds //= 2
                self.output_blocks.append(layers)
                idx += 1
        self.out_norm = p.norm("out.0")
        self.out_conv = p.conv3x3("out.2")

    # ------------------------------------------------------------------------------------------------------------
    def _mlp(self, v, mlp):
        (w0, b0), (w1, b1) = mlp
        return ops.linear(ops.linear(v.to(BF16), w0, b0, act="silu"), w1, b1)

    def _run(self, layers, h, st):
        for kind, m in layers:
            if kind == "res":
                h = m(h, st["F"], st["H"], st["W"], st["emb_silu"], st["B"], st["T"], st["part"])
            elif kind == "st":
                h = m(h, st["F"], st["H"] * st["W"], st["ctx_text"], st["ctx_img"])
            elif kind == "tt":
                h = m(h, st["B"], st["T"], st["H"] * st["W"], st["part"])
            elif kind == "conv":
                h, _, _ = ops.conv3x3(h, st["F"], st["H"], st["W"], *m)
            elif kind == "down":
                h, st["H"], st["W"] = ops.conv3x3(h, st["F"], st["H"], st["W"], *m, stride=2)
            elif kind == "up":
                h, st["H"], st["W"] = ops.conv3x3(h, st["F"], st["H"], st["W"], *m, upsample=True)
        return h

    part = None   # vc_b200.frame_parallel.FramePartition: this rank's frame slice (None / inactive = all frames here)
    trace = None  # set to a list to record (name, activation[F, S, C]) after every top-level block (tests/debug)

    def _rec(self, name, h, st):
        if self.trace is not None:
            self.trace.append((name, h.view(st["F"], st["H"], st["W"], -1).permute(0, 3, 1, 2).float()))

    def forward(self, x, timesteps, context, fs=None):
        """x [1, C_in, t, h, w] fp32; timesteps [1]; context [1, 77+256, 1024]; fs [1] -> [1, C_out, t, h, w] bf16.
        With an active `self.part` the input is still the full clip; the result holds this rank's frames only
        ([1, C_out, F_r, h, w]); DiffusionModelB200 gathers them.  Inference: no graph is recorded."""
        with torch.no_grad():
            return self._forward(x, timesteps, context, fs)

    def forward_with_grad(self, x, timesteps, context, fs=None):
        """Same network with the tape on: when `x.requires_grad`, every operator records its input-gradient
        (vc_b200.grad), so `y.backward(gradient=g, inputs=x)` yields dL/dx -- the call the guided sampler makes
        (ddim_guidance.py:264-265,309).  Under a frame partition the input is the full clip and x.grad comes back non-zero
        in this rank's frames only: the temporal layers' re-shardings and sharded GroupNorms record their adjoint
        exchanges, the caller sums x.grad over the ranks (vc_b200.guided.GuidedPlan)."""
        with torch.enable_grad():
            return self._forward(x, timesteps, context, fs)

    def _forward(self, x, timesteps, context, fs=None):
        b, cin, t, hh, ww = x.shape
        if b != 1:
            return torch.cat([self._forward(x[i:i + 1], timesteps[i:i + 1], context[i:i + 1], None if fs is None else fs[i:i + 1])
                              for i in range(b)])
        emb = self._mlp(timestep_embedding(timesteps, self.model_channels), self.time_embed)
        if self.fs_condition:
            if fs is None:
                fs = torch.tensor([self.default_fs] * b, dtype=torch.long, device=x.device)
            emb = emb + self._mlp(timestep_embedding(fs, self.model_channels), self.fps_embedding)
        emb_silu = torch.nn.functional.silu(emb)  # [1, 4*mc] bf16: the input of every ResBlock's emb_layers
        ctx = context.to(BF16)
        part = self.part if (self.part is not None and self.part.active) else None
        if part is not None:
            assert part.T == t, "frame partition was built for a different clip length"
            x = x[:, :, part.frame_slice()]
        fl = x.shape[2]
        st = dict(B=b, F=b * fl, T=t, H=hh, W=ww, emb_silu=emb_silu, ctx_text=ctx[:, :77].contiguous(),
                  ctx_img=ctx[:, 77:].contiguous(), part=part)
        h = x.permute(0, 2, 3, 4, 1).reshape(b * fl, hh * ww, cin).to(BF16).contiguous()
        hs = []
        for i, layers in enumerate(self.input_blocks):
            h = self._run(layers, h, st)
            self._rec(f"input_blocks.{i}", h, st)
            if i == 0 and self.init_attn is not None:
                h = self.init_attn(h, b, t, st["H"] * st["W"], part)
                self._rec("init_attn", h, st)
            hs.append(h)
        h = self._run(self.middle, h, st)
        self._rec("middle_block", h, st)
        for i, layers in enumerate(self.output_blocks):
            h = torch.cat([h, hs.pop()], dim=-1)
            h = self._run(layers, h, st)
            self._rec(f"output_blocks.{i}", h, st)
        # `h = h.type(x.dtype)`: the last norm/SiLU run in fp32, one rounding at the conv input (openaimodel3d.py:598-599)
        h = ops.groupnorm(h, *self.out_norm, st["F"], st["H"] * st["W"], eps=1e-5, silu=2)
        y, _, _ = ops.conv3x3(h, st["F"], st["H"], st["W"], *self.out_conv)
        return y.view(b, fl, st["H"], st["W"], -1).permute(0, 4, 1, 2, 3).contiguous()

    __call__ = forward


class _GraphedForward:
    """One CUDA graph of `UNetB200.forward` for fixed input shapes: ~1 300 kernel launches (and, under a frame-sharded
    plan, its NCCL exchanges) replayed by one cudaGraphLaunch.  A forward at the 576x1024 shape costs ~155 ms of Python +
    ctypes on the host for ~170 ms of kernels on one GPU -- hidden there, but with the frames sharded eight ways the
    kernels shrink to ~45 ms and the host becomes the step (strong-scaling efficiency 0.65 at N = 8 in round 1).
    Inputs are copied into static buffers; the output is cloned out (the next replay overwrites it)."""

    def __init__(self, unet, xc, t, cc, fs):
        self.xc, self.t, self.cc, self.fs = xc.clone(), t.clone(), cc.clone(), fs.clone()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=xc.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):  # lazy initialisation (function attributes, scratch caches) must not be captured
            for _ in range(2):
                unet(self.xc, self.t, self.cc, fs=self.fs)
        cur.wait_stream(side)
        torch.cuda.synchronize(xc.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = unet(self.xc, self.t, self.cc, fs=self.fs)

    def __call__(self, xc, t, cc, fs):
        self.xc.copy_(xc)
        self.t.copy_(t)
        self.cc.copy_(cc)
        self.fs.copy_(fs)
        self.graph.replay()
        return self.out.clone()


class _GraphedGrad:
    """The differentiated U-Net call of the guided sampler (forward WITH tape, backward to the input latent) as CUDA graphs.
    Eagerly a guided step issues ~8 300 launches and needs ~460 ms of Python to do so -- as long as its kernels run
    (profiles/r02_guided_kernel_breakdown.txt); ~7 000 of them are the two U-Net forwards and their backwards, whose shapes
    never change.  torch.cuda.make_graphed_callables captures one forward graph and one backward graph per slot (static
    inputs, saved activations in the graph's private pool).  A guided step holds two forwards (cond / uncond) before it
    runs their backwards, so there are two slots; a slot is taken by a forward and released when its backward starts (a
    hook on the output's gradient).  A third forward while both are taken, or any capture failure, runs eagerly."""

    def __init__(self, unet, slots=2):
        self.unet, self.max_slots, self.slots, self.error = unet, slots, [], None

    def _make(self, xc, t, cc, fs):
        unet = self.unet

        def fn(xc_, t_, cc_, fs_):
            return unet.forward_with_grad(xc_, t_, cc_, fs=fs_).float()

        sample = (xc.detach().clone().requires_grad_(True), t.clone(), cc.detach().clone(), fs.clone())
        return {"fn": torch.cuda.make_graphed_callables(fn, sample), "busy": False}

    def __call__(self, xc, t, cc, fs):
        slot = next((s for s in self.slots if not s["busy"]), None)
        if slot is None and len(self.slots) < self.max_slots and self.error is None:
            try:
                slot = self._make(xc, t, cc, fs)
                self.slots.append(slot)
            except Exception as ex:  # same kernels either way: keep running eagerly and say why
                import warnings
                self.error = repr(ex)[:300]
                warnings.warn("vc_b200: CUDA-graph capture of the differentiated U-Net failed, running it eagerly: " + self.error)
                torch.cuda.synchronize(xc.device)
        if slot is None:
            return self.unet.forward_with_grad(xc, t, cc, fs=fs).float()
        y = slot["fn"](xc, t, cc, fs)
        slot["busy"] = True

        def release(grad, slot=slot):
            slot["busy"] = False
            return None

        y.register_hook(release)
        return y


class DiffusionModelB200:
    """The slice of the reference's LatentDiffusion object the sampler needs: `apply_model` with the 'hybrid'
    conditioning of DiffusionWrapper.forward (lvdm/models/ddpm3d.py:1426-1443): channel-concat c_concat, cross-attend
    to cat(c_crossattn)."""

    def __init__(self, unet, schedule, plan=None, use_graph=None):
        """plan: vc_b200.frame_parallel.DenoisePlan (one process per GPU) or None for a single GPU.
        use_graph: replay the inference forward as a CUDA graph (default: GVD_UNET_GRAPH, on)."""
        import os
        self.unet, self.schedule, self.plan = unet, schedule, plan
        if plan is not None:
            unet.part = plan.part
        self.use_graph = (os.environ.get("GVD_UNET_GRAPH", "1") != "0") if use_graph is None else bool(use_graph)
        self._graphs, self.graph_error = {}, None
        # the guided sampler's differentiated call as CUDA graphs (GVD_GUIDED_GRAPH, on; single-GPU plans only: the
        # frame-sharded tape carries NCCL exchanges)
        self.use_grad_graph = os.environ.get("GVD_GUIDED_GRAPH", "1") != "0"
        self._grad_graphs = {}

    def _local(self, x, t, cond, fs):
        xc = torch.cat([x] + list(cond["c_concat"]), dim=1)
        cc = torch.cat(list(cond["c_crossattn"]), dim=1)
        if torch.is_grad_enabled() and x.requires_grad:  # the guided sampler differentiates through this call
            if (self.use_grad_graph and xc.is_cuda and xc.shape[0] == 1 and self.unet.trace is None
                    and (self.plan is None or self.plan.world == 1)):
                if fs is None:
                    fs = torch.full((1,), int(self.unet.default_fs), dtype=torch.long, device=xc.device)
                key = (tuple(xc.shape), tuple(cc.shape), xc.dtype, cc.dtype, t.dtype, fs.dtype)
                g = self._grad_graphs.get(key)
                if g is None:
                    g = self._grad_graphs[key] = _GraphedGrad(self.unet)
                return g(xc, t, cc, fs)
            return self.unet.forward_with_grad(xc, t, cc, fs=fs).float()
        if self.use_graph and xc.is_cuda and xc.shape[0] == 1 and self.unet.trace is None:
            if fs is None:
                fs = torch.full((1,), int(self.unet.default_fs), dtype=torch.long, device=xc.device)
            key = (tuple(xc.shape), tuple(cc.shape), xc.dtype, cc.dtype, t.dtype, fs.dtype)
            g = self._graphs.get(key)
            if g is None:
                try:
                    g = self._graphs[key] = _GraphedForward(self.unet, xc, t, cc, fs)
                except Exception as ex:  # same kernels either way: keep running eagerly and say why
                    import warnings
                    self.use_graph, self.graph_error = False, repr(ex)[:300]
                    warnings.warn("vc_b200: CUDA-graph capture of the U-Net forward failed, running it eagerly: " + self.graph_error)
                    torch.cuda.synchronize(xc.device)
                    return self.unet(xc, t, cc, fs=fs).float()
            return g(xc, t, cc, fs).float()
        return self.unet(xc, t, cc, fs=fs).float()

    def apply_model(self, x, t, cond, fs=None, **kwargs):
        y = self._local(x, t, cond, fs)
        if self.plan is None or self.plan.world == 1:
            return y
        if self.plan.cfg_ways == 1:
            return self.plan.gather_outputs(y)[0]
        raise RuntimeError("CFG-split plan: call apply_model_cfg (each rank evaluates one of cond / uncond)")

    def apply_model_cfg(self, x, t, cond, uncond, fs=None, **kwargs):
        """Both forwards of a classifier-free-guidance step (ddim.py:222-223) -> (e_cond, e_uncond), full clips on every
        rank.  Under a CFG-split plan each rank runs one of them on its frame slice."""
        plan = self.plan
        if plan is None or plan.world == 1:
            return self._local(x, t, cond, fs), self._local(x, t, uncond, fs)
        if plan.cfg_ways == 1:
            return plan.gather_outputs(self._local(x, t, cond, fs))[0], plan.gather_outputs(self._local(x, t, uncond, fs))[0]
        e_c, e_u = plan.gather_outputs(self._local(x, t, cond if plan.cfg_index == 0 else uncond, fs))
        return e_c, e_u
