"""Loss-guided DDIM step over the B200-native denoiser -- the call surface of the reference's
lvdm/models/samplers/ddim_guidance.py::DDIMSamplerGuidance.p_sample_ddim (:205-362, Algorithm 1 of the paper) for the
options the guidedvd pipeline uses (VC/utils_vc/diffusion_utils.py: v-prediction, CFG 7.5, guidance_rescale 0.7,
eta 1.0, dynamic rescale, `loss_guidance_fn` = utils/viewcrafter_wrapper.py::LossGuidance, recur_steps 1 or 2).

One guided step:
  1. x.requires_grad_(True); e_cond, e_uncond = U-Net(x) with the tape on (vc_b200.grad; both graphs are kept, like
     the reference keeps them -- rescale_noise_cfg couples the two outputs through their standard deviations);
  2. x_prev, pred_x0 = the fused DDIM update (gvd_ddim_step, same arithmetic as the plain sampler);
  3. per frame: decode pred_x0 through the VAE decoder with grad, evaluate the guidance loss against the 3DGS
     rendering, G_f = d loss / d pred_x0_f (/ numel unless mean_loss)                         (:282-302)
  4. dL/dx = gvd_ddim_pred_x0_vjp (explicit x term, cotangents of e_cond / e_uncond) + the U-Net backward of both
     forwards                                                                                  (:309-311)
  5. rho = rms(e_cond - e_uncond) * cfg / rms(dL/dx) * 0.2 * scale_guidance_weight; x_prev -= rho * dL/dx (:318-326)
  6. (recurrence, :334) x = sqrt(beta_t) x_prev + sqrt(1 - beta_t) N(0, 1), beta_t = a_t / a_prev.

`model` protocol beyond vc_b200.sampler.DDIMSampler's: `model.differentiable_decode_first_stage(z[1,4,n,h,w]) ->
[1,3,n,H,W]` (ddpm3d.py:674-675; vc_b200.vae.DecoderB200.differentiable_decode is the native one) and, optionally,
`model.guided_decode_frames` = how many frames to decode per call (default 1, the reference's loop; frames are a batch
dimension of the decoder, so any chunk size gives the same gradients).  `loss_guidance_fn` protocol: SURVEY.md 8b.

One process per GPU (`GuidedPlan`, SURVEY.md section 8e "Guided path"): the conditional and the unconditional U-Net
forward + backward are independent given the two cotangents, so the first half of the ranks evaluates `cond`, the second
half `uncond`, each half with its frames sharded as in the plain sampler's plan and the adjoint exchanges recorded for
the backward (one gather of the outputs before the coupled pred_x0 arithmetic, one all-reduce of dL/dx after the
backward); the 25 decoder forward + backward passes are independent frames and are dealt out over ALL ranks (one
gather of dL/dpred_x0 and of the decoded frames).  Noise is drawn on rank 0 and broadcast.  Everything else (fused
DDIM update, VJP, rho) is replicated -- it is a few passes over a 256 k-element latent.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .frame_parallel import split_sizes
from .sampler import DDIMSampler, _randn


class GuidedPlan:
    """World layout of one guided step (one process per GPU).

    U-Net forward + backward: the DenoisePlan layout of the plain sampler -- `cfg_ways` (2 when the world size is even)
    x `frame_ways` ranks; rank -> (branch: 0 = cond / 1 = uncond, frame slice).  Inside a branch the frames are sharded
    like at inference (FramePartition: all-to-all re-shard around the temporal layers, summed GroupNorm statistics),
    with the adjoint exchanges recorded for the backward (vc_b200.grad.ToPixels / ToFrames / GroupNormSharded).
    Decoder forward + backward: the frames dealt out over ALL ranks.
    Exchanges per step: one all-gather of the local U-Net outputs, one ragged gather of dL/dpred_x0 and of the decoded
    frames, one all-reduce of dL/dx (every (branch, frame slice) pair lives on exactly one rank)."""

    def __init__(self, n_frames, model=None):
        from .frame_parallel import DenoisePlan

        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.denoise = DenoisePlan(n_frames)
        self.part = self.denoise.part
        self.split_cfg = self.denoise.cfg_ways == 2
        self.branch = self.denoise.cfg_index if self.split_cfg else None
        self.frames = split_sizes(n_frames, self.world)       # decoder passes
        self.f0 = sum(self.frames[:self.rank])
        self.f1 = self.f0 + self.frames[self.rank]
        if model is not None:
            self.attach(model)

    def attach(self, model):
        """model: DiffusionModelB200.  Its U-Net runs frame-sharded from now on; the sampler finds the plan on the model."""
        model.unet.part = self.part
        model.plan = self.denoise      # un-guided steps on the same model take the plain sampler's sharded path
        model.guided_plan = self
        return self

    def unet_outputs(self, model, x, t, c, uc, **kwargs):
        """-> (graph-carrying local outputs [list, in cond/uncond order of what THIS rank evaluated], e_cond, e_uncond
        full clips without graph, identical on every rank)."""
        if self.split_cfg:
            mine = [model._local(x, t, c if self.branch == 0 else uc, kwargs.get("fs"))]
            e_c, e_u = self.denoise.gather_outputs(mine[0].detach().float().contiguous())
        else:  # odd world: both branches on this rank's frame slice
            mine = [model._local(x, t, c, kwargs.get("fs")), model._local(x, t, uc, kwargs.get("fs"))]
            e_c = self.denoise.gather_outputs(mine[0].detach().float().contiguous())[0]
            e_u = self.denoise.gather_outputs(mine[1].detach().float().contiguous())[0]
        return mine, e_c.contiguous(), e_u.contiguous()

    def backward(self, mine, de_c, de_u, x):
        """Cotangents of the full clips -> this rank's slices -> both tapes -> dL/dx summed over the world."""
        sl = self.part.frame_slice() if self.part.active else slice(None)
        cots = [de_c if self.branch == 0 else de_u] if self.split_cfg else [de_c, de_u]
        torch.autograd.backward(mine, [ct[:, :, sl].to(m.dtype).contiguous() for ct, m in zip(cots, mine)], inputs=[x])
        total = x.grad.detach().float().clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
        return total

    def randn(self, shape, device):
        z = torch.randn(shape, device=device)
        dist.broadcast(z, src=0)
        return z

    def gather_decoder_results(self, decoded, grads, pred_x0):
        """This rank's decoded frames and dL/dpred_x0 slices (lists of [1, C, n, H, W] chunks, possibly empty) -> the full
        clips on every rank.  A rank that owns no frame learns the image size from its peers (one 3-int all-reduce)."""
        dims = torch.tensor(list(decoded[0].shape[1:2]) + list(decoded[0].shape[3:]) if decoded else [0, 0, 0], dtype=torch.int64,
                            device=pred_x0.device)
        dist.all_reduce(dims, op=dist.ReduceOp.MAX)
        ci, H, W = (int(v) for v in dims.tolist())
        d_local = torch.cat(decoded, dim=2) if decoded else pred_x0.new_zeros(1, ci, 0, H, W)
        g_local = torch.cat(grads, dim=2).float() if grads else pred_x0.new_zeros(1, pred_x0.shape[1], 0, *pred_x0.shape[3:])
        return self.gather_frames(d_local), self.gather_frames(g_local).contiguous()

    def gather_frames(self, local, dim=2):
        """Concatenate per-rank frame slices (ragged: 25 frames over 8 ranks = 4,3,...,3) along `dim` on every rank."""
        fmax = max(self.frames)
        shape = list(local.shape)
        shape[dim] = fmax
        pad = torch.zeros(shape, dtype=local.dtype, device=local.device)
        pad.narrow(dim, 0, local.shape[dim]).copy_(local)
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(parts, pad)
        return torch.cat([p.narrow(dim, 0, n) for p, n in zip(parts, self.frames) if n > 0], dim=dim)


def _rms(t):
    return (t.float() * t.float()).mean().sqrt().item()


class DDIMSamplerGuidance(DDIMSampler):
    def p_sample_ddim(self, x, c, t, index, unconditional_guidance_scale=1., unconditional_conditioning=None,
                      guidance_rescale=0.0, temperature=1., noise=None, loss_guidance_fn=None, recur_noise=None, **kwargs):
        """noise / recur_noise: optional pre-drawn N(0,1) tensors (one per recurrence: lists or a single tensor) so a
        test can replay the reference's draws; drawn fresh otherwise, in the reference's order."""
        if loss_guidance_fn is None:
            return super().p_sample_ddim(x, c, t, index, unconditional_guidance_scale, unconditional_conditioning,
                                         guidance_rescale, temperature, noise, **kwargs)
        if x.shape[0] != 1:
            raise ValueError("guided sampling supports batch size 1 (ddim_guidance.py:246-247)")
        lg = loss_guidance_fn
        repeat = int(lg.recur_steps)
        if repeat not in (1, 2):
            raise ValueError("only support 1 or 2 recur steps (ddim_guidance.py:250-251)")
        sgw = 1.0
        if lg.scale_guidance_weight:
            sgw = lg.guidance_weight_fn(lg.current_train_iter)
        cfg = float(unconditional_guidance_scale)
        coef = self.ddim.coefficients(index, cfg, guidance_rescale, temperature)
        a_t, a_prev = np.float32(self.ddim.alphas[index]), np.float32(self.ddim.alphas_prev[index])
        beta_t = np.float32(a_t / a_prev)
        uc = unconditional_conditioning
        model = self.model
        n_frames = x.shape[2]
        gp = getattr(model, "guided_plan", None)
        if gp is not None and gp.world == 1:
            gp = None
        if gp is not None and uc is None:
            raise ValueError("the multi-GPU guided plan splits the cond / uncond pair: unconditional_conditioning is required")
        pick = lambda src, j: None if src is None else (src[j] if isinstance(src, (list, tuple)) else src)  # noqa: E731
        x_prev = pred_x0 = None
        for j in range(repeat):
            x = x.detach().float().requires_grad_(True)
            mine = None
            with torch.enable_grad():
                if gp is not None:   # this rank's (branch, frame slice) only
                    mine, e_cd, e_ud = gp.unet_outputs(model, x, t, c, uc, **kwargs)
                    e_c = e_u = None
                else:
                    e_c = model.apply_model(x, t, c, **kwargs)
                    e_u = model.apply_model(x, t, uc, **kwargs) if uc is not None else None
            nz = pick(noise, j)
            if nz is None:
                nz = gp.randn(x.shape, x.device) if gp is not None else _randn(model, x.shape, x.device)
            if mine is None:
                e_cd = e_c.detach().float().contiguous()
                e_ud = None if e_u is None else e_u.detach().float().contiguous()
            x_prev, pred_x0 = ops.ddim_step(x.detach().contiguous(), e_cd, e_ud, nz.float().contiguous(), coef)
            grads, decoded = [], []
            chunk = max(1, int(getattr(model, "guided_decode_frames", 1)))
            lo, hi = (0, n_frames) if gp is None else (gp.f0, gp.f1)   # this rank's frames of the decoder passes
            for f0 in range(lo, hi, chunk):
                f1 = min(hi, f0 + chunk)
                z = pred_x0[:, :, f0:f1].clone().requires_grad_(True)  # the decoder graph ends here (:285)
                with torch.enable_grad():
                    d_x0 = model.differentiable_decode_first_stage(z)
                    # frames are independent through the decoder, so one backward over sum_f loss_f / numel_f gives
                    # every frame the gradient the reference's one-frame-at-a-time loop gives it
                    total = None
                    for f in range(f0, f1):
                        loss_dict, numel = lg(d_x0[0][:, f - f0:f - f0 + 1], index, f, f + 1)
                        term = loss_dict["recon"] if lg.mean_loss else loss_dict["recon"] / numel
                        total = term if total is None else total + term
                    g = torch.autograd.grad(outputs=total, inputs=z)[0]
                grads.append(g.detach())
                decoded.append(d_x0.detach())
            if gp is None:
                d_all, G = torch.cat(decoded, dim=2), torch.cat(grads, dim=2).float().contiguous()
            else:
                d_all, G = gp.gather_decoder_results(decoded, grads, pred_x0)
            lg.save_pred_x0(d_all, index)
            dx, de_c, de_u = ops.ddim_pred_x0_vjp(e_cd, e_ud, G, coef)
            if mine is not None:
                guided = gp.backward(mine, de_c, de_u, x) + dx
            else:
                if e_u is None:
                    torch.autograd.backward([e_c], [de_c.to(e_c.dtype)], inputs=[x])
                else:
                    torch.autograd.backward([e_c, e_u], [de_c.to(e_c.dtype), de_u.to(e_u.dtype)], inputs=[x])
                guided = x.grad.detach().float() + dx
            x.grad = None
            tmp_s = _rms(guided)
            rho = 0.0
            if tmp_s != 0:
                corr = _rms(e_cd - e_ud) if e_ud is not None else 0.0
                rho = corr * cfg / tmp_s * (0.2 * sgw)
            x_prev = x_prev - rho * guided
            rz = pick(recur_noise, j)
            if rz is None:
                rz = gp.randn(x.shape, x.device) if gp is not None else _randn(model, x.shape, x.device)
            x = float(np.sqrt(beta_t)) * x_prev + float(np.sqrt(np.float32(1) - beta_t)) * rz
        return x_prev.detach(), pred_x0.detach()
