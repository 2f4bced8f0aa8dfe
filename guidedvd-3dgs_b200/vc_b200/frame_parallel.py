"""Multi-GPU execution plan of the denoiser (diffusion half): one process per GPU.

Two levels of natural sharding:
  * CFG pair -- the conditional and unconditional U-Net forwards of a DDIM step are independent (ddim.py:222-223): with
    an even world size the first half of the ranks evaluates `cond`, the second half `uncond`.
  * frames   -- every spatial layer treats the 25 frames as a batch (openaimodel3d.py:566), so inside one CFG half the
    frames are split into contiguous slices, one per rank.  The temporal layers (TemporalConvBlock, TemporalTransformer)
    mix all frames of ONE pixel, so around them the activation is re-sharded frames -> pixels with one all-to-all and
    back with another (the exchange step of this path); their GroupNorms take statistics over all frames and pixels, so
    the shards add up (sum, sum of squares) per group with a 64-float all-reduce.
    The all-to-all moves each activation once (a rank keeps 1/G of it), against 2 x (K and V) x (G-1)/G for an all-gather
    of the temporal keys/values and no +-1-frame halos for the (3,1,1) convolutions, which is why it is used here.
The per-step result exchange is one all-gather of the local output frames (3.7 MB of latent at 72x128).

Everything here is layout + collectives (torch.distributed over NCCL on the GPUs, gloo in the CPU tests); no arithmetic.
"""
import torch
import torch.distributed as dist


def split_sizes(n, parts):
    """Contiguous, near-equal split of n items: the first n % parts shards get one more (25 frames / 4 -> 7,6,6,6)."""
    return [n // parts + (1 if i < n % parts else 0) for i in range(parts)]


class FramePartition:
    """Frame <-> pixel re-sharding of channels-last activations inside one process group.

    frame layout: x[F_r, S, C]  (this rank's frames, all pixels)
    pixel layout: x[T,  S_r, C] (all frames, this rank's pixels)
    """

    def __init__(self, T, group=None):
        self.group = group
        self.size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.T = T
        self.frames = split_sizes(T, self.size)
        self.f0 = sum(self.frames[:self.rank])
        self.F = self.frames[self.rank]

    @property
    def active(self):
        return self.size > 1

    def frame_slice(self):
        return slice(self.f0, self.f0 + self.F)

    def pixels(self, S):
        return split_sizes(S, self.size)

    def to_pixels(self, x):
        """x[F_r, S, C] -> [T, S_r, C]."""
        if not self.active:
            return x
        if torch.is_grad_enabled() and x.requires_grad:  # guided sampler: record the adjoint all-to-all (vc_b200.grad)
            from .grad import ToPixels
            return ToPixels.apply(x, self)
        Fr, S, Cc = x.shape
        px = self.pixels(S)
        Sr = px[self.rank]
        if S % self.size == 0:
            send = x.view(Fr, self.size, Sr, Cc).transpose(0, 1).contiguous()  # one pack copy: [dest][F_r, S_r, C]
        else:
            offs = [sum(px[:j]) for j in range(self.size)]
            send = torch.cat([x[:, offs[j]:offs[j] + px[j], :].reshape(-1) for j in range(self.size)])
        out = torch.empty(self.T, Sr, Cc, dtype=x.dtype, device=x.device)
        dist.all_to_all_single(out.view(-1), send.view(-1), [f * Sr * Cc for f in self.frames], [Fr * p * Cc for p in px],
                               group=self.group)
        return out  # blocks arrive in rank order == frame order: already [T, S_r, C]

    def to_frames(self, y, S):
        """y[T, S_r, C] -> [F_r, S, C] (S = total pixels)."""
        if not self.active:
            return y
        if torch.is_grad_enabled() and y.requires_grad:
            from .grad import ToFrames
            return ToFrames.apply(y, self, S)
        T, Sr, Cc = y.shape
        px = self.pixels(S)
        Fr = self.F
        recv = torch.empty(Fr * S * Cc, dtype=y.dtype, device=y.device)
        dist.all_to_all_single(recv, y.reshape(-1), [Fr * p * Cc for p in px], [f * Sr * Cc for f in self.frames],
                               group=self.group)
        if S % self.size == 0:
            return recv.view(self.size, Fr, Sr, Cc).transpose(0, 1).reshape(Fr, S, Cc)  # one unpack copy
        parts, o = [], 0
        for p in px:
            parts.append(recv[o:o + Fr * p * Cc].view(Fr, p, Cc))
            o += Fr * p * Cc
        return torch.cat(parts, dim=1)

    def sum_stats(self, stats):
        if self.active:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.group)
        return stats


class DenoisePlan:
    """World layout for one DDIM step: `cfg_ways` (1 or 2) x `frame_ways` ranks."""

    def __init__(self, T, cfg_split=True):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.T = T
        self.cfg_ways = 2 if (cfg_split and self.world % 2 == 0) else 1
        self.frame_ways = self.world // self.cfg_ways
        self.cfg_index = self.rank // self.frame_ways  # 0: this rank evaluates `cond`, 1: `uncond`
        group = None
        if self.world > 1 and self.cfg_ways > 1:
            # every rank creates every group, in the same order (torch.distributed contract)
            for c in range(self.cfg_ways):
                g = dist.new_group(list(range(c * self.frame_ways, (c + 1) * self.frame_ways)))
                if c == self.cfg_index:
                    group = g
        self.part = FramePartition(T, group)  # a 1-rank group (world 2 = pure CFG split) is an inactive partition

    def gather_outputs(self, y_local):
        """y_local [1, C, F_r, h, w] (this rank's frames of its CFG half) -> list of cfg_ways full tensors [1, C, T, h, w],
        identical on every rank: one all-gather of frame-padded blocks over the whole world."""
        if self.world == 1:
            return [y_local]
        _, Cc, Fr, h, w = y_local.shape
        fmax = max(self.part.frames)
        blk = torch.zeros(fmax, Cc, h, w, dtype=y_local.dtype, device=y_local.device)
        blk[:Fr] = y_local[0].transpose(0, 1)
        allb = torch.empty(self.world, fmax, Cc, h, w, dtype=y_local.dtype, device=y_local.device)
        dist.all_gather_into_tensor(allb.view(-1), blk.view(-1))
        outs = []
        for c in range(self.cfg_ways):
            fr = [allb[c * self.frame_ways + j, :self.part.frames[j]] for j in range(self.frame_ways)]
            outs.append(torch.cat(fr, dim=0).transpose(0, 1).unsqueeze(0).contiguous())
        return outs
