"""View-parallel rasterization across the GPUs of one box.

Training views are independent units: every rank keeps the full Gaussian set, renders its own views
(forward + backward, no communication inside the step) and the per-Gaussian gradients are summed once per
step with a single all-reduce over one flat buffer (62 floats per Gaussian: means 3, SH 48, opacity 1,
scale 3, rotation 4, plus means2D 3).  One process per GPU, `torch.distributed` (NCCL over NVLink on the
GPU box; gloo in the CPU tests).
"""
from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def shard_views(num_views: int, world_size: int, rank: int) -> List[int]:
    """Contiguous, balanced split: the first (num_views % world_size) ranks get one extra view."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(num_views, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def allreduce_gradients(grads: Sequence[torch.Tensor], group=None, average: bool = False) -> None:
    """In-place sum (or mean) of a list of gradient tensors over the ranks, as ONE collective."""
    grads = [g for g in grads if g is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    # Fast path: the rasterizer hands out all its gradients as views of one flat buffer -> reduce it in place.
    base = grads[0]._base
    if base is not None and base.dim() == 1 and all(g._base is base for g in grads):
        lo = min(g.storage_offset() for g in grads)
        hi = max(g.storage_offset() + g.numel() for g in grads)
        seg = base[lo - base.storage_offset():hi - base.storage_offset()]
        dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=group)
        if average:
            seg /= dist.get_world_size(group)
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def accumulate_views(per_view_grads: Iterable[Sequence[torch.Tensor]]) -> List[torch.Tensor]:
    """Sum the gradient lists of the views one rank rendered (before the cross-rank all-reduce)."""
    total = None
    for gs in per_view_grads:
        if total is None:
            total = [g.clone() for g in gs]
        else:
            for t, g in zip(total, gs):
                t.add_(g)
    return total or []
