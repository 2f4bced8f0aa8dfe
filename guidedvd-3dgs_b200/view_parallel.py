"""View-parallel rasterization across the GPUs of one box.

Training views are independent units: every rank keeps the full Gaussian set, renders its own views
(forward + backward, no communication inside the step) and the per-Gaussian gradients are summed once per
step with a single all-reduce over one flat buffer (62 floats per Gaussian: means 3, SH 48, opacity 1,
scale 3, rotation 4, plus means2D 3).  One process per GPU, `torch.distributed` (NCCL over NVLink on the
GPU box; gloo in the CPU tests).
"""
import ctypes as C
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


class _RawCuda:
    """Exposes a raw device allocation to torch (zero-copy) through the CUDA array interface."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class GradientExchange:
    """The gradient sum of the view-parallel step as ONE hand-written kernel over NVLink (include/gvd_exchange.h): each
    rank owns an exchange buffer the rasterizer backward writes its gradients into
    (diff_gaussian_rasterization.set_gradient_buffer); `allreduce()` sums it in place across the ranks of the world.

    Two ways to obtain buffers every rank can address:
      * "nvls" (preferred): torch.distributed._symmetric_memory allocates the buffers and binds them to ONE multicast
        object; the kernel then sums a slice inside the NVSwitch with `multimem.ld_reduce` and writes all replicas with
        one `multimem.st`.  torch is plumbing here (allocation + handle exchange); the collective is the library's kernel.
      * "peer": cudaMalloc + CUDA IPC handles carried by one torch.distributed all_gather; direct peer loads/stores.
    Default: "nvls" from 4 ranks on, "peer" below (measured at N = 2 on B200: 0.83 ms per C2 step with peer loads/stores
    against 0.97 ms through the switch -- with one peer there is nothing for the in-switch reduction to save);
    GVD_EXCHANGE=peer|nvls forces one; "peer" is also the fallback when multicast is unavailable.  CUDA only."""

    def __init__(self, n_floats: int, device):
        import os

        import gvd_native as _n

        self.lib = _n.raster()
        self._n = _n
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if not (2 <= self.world <= _n.EXCHANGE_MAX_RANKS):
            raise ValueError("GradientExchange needs 2..8 ranks")
        self.device = torch.device(device)
        self.n_floats = (int(n_floats) + 3) // 4 * 4
        self.payload = self.n_floats * 4
        self.multicast, self.mode, self._symm, self.why_not_nvls = None, "peer", None, None
        want = os.environ.get("GVD_EXCHANGE", "nvls" if self.world >= 4 else "peer")
        ok = torch.zeros(1, device=self.device)
        if want != "peer":
            try:
                self._init_symmetric()
                ok.fill_(1.0)
            except Exception as ex:  # no multicast on this box / this torch: every rank must take the same route
                self.why_not_nvls = repr(ex)[:200]
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) < 1.0:
            if want != "peer" and self.why_not_nvls is None:
                self.why_not_nvls = "a peer rank could not map the multicast object"
            self._symm, self.multicast = None, None
            self._init_ipc()
        self.epoch = 0
        dist.barrier()  # every mapping exists before anybody launches

    def _init_symmetric(self):
        import torch.distributed._symmetric_memory as symm_mem

        flag_floats = self._n.EXCHANGE_FLAG_BYTES // 4
        with torch.cuda.device(self.device):
            t = symm_mem.empty(self.n_floats + flag_floats, dtype=torch.float32, device=self.device)
            t.zero_()
            hdl = symm_mem.rendezvous(t, dist.group.WORLD)
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        if not mc:
            raise RuntimeError("symmetric memory has no multicast mapping on this system")
        self._symm = (t, hdl)
        self.ptrs = [int(p) for p in hdl.buffer_ptrs]
        self.multicast = mc
        self.buffer = t[:self.n_floats]
        self.mode = "nvls"
        torch.cuda.synchronize(self.device)

    def _init_ipc(self):
        _n = self._n
        ptr, handle = C.c_void_p(), C.create_string_buffer(_n.EXCHANGE_HANDLE_BYTES)
        with torch.cuda.device(self.device):
            self._check(self.lib.gvd_exchange_alloc(self.payload, C.byref(ptr), handle), "gvd_exchange_alloc")
            handles = [None] * self.world
            dist.all_gather_object(handles, handle.raw)
            self.ptrs = []
            for q, h in enumerate(handles):
                if q == self.rank:
                    self.ptrs.append(ptr.value)
                else:
                    pp = C.c_void_p()
                    self._check(self.lib.gvd_exchange_open(h, C.byref(pp)), "gvd_exchange_open")
                    self.ptrs.append(pp.value)
        self.buffer = torch.as_tensor(_RawCuda(ptr.value, self.n_floats), device=self.device)
        self.mode = "peer"

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed: " + self._n.last_error(self.lib))

    def owns(self, t: torch.Tensor) -> bool:
        lo = self.buffer.data_ptr()
        return t.is_cuda and lo <= t.data_ptr() and t.data_ptr() + t.numel() * t.element_size() <= lo + self.payload

    def allreduce(self, n_floats: int = None) -> torch.Tensor:
        """Sums the first n_floats (default: all) of every rank's buffer in place; stream-ordered on the current stream."""
        a = self._n.ExchangeArgs()
        a.world, a.rank = self.world, self.rank
        for q, pq in enumerate(self.ptrs):
            a.bufs[q] = pq
        a.payload_bytes = self.payload
        a.n_floats = self.n_floats if n_floats is None else (int(n_floats) + 3) // 4 * 4
        self.epoch += 1
        a.epoch = self.epoch
        a.multicast = self.multicast
        with torch.cuda.device(self.device):
            self._check(self.lib.gvd_exchange_allreduce_sum(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                        "gvd_exchange_allreduce_sum")
        return self.buffer

    def close(self):
        if getattr(self, "ptrs", None) is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier()
        if self._symm is not None:   # symmetric memory: torch owns the allocation and its mappings
            self.buffer, self._symm, self.ptrs = None, None, None
            return
        for q, pq in enumerate(self.ptrs):
            if q != self.rank:
                self.lib.gvd_exchange_close(pq)
        dist.barrier()
        self.buffer = None
        self.lib.gvd_exchange_free(self.ptrs[self.rank])
        self.ptrs = None


def allreduce_gradients(grads: Sequence[torch.Tensor], group=None, average: bool = False, exchange: GradientExchange = None,
                        leaves=None, views=None) -> None:
    """In-place sum (or mean) of the gradients of one step over the ranks, as ONE collective.

    Peer-memory path (`exchange`, `leaves`, `views`): the rasterizer backward wrote this rank's gradients into the
    exchange buffer (diff_gaussian_rasterization.set_gradient_buffer); `views` = diff_gaussian_rasterization.
    gradient_views(device) names the slices.  The buffer is summed across the ranks by one launch of the peer-memory
    kernel and every `leaves[name].grad` is then pointed at its slice of the buffer -- no packing, no copy back
    (autograd itself does not adopt gradients that are views of a larger buffer, it copies them).
    Otherwise: NCCL/gloo, one all-reduce over the flat base buffer when all gradients are views of one, else one per
    tensor."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    if exchange is not None and leaves is not None and views is not None:
        # `leaves` must be the tensors that were handed to the rasterizer themselves (the slices ARE their gradients).
        # GaussianModel's parameters go through exp / sigmoid / normalize first (or through rasterize_gaussians_raw, which
        # returns raw-parameter gradients): pass the rasterizer's direct inputs here, not upstream parameters.
        for name, leaf in leaves.items():
            v = views.get(name)
            if v is not None and leaf is not None and v.numel() != leaf.numel():
                raise ValueError(f"allreduce_gradients: leaf '{name}' has {leaf.numel()} elements, its gradient slice {v.numel()}: "
                                 "leaves must be the rasterizer's direct inputs")
        n = int(views["_floats"])
        exchange.allreduce(n)
        if average:
            exchange.buffer[:n] /= exchange.world
        for name, leaf in leaves.items():
            v = views.get(name)
            if v is not None and leaf is not None:
                leaf.grad = v if v.shape == leaf.shape else v.view_as(leaf)
        return
    grads = [g for g in grads if g is not None]
    if not grads:
        return
    # the rasterizer hands out all its gradients as views of one flat buffer -> reduce it in place
    base = grads[0]._base
    if base is not None and base.dim() == 1 and all(g._base is base for g in grads):
        lo = min(g.storage_offset() for g in grads)
        hi = max(g.storage_offset() + g.numel() for g in grads)
        seg = base[lo - base.storage_offset():hi - base.storage_offset()]
        dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=group)
        if average:
            seg /= dist.get_world_size(group)
        return
    for g in grads:
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
        if average:
            g /= dist.get_world_size(group)


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
