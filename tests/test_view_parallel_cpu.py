"""CPU (gloo, world_size 2): the host logic of the N>1 path -- view sharding and the single gradient all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_views_balanced():
    import view_parallel as vp

    for n in (0, 1, 6, 8, 25):
        for w in (1, 2, 4, 8):
            parts = [vp.shard_views(n, w, r) for r in range(w)]
            assert sum(parts, []) == list(range(n))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1
    assert vp.shard_views(25, 8, 0) == [0, 1, 2, 3] and len(vp.shard_views(25, 8, 7)) == 3


def _worker(rank, world, port, q):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "guidedvd-3dgs_b200"))
    import view_parallel as vp

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    P = 1000
    views = vp.shard_views(5, world, rank)
    per_view = [[torch.randn(P, 3, generator=torch.Generator().manual_seed(10 * v + k)) for k in range(3)] for v in views]
    grads = vp.accumulate_views(per_view)
    vp.allreduce_gradients(grads)
    expect = [sum(torch.randn(P, 3, generator=torch.Generator().manual_seed(10 * v + k)) for v in range(5)) for k in range(3)]
    ok = all(torch.allclose(a, b, atol=1e-5) for a, b in zip(grads, expect))

    # the three collective shapes of allreduce_gradients: (a) views of one flat buffer -> a single all-reduce on it,
    # (b) the exchange path -- sum the caller-owned buffer, then point every leaf.grad at its slice (a CPU stand-in for
    # GradientExchange keeps the host logic testable without GPUs), (c) averaging
    import diff_gaussian_rasterization as dgr

    Pg, M = 37, 16
    flat = torch.full((dgr.gradient_buffer_floats(Pg, M),), float(rank + 1))
    v = dgr._C._grad_views(flat[:sum((Pg * w + 3) // 4 * 4 for w in (3, 48, 1, 3, 4, 0, 0, 3))], Pg, M, True, True, False, False)
    vp.allreduce_gradients([t for t in v if t is not None])
    ok &= bool((v[1] == 3.0).all() and (v[5] == 3.0).all() and (v[0] == 3.0).all())  # 1 + 2 on both ranks

    class FakeExchange:
        def __init__(self, n):
            self.buffer, self.world, self.calls = torch.full((n,), float(rank + 1)), world, []

        def allreduce(self, n):
            self.calls.append(n)
            dist.all_reduce(self.buffer[:n])

    ex = FakeExchange(dgr.gradient_buffer_floats(Pg, M))
    total = sum((Pg * w + 3) // 4 * 4 for w in (3, 48, 1, 3, 4, 0, 0, 3))
    m2, m3, op, col, cov, sh, sc, rot = dgr._C._grad_views(ex.buffer[:total], Pg, M, True, True, False, False)
    views_named = {"means2D": m2, "means3D": m3, "opacities": op, "colors_precomp": col, "cov3D_precomp": cov, "shs": sh,
                   "scales": sc, "rotations": rot, "_floats": total}
    leaves = {"means3D": torch.zeros(Pg, 3, requires_grad=True), "shs": torch.zeros(Pg, M, 3, requires_grad=True),
              "opacities": torch.zeros(Pg, 1, requires_grad=True), "scales": torch.zeros(Pg, 3, requires_grad=True),
              "rotations": torch.zeros(Pg, 4, requires_grad=True), "means2D": torch.zeros(Pg, 3, requires_grad=True)}
    for leaf in leaves.values():
        leaf.grad = torch.full_like(leaf, -7.0)  # what autograd's copy would have left there
    vp.allreduce_gradients(None, exchange=ex, leaves=leaves, views=views_named, average=True)
    ok &= ex.calls == [total]
    for name, leaf in leaves.items():
        ok &= leaf.grad.shape == leaf.shape and bool((leaf.grad == 1.5).all())                # (1 + 2) / 2
        ok &= leaf.grad.untyped_storage().data_ptr() == ex.buffer.untyped_storage().data_ptr()  # a slice of the buffer
    q.put((rank, ok, views))
    dist.destroy_process_group()


def test_allreduce_gradients_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] and res[1][1]
    assert res[0][2] == [0, 1, 2] and res[1][2] == [3, 4]
