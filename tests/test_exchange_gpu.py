"""GPU (needs >= 2 devices; skipped on a 1-GPU box): the NVLink peer-memory gradient sum (include/gvd_exchange.h)
against the sum computed on the host, through view_parallel.GradientExchange.  Integer-valued floats make the
expected result exact whatever the order of additions."""
import os
import socket
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "guidedvd-3dgs_b200"))
    import view_parallel as vp

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ok = True
    n = 1_000_003
    ex = vp.GradientExchange(n, dev)
    for it, nf in enumerate((ex.n_floats, 4, 4096 * 3 + 4, ex.n_floats)):  # full buffer, tiny, ragged slice, repeated epochs
        vals = [torch.randint(-1000, 1000, (ex.n_floats,), generator=torch.Generator().manual_seed(31 * it + r)).float()
                for r in range(world)]
        ex.buffer.copy_(vals[rank])
        ex.allreduce(nf)
        torch.cuda.synchronize()
        want = vals[rank].clone()
        want[:nf] = sum(v[:nf] for v in vals)
        ok &= bool(torch.equal(ex.buffer.cpu(), want))  # summed prefix exact, the rest of the buffer untouched
    ex.close()
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs with peer access")
def test_peer_memory_allreduce_matches_host_sum():
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
