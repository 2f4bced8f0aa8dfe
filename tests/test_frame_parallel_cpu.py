"""CPU (gloo): host logic of the frame-sharded denoiser -- frame<->pixel re-sharding, statistics sum, output gather and
the CFG x frame world layout (vc_b200/frame_parallel.py).  The arithmetic kernels need a GPU; these tests pin the layout
and the collectives, which is everything that differs between 1 and N ranks."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_sizes():
    sys.path.insert(0, os.path.join(ROOT, "guidedvd-3dgs_b200"))
    from vc_b200.frame_parallel import split_sizes

    assert split_sizes(25, 4) == [7, 6, 6, 6]
    assert split_sizes(25, 8) == [4, 3, 3, 3, 3, 3, 3, 3]
    assert split_sizes(144, 8) == [18] * 8
    assert split_sizes(3, 4) == [1, 1, 1, 0]


def _worker(rank, world, port, q, cfg_split):
    sys.path.insert(0, os.path.join(ROOT, "guidedvd-3dgs_b200"))
    from vc_b200.frame_parallel import DenoisePlan

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    T, Cc = 7, 4
    plan = DenoisePlan(T, cfg_split=cfg_split)
    part = plan.part
    for S in (12, 10):  # divisible and ragged pixel counts
        full = torch.randn(T, S, Cc, generator=torch.Generator().manual_seed(5 + S + 100 * plan.cfg_index))
        mine = full[part.frame_slice()].contiguous()
        if part.active:
            px = part.pixels(S)
            s0 = sum(px[:part.rank])
            y = part.to_pixels(mine)
            ok &= torch.equal(y, full[:, s0:s0 + px[part.rank]])
            ok &= torch.equal(part.to_frames(y.contiguous(), S), mine)
            st = torch.tensor([float(mine.sum()), float((mine * mine).sum())])
            part.sum_stats(st)
            ok &= bool(torch.allclose(st, torch.tensor([float(full.sum()), float((full * full).sum())]), rtol=1e-5))
        else:
            ok &= part.to_pixels(mine) is mine
    # per-step result exchange: every rank ends with the full clip of every CFG half
    h = w = 3
    halves = [torch.randn(1, 2, T, h, w, generator=torch.Generator().manual_seed(77 + c)) for c in range(plan.cfg_ways)]
    outs = plan.gather_outputs(halves[plan.cfg_index][:, :, part.frame_slice()].contiguous())
    ok &= len(outs) == plan.cfg_ways and all(torch.equal(a, b) for a, b in zip(outs, halves))
    q.put((rank, bool(ok), plan.cfg_ways, plan.frame_ways, plan.cfg_index, part.frames))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,cfg_split,expect", [(2, True, (2, 1)), (2, False, (1, 2)), (3, True, (1, 3)), (4, True, (2, 2)), (8, True, (2, 4))])
def test_frame_partition_gloo(world, cfg_split, expect):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, cfg_split)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert (res[0][2], res[0][3]) == expect
    assert [r[4] for r in res] == [i // expect[1] for i in range(world)]
