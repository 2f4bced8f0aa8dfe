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
    n = 1_300_003  # also holds the 62 floats per Gaussian of the 20000-Gaussian scene below
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
    # the rasterizer backward writes into the exchange buffer; after the sum every leaf.grad is a slice of that buffer
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import diff_gaussian_rasterization as dgr
    import synth

    sc = synth.synth_scene(20000, 7, device=dev)
    cam = synth.synth_camera(8 + rank, 128, 128, device=dev)
    dgr.set_gradient_buffer(ex.buffer)
    leaves = {k: sc[k].detach().clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
    st = dgr.GaussianRasterizationSettings(
        image_height=128, image_width=128, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"], bg=torch.zeros(3, device=dev),
        scale_modifier=1.0, viewmatrix=cam["viewmatrix"], projmatrix=cam["projmatrix"], sh_degree=3, campos=cam["campos"],
        prefiltered=False, debug=False, confidence=sc["confidence"])
    for _ in range(2):  # second round: the cached views of the buffer are reused
        for v in list(leaves.values()) + [m2d]:
            v.grad = None
        color, radii, depth, alpha = dgr.GaussianRasterizer(st)(
            means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"],
            scales=leaves["scales"], rotations=leaves["rotations"])
        (color.sum() + depth.sum()).backward()
        want = {k: v.grad.clone() for k, v in leaves.items()}
        for k in want:
            dist.all_reduce(want[k])
        vp.allreduce_gradients(None, exchange=ex, leaves=dict(leaves, means2D=m2d), views=dgr.gradient_views(dev))
        torch.cuda.synchronize()
        for k, v in leaves.items():
            ok &= bool(ex.owns(v.grad)) and bool(torch.allclose(v.grad, want[k], rtol=1e-6, atol=1e-7))
    dgr.set_gradient_buffer(None)
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
