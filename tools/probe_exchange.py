"""N-rank probe of the peer-memory gradient sum: correctness against NCCL, timing of both, autograd aliasing.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/probe_exchange.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "guidedvd-3dgs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import view_parallel as vp  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
out = {"world": world}
P = 500_000
n = P * 65 + 64
ex = vp.GradientExchange(n, dev)
g = torch.Generator(device=dev).manual_seed(100 + rank)
src = torch.randn(ex.n_floats, device=dev, generator=g)
ex.buffer.copy_(src)
ref = src.clone()
dist.all_reduce(ref)
ex.allreduce()
torch.cuda.synchronize()
out["max_abs_diff_vs_nccl"] = float((ex.buffer - ref).abs().max())
# all ranks hold bit-identical sums
chk = ex.buffer.double().sum().reshape(1)
lst = [torch.empty_like(chk) for _ in range(world)]
dist.all_gather(lst, chk)
out["identical_across_ranks"] = bool(all(float(x) == float(lst[0]) for x in lst))


def timed(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t), 4)


for mb in (124, 30, 4):
    nf = mb * 250_000
    out[f"peer_kernel_ms_{mb}MB"] = timed(lambda: ex.allreduce(nf))
    tmp = torch.zeros(nf, device=dev)
    out[f"nccl_ms_{mb}MB"] = timed(lambda: dist.all_reduce(tmp))

# does autograd keep the rasterizer's gradient views (so the in-place sum is what .grad sees)?
import diff_gaussian_rasterization as dgr  # noqa: E402
import synth  # noqa: E402
sc = synth.synth_scene(20000, 7, device=dev)
cam = synth.synth_camera(8 + rank, 128, 128, device=dev)
dgr.set_gradient_buffer(ex.buffer)
leaves = {k: sc[k].detach().clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
st = dgr.GaussianRasterizationSettings(image_height=128, image_width=128, tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"],
                                       bg=torch.zeros(3, device=dev), scale_modifier=1.0, viewmatrix=cam["viewmatrix"],
                                       projmatrix=cam["projmatrix"], sh_degree=3, campos=cam["campos"], prefiltered=False,
                                       debug=False, confidence=sc["confidence"])
color, radii, depth, alpha = dgr.GaussianRasterizer(st)(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                                         shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
color.sum().backward()
out["autograd_adopts_buffer_views"] = {k: bool(ex.owns(v.grad)) for k, v in leaves.items()}
local = {k: v.grad.clone() for k, v in leaves.items()}
vp.allreduce_gradients(None, exchange=ex, leaves=dict(leaves, means2D=m2d), views=dgr.gradient_views(dev))
out["grads_live_in_exchange_buffer"] = {k: bool(ex.owns(v.grad)) for k, v in leaves.items()}
for k in local:
    dist.all_reduce(local[k])
torch.cuda.synchronize()
out["grad_sum_max_abs_diff_vs_nccl"] = max(float((leaves[k].grad - local[k]).abs().max()) for k in local)
dgr.set_gradient_buffer(None)

# torch symmetric memory availability (would add an NVLS/multicast variant)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD.group_name)
    out["symm_mem"] = {"ok": True, "multicast": bool(h.has_multicast_support(torch._C._distributed_c10d._DeviceType.CUDA if False else "cuda", dev.index)) if False else int(h.multicast_ptr != 0)}
except Exception as e:  # noqa: BLE001
    out["symm_mem"] = {"ok": False, "error": repr(e)[:200]}
ex.close()
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
