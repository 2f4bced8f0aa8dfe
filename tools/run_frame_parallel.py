"""N-rank check of the frame-sharded denoiser on real GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/run_frame_parallel.py [--h 40 --w 64] [--cfg-split 0|1]
Every rank builds the same seeded U-Net, runs (a) the single-GPU forward pair and (b) the CFG-split x frame-sharded pair
and compares them; then times both.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "guidedvd-3dgs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--t", type=int, default=25)
    ap.add_argument("--h", type=int, default=72)
    ap.add_argument("--w", type=int, default=128)
    ap.add_argument("--mc", type=int, default=320)
    ap.add_argument("--cfg-split", type=int, default=1)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import unet_ref
    from vc_b200.frame_parallel import DenoisePlan
    from vc_b200.schedule import ModelSchedule
    from vc_b200.unet import DiffusionModelB200, UNetB200

    ref, cfg = unet_ref.build_reference_unet(model_channels=a.mc, device="cpu")
    unet = UNetB200(ref.state_dict(), device=dev, **cfg)
    del ref
    x, cc, ctx, ctx_uc = unet_ref.synth_inputs(a.t, a.h, a.w, device=dev)
    fs = torch.tensor([10], device=dev)
    ts = torch.tensor([999], device=dev)
    cond = {"c_concat": [cc], "c_crossattn": [ctx]}
    uc = {"c_concat": [cc], "c_crossattn": [ctx_uc]}
    single = DiffusionModelB200(unet, ModelSchedule())
    e_c0, e_u0 = single.apply_model_cfg(x, ts, cond, uc, fs=fs)
    plan = DenoisePlan(a.t, cfg_split=bool(a.cfg_split))
    sharded = DiffusionModelB200(unet, ModelSchedule(), plan=plan)
    e_c1, e_u1 = sharded.apply_model_cfg(x, ts, cond, uc, fs=fs)

    def rel(p, q):
        return float((p - q).norm() / q.norm())
    errs = (rel(e_c1, e_c0), rel(e_u1, e_u0))

    def timed(model):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            model.apply_model_cfg(x, ts, cond, uc, fs=fs)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / a.reps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    ms_sharded = timed(sharded)
    unet.part = None
    ms_single = timed(single)
    worst = torch.tensor(list(errs), device=dev)
    if world > 1:
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"world": world, "cfg_ways": plan.cfg_ways, "frame_ways": plan.frame_ways, "latent": [a.t, a.h, a.w],
                          "rel_l2_cond": float(worst[0]), "rel_l2_uncond": float(worst[1]), "ms_pair_single_gpu": round(ms_single, 2),
                          "ms_pair_sharded": round(ms_sharded, 2), "speedup": round(ms_single / ms_sharded, 3)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
