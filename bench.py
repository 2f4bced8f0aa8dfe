#!/usr/bin/env python
"""bench.py -- views/s of the differentiable Gaussian rasterizer (fwd+bwd) on B200.

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W            our sm_100a path
  python bench.py --impl reference --gpus N ...            the unmodified reference (oracle/_ref)
One JSON line on rank 0.  A "step" is one view: rasterizer forward + backward through the
`GaussianRasterizer` autograd boundary with fixed random cotangents.
N>1: one process per GPU (torchrun), each rank renders its own views of the replicated scene
(weak scaling) and the per-Gaussian gradients are summed once per step by the library's NVLink kernel
(in-switch NVLS reduction when a multicast mapping is available); the reference arm uses NCCL.
Secondary blocks inside the same line: `denoise` (BASELINE configs[2]), `guided` (configs[3] shape, the dominant
cost of train_guidedvd.py), `c5` (configs[4], N > 1).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "guidedvd-3dgs_b200")
for p in (PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

if "reference" in sys.argv:
    # The reference's guided step keeps 25 decoder graphs alive (172 GB of the 180 GB: bench line `guided.peak_mem_gb`).  At that
    # fill level a fragmented caching allocator falls into cudaMalloc retries -- one box measured 15 s per step where
    # others measured 2.3-3.2 s.  Expandable segments give the reference arm its best case.
    os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import view_parallel  # noqa: E402

NCY = 4  # cameras per rank, visited round-robin
SIZING = {"sync": "sync: speculative buffers, R/V validated before the forward returns (default)",
          "exact": "exact: library waits for R/V, exact callbacks",
          "defer": "defer: validation moved to the backward (opt-in)"}.get(
              {"1": "defer", "0": "exact"}.get(os.environ.get("GVD_SPECULATE", "sync"), os.environ.get("GVD_SPECULATE", "sync")), "sync")

WORKLOADS = {
    # name: (P, W, H, seed, sh_degree, description)
    "C2": (500_000, 640, 480, 20260002, 3,
           "BASELINE.json configs[1]: Replica office_3-like synthetic scene, 500k Gaussians, 640x480, SH deg 3, "
           "rasterizer fwd+bwd only"),
    "C4": (800_000, 1600, 1066, 20260004, 3, "ScanNet++-like 800k Gaussians 1600x1066"),
    "C5": (2_000_000, 640, 480, 20260005, 3, "Replica room_0-like 2M Gaussians, 640x480 (one view per rank)"),
    "small": (50_000, 320, 240, 20260001, 3, "smoke-size"),
}


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons of this rank's GPU (NVML, in-process) while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    def run(self):
        if self.h is None:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, int(r)))
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if self.h is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        sm = sorted(s[0] for s in self.samples)
        bits = 0
        for _, r in self.samples:
            bits |= r
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons = sorted(n for n, b in names.items() if bits & b)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(self.samples)}


def load_impl(impl):
    if impl == "ours":
        import diff_gaussian_rasterization as pkg
        import gvd_native
        gvd_native.raster()  # fail loudly if the CUDA library is missing
        return pkg
    import refload
    pkg = refload.ref_dgr()
    return pkg


def make_step(pkg, sc, cam, bg, D):
    """Returns step(cot) -> (color, leaves) running fwd+bwd through the public API."""
    leaves = {k: sc[k].detach().clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
    conf = sc["confidence"]

    def step(cot, viewmatrix, projmatrix, campos, after_forward=None):
        settings = pkg.GaussianRasterizationSettings(
            image_height=cam["height"], image_width=cam["width"], tanfovx=cam["tanfovx"], tanfovy=cam["tanfovy"],
            bg=bg, scale_modifier=1.0, viewmatrix=viewmatrix, projmatrix=projmatrix, sh_degree=D, campos=campos,
            prefiltered=False, debug=False, confidence=conf)
        rast = pkg.GaussianRasterizer(raster_settings=settings)
        for v in leaves.values():
            v.grad = None
        means2D.grad = None
        color, radii, depth, alpha = rast(means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"],
                                          shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"])
        if after_forward is not None:
            after_forward()
        torch.autograd.backward([color, depth, alpha], [cot[0:3], cot[3:4], cot[4:5]])
        return color, radii

    return step, leaves, means2D


def grad_flat(leaves):
    return [v.grad for v in leaves.values() if v.grad is not None]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--spinup-seconds", type=float, default=1.5, help="untimed load before the warm-up steps (both arms)")
    ap.add_argument("--no-denoise", action="store_true", help="skip the secondary DDIM denoise-steps/s (and guided) measurement")
    ap.add_argument("--no-guided", action="store_true", help="skip the guided-step block of the secondary measurement")
    ap.add_argument("--no-c5", action="store_true", help="N > 1: skip the configs[4] block (2 M Gaussians, 496 MB gradient sum)")
    ap.add_argument("--ref-device", default="gpu", choices=["gpu", "cpu"],
                    help="reference arm: compiled reference CUDA on the GPU (default) or the C oracle port on host cores")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    K, Wm = args.steps, max(args.warmup, 3)
    P, W, H, seed, D, desc = WORKLOADS[args.workload]

    if args.impl == "reference" and args.ref_device == "cpu":
        if rank == 0:
            print(json.dumps(cpu_reference_line(args, P, W, H, seed, D, desc, K, Wm)))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    pkg = load_impl(args.impl)
    if pkg is None:
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (python oracle/build_ref.py)"}))
        return

    import synth
    sc = synth.synth_scene(P, seed, device=dev)
    # NCY different views per rank, visited round-robin: no two consecutive steps render the same camera (view 0 of rank 0
    # is round 1's single camera, seed + 1)
    cams = [synth.synth_camera(seed + 1 + rank + 101 * j, W, H, device=dev) for j in range(NCY)]
    cam = cams[0]
    bg = torch.zeros(3, device=dev)
    g = torch.Generator().manual_seed(seed + 2 + rank)
    cot_host = torch.randn(5, H, W, generator=g).pin_memory()
    cam_hosts = [torch.cat([c["viewmatrix"].flatten(), c["projmatrix"].flatten(), c["campos"].flatten()]).cpu() for c in cams]
    cot_dev = cot_host.to(dev)
    step, leaves, means2D = make_step(pkg, sc, cam, bg, D)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # N > 1: the backward writes its gradients straight into this rank's peer-mapped exchange buffer and the sum over
    # ranks is one launch of the library's NVLink peer-memory kernel (include/gvd_exchange.h); the reference arm, which
    # has no such hook, sums its gradients with NCCL.
    exchange, exchange_check = None, None
    if world > 1 and args.impl == "ours":
        exchange = view_parallel.GradientExchange(pkg.gradient_buffer_floats(P), dev)
        exchange_check = check_exchange(exchange, dev, world)  # one-shot, before anything is timed: the sum against NCCL's
        pkg.set_gradient_buffer(exchange.buffer)

    named = dict(leaves, means2D=means2D)

    def sum_gradients():
        if exchange is not None:
            view_parallel.allreduce_gradients(None, exchange=exchange, leaves=named, views=pkg.gradient_views(dev))
        else:
            view_parallel.allreduce_gradients(grad_flat(leaves))

    def taken():
        # a bare step outside the timed loops (spin-up, workload facts, stage timers): mark the exchange buffer's gradients
        # as consumed, or the rasterizer's aliasing guard gives every later backward fresh memory (and warns once)
        if exchange is not None:
            pkg.gradient_views(dev)

    res_state = {"k": 0}

    def resident_step():
        c = cams[res_state["k"] % NCY]
        res_state["k"] += 1
        step(cot_dev, c["viewmatrix"], c["projmatrix"], c["campos"])
        if world > 1:
            sum_gradients()

    # End-to-end step: this step's inputs (camera matrices + cotangent images, one pinned host buffer) are uploaded
    # every step. Like a prefetching data loader, the upload of step k+1 is issued on a copy stream while step k
    # computes (two device slots). The step's result -- the scalar a trainer reads back every iteration
    # (train_baseline.py:88), one reduction kernel -- travels to pinned host memory on the same side stream and is READ
    # `LAG` steps later (asynchronous loss logging), so the host never idles the GPU; every step's value is read inside
    # the timed region (the last ones by e2e_flush).
    LAG = 2
    NCAM = 64  # camera block padded to 256 bytes so the cotangent images stay 256-byte aligned
    in_hosts = [torch.cat([ch, torch.zeros(NCAM - ch.numel()), cot_host.flatten()]).pin_memory() for ch in cam_hosts]
    in_host = in_hosts[0]
    copy_stream = torch.cuda.Stream(device=dev)
    dev_slots = [torch.empty_like(in_host, device=dev) for _ in range(2)]
    h2d_done = [torch.cuda.Event() for _ in range(2)]
    slot_free = [torch.cuda.Event() for _ in range(2)]
    loss_pinned = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(LAG + 1)]
    loss_ready = [torch.cuda.Event() for _ in range(LAG + 1)]
    loss_events = [torch.cuda.Event() for _ in range(LAG + 1)]
    e2e_state = {"k": 0, "last": None, "staged": 0, "pending": None, "fwd_recorded": False}
    # The 6 MB upload of the next step's inputs is timed to run under the current step's BACKWARD. Anything the copy
    # engines serve inside the compute stream (cudaMemsetAsync: CUB's radix-sort counters in the forward; formerly also
    # the backward's accumulator clear and the 4-byte read-back of R, both now done by kernels) queues behind a
    # transfer in flight: with the upload running under the forward a step took 625 us, under the backward 589 us
    # (device-resident inputs: 571 us). GVD_E2E_H2D_AT=start restores the earlier placement (A/B knob).
    H2D_AT_BWD = os.environ.get("GVD_E2E_H2D_AT", "bwd") != "start"
    fwd_event = torch.cuda.Event()

    def side_stream_work(k):
        """One visit to the copy stream per step: result of step k-1 to the host, inputs of step k+1 to the device."""
        with torch.cuda.stream(copy_stream):
            pend = e2e_state["pending"]
            if pend is not None:
                j, loss = pend
                copy_stream.wait_event(loss_ready[j % (LAG + 1)])
                loss_pinned[j % (LAG + 1)].copy_(loss, non_blocking=True)
                loss.record_stream(copy_stream)
                loss_events[j % (LAG + 1)].record(copy_stream)
                e2e_state["pending"] = None
            if k is not None and e2e_state["staged"] <= k:
                copy_stream.wait_event(slot_free[k % 2])  # the step that last used this slot has finished
                if H2D_AT_BWD and e2e_state["fwd_recorded"]:
                    copy_stream.wait_event(fwd_event)      # ... and the current step has reached its backward
                dev_slots[k % 2].copy_(in_hosts[k % NCY], non_blocking=True)
                h2d_done[k % 2].record(copy_stream)
                e2e_state["staged"] = k + 1

    def read_back(j):
        loss_events[j % (LAG + 1)].synchronize()
        e2e_state["last"] = float(loss_pinned[j % (LAG + 1)])

    def e2e_step():
        k = e2e_state["k"]
        if e2e_state["staged"] <= k:
            side_stream_work(k)
        def prefetch():  # D2H of the previous result + H2D of the next inputs overlap this step's kernels
            if H2D_AT_BWD:
                fwd_event.record()
                e2e_state["fwd_recorded"] = True
            side_stream_work(k + 1)

        if not H2D_AT_BWD:
            prefetch()
        buf = dev_slots[k % 2]
        torch.cuda.current_stream().wait_event(h2d_done[k % 2])
        cot = buf[NCAM:].view(5, H, W)
        color, radii = step(cot, buf[0:16].view(4, 4), buf[16:32].view(4, 4), buf[32:35],
                            after_forward=prefetch if H2D_AT_BWD else None)
        if world > 1:
            sum_gradients()
        slot_free[k % 2].record()
        loss = torch.dot(color.detach().view(-1), cot[0:3].reshape(-1))
        loss_ready[k % (LAG + 1)].record()
        e2e_state["pending"] = (k, loss)
        if k >= LAG:
            read_back(k - LAG)
        e2e_state["k"] = k + 1

    def e2e_flush():
        k = e2e_state["k"]
        side_stream_work(None)
        for j in range(max(0, k - LAG), k):
            read_back(j)

    def timed(fn, n, flush=None):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        if flush is not None:
            flush()
        e1.record()
        sync_all()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), wall

    # untimed: let the caching allocator reach its steady state and the GPU leave its idle clocks (a fresh box needs
    # about a second of load before the SM clock settles: the first 0.4 s of work measured 25 % slow), then the W warm-up
    # steps proper. The spin-up is time-based, so the number of iterations differs from rank to rank: it runs the LOCAL
    # step only -- no collective may sit inside a loop whose trip count is not the same on every rank.
    t_spin = time.perf_counter()
    while True:
        for _ in range(20):
            step(cot_dev, cam["viewmatrix"], cam["projmatrix"], cam["campos"])
            taken()
        torch.cuda.synchronize()
        if time.perf_counter() - t_spin > args.spinup_seconds:
            break
    sync_all()
    for _ in range(5):
        resident_step()
    for _ in range(Wm):
        resident_step()
        e2e_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_res, wall_res = timed(resident_step, K)
    ms_e2e, wall_e2e = timed(e2e_step, K, e2e_flush)
    clocks = sampler.stop()

    # workload facts (same for both arms): R and visible count, averaged over the camera cycle
    Rs, vs = [], []
    for c in cams:
        color, radii = step(cot_dev, c["viewmatrix"], c["projmatrix"], c["campos"])
        taken()
        Rs.append(int(getattr(color.grad_fn, "num_rendered", 0) or 0))
        vs.append(int((radii > 0).sum().item()))
    R, visible = sum(Rs) // NCY, sum(vs) // NCY
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    HWp = H * W

    roofline = None
    stage_ms = None
    if args.impl == "ours":
        import gvd_native
        lib = gvd_native.raster()
        # stage timers: CUDA events recorded inside the library on the launching stream.  Five chunks of the camera cycle,
        # per-stage MEDIAN of the chunk means: one slow step (an allocator refill, a clock dip) cannot move the figure
        nprof = max(min(K, 48) // NCY // 4, 1) * NCY
        chunks = []
        for _ in range(5):
            lib.gvd_raster_profile_enable(1)
            for i in range(nprof):
                c = cams[i % NCY]
                step(cot_dev, c["viewmatrix"], c["projmatrix"], c["campos"])
                taken()
            torch.cuda.synchronize()
            st = gvd_native.RasterStageTimes()
            lib.gvd_raster_profile_read(C.byref(st))
            chunks.append([st.ms[i] / nprof for i in range(len(gvd_native.STAGE_NAMES))])
        lib.gvd_raster_profile_enable(0)
        stage_ms = {n: sorted(ch[i] for ch in chunks)[len(chunks) // 2] for i, n in enumerate(gvd_native.STAGE_NAMES)}  # per step
        # algorithmic bytes per stage, SURVEY.md section 8(d). The reference's scan + duplicateWithKeys + 64-bit
        # radix sort + identifyTileRanges (8P + 12R + 12R*2*passes + 8R+8T) are replaced here by a depth sort of the
        # Gaussians and a counting sort on the tile id; the figures below are the bytes THESE stages must move.
        vchunks = (visible + 63) // 64
        alg = {
            "preprocess": P * (12 + 12 + 16 + 4 + 12 * (D + 1) ** 2) + P * 12 + visible * 64,
            "depth_sort": P * 8 + visible * 12 + visible * 8 * 2 * 4,   # compaction + four (key, id) passes
            "bin_count": visible * 16 + vchunks * tiles * 4 * 3 + tiles * 12,
            "bin_fill": visible * 16 + vchunks * tiles * 4 + 4 * R,
            "export_keys": 0,
            "render_fwd": 44 * R + 24 * HWp + 8 * tiles,
            "render_bwd": 44 * R + 28 * HWp + 8 * tiles + 80 * R,
            # the dense zero fill of the outputs (248 B x P) now rides in render_bwd; this kernel touches visible rows only
            "gaussian_bwd": visible * (12 + 4 + 24 + 16 + 12 * (D + 1) ** 2 + 3 + 16 + 12 + 4) + visible * (12 + 12 + 4 + 12 * 16 + 12 + 16),
        }
        top = max(stage_ms, key=lambda k: stage_ms[k])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        ach = alg[top] / (stage_ms[top] * 1e-3) / 1e9 if stage_ms[top] > 0 else 0.0
        # dram__bytes_read.sum + dram__bytes_write.sum per launch, read from this round's committed summary of the
        # `ncu --set full` capture (profiles/r02_ncu_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep)
        ncu_traffic, traffic_src = {}, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
            ncu_traffic, traffic_src = tj.get(args.workload, {}), tj.get("source")
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": top, "achieved": round(ach, 2), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": ncu_traffic.get(top),
                    "traffic_source": traffic_src if top in ncu_traffic else None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                    "algorithmic_bytes": alg[top],
                    "note": "render kernels are fp32-ALU/SFU/atomic bound, not HBM bound (SURVEY.md 0.5); see DESIGN.md",
                    "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
                    "stage_gbs": {k: round(alg[k] / (v * 1e-3) / 1e9, 1) if v > 0 else None for k, v in stage_ms.items()},
                    "traversed_note": "render_* GB/s use SURVEY 8d's whole-list byte count (44 B x R); tiles stop early and "
                                      "actually walk only a few % of their lists"}

    exchange_mode = None
    if exchange is not None:
        exchange_mode = ("NVLS: multimem.ld_reduce + multimem.st through the NVSwitch" if exchange.mode == "nvls" else
                         "peer ld/st over NVLink" + (f" (no multicast: {exchange.why_not_nvls})" if exchange.why_not_nvls else ""))
        pkg.set_gradient_buffer(None)
        exchange.close()
    emitted = []

    def emit(denoise_result):
        """Rank 0: build and print THE json line (once)."""
        if emitted:
            return
        emitted.append(1)
        views_per_s = world * K / (ms_res * 1e-3)
        e2e_views_per_s = world * K / (ms_e2e * 1e-3)
        h2d = in_host.numel() * 4
        ws_mb = (P * 236 + P * 248 + R * (12 * 2 + 48 + 12) + HWp * 48) / 1e6
        line = {
            "metric": "3DGS train-step views/sec (rasterizer fwd+bwd)", "value": round(views_per_s, 2), "unit": "views/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": round(ms_res / K, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": args.impl,
            "config": {"workload": args.workload, "description": desc, "P": P, "width": W, "height": H, "sh_degree": D,
                       "num_rendered": R, "visible": visible, "tiles": tiles, "cameras_per_rank": NCY,
                       "sizing": SIZING if args.impl == "ours" else "reference (cudaMemcpy + sync per frame)",
                       "exchange": (exchange_mode if world > 1 and args.impl == "ours" else None),
                       "l2_policy": f"inputs larger than L2: per-step working set ~{ws_mb:.0f} MB > 126 MB; {NCY} cameras visited "
                                    "round-robin, so consecutive steps do not share a view",
                       "parallelism": f"view-parallel dp{world}" + ((" + gradient sum of 59 floats/Gaussian per step: " + ("one NVLink peer-memory kernel (gvd_exchange_allreduce_sum)"
                                       if args.impl == "ours" else "NCCL all-reduce")) if world > 1 else "")},
            "e2e": {"value": round(e2e_views_per_s, 2), "unit": "views/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / K, 4), "wall_ms_per_step": round(wall_e2e / K * 1e3, 4)},
            "gpu_launches": ((14 + (1 if world > 1 else 0)) * K * 2) if args.impl == "ours" else 0,
            "gpu_launches_note": "own kernels in the two timed regions, per step: preprocess, compact, sort_pass x4, bin_count, "
                                 "bin_prefix, bin_ranges, bin_fill, render_fwd, zero_fill, render_bwd, gaussian_bwd = 14, + "
                                 "grad_allreduce_kernel at N > 1 (no library kernel is left in the rasterizer; the cuBLAS dot of "
                                 "the e2e result is not counted)",
            "clocks": clocks,
        }
        if roofline:
            line["roofline"] = roofline
        if exchange_check is not None:
            line["exchange_check"] = exchange_check
        if denoise_result is not None:
            for k in ("guided", "c5", "train_step"):
                if isinstance(denoise_result, dict) and k in denoise_result:
                    line[k] = denoise_result.pop(k)
            if denoise_result:
                line["denoise"] = denoise_result
        if args.impl == "reference":
            line["cpu_baseline"] = {"value": line["value"], "unit": "views/s", "cores": 0, "kind": "reference",
                                    "sample": "the reference has no CPU rasterizer: this arm is its own CUDA code "
                                              "(oracle/_ref, compiled unmodified for sm_100a) on the same B200, full workload"}
        elif world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(P, W, H, seed, D)
        print(json.dumps(line))
        sys.stdout.flush()

    c5_result = None
    if world > 1 and not args.no_c5 and args.workload == "C2":
        try:
            del sc, leaves, means2D
            torch.cuda.empty_cache()
            c5_result = c5_bench(pkg, args.impl, dev, world, rank)
        except Exception as ex:  # a secondary block must never take the headline line down
            c5_result = {"error": repr(ex)[:300]}
        sc = leaves = means2D = None

    denoise_result = None
    if not args.no_denoise and (world == 1 or args.impl == "ours"):
        # every rank takes part: at N > 1 the DDIM step is CFG-split x frame-sharded (vc_b200/frame_parallel.py)
        guard = None
        if world > 1:
            # a peer that died (or a mismatched collective) would park this rank inside NCCL until its watchdog aborts
            # the process -- and the headline line with it: after DENOISE_DEADLINE_S print what was measured and leave
            def bail():
                if rank == 0:
                    emit(dict({"error": f"secondary metric did not finish within {DENOISE_DEADLINE_S} s at N = {world}"},
                              **({"c5": c5_result} if c5_result else {})))
                sys.stdout.flush()
                os._exit(0)
            guard = threading.Timer(DENOISE_DEADLINE_S, bail)
            guard.daemon = True
            guard.start()
        try:
            sc = leaves = means2D = None
            torch.cuda.empty_cache()
            denoise_result = diffusion_bench(args.impl, dev, world=world, guided=not args.no_guided)
            if guard is not None:
                guard.cancel()
        except Exception as ex:  # the secondary metric must never take the headline line down
            denoise_result = {"error": repr(ex)[:300]}
            if world > 1:
                # a rank that failed alone would leave its peers waiting inside a collective until the NCCL watchdog
                # kills the job -- and the headline line with it.  Print what was measured and leave.
                if rank == 0:
                    emit(dict(denoise_result, **({"c5": c5_result} if c5_result else {})))
                sys.stdout.flush()
                os._exit(0)
    if c5_result is not None:
        denoise_result = dict(denoise_result or {}, c5=c5_result)
    train_result = None
    if world == 1 and not args.no_denoise and args.workload == "C2":
        # whole training iteration (render + L1/SSIM loss + backward + densification statistics + Adam), SURVEY 8 row f3:
        # ours with the activations folded into the rasterizer kernels; the reference arm = its rasterizer + its loss /
        # optimizer statements in torch (tools/bench_train_step.py)
        try:
            sc = leaves = means2D = None
            torch.cuda.empty_cache()
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_train_step
            train_result = bench_train_step.run("folded" if args.impl == "ours" else "reference", P, W, H, seed, D, 100)
        except Exception as ex:  # a secondary block must never take the headline line down
            train_result = {"error": repr(ex)[:300]}
        denoise_result = dict(denoise_result or {}, train_step=train_result)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    emit(denoise_result)
    if world > 1:
        dist.destroy_process_group()


DENOISE_DEADLINE_S = 480


def check_exchange(exchange, dev, world):
    """Correctness of the gradient sum at this N, visible to whoever reads the line: random payload, the library's kernel
    against NCCL's all-reduce of the same data, and whether every rank ended with the same bits."""
    n = exchange.n_floats
    g = torch.Generator(device=dev).manual_seed(4242 + dist.get_rank())
    src = torch.randn(n, device=dev, generator=g)
    ref = src.clone()
    dist.all_reduce(ref)
    exchange.buffer.copy_(src)
    exchange.allreduce()
    torch.cuda.synchronize()
    diff = float((exchange.buffer - ref).abs().max())
    scale = float(ref.abs().max())
    chk = exchange.buffer.view(torch.int32).to(torch.int64).sum().reshape(1)
    lst = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    return {"max_abs_diff_vs_nccl": diff, "max_abs_value": scale, "identical_across_ranks": bool(all(int(x) == int(lst[0]) for x in lst)),
            "floats": n, "mode": exchange.mode}


def c5_bench(pkg, impl, dev, world, rank, steps=40, warm=5):
    """BASELINE.json configs[4]: P = 2 M Gaussians, one 640x480 view per rank per step, the 496 MB gradient sum every step."""
    import synth
    P, W, H, seed, D, desc = WORKLOADS["C5"]
    sc = synth.synth_scene(P, seed, device=dev)
    cams = [synth.synth_camera(seed + 1 + rank + 101 * j, W, H, device=dev) for j in range(NCY)]
    bg = torch.zeros(3, device=dev)
    cot = torch.randn(5, H, W, generator=torch.Generator().manual_seed(seed + 2 + rank)).to(dev)
    step, leaves, means2D = make_step(pkg, sc, cams[0], bg, D)
    exchange = None
    if impl == "ours":
        exchange = view_parallel.GradientExchange(pkg.gradient_buffer_floats(P), dev)
        pkg.set_gradient_buffer(exchange.buffer)
    named = dict(leaves, means2D=means2D)

    def one(k):
        c = cams[k % NCY]
        step(cot, c["viewmatrix"], c["projmatrix"], c["campos"])
        if exchange is not None:
            view_parallel.allreduce_gradients(None, exchange=exchange, leaves=named, views=pkg.gradient_views(dev))
        else:
            view_parallel.allreduce_gradients(grad_flat(leaves))

    for k in range(warm):
        one(k)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        one(k)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    mode = exchange.mode if exchange is not None else "nccl"
    if exchange is not None:
        pkg.set_gradient_buffer(None)
        exchange.close()
    return {"metric": "3DGS train-step views/sec (rasterizer fwd+bwd)", "value": round(world * steps / (ms.item() * 1e-3), 2), "unit": "views/s",
            "ms_per_step": round(ms.item() / steps, 4), "steps": steps, "n_gpus": world,
            "config": {"workload": "C5", "description": desc, "P": P, "width": W, "height": H, "views_per_step": world,
                       "gradient_sum_MB_per_step": round(pkg.gradient_buffer_floats(P) * 4 / 1e6, 1) if impl == "ours" else round(P * 59 * 4 / 1e6, 1),
                       "exchange": mode}}


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def _ref_latent_model(ref, dev, decoder=None):
    """The slice of the reference LatentDiffusion its samplers touch (ddpm3d.py:123-151,239-251,519-527,674-675), with the
    schedule built by the reference's own helpers -- the reference class itself needs pytorch_lightning.  Test harness
    code (tests/test_guided_cpu.py::_reference_sampler) moved to the device."""
    import numpy as np
    import unet_ref
    if unet_ref.REF_VC not in sys.path:
        sys.path.insert(0, unet_ref.REF_VC)
    from lvdm.models.utils_diffusion import make_beta_schedule, rescale_zero_terminal_snr

    class Wrapper(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.diffusion_model = m

    class Model:
        parameterization = "v"
        use_dynamic_rescale = True
        num_timesteps = 1000

        def __init__(self):
            betas = rescale_zero_terminal_snr(make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012))
            ac = np.cumprod(1. - betas, axis=0)
            t32 = lambda a: torch.tensor(a, dtype=torch.float32, device=dev)  # noqa: E731
            self.device = dev
            self.betas, self.alphas_cumprod = t32(betas), t32(ac)
            self.alphas_cumprod_prev = t32(np.append(1., ac[:-1]))
            self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod = t32(np.sqrt(ac)), t32(np.sqrt(1. - ac))
            self.scale_arr = t32(np.concatenate((np.linspace(1.0, 0.3, 400), np.full(1000, 0.3))))
            self.model, self.first_stage_model = Wrapper(ref), decoder

        def apply_model(self, x, t, c, fs=None, **kw):
            return ref(torch.cat([x] + c["c_concat"], 1), t, context=torch.cat(c["c_crossattn"], 1), fs=fs)

        def predict_start_from_z_and_v(self, x_t, t, v):
            return self.sqrt_alphas_cumprod[t].view(-1, 1, 1, 1, 1) * x_t - self.sqrt_one_minus_alphas_cumprod[t].view(-1, 1, 1, 1, 1) * v

        def predict_eps_from_z_and_v(self, x_t, t, v):
            return self.sqrt_alphas_cumprod[t].view(-1, 1, 1, 1, 1) * v + self.sqrt_one_minus_alphas_cumprod[t].view(-1, 1, 1, 1, 1) * x_t

        def differentiable_decode_first_stage(self, z, **kw):
            return decoder(z)

    return Model()


def _timed_steps(fn, steps, warm, dev, world):
    for i in range(warm):
        fn(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.reset_peak_memory_stats()
    e0.record()
    for i in range(steps):
        fn(warm + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:  # the step ends when the slowest rank has its latent: max over ranks
        tmax = torch.tensor([ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    return ms


def diffusion_bench(impl, dev, world=1, guided=True, t=25, h=72, w=128):
    """Secondary metrics of BASELINE.json on the full-size ViewCrafter U-Net (1.44 B parameters, seeded random weights,
    SURVEY.md section 8d), built once and used by both blocks:
      denoise  DDIM denoise-steps/s at configs[2] (25 frames, 576x1024 -> latent 72x128; cond + uncond U-Net forward + sampler
               update per step).  'ours' = vc_b200 (tcgen05 GEMM / flash attention, fused update); 'reference' = the reference
               UNetModel under torch.autocast(bfloat16) driven by the reference's OWN DDIMSampler.p_sample_ddim
               (lvdm/models/samplers/ddim.py:206-280) -- no library of this repository is loaded in that arm.
      guided   guided DDIM steps/s at the configs[3] shape (25 frames, latent 40x64 -> 320x512 images): two U-Net forwards
               with the tape, 25 VAE decodes with the tape, the guidance loss, both backward passes, the update
               (ddim_guidance.py:259-337).  N = 1, and N = 2 for 'ours' (cfg-split GuidedPlan)."""
    import unet_ref
    if not unet_ref.ref_available():
        return {"unavailable": "oracle/_ref/ViewCrafter not installed (python oracle/build_ref.py vc)"}
    ref, cfg = unet_ref.build_reference_unet(model_channels=320, device=dev)
    x, cc, ctx, ctx_uc = unet_ref.synth_inputs(t, h, w, device=dev)
    fs = torch.tensor([10], device=dev)
    cond = {"c_concat": [cc], "c_crossattn": [ctx]}
    uc = {"c_concat": [cc], "c_crossattn": [ctx_uc]}
    noise = torch.randn(x.shape, generator=torch.Generator().manual_seed(7)).to(dev)
    flops = 2 * 82.76e12
    ref_cpu, plan, unet = None, None, None
    if impl == "ours":
        from vc_b200.sampler import DDIMSampler
        from vc_b200.schedule import ModelSchedule
        from vc_b200.unet import DiffusionModelB200, UNetB200
        if world > 1:
            from vc_b200.frame_parallel import DenoisePlan
            plan = DenoisePlan(t)
        unet = UNetB200(ref.state_dict(), device=dev, **cfg)
        model = DiffusionModelB200(unet, ModelSchedule(), plan=plan)
        if world == 1 and os.environ.get("GVD_BENCH_CPU_UNET", "1") == "1":
            try:
                ref_cpu = ref.cpu()  # kept for the CPU row below (BASELINE.md CPU row 5); leaves the GPU
            except Exception:        # the CPU row is a reported extra: never let it cost the GPU measurement
                ref_cpu = None
        del ref
        ref = None
        torch.cuda.empty_cache()
        sampler = DDIMSampler(model)
        sampler.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0)

        def one(i):
            index = 49 - (i % 40)
            ts = torch.full((1,), int(sampler.ddim_timesteps[index]), device=dev, dtype=torch.long)
            return sampler.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5,
                                         unconditional_conditioning=uc, guidance_rescale=0.7, noise=noise, fs=fs)
        steps, warm = 6, 3
    else:
        if unet_ref.REF_VC not in sys.path:
            sys.path.insert(0, unet_ref.REF_VC)
        import lvdm.models.samplers.ddim as ddim_mod
        ddim_mod.DDIMSampler.register_buffer = lambda self, name, attr: setattr(self, name, attr.to(dev) if torch.is_tensor(attr) else attr)
        ddim_mod.noise_like = lambda shape, device, repeat=False: noise
        sampler = ddim_mod.DDIMSampler(_ref_latent_model(ref, dev))
        sampler.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0, verbose=False)

        def one(i):
            index = 49 - (i % 40)
            ts = torch.full((1,), int(sampler.ddim_timesteps[index]), device=dev, dtype=torch.long)
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                return sampler.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5,
                                             unconditional_conditioning=uc, guidance_rescale=0.7, fs=fs)
        steps, warm = 3, 1
    ms = _timed_steps(one, steps, warm, dev, world)
    res = _denoise_line(ms, steps, warm, t, h, w, world, plan, flops)
    if guided and (world == 1 or (impl == "ours" and world == 2)):
        try:
            res["guided"] = guided_bench(impl, dev, world, ref, unet, cfg)
        except Exception as ex:
            res["guided"] = {"error": repr(ex)[:300]}
    if ref_cpu is not None:
        res["cpu_baseline"] = cpu_unet_row(ref_cpu, flops)
    return res


def guided_bench(impl, dev, world, ref, unet, cfg, t=25, h=40, w=64, decode_frames=5):
    """See diffusion_bench.  `ref` = the reference UNetModel (reference arm), `unet` = the UNetB200 built from it (ours)."""
    import test_guided_cpu as tg   # the LossGuidance stand-in (SURVEY.md 8b protocol); reference-sampler harness
    import test_vae_cpu as tv      # the reference VAE decoder with seeded weights
    import unet_ref
    vae = tv.RefFirstStage(ch=128).to(dev).eval()
    x, cc, ctx, ctx_uc = unet_ref.synth_inputs(t, h, w, device=dev)
    cond, uc = {"c_concat": [cc], "c_crossattn": [ctx]}, {"c_concat": [cc], "c_crossattn": [ctx_uc]}
    fs = torch.tensor([10], device=dev)
    g = torch.Generator().manual_seed(123)
    targets = [(torch.rand(3, 8 * h, 8 * w, generator=g) * 2 - 1).to(dev) for _ in range(t)]
    masks = [(torch.rand(1, 8 * h, 8 * w, generator=g) > 0.3).float().to(dev) for _ in range(t)]
    index = 30
    # 2 U-Net forwards + their input-gradient (2x a forward) + 25 decoder forward + latent-gradient
    flops = 2 * 20.19e12 * 3 + t * 1.56e12 * 3
    lg = tg.StubGuidance(targets, masks, 1)
    plan_txt = "single GPU"
    if impl == "ours":
        from vc_b200.guided import DDIMSamplerGuidance, GuidedPlan
        from vc_b200.schedule import ModelSchedule
        from vc_b200.unet import DiffusionModelB200
        from vc_b200.vae import DecoderB200
        model = DiffusionModelB200(unet, ModelSchedule())
        dec = DecoderB200(vae.state_dict(), device=dev, scale_factor=tv.SCALE)
        del vae
        model.differentiable_decode_first_stage = dec.differentiable_decode
        model.guided_decode_frames = decode_frames
        s = DDIMSamplerGuidance(model)
        s.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0)
        if world > 1:
            gp = GuidedPlan(t, model)
            plan_txt = f"cfg{gp.denoise.cfg_ways} x frames{gp.denoise.frame_ways}, decoder frames dealt over {world} ranks"
        ts = torch.full((1,), int(s.ddim_timesteps[index]), dtype=torch.long, device=dev)

        def one(i):
            s.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5, unconditional_conditioning=uc,
                            guidance_rescale=0.7, fs=fs, loss_guidance_fn=lg)
        steps, warm = 3, 1
    else:
        class PerFrame(torch.nn.Module):  # decode_core's per-frame loop (ddpm3d.py:646-668)
            def __init__(self):
                super().__init__()
                self.vae = vae

            def forward(self, z):
                return torch.stack([self.vae(z[:, :, f])[0] for f in range(z.shape[2])], dim=1).unsqueeze(0)

        if unet_ref.REF_VC not in sys.path:
            sys.path.insert(0, unet_ref.REF_VC)
        import lvdm.models.samplers.ddim_guidance as dg
        dg.DDIMSamplerGuidance.register_buffer = lambda self, name, attr: setattr(self, name, attr.to(dev) if torch.is_tensor(attr) else attr)
        # third_party/ViewCrafter/viewcrafter.py:322 sets unet_config.use_checkpoint = True before it builds the model: the
        # reference's guided step recomputes every ResBlock / transformer block in its backward instead of keeping their
        # activations.  Without it this arm sat at 172 of 180 GB and, on the boxes of the round's last runs, in allocator
        # retries (15 s per step instead of 2.3-3.2 s).
        for mod in ref.modules():
            if isinstance(getattr(mod, "use_checkpoint", None), bool):
                mod.use_checkpoint = True
            if isinstance(getattr(mod, "checkpoint", None), bool):
                mod.checkpoint = True
        s = dg.DDIMSamplerGuidance(_ref_latent_model(ref, dev, PerFrame()))
        s.make_schedule(50, ddim_discretize="uniform_trailing", ddim_eta=1.0, verbose=False)
        ts = torch.full((1,), int(s.ddim_timesteps[index]), dtype=torch.long, device=dev)

        def one(i):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                s.p_sample_ddim(x, cond, ts, index=index, unconditional_guidance_scale=7.5, unconditional_conditioning=uc,
                                guidance_rescale=0.7, fs=fs, loss_guidance_fn=lg)
        steps, warm = 2, 1
    torch.cuda.empty_cache()
    ms = _timed_steps(one, steps, warm, dev, world)
    peak = float(_peaks().get("bf16_tflops_sustained", 1397.1)) * world
    return {"metric": "guided DDIM steps/sec", "impl": impl, "value": round(1e3 / ms, 4), "unit": "steps/s", "ms_per_step": round(ms, 1),
            "steps": steps, "warmup": warm, "dtype": "bf16", "tflops_per_s": round(flops / ms / 1e9, 1),
            "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1), "scaling": "strong",
            "roofline": {"bound": "tensor", "achieved": round(flops / ms / 1e9, 1), "peak": peak, "unit": "TFLOP/s",
                         "frac": round(flops / ms / 1e9 / peak, 4), "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained"},
            "config": {"workload": "C4 guided step", "frames": t, "latent": [h, w], "cfg": 7.5, "recur_steps": 1, "n_gpus": world,
                       "unet_params_M": 1438.9, "vae_ch": 128, "decode_frames_per_call": decode_frames if impl == "ours" else 1,
                       "parallelism": plan_txt}}


def cpu_unet_row(ref_cpu, step_flops, frames=2, h=40, w=56):
    """BASELINE.md CPU row 5: the reference UNetModel in fp32 on the host cores.  A bounded sample -- one forward at the
    reference's training resolution (latent 40x56) with `frames` frames instead of 25 -- turned into a FLOP-scaled
    denoise-steps/s estimate for the benchmarked shape, and labelled as an estimate."""
    try:
        g = torch.Generator().manual_seed(11)
        xc = torch.randn(1, 8, frames, h, w, generator=g)
        ctxc = torch.randn(1, 333, 1024, generator=g)
        ts, fs = torch.tensor([481]), torch.tensor([10])
        best = 1e30
        with torch.no_grad():
            for _ in range(2):
                t0 = time.perf_counter()
                ref_cpu(xc, ts, context=ctxc, fs=fs)
                best = min(best, time.perf_counter() - t0)
        fl = 17.59e12 * frames / 25.0
        return {"value": round(fl / best / step_flops, 6), "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "reference",
                "cpu_tflops_per_s": round(fl / best / 1e12, 3),
                "sample": f"reference UNetModel fp32 on the host, one forward at [1,8,{frames},{h},{w}] ({fl / 1e12:.2f} TFLOP, best of 2: "
                          f"{best:.2f} s); value = FLOP-scaled ESTIMATE of denoise-steps/s at the benchmarked shape (a step = "
                          f"{step_flops / 1e12:.1f} TFLOP), not a measured step"}
    except Exception as ex:
        return {"value": None, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": "reference", "sample": f"unavailable: {ex!r}"[:200]}


def _denoise_line(ms, steps, warm, t, h, w, world, plan, flops):
    pk = _peaks()
    peak = float(pk.get("bf16_tflops_sustained", 1397.1)) * world
    return {"metric": "DDIM denoise-steps/sec", "value": round(1e3 / ms, 4), "unit": "steps/s", "ms_per_step": round(ms, 2),
            "steps": steps, "warmup": warm, "dtype": "bf16", "config": {"workload": "C3", "frames": t, "latent": [h, w], "cfg": 7.5,
            "ddim_steps": 50, "unet_params_M": 1438.9, "n_gpus": world,
            "parallelism": "single GPU" if plan is None else f"cfg{plan.cfg_ways} x frames{plan.frame_ways} (all-to-all re-shard "
            "around temporal layers, all-gather of output frames)"}, "scaling": "strong", "tflops_per_s": round(flops / (ms * 1e-3) / 1e12, 1),
            "roofline": {"bound": "tensor", "achieved": round(flops / (ms * 1e-3) / 1e12, 1), "peak": peak, "unit": "TFLOP/s",
                         "frac": round(flops / (ms * 1e-3) / 1e12 / peak, 4),
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if pk else "fallback 1397.1 TFLOP/s"}}


def cpu_baseline(P, W, H, seed, D, budget_s=20.0):
    """Times the C restatement (oracle/) on the host: a bounded sample of the same workload."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import raster_oracle
        cb = raster_oracle.timed_sample(P, W, H, seed, D, budget_s=budget_s)
    except Exception as ex:  # the baseline is a reported number, never the product path
        cb = {"value": None, "unit": "views/s", "cores": 1, "kind": "port", "sample": f"unavailable: {ex!r}"}
    cb["also"] = cpu_point_render_rows()
    return cb


def cpu_point_render_rows():
    """BASELINE.md CPU row 4 (configs[0]): the numpy z-buffer point projection (scene/pcd2img.py:4-70, restated in
    oracle/pcd2img_oracle.py and pinned to the reference's outputs) timed on this host at 1 000 points / 128x128 and at
    500 000 points / 640x480.  (The reference's pytorch3d point renderer is not installable offline.)"""
    try:
        import pcd2img_oracle as po
        out = {"what": "scene/pcd2img.py::project_point_cloud_to_image restated in numpy (single thread), best of 5",
               "host_cores": os.cpu_count()}
        for name, (n, w, h) in (("c1_1k_128x128_ms", (1000, 128, 128)), ("point_render_500k_640x480_ms", (500000, 640, 480))):
            pts, col, K, E = po.synth_case(n, w, h, seed=0)
            best = 1e30
            for _ in range(5):
                t0 = time.perf_counter()
                po.project_point_cloud_to_image(pts, col, K, E, w, h)
                best = min(best, time.perf_counter() - t0)
            out[name] = round(best * 1e3, 3)
        return out
    except Exception as ex:
        return {"unavailable": repr(ex)[:200]}


def cpu_reference_line(args, P, W, H, seed, D, desc, K, Wm):
    cb = cpu_baseline(P, W, H, seed, D, budget_s=60.0)
    return {"metric": "3DGS train-step views/sec (rasterizer fwd+bwd)", "value": cb.get("value"), "unit": "views/s",
            "n_gpus": 1, "steps": K, "warmup": Wm, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": args.workload, "description": desc}, "cpu_baseline": cb,
            "e2e": {"value": cb.get("value"), "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


if __name__ == "__main__":
    main()
