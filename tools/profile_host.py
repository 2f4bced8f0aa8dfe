"""Where the HOST time of one rasterizer step goes (Python + autograd + ctypes): cProfile over N fwd+bwd steps through the
public API, device work left asynchronous.  bench.py's end-to-end figure is bounded by this when it exceeds the GPU time."""
import cProfile, os, pstats, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "guidedvd-3dgs_b200")); sys.path.insert(0, os.path.join(HERE, "..", "tests")); sys.path.insert(0, os.path.join(HERE, ".."))
import torch
import bench, synth
import diff_gaussian_rasterization as pkg
dev = torch.device("cuda", 0)
P, W, H, seed, D, _ = bench.WORKLOADS["C2"]
sc = synth.synth_scene(P, seed, device=dev)
cam = synth.synth_camera(seed + 1, W, H, device=dev)
bg = torch.zeros(3, device=dev)
cot = torch.randn(5, H, W, device=dev)
step, leaves, m2d = bench.make_step(pkg, sc, cam, bg, D)
for _ in range(50):
    step(cot, cam["viewmatrix"], cam["projmatrix"], cam["campos"])
torch.cuda.synchronize()
N = 400
# (a) host time per step with the GPU kept far behind (tiny image -> the device is never the bottleneck is NOT what we want:
# measure at the real size, but time only the host side: total wall of issuing N steps, then the sync separately)
t0 = time.perf_counter()
for _ in range(N):
    step(cot, cam["viewmatrix"], cam["projmatrix"], cam["campos"])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"issue {1e6 * (t1 - t0) / N:.1f} us/step, drain {1e3 * (t2 - t1):.2f} ms  (issue ~ max(host, device))")
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    step(cot, cam["viewmatrix"], cam["projmatrix"], cam["campos"])
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
