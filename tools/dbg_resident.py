import sys, time, os
sys.path.insert(0,'tests'); sys.path.insert(0,'guidedvd-3dgs_b200'); sys.path.insert(0,'.')
import torch, synth, bench
import diff_gaussian_rasterization as pkg
P,W,H,seed,D,_ = bench.WORKLOADS['C2']
dev=torch.device('cuda:0')
sc = synth.synth_scene(P, seed, device=dev); cam = synth.synth_camera(seed+1, W, H, device=dev); bg=torch.zeros(3,device=dev)
cot = torch.randn(5,H,W,device=dev)
step, leaves, m2 = bench.make_step(pkg, sc, cam, bg, D)
def run(n, sync):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n):
        step(cot, cam['viewmatrix'], cam['projmatrix'], cam['campos'])
        if sync: torch.cuda.synchronize()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
for _ in range(5): step(cot, cam['viewmatrix'], cam['projmatrix'], cam['campos'])
st0=torch.cuda.memory_stats()
for rep in range(3):
    print('nosync', round(run(100, False),4), 'sync', round(run(100, True),4))
st1=torch.cuda.memory_stats()
print({k:(st1[k]-st0[k]) for k in ('num_device_alloc','num_device_free','num_alloc_retries')}, 'reserved MB', st1['reserved_bytes.all.current']/1e6, 'alloc MB', st1['allocated_bytes.all.peak']/1e6)
