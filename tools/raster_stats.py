import sys, ctypes as C
sys.path.insert(0,'tests'); sys.path.insert(0,'guidedvd-3dgs_b200')
import torch, synth, parity_raster as pr, gvd_native
import diff_gaussian_rasterization as ours
for cfg in ['C2','C4']:
    P,W,H,seed = synth.CONFIGS[cfg]
    sc, cam, cot, bg, D = pr.make_inputs(P,W,H,seed,3)
    o = pr.run(ours, sc, cam, cot, bg, D, backward=False)
    R = o['num_rendered']; lib = gvd_native.raster(); L = gvd_native.RasterLayout(); lib.gvd_raster_layout(P,R,W,H,C.byref(L))
    packed = o['binning'][L.bin_packed:L.bin_packed+48*R].view(torch.int32).view(R,12)
    mask = packed[:,11] & 0xff
    v = pr.ours_views(o,P,W,H)
    lens = (v['ranges'][:,1]-v['ranges'][:,0]).float()
    nc = v['n_contrib'].view(H,W).float()
    tiles_x=(W+15)//16
    # per tile max n_contrib
    import torch.nn.functional as F
    ncp = F.pad(nc,(0,(16-W%16)%16,0,(16-H%16)%16))
    tmax = ncp.view(ncp.shape[0]//16,16,ncp.shape[1]//16,16).amax(dim=(1,3)).flatten()
    pop = torch.zeros_like(mask)
    for b in range(8): pop += (mask>>b)&1
    print(cfg, 'R',R,'visible',int((o['radii']>0).sum()),'tiles',lens.numel(),'avg len',lens.mean().item(),'max len',lens.max().item())
    print('  mask==0 frac', (mask==0).float().mean().item(), 'avg popcount', pop.float().mean().item(), 'popcount hist', torch.bincount(pop.long(), minlength=9).tolist())
    print('  sum over tiles of max n_contrib / R =', (tmax.sum()/R).item(), ' mean n_contrib/len', (nc.mean()/lens.mean()).item())
    # warp-level: fraction of (warp, entry) pairs visited in fwd approx: entries up to tile max n_contrib with mask bit
    radii=o['radii'][o['radii']>0].float(); print('  radius px: mean',radii.mean().item(),'median',radii.median().item(),'p90',radii.quantile(0.9).item())
    tt = v['tiles_touched'][v['tiles_touched']>0].float(); print('  tiles/gaussian mean',tt.mean().item(),'median',tt.median().item(),'max',tt.max().item())
