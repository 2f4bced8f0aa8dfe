import sys; sys.path.insert(0,"tests"); sys.path.insert(0,"guidedvd-3dgs_b200")
import torch, refload
from simple_knn._C import distCUDA2
ref = refload.ref_knn()
for P in (2,3,4):
    g = torch.Generator().manual_seed(77+P); pts = torch.randn(P,3,generator=g).cuda()
    print(P, 'ours', distCUDA2(pts)); print(P, 'ref ', ref.distCUDA2(pts))
    print(torch.cdist(pts,pts)**2)
