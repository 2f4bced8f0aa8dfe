"""Drop-in for the reference's pybind module `simple_knn._C` (simple-knn/ext.cpp:15-17):

    distCUDA2(points[P,3] float32 cuda) -> (meanDist2[P] float32, nearestIdx[P,3] int32)

Mean squared distance to the 3 nearest other points and their indices (simple-knn/spatial.cu:15-27),
computed by the sm_100a library behind include/gvd_knn.h.  No fallback path exists.
"""
import ctypes as C
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import gvd_native as _n  # noqa: E402


def distCUDA2(points):
    lib = _n.knn()
    if points.dim() != 2 or points.size(1) != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    if not points.is_cuda or points.dtype != torch.float32:
        raise RuntimeError("points must be a float32 CUDA tensor")
    pts = points.detach().contiguous()
    P = pts.size(0)
    dev = pts.device
    means = torch.zeros(P, dtype=torch.float32, device=dev)
    idx = torch.zeros(P, 3, dtype=torch.int32, device=dev)
    if P == 0:
        return means, idx
    nbytes = int(lib.gvd_knn3_tmp_bytes(P))
    tmp = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.gvd_knn3(P, pts.data_ptr(), means.data_ptr(), idx.data_ptr(), tmp.data_ptr(), nbytes,
                          C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError("gvd_knn3 failed: " + (lib.gvd_knn_last_error() or b"").decode())
    return means, idx
