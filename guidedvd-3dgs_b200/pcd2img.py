"""Drop-in for the reference's `scene/pcd2img.py` (project_point_cloud_to_image, :4-70): same name, arguments and
return values (numpy in, numpy out), computed by the sm_100a z-buffer kernels behind include/gvd_points.h.

    from pcd2img import project_point_cloud_to_image        # guidedvd-3dgs_b200/ on sys.path
    image, mask = project_point_cloud_to_image(points, colors, K, E, width, height, near=0.1, far=1000.0)

`project_points_cuda` is the device-resident form (torch CUDA tensors in and out, stream-ordered, no synchronisation).
There is no CPU path: a missing library or GPU raises.
"""
import ctypes as C

import numpy as np
import torch

import gvd_native as _n


def project_points_cuda(points, colors, intrinsics, extrinsics, width, height, near=0.1, far=1000.0):
    """points [n,3] float64 cuda, colors [n,3] uint8 cuda; intrinsics [3,3], extrinsics [4,4] host (numpy / CPU tensor)
    -> image [H,W,3] uint8 cuda, mask [H,W] uint8 cuda."""
    lib = _n.points()
    if points.dim() != 2 or points.shape[1] != 3 or colors.shape != points.shape:
        raise ValueError("points and colors must both be [n, 3]")
    dev = points.device
    pts = points.to(torch.float64).contiguous()
    col = colors.to(torch.uint8).contiguous()
    Kh = np.ascontiguousarray(np.asarray(intrinsics, dtype=np.float64).reshape(3, 3))
    Eh = np.ascontiguousarray(np.asarray(extrinsics, dtype=np.float64).reshape(4, 4))
    image = torch.empty(height, width, 3, dtype=torch.uint8, device=dev)
    mask = torch.empty(height, width, dtype=torch.uint8, device=dev)
    nbytes = int(lib.gvd_point_project_scratch_bytes(int(width), int(height)))
    scratch = torch.empty(nbytes // 8 + 1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.gvd_point_project(pts.data_ptr(), col.data_ptr(), pts.shape[0], Kh.ctypes.data_as(C.c_void_p),
                                   Eh.ctypes.data_as(C.c_void_p), int(width), int(height), float(near), float(far),
                                   image.data_ptr(), mask.data_ptr(), scratch.data_ptr(), nbytes,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError("gvd_point_project failed: " + (lib.gvd_points_last_error() or b"").decode())
    return image, mask


def project_point_cloud_to_image(point_cloud, colors, intrinsics, extrinsics, width, height, near=0.1, far=1000.0):
    """scene/pcd2img.py:4-70: (N,3) points, (N,3) uint8 colours, (3,3) K, (4,4) E -> (H,W,3) uint8 image, (H,W) uint8 mask."""
    dev = torch.device("cuda", torch.cuda.current_device())
    pts = torch.from_numpy(np.ascontiguousarray(np.asarray(point_cloud, dtype=np.float64)).reshape(-1, 3)).to(dev)
    col = torch.from_numpy(np.ascontiguousarray(np.asarray(colors).astype(np.uint8)).reshape(-1, 3)).to(dev)
    image, mask = project_points_cuda(pts, col, intrinsics, extrinsics, width, height, near, far)
    return image.cpu().numpy(), mask.cpu().numpy()
