"""CPU restatement of the reference's point-cloud z-buffer projection -- TEST INFRASTRUCTURE (oracle): only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import it; the product (guidedvd-3dgs_b200/pcd2img.py)
never does.

Follows /root/reference/scene/pcd2img.py::project_point_cloud_to_image (:4-70) step by step in float64 numpy:
  :26-31  homogeneous points, extrinsics, near/far filter on camera z (strict inequalities)
  :37-40  intrinsics, perspective divide by the THIRD ROW of K (not by z)
  :43-45  np.round (half to even) and cast to int
  :48-53  bounds filter
  :56-69  nearest point per pixel: the reference argsorts by z and takes np.unique's first occurrence; stated here as
          "minimum z per pixel, lowest original index among exact ties" (the reference's quicksort argsort leaves
          exact-z ties unspecified; none occur in the pinned fixtures).
Pinned: tests/golden/pcd2img_*.npz are outputs of the reference function itself, generated in this container by
tests/make_golden_pcd2img.py (tests/test_pcd2img_cpu.py compares bit for bit).
"""
import numpy as np


def synth_case(n_points, width, height, seed=0, spread=1.0):
    """Seeded inputs shared by the golden generator, the tests and bench.py: a blob of points in front of a camera that
    looks down +z, part of it behind the near plane and outside the frustum, colours uint8."""
    rng = np.random.default_rng(seed)
    pts = rng.normal(size=(n_points, 3)) * np.array([1.2, 0.9, 1.5]) * spread + np.array([0.0, 0.0, 2.0])
    colors = rng.integers(0, 256, size=(n_points, 3), dtype=np.uint8)
    f = 0.8 * width
    K = np.array([[f, 0.0, width / 2.0], [0.0, f, height / 2.0], [0.0, 0.0, 1.0]])
    ang = 0.2
    R = np.array([[np.cos(ang), 0.0, np.sin(ang)], [0.0, 1.0, 0.0], [-np.sin(ang), 0.0, np.cos(ang)]])
    E = np.eye(4)
    E[:3, :3] = R
    E[:3, 3] = [0.1, -0.05, 0.3]
    return pts, colors, K, E


def project_point_cloud_to_image(point_cloud, colors, intrinsics, extrinsics, width, height, near=0.1, far=1000.0):
    image = np.zeros((height, width, 3), dtype=np.uint8)
    mask = np.zeros((height, width), dtype=np.uint8)
    n = point_cloud.shape[0]
    if n == 0:
        return image, mask
    hom = np.hstack((np.asarray(point_cloud, dtype=np.float64), np.ones((n, 1))))
    cam = (np.asarray(extrinsics, dtype=np.float64) @ hom.T).T
    keep = (cam[:, 2] > near) & (cam[:, 2] < far)
    idx = np.nonzero(keep)[0]
    cam = cam[keep]
    img = (np.asarray(intrinsics, dtype=np.float64) @ cam[:, :3].T).T
    with np.errstate(divide="ignore", invalid="ignore"):
        u = np.round(img[:, 0] / img[:, 2])
        v = np.round(img[:, 1] / img[:, 2])
    ok = np.isfinite(u) & np.isfinite(v) & (u >= 0) & (u < width) & (v >= 0) & (v < height)
    u, v, z, idx = u[ok].astype(np.int64), v[ok].astype(np.int64), cam[ok, 2], idx[ok]
    if u.size == 0:
        return image, mask
    pix = v * width + u
    order = np.lexsort((idx, z, pix))          # by pixel, then depth, then original index
    first = np.ones(order.size, dtype=bool)
    first[1:] = pix[order][1:] != pix[order][:-1]
    win = order[first]
    image[v[win], u[win]] = np.asarray(colors)[idx[win]]
    mask[v[win], u[win]] = 1
    return image, mask
