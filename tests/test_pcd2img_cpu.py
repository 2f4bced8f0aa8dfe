"""Row a22 (BASELINE.json configs[0]): the point-cloud z-buffer projection.
  * oracle/pcd2img_oracle.py against the golden outputs of the REFERENCE function (scene/pcd2img.py:4-70), bit for bit;
  * the CUDA source csrc/point_project.cu executed on the host (tests/cuda_emu) through its C-ABI entry point against the
    oracle and the goldens, bit for bit, plus the edge cases the routine has (no points, everything culled, one point,
    many points on one pixel, exact-depth ties, out-of-frustum points, degenerate divide);
  * the shared library exports what include/gvd_points.h declares and validates its arguments before any CUDA call."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "guidedvd-3dgs_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"),
          os.path.join(ROOT, "tests", "cuda_emu")):
    if p not in sys.path:
        sys.path.insert(0, p)

import pcd2img_oracle as po  # noqa: E402
from make_golden_pcd2img import CASES  # noqa: E402


def _golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", f"pcd2img_{name}.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name):
    n, w, h, seed, spread, near, far = CASES[name]
    g = _golden(name)
    image, mask = po.project_point_cloud_to_image(*po.synth_case(n, w, h, seed, spread), w, h, near, far)
    assert np.array_equal(image, g["image"]) and np.array_equal(mask, g["mask"])
    assert image.dtype == np.uint8 and mask.dtype == np.uint8 and int(mask.sum()) > 100


@pytest.fixture(scope="module")
def emu():
    import build_emu

    lib = C.CDLL(build_emu.build("point_project"))
    lib.gvd_points_last_error.restype = C.c_char_p
    lib.gvd_point_project_scratch_bytes.restype = C.c_size_t
    lib.gvd_point_project_scratch_bytes.argtypes = [C.c_int, C.c_int]
    lib.gvd_point_project.restype = C.c_int
    lib.gvd_point_project.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double,
                                      C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    return lib


def _run(lib, pts, col, K, E, w, h, near=0.1, far=1000.0):
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
    col = np.ascontiguousarray(col, dtype=np.uint8).reshape(-1, 3)
    K, E = np.ascontiguousarray(K, dtype=np.float64), np.ascontiguousarray(E, dtype=np.float64)
    image = np.full((h, w, 3), 7, dtype=np.uint8)   # poisoned: the call must write every pixel
    mask = np.full((h, w), 7, dtype=np.uint8)
    nb = lib.gvd_point_project_scratch_bytes(w, h)
    scratch = np.zeros(nb // 8 + 1, dtype=np.int64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rc = lib.gvd_point_project(p(pts) if pts.size else None, p(col) if col.size else None, pts.shape[0], p(K), p(E), w, h, near, far,
                               p(image), p(mask), p(scratch), nb, None)
    assert rc == 0, lib.gvd_points_last_error()
    return image, mask


@pytest.mark.parametrize("name", sorted(CASES))
def test_kernels_match_golden_and_oracle(emu, name):
    n, w, h, seed, spread, near, far = CASES[name]
    pts, col, K, E = po.synth_case(n, w, h, seed, spread)
    image, mask = _run(emu, pts, col, K, E, w, h, near, far)
    g = _golden(name)
    assert np.array_equal(image, g["image"]) and np.array_equal(mask, g["mask"])


def test_kernel_edge_cases(emu):
    K = np.array([[50.0, 0, 16], [0, 50.0, 12], [0, 0, 1]])
    E = np.eye(4)
    w, h = 32, 24
    # no points at all / nothing survives the depth filter
    for pts in (np.zeros((0, 3)), np.array([[0.0, 0, 0.05], [0, 0, -3.0], [0, 0, 5000.0]])):
        image, mask = _run(emu, pts, np.full((len(pts), 3), 200, np.uint8), K, E, w, h)
        assert not image.any() and not mask.any()
    # one point in the centre
    image, mask = _run(emu, [[0.0, 0, 2.0]], [[10, 20, 30]], K, E, w, h)
    assert mask.sum() == 1 and tuple(image[12, 16]) == (10, 20, 30)
    # many points on one pixel: the nearest wins whatever its position in the list; exact ties -> lowest index
    pts = np.array([[0.0, 0, 3.0], [0, 0, 1.5], [0, 0, 2.0], [0, 0, 1.5]])
    col = np.array([[1, 1, 1], [2, 2, 2], [3, 3, 3], [4, 4, 4]], dtype=np.uint8)
    image, mask = _run(emu, pts, col, K, E, w, h)
    assert mask.sum() == 1 and tuple(image[12, 16]) == (2, 2, 2)
    oi, om = po.project_point_cloud_to_image(pts, col, K, E, w, h)
    assert np.array_equal(image, oi) and np.array_equal(mask, om)
    # strict near/far, half-to-even rounding at pixel boundaries, points just outside the image
    pts = np.array([[0.0, 0, 0.1], [0.01, 0, 1.0], [0.03, 0, 1.0], [-0.33, 0, 1.0], [0.3199, 0, 1.0], [0, 0.25, 1.0]])
    col = (np.arange(18, dtype=np.uint8).reshape(6, 3) + 1)
    image, mask = _run(emu, pts, col, K, E, w, h)
    oi, om = po.project_point_cloud_to_image(pts, col, K, E, w, h)
    assert np.array_equal(image, oi) and np.array_equal(mask, om)
    # third row of K that makes the divide degenerate for some points (0/0, x/0): rejected like numpy's non-finite pixels
    Kd = np.array([[50.0, 0, 16], [0, 50.0, 12], [1.0, 0, 0]])
    pts = np.array([[0.0, 0, 2.0], [1.0, 0.5, 2.0], [-1.0, 0, 2.0]])
    image, mask = _run(emu, pts, col[:3], Kd, E, w, h)
    oi, om = po.project_point_cloud_to_image(pts, col[:3], Kd, E, w, h)
    assert np.array_equal(image, oi) and np.array_equal(mask, om)


def test_kernels_random_sweep_against_oracle(emu):
    rng = np.random.default_rng(5)
    for trial in range(6):
        w, h = int(rng.integers(3, 70)), int(rng.integers(3, 50))
        n = int(rng.integers(1, 4000))
        pts, col, K, E = po.synth_case(n, w, h, seed=100 + trial, spread=float(rng.uniform(0.3, 3.0)))
        near, far = float(rng.uniform(0.05, 1.0)), float(rng.uniform(2.0, 50.0))
        image, mask = _run(emu, pts, col, K, E, w, h, near, far)
        oi, om = po.project_point_cloud_to_image(pts, col, K, E, w, h, near, far)
        assert np.array_equal(image, oi) and np.array_equal(mask, om), trial


def test_library_exports_and_argument_validation():
    import gvd_native

    lib = gvd_native.points()
    txt = open(os.path.join(ROOT, "include", "gvd_points.h")).read()
    declared = sorted(set(re.findall(r"GVD_POINTS_API\s+[\w\s\*]+?\b(gvd_\w+)\s*\(", txt)))
    assert set(declared) == set(gvd_native.POINTS_SYMBOLS) and len(declared) == 3
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.gvd_point_project_scratch_bytes(640, 480) == 640 * 480 * 12
    assert lib.gvd_point_project_scratch_bytes(0, 480) == 0
    # validation happens before any CUDA call (this container has no GPU)
    buf = (C.c_char * 64)()
    assert lib.gvd_point_project(None, None, 0, None, None, 4, 4, 0.1, 10.0, buf, buf, buf, 64, None) == 2
    assert b"null" in lib.gvd_points_last_error()
    assert lib.gvd_point_project(buf, buf, 1, buf, buf, 0, 4, 0.1, 10.0, buf, buf, buf, 64, None) == 2
    assert lib.gvd_point_project(buf, buf, 1, buf, buf, 4, 4, 0.1, 10.0, buf, buf, buf, 8, None) == 2
    assert b"scratch" in lib.gvd_points_last_error()
