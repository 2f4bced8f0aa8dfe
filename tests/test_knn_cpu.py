"""CPU: the kNN oracle against a plain O(P^2) numpy search, and the kNN C-ABI library's exported symbols."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_knn_oracle_vs_bruteforce():
    import knn_oracle

    rng = np.random.default_rng(3)
    pts = rng.normal(size=(700, 3)).astype(np.float32)
    mean, idx = knn_oracle.knn3(pts)
    d = ((pts[:, None, :].astype(np.float64) - pts[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d, np.inf)
    order = np.argsort(d, 1)[:, :3]
    assert (np.sort(order, 1) == np.sort(idx, 1)).all()
    np.testing.assert_allclose(mean, np.take_along_axis(d, order, 1).mean(1), rtol=1e-5)


def test_knn_oracle_small_counts():
    import knn_oracle

    for P in (1, 2, 3):
        mean, idx = knn_oracle.knn3(np.eye(3, dtype=np.float32)[:P])
        assert mean.shape == (P,) and (mean > 1e37).all() and idx.shape == (P, 3)


def test_knn_library_exports():
    import gvd_native

    lib = gvd_native.knn()
    txt = open(os.path.join(ROOT, "include", "gvd_knn.h")).read()
    declared = sorted(set(re.findall(r"GVD_KNN_API\s+[\w\s\*]+?\b(gvd_\w+)\s*\(", txt)))
    assert set(declared) == set(gvd_native.KNN_SYMBOLS)
    for s in declared:
        assert hasattr(lib, s)
    assert lib.gvd_knn3_tmp_bytes(1000) >= 1000 * (16 + 16)
    assert lib.gvd_knn3(0, None, None, None, None, 0, None) == 0
    assert lib.gvd_knn3(10, None, None, None, None, 0, None) != 0


def test_knn_kernels_on_the_host_match_oracle():
    """csrc/knn.cu executed on the host (tests/cuda_emu: bounding box, Morton codes, CUB sort stand-in, leaf / node boxes,
    the warp-per-query 3-NN walk with its REDUX k-select) through the C ABI: the same neighbour sets and mean squared
    distances as the oracle (and, through it, as the reference's distCUDA2), incl. duplicate points and tiny clouds."""
    import ctypes as C
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))
    import build_emu
    import knn_oracle

    lib = C.CDLL(build_emu.build("knn"))
    lib.gvd_knn_last_error.restype = C.c_char_p
    lib.gvd_knn3_tmp_bytes.restype = C.c_size_t
    lib.gvd_knn3_tmp_bytes.argtypes = [C.c_int]
    lib.gvd_knn3.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    rng = np.random.default_rng(11)
    clouds = [rng.normal(size=(1500, 3)).astype(np.float32),
              np.concatenate([rng.uniform(-1, 1, size=(300, 3)), rng.normal(size=(40, 3)) * 1e-3 + 5.0]).astype(np.float32),
              rng.normal(size=(5, 3)).astype(np.float32)]
    clouds[1][10:14] = clouds[1][9]      # exact duplicates: distance 0 neighbours
    for pts in clouds:
        P = pts.shape[0]
        mean = np.zeros(P, np.float32)
        idx = np.zeros((P, 3), np.int32)
        nb = lib.gvd_knn3_tmp_bytes(P)
        tmp = np.zeros(nb // 8 + 32, np.int64)
        base = tmp.ctypes.data + (-tmp.ctypes.data) % 256
        rc = lib.gvd_knn3(P, pts.ctypes.data, mean.ctypes.data, idx.ctypes.data, base, nb, None)
        assert rc == 0, lib.gvd_knn_last_error()
        o_mean, o_idx = knn_oracle.knn3(pts)
        np.testing.assert_allclose(mean, o_mean, rtol=1e-6, atol=1e-12)
        d = ((pts[:, None, :].astype(np.float64) - pts[None, :, :]) ** 2).sum(-1)
        np.fill_diagonal(d, np.inf)
        got = np.sort(np.take_along_axis(d, idx.astype(np.int64), 1), 1)
        want = np.sort(np.take_along_axis(d, o_idx.astype(np.int64), 1), 1)
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-12)   # same distances (ties may pick a different twin)
