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
