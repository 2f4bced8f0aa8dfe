"""CPU restatement of the reference 3-NN op -- TEST INFRASTRUCTURE (checker only).

Follows submodules/simple-knn/simple_knn.cu:131-148 (updateKBest: strict '>' insertion, so on equal
distances the candidate met first wins) and :150-190 (boxMeanDist: exact 3 nearest OTHER points, result
(d0+d1+d2)/3 with d ascending, float32).  The reference's Morton/box traversal only prunes; the result is
the exact 3-NN set, which is what this brute-force / KD-tree restatement computes.  Squared distances use
nvcc's contraction of dx*dx + dy*dy + dz*dz (fma(dz,dz, fma(dx,dx, dy*dy))), emulated in float64->float32.

Parity status: pinned against the compiled reference (oracle/_ref/simple_knn) on a B200 by
tests/test_knn_gpu.py; the reference ships no tests of its own.
"""
import numpy as np


def _d2(p, q):
    d = (p.astype(np.float32) - q.astype(np.float32)).astype(np.float32)
    dx, dy, dz = d[..., 0].astype(np.float64), d[..., 1].astype(np.float64), d[..., 2].astype(np.float64)
    t = np.float32(dy * dy).astype(np.float64)
    t = np.float32(dx * dx + t).astype(np.float64)
    return np.float32(dz * dz + t)


def knn3(points):
    """points [P,3] float32 -> (mean_d2 [P] float32, idx [P,3] int32). O(P log P) via scipy KD-tree for the
    candidate set (k = 8 to survive float32 ties), exact float32 re-ranking on top."""
    from scipy.spatial import cKDTree

    pts = np.ascontiguousarray(points, dtype=np.float32)
    P = pts.shape[0]
    FLT_MAX = np.float32(np.finfo(np.float32).max)
    best = np.full((P, 3), FLT_MAX, np.float32)
    bidx = np.zeros((P, 3), np.int32)
    if P > 1:
        k = min(P, 9)
        _, cand = cKDTree(pts.astype(np.float64)).query(pts.astype(np.float64), k=k)
        cand = cand.reshape(P, k)
        d = _d2(pts[cand], pts[:, None, :])
        d[cand == np.arange(P)[:, None]] = np.inf
        if k < 3:  # fewer than 3 other points: pad with "no neighbour"
            d = np.concatenate([d, np.full((P, 3 - k), np.inf, d.dtype)], 1)
            cand = np.concatenate([cand, np.zeros((P, 3 - k), cand.dtype)], 1)
        order = np.argsort(d, axis=1, kind="stable")[:, :3]
        dd = np.take_along_axis(d, order, 1)
        ii = np.take_along_axis(cand, order, 1)
        ok = np.isfinite(dd)
        best[ok] = dd[ok]
        bidx[ok] = ii[ok]
    with np.errstate(over="ignore"):
        mean = ((best[:, 0] + best[:, 1]) + best[:, 2]) / np.float32(3.0)
    return mean.astype(np.float32), bidx
