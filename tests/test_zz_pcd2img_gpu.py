"""GPU parity of the point-cloud z-buffer projection (row a22, include/gvd_points.h) through the drop-in
`pcd2img.project_point_cloud_to_image`: bit-exact against the golden outputs of the reference function
(tests/golden/pcd2img_*.npz, scene/pcd2img.py:4-70) and against the oracle at BASELINE's larger "point render" size
(500 000 points, 640x480), plus the edge cases.  Integer / byte outputs: the bar is equality.

Green on B200 since round 2 (profiles/r02_first_hw_run.txt); the same source also runs on the host (tests/test_pcd2img_cpu.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("name", ["c1", "dense", "wide"])
def test_matches_reference_golden(name):
    import pcd2img
    import pcd2img_oracle as po
    from make_golden_pcd2img import CASES

    n, w, h, seed, spread, near, far = CASES[name]
    g = np.load(os.path.join(ROOT, "tests", "golden", f"pcd2img_{name}.npz"))
    image, mask = pcd2img.project_point_cloud_to_image(*po.synth_case(n, w, h, seed, spread), w, h, near, far)
    assert image.dtype == np.uint8 and mask.dtype == np.uint8
    assert np.array_equal(image, g["image"]) and np.array_equal(mask, g["mask"])


def test_point_render_size_and_edges():
    import pcd2img
    import pcd2img_oracle as po

    pts, col, K, E = po.synth_case(500_000, 640, 480, seed=3, spread=1.0)
    image, mask = pcd2img.project_point_cloud_to_image(pts, col, K, E, 640, 480)
    oi, om = po.project_point_cloud_to_image(pts, col, K, E, 640, 480)
    assert np.array_equal(image, oi) and np.array_equal(mask, om)
    # deterministic run to run (atomicMin on keys, then on indices: no race decides a pixel)
    image2, _ = pcd2img.project_point_cloud_to_image(pts, col, K, E, 640, 480)
    assert np.array_equal(image, image2)
    image, mask = pcd2img.project_point_cloud_to_image(np.zeros((0, 3)), np.zeros((0, 3), np.uint8), K, E, 64, 48)
    assert image.shape == (48, 64, 3) and not image.any() and not mask.any()
    with pytest.raises(ValueError):
        import torch
        pcd2img.project_points_cuda(torch.zeros(4, 2, device="cuda", dtype=torch.float64), torch.zeros(4, 3, device="cuda", dtype=torch.uint8), K, E, 8, 8)
