"""Generate tests/golden/pcd2img_*.npz by running the REFERENCE function (scene/pcd2img.py:4-70) in this container.
Inputs are seeded (oracle/pcd2img_oracle.py::synth_case), so only the outputs are stored.
Run: python tests/make_golden_pcd2img.py   (needs /root/reference; not needed on the GPU box)."""
import hashlib
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pcd2img_oracle as po  # noqa: E402

CASES = {  # name: (points, width, height, seed, spread, near, far)
    "c1": (1000, 128, 128, 0, 1.0, 0.1, 1000.0),          # BASELINE.json configs[0]
    "dense": (60000, 96, 64, 1, 0.6, 0.5, 3.0),           # ~10 points per pixel: the z-buffer decides almost every pixel
    "wide": (5000, 200, 120, 2, 3.0, 0.1, 1000.0),        # most points outside the image or behind the camera
}


def main():
    spec = importlib.util.spec_from_file_location("ref_pcd2img", "/root/reference/scene/pcd2img.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for name, (n, w, h, seed, spread, near, far) in CASES.items():
        pts, col, K, E = po.synth_case(n, w, h, seed, spread)
        image, mask = ref.project_point_cloud_to_image(pts, col, K, E, w, h, near, far)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"pcd2img_{name}.npz"), image=image, mask=mask,
                            params=np.asarray([n, w, h, seed, spread, near, far], dtype=np.float64))
        print(name, "covered pixels", int(mask.sum()), "sha256", hashlib.sha256(image.tobytes()).hexdigest()[:16])


if __name__ == "__main__":
    main()
