"""TEST INFRASTRUCTURE ONLY -- small golden vectors from the UNMODIFIED reference that the other generators do not cover.

    python oracle/gen_golden_misc.py          # rewrites tests/golden/corners_kitti_order.npz

corners_kitti_order: get_corners_of_cuboid(..., iou_3d_convention=False) (lib/math_3d.py:405-426) for N = 4 boxes -- the one
batch size besides 1 for which the reference's torch branch broadcasts (`corners[:, 0, [1,2,3,4]] = l3d`, :422)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402


def main():
    ref = ref_shim.load()
    rng = np.random.default_rng(21)
    b7 = np.stack([rng.uniform(-20, 20, 4), rng.uniform(0.5, 2, 4), rng.uniform(5, 60, 4), rng.uniform(1.4, 1.9, 4),
                   rng.uniform(1.3, 1.8, 4), rng.uniform(3, 5, 4), rng.uniform(-np.pi, np.pi, 4)], 1).astype(np.float32)
    t = [torch.from_numpy(b7[:, i].copy()) for i in range(7)]
    # NOTE the reference's `= l3d` (no unsqueeze) assigns l3d[j] to CORNER j of every box when N == 4, not box i's length to its
    # corners; with identical dimensions for the 4 boxes both readings coincide, so the fixture uses one (w, h, l) for all four
    for i in (3, 4, 5):
        t[i][:] = t[i][0]
        b7[:, i] = b7[0, i]
    out = ref.math_3d.get_corners_of_cuboid(*t, iou_3d_convention=False).numpy()
    path = os.path.join(ROOT, "tests", "golden", "corners_kitti_order.npz")
    np.savez_compressed(path, boxes7=b7, corners=out.astype(np.float32))
    print("wrote", path)


if __name__ == "__main__":
    main()
