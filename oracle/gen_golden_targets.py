"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/targets_overlaps.npz and iou2d_f64.npz by running the UNMODIFIED reference
(lib/core.py iou / iou_ign, numpy branch, exactly as lib/rpn_util.py:439-461 compute_targets calls them) on CPU in the
build container.  Inputs mirror the call site's dtypes: rois float32 (network anchors), ground truths float64.

    python oracle/gen_golden_targets.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402


def scene(seed, m, g, gi):
    rng = np.random.default_rng(seed)
    c = rng.uniform(0, 1200, (m, 2)); wh = rng.uniform(8, 300, (m, 2))
    rois = np.concatenate([c - wh / 2, c + wh / 2, rng.uniform(0, 1, (m, 1))], 1).astype(np.float32)   # x1,y1,x2,y2,tracker
    gc = rng.uniform(100, 1100, (g, 2)); gwh = rng.uniform(20, 250, (g, 2))
    gts = np.concatenate([gc - gwh / 2, gc + gwh / 2], 1)                                             # float64
    ic = rng.uniform(100, 1100, (gi, 2)); iwh = rng.uniform(20, 400, (gi, 2))
    ign = np.concatenate([ic - iwh / 2, ic + iwh / 2], 1)
    if m > 10 and g > 1:
        rois[3, :4] = gts[1].astype(np.float32)          # a near-perfect match
        rois[5, :4] = rois[9, :4]                        # duplicates -> argmax ties
    return rois, gts, ign


def main():
    ref = ref_shim.load()
    out = {}
    for tag, (seed, m, g, gi) in {"small": (1, 50, 3, 2), "one_gt": (2, 333, 1, 1), "big": (3, 4000, 9, 4)}.items():
        rois, gts, ign = scene(seed, m, g, gi)
        ols = ref.core.iou(rois[:, :4], gts)                        # lib/rpn_util.py:448 (numpy branch, float32 x float64)
        ols_ign = ref.core.iou_ign(rois[:, :4], ign)                # :439
        out.update({tag + "_rois": rois, tag + "_gts": gts, tag + "_ign": ign, tag + "_ols": ols, tag + "_ols_ign": ols_ign,
                    tag + "_ols_max": np.amax(ols, axis=1), tag + "_targets": np.argmax(ols, axis=1),
                    tag + "_gt_best_rois": np.argmax(ols, axis=0), tag + "_gt_best_ols": np.amax(ols, axis=0),
                    tag + "_ols_ign_max": np.amax(ols_ign, axis=1)})
        print(tag, ols.dtype, ols.shape, ols_ign.dtype)
    path = os.path.join(ROOT, "tests", "golden", "targets_overlaps.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    # iou / intersect on float64 and mixed float32 / float64 numpy inputs (the dtypes of lib/rpn_util.py:448 and :1295)
    rng = np.random.default_rng(64)

    def boxes(m, dt):
        c = rng.uniform(0, 500, (m, 2)); wh = rng.uniform(5, 200, (m, 2))
        return np.concatenate([c - wh / 2, c + wh / 2], 1).astype(dt)
    g = {"a64": boxes(57, np.float64), "b64": boxes(57, np.float64), "a32": boxes(57, np.float32), "b32": boxes(57, np.float32)}
    g["a64"][3] = g["b64"][5]                                                      # an exact match and a zero-area pair (0 / 0)
    g["a64"][7, 2:] = g["a64"][7, :2]; g["b64"][7] = g["a64"][7]
    for ka, kb in (("a64", "b64"), ("a32", "b64"), ("a64", "b32")):
        for mode in ("combinations", "list"):
            with np.errstate(invalid="ignore", divide="ignore"):
                g["iou_%s_%s_%s" % (ka, kb, mode)] = ref.core.iou(g[ka], g[kb], mode=mode)
                g["inter_%s_%s_%s" % (ka, kb, mode)] = ref.core.intersect(g[ka], g[kb], mode=mode)
    path = os.path.join(ROOT, "tests", "golden", "iou2d_f64.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
