"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/loss_branch_ref.npz by running the UNMODIFIED reference's
RPN_3D_loss.forward (lib/loss/rpn_3d.py:162-1409) on CPU in the build container (oracle/ref_harness.py, stock arm,
`--fake-cuda`: the method hard-codes `.cuda()`; the shim maps those calls to the CPU, no reference source is edited).

    python oracle/gen_golden_loss.py            # rewrites tests/golden/loss_branch_ref.npz

This pins rows a11 / a12 / (f)1 of SURVEY.md section 8 against the reference's own run: per case and image the fixture
holds what entered the GrooMeD branch (foreground anchors, their scores, 2D boxes, 7-DoF boxes, ground truths) and what
the reference produced (the top-500 selection, rescored scores, keep lists, targets after NMS, the after-NMS loss and
its gradient wrt the network's acceptance logits).  All loss terms except the after-NMS loss are switched off through
the reference's own lambdas so that the recorded gradient is that term's alone; the live stock-vs-installed comparison
on the GPU (tests/test_gpu_reference_dropin.py) runs the full default config.
Only foreground rows are stored (the branch reads nothing else), so the fixture stays small.
"""
import json
import os
import subprocess
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "loss_branch_ref.npz")

ONLY_AFTER_NMS = dict(cls_2d_lambda=0, iou_2d_lambda=0, bbox_2d_lambda=0, bbox_3d_lambda=0, bbox_un_dynamic=False,
                      use_acceptance_prob_in_regression_loss=False, after_nms_lambda=1.0)
# name -> (scene arguments, config overrides on top of scripts/config/groumd_nms.py)
CASES = {
    "2d_rank":       (dict(seed=0, batch=2, feat="24x80"), dict()),                                  # image 0 has > 500 foreground anchors
    "3d_rank":       (dict(seed=1, batch=2, feat="16x56"), dict(overlap_in_nms="3d")),
    "product_rank":  (dict(seed=2, batch=2, feat="16x56"), dict(overlap_in_nms="product")),
    "2d_projected":  (dict(seed=3, batch=2, feat="16x56"), dict(diff_nms_boxes_2d="projected")),
    "2d_classify":   (dict(seed=4, batch=2, feat="16x56"), dict(after_nms_loss_mode="classify")),
    "2d_regress":    (dict(seed=5, batch=2, feat="16x56"), dict(after_nms_loss_mode="regress")),
    "2d_rank_all":   (dict(seed=6, batch=3, feat="16x56"), dict(rank_boxes_of_all_images_at_once=True)),
    "2d_sigmoidal_nomask": (dict(seed=7, batch=2, feat="16x56"), dict(diff_nms_pruning_method="sigmoidal", diff_nms_mask_group_boxes=False,
                                                                       diff_nms_group_size=20)),
}


def run_case(scene_args, overrides):
    ov = dict(ONLY_AFTER_NMS)
    ov.update(overrides)
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "run.npz")
        cmd = [sys.executable, "-m", "oracle.ref_harness", "--arm", "stock", "--fake-cuda", "--out", out, "--seed", str(scene_args["seed"]),
               "--batch", str(scene_args["batch"]), "--feat", scene_args["feat"]]
        for k, v in ov.items():
            cmd += ["--set", "%s=%s" % (k, json.dumps(v))]
        subprocess.run(cmd, check=True, cwd=ROOT, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        d = np.load(out)
        return {k: d[k] for k in d.files}, ov


def compact(name, run, ov, scene_args):
    """Keep the foreground rows and the reference's outputs of one run."""
    g = {}
    p = name + "/"
    g[p + "overrides"] = np.frombuffer(json.dumps(ov).encode(), dtype=np.uint8)
    g[p + "scene"] = np.frombuffer(json.dumps(scene_args).encode(), dtype=np.uint8)
    g[p + "loss"] = run["loss"]
    n = int(run["n_nms"][0])
    g[p + "n_images"] = np.array([n])
    acc_grad = run["grad_acc_logit"]
    for i in range(n):
        q = "%simg%d_" % (p, i)
        img = int(run["nms%d_img_index" % i])
        fg = run["nms%d_fg_inds" % i]
        g[q + "img_index"] = np.array([img])
        g[q + "fg_inds"] = fg
        for k in ("scores_to_nms_fg", "coords_2d_fg", "boxes7_fg", "gts_val", "gts_3d", "fg_index_for_nms", "prob", "valid", "invalid"):
            g[q + k] = run["nms%d_%s" % (i, k)]
        g[q + "iou_in_crc32"] = np.array([zlib.crc32(np.ascontiguousarray(run["nms%d_iou_in" % i]).tobytes())], dtype=np.uint64)
        g[q + "targets_fg"] = run["full_targets_after_nms"][img][fg]
        g[q + "scores_after_fg"] = run["full_scores_after_nms"][img][fg]
        g[q + "grad_acc_logit_fg"] = acc_grad[img][fg]
        g[q + "grad_scores_fg"] = run["grad_acc_prob"][img][fg]
    rest = np.ones(acc_grad.shape, bool)
    for i in range(n):
        rest[int(run["nms%d_img_index" % i]), run["nms%d_fg_inds" % i]] = False
    assert not acc_grad[rest].any(), "the after-NMS loss reaches only foreground anchors"
    return g


def main():
    from groomed_nms_b200 import synthetic
    allg = {}
    for name, (sa, ov) in CASES.items():
        run, full_ov = run_case(sa, ov)
        H, W = [int(x) for x in sa["feat"].split("x")]
        sc = synthetic.c5_scene(seed=sa["seed"], batch=sa["batch"], feat_size=(H, W))
        g = compact(name, run, full_ov, sa)
        g[name + "/p2"] = sc["p2"]
        allg.update(g)
        n = int(run["n_nms"][0])
        print("%-22s loss %.6f  images %d  fg %s  nms %s  targets %s" % (
            name, float(run["loss"][0]), n, [len(run["nms%d_fg_inds" % i]) for i in range(n)],
            [len(run["nms%d_prob" % i]) for i in range(n)], [int((g["%s/img%d_targets_fg" % (name, i)] == 1).sum()) for i in range(n)]))
    np.savez_compressed(OUT, **allg)
    print("wrote %s %.1f KiB" % (OUT, os.path.getsize(OUT) / 1024.0))


if __name__ == "__main__":
    main()
