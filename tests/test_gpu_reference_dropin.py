"""The drop-in claim, checked live on the GPU: the UNMODIFIED reference's RPN_3D_loss.forward + backward
(baseline/_ref/lib/loss/rpn_3d.py, staged by tools/stage_reference.py) is run twice on the same synthetic C5 scene --
stock, and with `groomed_nms_b200.install()` active (every `lib.groomed_nms` / `lib.nms` / `lib.nms_others` /
`lib.loss.aploss` import and the overlap functions of `lib.core` resolve to the sm_100a kernels) -- one process per
arm (oracle/ref_harness.py), and everything the GrooMeD branch consumes and produces is compared.

Bars: indices (top-500 selection, keep lists, targets after NMS) exact; overlaps from identical boxes bit-exact;
rescored scores, losses and gradients 1e-5 relative (BASELINE.json north_star)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from gpu_util import bits_equal

pytestmark = pytest.mark.gpu

STAGED = os.path.join(ROOT, "baseline", "_ref", "lib", "loss", "rpn_3d.py")


def _run(arm, out, extra):
    cmd = [sys.executable, "-m", "oracle.ref_harness", "--arm", arm, "--out", out] + extra
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, "ref_harness %s failed:\n%s" % (arm, r.stdout[-4000:])
    d = np.load(out)
    return {k: d[k] for k in d.files}


CONFIGS = {
    # gradient reaches the raw 3D box parameters through get_corners_of_cuboid and iou3d_approximate (the acceptance-probability
    # target is not detached, lib/loss/rpn_3d.py:663-679 with :1060): install() must not drop it
    "accprob_regress_all": ["--seed", "8", "--batch", "2", "--feat", "16x56", "--set", "acceptance_prob_lambda=1.0", "--set",
                            "acceptance_prob_mode=\"regress\"", "--set", "boxes_for_acceptance_prob=\"all\""],
    "default_2d": ["--seed", "0", "--batch", "2", "--feat", "24x80"],
    "product": ["--seed", "2", "--batch", "2", "--feat", "16x56", "--set", "overlap_in_nms=\"product\""],
    "3d_nomask_sigmoidal": ["--seed", "7", "--batch", "2", "--feat", "16x56", "--set", "overlap_in_nms=\"3d\"", "--set",
                            "diff_nms_mask_group_boxes=false", "--set", "diff_nms_pruning_method=\"sigmoidal\"", "--set", "diff_nms_group_size=20"],
    "classify_projected": ["--seed", "4", "--batch", "2", "--feat", "16x56", "--set", "after_nms_loss_mode=\"classify\"", "--set",
                           "diff_nms_boxes_2d=\"projected\""],
}


@pytest.mark.skipif(not os.path.isfile(STAGED), reason="reference not staged (python tools/stage_reference.py in the build container)")
@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_reference_loss_stock_vs_installed(name, tmp_path):
    extra = CONFIGS[name]
    stock = _run("stock", str(tmp_path / "stock.npz"), extra)
    ours = _run("installed", str(tmp_path / "installed.npz"), extra)
    n = int(stock["n_nms"][0])
    assert n == int(ours["n_nms"][0]) and n > 0
    two_d = "2d" in name
    for i in range(n):
        k = "nms%d_" % i
        assert np.array_equal(stock[k + "fg_index_for_nms"], ours[k + "fg_index_for_nms"])          # lib/loss/rpn_3d.py:731-737
        assert bits_equal(stock[k + "scores_in"], ours[k + "scores_in"])
        if two_d:
            assert bits_equal(stock[k + "iou_in"], ours[k + "iou_in"])                                # same boxes -> same bits
        else:       # corners come from cos/sin + bmm (torch) vs one kernel: ulps apart, so are the overlaps
            assert np.allclose(stock[k + "iou_in"], ours[k + "iou_in"], rtol=1e-4, atol=1e-5)
        assert np.array_equal(stock[k + "valid"], ours[k + "valid"])
        assert sorted(stock[k + "invalid"].tolist()) == sorted(ours[k + "invalid"].tolist())
        assert np.allclose(stock[k + "prob"], ours[k + "prob"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(stock["full_targets_after_nms"], ours["full_targets_after_nms"])            # :801-825
    assert np.allclose(stock["full_scores_after_nms"], ours["full_scores_after_nms"], rtol=1e-5, atol=1e-6)
    assert np.isclose(stock["loss"][0], ours["loss"][0], rtol=1e-5)
    for key in stock:
        if key.startswith("stat_"):
            assert np.allclose(stock[key], ours[key], rtol=1e-4, atol=1e-6), key
    for key in ("grad_acc_logit", "grad_acc_prob", "grad_cls", "grad_bbox_2d", "grad_bbox_3d"):
        a, b = stock[key], ours[key]
        assert np.allclose(a, b, rtol=1e-4, atol=1e-6 * max(1e-12, np.abs(a).max())), key
    assert np.abs(stock["grad_acc_logit"]).max() > 0


DETECT_CONFIGS = {
    "groomed_2d": [],
    "groomed_product": ["--set", "overlap_in_nms=\"product\""],
    "classical": ["--set", "use_nms_in_loss=false"],
}


@pytest.mark.skipif(not os.path.isfile(STAGED), reason="reference not staged (python tools/stage_reference.py in the build container)")
@pytest.mark.parametrize("name", sorted(DETECT_CONFIGS))
def test_reference_im_detect_3d_stock_vs_installed(name, tmp_path):
    """The inference call site (lib/rpn_util.py:1258-1341 inside im_detect_3d): the reference's own function, stock (its numpy /
    CPU-torch NMS; py_cpu_nms standing in for the unbuilt Cython gpu_nms) and with install() active (iou, corners,
    iou3d_approximate, differentiable_nms and gpu_nms on the sm_100a kernels): the same detections survive, in the same order."""
    extra = ["--detect", "--seed", "31", "--batch", "1", "--feat", "24x80"] + DETECT_CONFIGS[name]
    stock = _run("stock", str(tmp_path / "stock.npz"), extra)
    ours = _run("installed", str(tmp_path / "installed.npz"), extra)
    assert np.array_equal(stock["pre_aboxes"], ours["pre_aboxes"])
    assert np.array_equal(stock["keep"], ours["keep"]) and len(stock["keep"]) > 0
    assert np.array_equal(stock["aboxes_out"], ours["aboxes_out"])
