"""GPU parity: the overlap part of compute_targets (lib/rpn_util.py:439-461) against golden
vectors produced by the UNMODIFIED reference (oracle/gen_golden_targets.py) and against the oracle on larger inputs."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
KEYS = ("ols", "ols_max", "targets", "gt_best_rois", "gt_best_ols", "ols_ign_max")


@pytest.mark.parametrize("tag", ["small", "one_gt", "big"])
def test_targets_overlaps_match_the_reference_bit_for_bit(tag):
    from groomed_nms_b200.lib.rpn_util import targets_overlaps
    g = load_golden("targets_overlaps")
    out = targets_overlaps(g[tag + "_rois"], g[tag + "_gts"], g[tag + "_ign"])
    for k in KEYS:
        assert np.array_equal(out[k], g[tag + "_" + k]), k


def test_targets_overlaps_full_anchor_grid_vs_oracle():
    """69 120 anchors (24 x 80 x 36, the reference's feature map) x 12 ground truths, ties, an anchor equal to a GT,
    float64 rois as well (then area_a is float64 too), no ignore regions."""
    from groomed_nms_b200.lib.rpn_util import targets_overlaps
    from oracle import groomed_oracle as O
    rng = np.random.default_rng(9)
    M, G = 69120, 12
    c = rng.uniform(0, 1280, (M, 2)) * [1.0, 0.3]; wh = rng.uniform(16, 400, (M, 2))
    rois = np.concatenate([c - wh / 2, c + wh / 2, rng.uniform(0, 1, (M, 1))], 1).astype(np.float32)
    gc = rng.uniform(100, 1100, (G, 2)) * [1.0, 0.3]; gwh = rng.uniform(30, 300, (G, 2))
    gts = np.concatenate([gc - gwh / 2, gc + gwh / 2], 1)
    rois[1000, :4] = gts[4].astype(np.float32)
    rois[2000:2004, :4] = rois[1000, :4]                              # equal maxima: the first one wins
    for r in (rois, rois.astype(np.float64)):
        out = targets_overlaps(r, gts, np.zeros((0, 4)))
        want = O.targets_overlaps(r, gts, np.zeros((0, 4)))
        for k in KEYS:
            assert np.array_equal(out[k], want[k]), k
        assert out["gt_best_rois"][4] == 1000
    out = targets_overlaps(rois, np.zeros((0, 4)), gts[:3])           # only ignore regions
    want = O.targets_overlaps(rois, np.zeros((0, 4)), gts[:3])
    assert set(out) == {"ols_ign_max"} and np.array_equal(out["ols_ign_max"], want["ols_ign_max"])
