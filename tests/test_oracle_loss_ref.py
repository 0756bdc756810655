"""CPU: the composed loss-branch oracle (oracle/loss_branch_oracle.py) against what the UNMODIFIED reference's
RPN_3D_loss.forward produced on the same inputs (tests/golden/loss_branch_ref.npz, oracle/gen_golden_loss.py).
This is what pins rows a11 / a12 of SURVEY.md section 8: selection and keep indices exact, rescored scores 1e-5."""
import numpy as np
import pytest

import loss_golden
from oracle import groomed_oracle as O
from oracle import loss_branch_oracle as LO

CASES = {c.name: c for c in loss_golden.cases()}


def _branch(case, im):
    c = case.conf()
    n = len(im["fg_inds"])
    kw = dict(overlap_in_nms=c.get("overlap_in_nms", "2d"), nms_thres=c["nms_thres"], temperature=c["diff_nms_temperature"],
              valid_thr=c.get("diff_nms_valid_box_prob_threshold", 0.3), group_size=c["diff_nms_group_size"], beta=c["best_target_box_beta"],
              pruning_method=c["diff_nms_pruning_method"], mask_group_boxes=c["diff_nms_mask_group_boxes"],
              boxes_2d=c["diff_nms_boxes_2d"], p2=case.p2, scale_factor=1.0)
    return LO.branch_image(im["scores_to_nms_fg"], np.arange(n), im["boxes7_fg"], im["coords_2d_fg"], im["gts_val"].astype(np.float32),
                           im["gts_3d"].astype(np.float32), **kw)


@pytest.mark.parametrize("name", sorted(CASES))
def test_branch_oracle_matches_reference_run(name):
    case = CASES[name]
    for im in case.images:
        o = _branch(case, im)
        fg = im["fg_inds"]
        assert fg[o["fg_index_for_nms"]].tolist() == im["fg_index_for_nms"].tolist()          # lib/loss/rpn_3d.py:731-737
        assert np.allclose(o["scores_after_nms"], im["prob"], rtol=1e-5, atol=1e-6)           # :791
        assert o["fwd"]["valid"].tolist() == im["valid"].tolist()
        want_t = np.zeros(len(fg), np.float32)
        want_t[o["best"]] = 1
        assert np.array_equal(want_t, im["targets_fg"])                                       # :801-825
        sa = np.zeros(len(fg), np.float32)
        sa[o["fg_index_for_nms"]] = o["scores_after_nms"]
        assert np.allclose(sa, im["scores_after_fg"], rtol=1e-5, atol=1e-6)                   # :793


@pytest.mark.parametrize("name", [n for n in sorted(CASES) if CASES[n].conf()["after_nms_loss_mode"] == "rank"
                                  and not CASES[n].conf()["rank_boxes_of_all_images_at_once"]])
def test_rank_loss_oracle_matches_reference_run(name):
    """lib/loss/rpn_3d.py:1117-1137: per-image AP loss over the foreground anchors, mean over images, times lambda;
    its gradient wrt the scores that entered NMS goes back through the NMS backward."""
    case = CASES[name]
    lam = case.conf()["after_nms_lambda"]
    tot, cnt = 0.0, 0
    for im in case.images:
        l, _ = O.aploss(im["scores_after_fg"], im["targets_fg"])
        tot += float(l)
        cnt += 1
    assert np.isclose(lam * tot / cnt, case.loss, rtol=1e-5, atol=1e-7)
    for im in case.images:
        o = _branch(case, im)
        _, g_ap = O.aploss(im["scores_after_fg"], im["targets_fg"])
        up = (lam * g_ap / cnt).astype(np.float32)[o["fg_index_for_nms"]]                    # d loss / d prob, NMS order
        gs, _ = O.differentiable_nms_backward(o["fwd"], up, need_grad_iou=False)
        want = np.zeros(len(im["fg_inds"]), np.float32)
        want[o["fg_index_for_nms"]] = gs
        assert np.allclose(want, im["grad_scores_fg"], rtol=1e-4, atol=1e-6 * max(1e-9, np.abs(im["grad_scores_fg"]).max()))
