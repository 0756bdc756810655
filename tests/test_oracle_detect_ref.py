"""CPU: the composed inference-site oracle (oracle/loss_branch_oracle.inference_site) against what the UNMODIFIED
reference's im_detect_3d returned on the same detections (tests/golden/inference_site_ref.npz) -- pins SURVEY.md
section 8(f) rank 2: kept rows and keep indices exact."""
import numpy as np
import pytest

import detect_golden
from oracle import loss_branch_oracle as LO

CASES = detect_golden.cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_inference_site_oracle_matches_reference_run(name):
    c = CASES[name]
    conf = c["conf"]
    n = len(c["pre_aboxes"])
    ab = c["pre_aboxes"]
    rows, keep = LO.inference_site(ab[:, :4], ab[:, 4], c["pre_coords_3d"][:n], c["pre_coords_3d_raw"][:n], c["pre_cls_pred"][:n].astype(np.float32),
                                   c["pre_tracker"][:n].astype(np.float32), bool(conf["use_nms_in_loss"]),
                                   overlap_in_nms=conf.get("overlap_in_nms", "2d"), nms_thres=conf["nms_thres"], topn_pre=conf["nms_topN_pre"],
                                   temperature=conf["diff_nms_temperature"], group_size=conf["diff_nms_group_size"],
                                   mask_group_boxes=conf["diff_nms_mask_group_boxes"])
    assert keep.tolist() == c["keep"].tolist()
    assert np.array_equal(rows, c["aboxes_out"].astype(np.float32))
