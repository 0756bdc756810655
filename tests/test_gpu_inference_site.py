"""GPU parity: NMS at the reference's inference call site (lib/rpn_util.py:1258-1341)
against the oracle composed from the golden-pinned pieces."""
import numpy as np
import pytest
import torch

from gpu_util import cuda
from test_gpu_loss_branch import _scene

pytestmark = pytest.mark.gpu


def _detections(seed, n):
    scores, _, b7, c2, _, _ = _scene(seed, n_anchor=n, n_fg=min(10, n), n_gt=6)
    rng = np.random.default_rng(seed + 100)
    coords_3d = rng.standard_normal((n, 11)).astype(np.float32)
    cls_pred = rng.integers(1, 4, n).astype(np.float32)
    tracker = np.arange(n, dtype=np.float32)
    return c2, scores, coords_3d, b7, cls_pred, tracker


@pytest.mark.parametrize("overlap", ["2d", "3d", "product"])
@pytest.mark.parametrize("n", [4000, 320])
def test_groomed_nms_at_the_inference_site(overlap, n):
    from groomed_nms_b200 import ops
    from groomed_nms_b200.lib.rpn_util import nms_after_detection
    from oracle import loss_branch_oracle as LO
    c2, scores, c3, raw, cls_pred, trk = _detections(21, n)
    conf = dict(use_nms_in_loss=True, overlap_in_nms=overlap, nms_thres=0.4, nms_topN_pre=3000, diff_nms_temperature=0.1)
    out, keep = nms_after_detection(c2, scores, c3, raw, cls_pred, trk, conf)
    assert isinstance(out, np.ndarray) and out.shape[1] == 5 + 1 + 11 + 1
    order = np.argsort(-scores, kind="stable")[:500]
    corners = ops.corners_from_boxes7(cuda(raw[order])).cpu().numpy()
    want, want_keep = LO.inference_site(c2, scores, c3, raw, cls_pred, trk, True, overlap_in_nms=overlap, temperature=0.1, corners=corners)
    assert keep.tolist() == want_keep.tolist()
    assert np.array_equal(out, want)
    # tensors in -> tensors out, on the device, same rows
    out_t, keep_t = nms_after_detection(cuda(c2), cuda(scores), cuda(c3), cuda(raw), cuda(cls_pred), cuda(trk), conf)
    assert out_t.is_cuda and np.array_equal(out_t.cpu().numpy(), out) and keep_t.cpu().tolist() == keep.tolist()


@pytest.mark.parametrize("n", [5000, 700, 1])
def test_classical_nms_at_the_inference_site(n):
    from groomed_nms_b200.lib.rpn_util import nms_after_detection
    from oracle import loss_branch_oracle as LO
    c2, scores, c3, raw, cls_pred, trk = _detections(22, n)
    conf = dict(use_nms_in_loss=False, nms_thres=0.4, nms_topN_pre=3000)
    out, keep = nms_after_detection(c2, scores, c3, raw, cls_pred, trk, conf)
    want, want_keep = LO.inference_site(c2, scores, c3, raw, cls_pred, trk, False)
    assert keep.tolist() == want_keep.tolist()
    assert np.array_equal(out, want)


def test_no_detections():
    from groomed_nms_b200.lib.rpn_util import nms_after_detection
    z = np.zeros((0, 4), np.float32)
    out, keep = nms_after_detection(z, np.zeros((0,), np.float32), np.zeros((0, 11), np.float32), np.zeros((0, 7), np.float32),
                                    np.zeros((0,), np.float32), np.zeros((0,), np.float32), dict(use_nms_in_loss=True))
    assert out.shape == (0, 18) and keep.shape == (0,)


import detect_golden  # noqa: E402

REF_CASES = detect_golden.cases()


@pytest.mark.parametrize("name", sorted(REF_CASES))
def test_inference_site_matches_reference_run(name):
    """nms_after_detection against what the UNMODIFIED reference's im_detect_3d returned on the same detections
    (tests/golden/inference_site_ref.npz, oracle/gen_golden_detect.py): kept rows and keep indices exact."""
    from groomed_nms_b200.lib.rpn_util import nms_after_detection
    c = REF_CASES[name]
    n = len(c["pre_aboxes"])
    ab = c["pre_aboxes"]
    out, keep = nms_after_detection(ab[:, :4], ab[:, 4], c["pre_coords_3d"][:n], c["pre_coords_3d_raw"][:n], c["pre_cls_pred"][:n].astype(np.float32),
                                    c["pre_tracker"][:n].astype(np.float32), c["conf"])
    assert keep.tolist() == c["keep"].tolist()
    assert np.array_equal(out, c["aboxes_out"].astype(np.float32))


@pytest.mark.parametrize("overlap", ["2d", "product"])
def test_float64_detections_follow_the_reference_dtypes(overlap):
    """float64 detections (what numpy promotion hands the reference at lib/rpn_util.py:1295): float64 score order, 2D IoUs in
    float64 rounded to float32 where differentiable_nms converts its inputs, float64 rows out."""
    from groomed_nms_b200 import ops
    from groomed_nms_b200.lib.rpn_util import nms_after_detection
    from oracle import loss_branch_oracle as LO
    c2, scores, c3, raw, cls_pred, trk = _detections(23, 1500)
    c2 = c2.astype(np.float64) * (1.0 + 1e-9); scores = scores.astype(np.float64)
    scores[40] = scores[41] * (1.0 + 1e-12)                                  # equal in float32, ordered in float64
    conf = dict(use_nms_in_loss=True, overlap_in_nms=overlap, nms_thres=0.4, nms_topN_pre=3000, diff_nms_temperature=0.1)
    out, keep = nms_after_detection(c2, scores, c3, raw, cls_pred, trk, conf)
    order = np.argsort(-scores, kind="stable")[:500]
    corners = ops.corners_from_boxes7(cuda(raw[order])).cpu().numpy()
    want, want_keep = LO.inference_site(c2, scores, c3, raw, cls_pred, trk, True, overlap_in_nms=overlap, temperature=0.1, corners=corners)
    assert out.dtype == np.float64 and keep.tolist() == want_keep.tolist() and np.array_equal(out, want)
