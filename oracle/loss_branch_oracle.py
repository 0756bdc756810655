"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the GrooMeD branch of the detection loss
(reference lib/loss/rpn_3d.py:731-825 and :1117-1137, "rank" mode).

Pinned against the reference's own run: tests/golden/loss_branch_ref.npz holds what the UNMODIFIED RPN_3D_loss.forward
consumed and produced on 8 synthetic cases (oracle/gen_golden_loss.py runs it in the build container with its `.cuda()`
calls mapped to the CPU by oracle/ref_shim.fake_cuda); tests/test_oracle_loss_ref.py checks this restatement against
it (selection / keep / target indices exact, scores 1e-5, loss and score gradients).  Every step below is a composition
of functions pinned by their own golden vectors (oracle.groomed_oracle: corners, iou, iou3d_approximate,
differentiable_nms, aploss), in the order of the reference lines cited."""
import numpy as np

from . import groomed_oracle as O

F32 = np.float32


def branch_image(scores, fg_inds, boxes7, coords_2d, gts_2d, gts_3d, overlap_in_nms="2d", nms_thres=0.4, temperature=0.1,
                 valid_thr=0.3, group_size=100, beta=0.3, max_boxes=500, corners_b1=None, pruning_method="linear",
                 group_boxes=True, mask_group_boxes=True, boxes_2d="normal", p2=None, scale_factor=1.0):
    """-> dict(fg_index_for_nms, scores_after_nms (in that order), best (anchor ids), fwd (oracle nms state)).
    corners_b1: optionally the CUDA path's corners (parity is defined from the corners onward)."""
    fg_scores = scores[fg_inds]
    order = O.stable_sort_desc(fg_scores)                                              # :731
    n = min(max_boxes, order.shape[0])
    fg_idx = fg_inds[order[:n]]                                                        # :737
    b7 = boxes7[fg_idx].astype(F32)
    if corners_b1 is None:
        corners_b1 = O.get_corners_of_cuboid(*[b7[:, i] for i in range(7)])            # :746-752
    corners_b1 = corners_b1.astype(F32).copy()
    box2d = coords_2d[fg_idx].astype(F32)
    if boxes_2d == "projected":                                                        # :756-768,774
        nms_box2d = O.projected_boxes_from_corners(p2, corners_b1, scale_factor)
    else:
        nms_box2d = box2d
    iou2d = O.iou(nms_box2d, nms_box2d)                                                # :772-774
    if overlap_in_nms == "2d":
        ov = iou2d
    else:
        _, g3 = O.iou3d_approximate(corners_b1, corners_b1, "combinations", "generalized")   # :780
        corners_b1 = O.mutated_corners(corners_b1)                                     # the reference's in-place Y<-Z
        ov3 = (F32(0.5) * (F32(1) + g3)).astype(F32)                                   # :781
        ov = ov3 if overlap_in_nms == "3d" else (iou2d * ov3).astype(F32)              # :783,786
    fwd = O.differentiable_nms(scores[fg_idx], ov, nms_threshold=nms_thres, temperature=temperature, pruning_method=pruning_method,
                               valid_box_prob_threshold=valid_thr, group_size=group_size, group_boxes=group_boxes,
                               mask_group_boxes=mask_group_boxes, dense=not (group_boxes and mask_group_boxes))   # :791
    gt7 = np.stack([gts_3d[:, 7], gts_3d[:, 8], gts_3d[:, 9], gts_3d[:, 3], gts_3d[:, 4], gts_3d[:, 5], gts_3d[:, 10]], 1).astype(F32)
    corners_b2 = O.get_corners_of_cuboid(*[gt7[:, i] for i in range(7)])               # :804-811
    _, g3gt = O.iou3d_approximate(corners_b1, corners_b2, "combinations", "generalized")   # :813
    iou2gt = O.iou(box2d, gts_2d[:, :4].astype(F32))                                   # :814
    score_gt = (iou2gt * (F32(0.5) * (F32(1) + g3gt)).astype(F32)).astype(F32)         # :817 (commuted product, see test)
    mx = score_gt.argmax(axis=0)                                                       # :818
    keep = score_gt[mx, np.arange(score_gt.shape[1])] > F32(beta)                      # :820
    return dict(fg_index_for_nms=fg_idx, scores_after_nms=fwd["prob"], best=fg_idx[mx[keep]], fwd=fwd, score_gt=score_gt)


def after_nms_rank_loss(scores_after, targets_after, weights, lam=1.0):
    """:1117-1137 rank mode, per image mean -> (loss, grad wrt scores_after [B,A])."""
    B = scores_after.shape[0]
    grad = np.zeros_like(scores_after, dtype=F32)
    tot, cnt = F32(0), 0
    for b in range(B):
        act = weights[b] > 0
        if act.sum() > 0:
            cnt += 1
            l, g = O.aploss(scores_after[b, act], targets_after[b, act])
            tot += l
            grad[b, act] = g
    if cnt:
        tot, grad = tot / cnt, grad / cnt
    return F32(lam) * tot, (F32(lam) * grad).astype(F32)


def inference_site(coords_2d, scores, coords_3d, coords_3d_raw, cls_pred, tracker, use_diff, overlap_in_nms="2d", nms_thres=0.4,
                   topn_pre=3000, temperature=1.0, valid_thr=0.3, group_size=100, max_boxes=500, corners=None, mask_group_boxes=True):
    """lib/rpn_util.py:1258-1341 composed from the pinned pieces -> (aboxes rows as the reference stacks them, keep_inds).
    Pinned against the reference's own im_detect_3d by tests/golden/inference_site_ref.npz (tests/test_oracle_detect_ref.py).
    corners: optionally the CUDA path's corners of the first 500 sorted boxes (parity is defined from the corners onward)."""
    f64 = coords_2d.dtype == np.float64              # float64 detections: float64 score order and float64 2D IoUs, rounded to float32
    #                                                  only where differentiable_nms converts its inputs (lib/groomed_nms.py:34-36)
    order = (np.argsort(-scores, kind="stable") if f64 else O.stable_sort_desc(scores.astype(F32)))[:min(topn_pre, len(scores))]   # :1260-1290
    if use_diff:                                                                         # :1293-1320
        sel = order[:max_boxes]
        box2d = coords_2d[sel].astype(F32)
        iou2d = O.overlap2d_f64(coords_2d[sel], coords_2d[sel]) if f64 else O.iou(box2d, box2d)
        if overlap_in_nms == "2d":
            ov = iou2d.astype(F32)
        else:
            if corners is None:
                r = coords_3d_raw[sel].astype(F32)
                corners = O.get_corners_of_cuboid(*[r[:, i] for i in range(7)])
            _, g3 = O.iou3d_approximate(corners.astype(F32).copy(), corners.astype(F32).copy(), "combinations", "generalized")
            ov3 = (F32(0.5) * (F32(1) + g3)).astype(F32)
            ov = ov3 if overlap_in_nms == "3d" else (iou2d * ov3).astype(F32)
        fwd = O.differentiable_nms(scores[sel].astype(F32), ov, nms_threshold=nms_thres, temperature=temperature,
                                   valid_box_prob_threshold=valid_thr, group_size=group_size, mask_group_boxes=mask_group_boxes,
                                   dense=not mask_group_boxes)
        keep = fwd["valid"]
        base = sel
    else:                                                                                # :1334
        dets = np.concatenate([coords_2d[order], scores[order, None]], 1).astype(F32)
        keep = np.asarray(O.hard_nms(dets, nms_thres, shift=1.0), dtype=np.int64)
        base = order
    rows = base[keep]
    out = np.concatenate([coords_2d[rows], scores[rows, None], cls_pred[rows, None], coords_3d[rows], tracker[rows, None]], 1)
    return out.astype(np.float64 if f64 else F32), keep
