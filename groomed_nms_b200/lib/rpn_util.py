"""NMS at the inference call site of the reference (lib/rpn_util.py:1258-1341, inside im_detect_3d), device resident.

The reference sorts the detections by score on the host with numpy, truncates to nms_topN_pre, then either
  * (use_nms_in_loss) takes the first 500, builds the 2D / 3D / product overlap matrix with numpy + a GPU round trip,
    runs differentiable_nms ON THE CPU (lib/groomed_nms.py:34-36) and keeps `keep_inds` only (:1292-1320), or
  * calls the Cython/CUDA gpu_nms (:1334),
and finally stacks [x1,y1,x2,y2,score, cls, coords_3d..., tracker] for the kept rows (:1337-1340).

nms_after_detection() does the same on the GPU without leaving it: one stable score sort, the fused GrooMeD-NMS forward
from boxes (no N x N matrix for the 2d / 3d overlaps) or the bitmask hard NMS, one gather of the kept rows.  It is the
replacement a maintainer drops in for :1260-1340; numpy in -> numpy out like the reference."""
import numpy as np
import torch

from .. import _lib, ops
from ._util import device


def _to_dev(x, dev, dtype=torch.float32):
    if torch.is_tensor(x):
        return x.to(dev).to(dtype)
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev).to(dtype)


def nms_after_detection(coords_2d, scores, coords_3d, coords_3d_raw, cls_pred, tracker, rpn_conf, use_differentiable_nms=None,
                        max_boxes=500):
    """coords_2d [A,4], scores [A], coords_3d [A,C], coords_3d_raw [A,>=7] (x,y,z,w,h,l,ry first), cls_pred [A], tracker [A]
    (numpy arrays or tensors); rpn_conf: mapping / EasyDict with nms_thres, nms_topN_pre and the diff_nms_* keys the
    reference reads at lib/rpn_util.py:1056-1063 (same defaults).
    Returns (aboxes [n_keep, 5 + 1 + C + 1] in the reference's column layout and row order, keep_inds into the
    score-sorted, truncated detections) -- numpy if the inputs were numpy, tensors on the inputs' device otherwise."""
    def get(key, default):
        return rpn_conf[key] if key in rpn_conf else default
    as_numpy = not torch.is_tensor(scores)
    dev = scores.device if torch.is_tensor(scores) and scores.is_cuda else device()
    use_diff = bool(get('use_nms_in_loss', False)) if use_differentiable_nms is None else bool(use_differentiable_nms)
    overlap_in_nms = get('overlap_in_nms', "2d")
    # the reference's detections are float64 numpy arrays here: it orders them by their float64 scores and thresholds float64
    # 2D IoUs rounded to float32 (differentiable_nms converts its inputs, lib/groomed_nms.py:34-36).  Same here when given float64.
    f64_in = (coords_2d.dtype == torch.float64) if torch.is_tensor(coords_2d) else (np.asarray(coords_2d).dtype == np.float64)
    sc_key = _to_dev(scores, dev, torch.float64).reshape(-1) if f64_in else None
    c2_64 = _to_dev(coords_2d, dev, torch.float64) if f64_in else None
    c2 = _to_dev(coords_2d, dev); sc = _to_dev(scores, dev).reshape(-1)
    c3 = _to_dev(coords_3d, dev); c3r = _to_dev(coords_3d_raw, dev)
    cls = _to_dev(cls_pred, dev).reshape(-1, 1); trk = _to_dev(tracker, dev).reshape(-1, 1)
    A = sc.shape[0]
    if A == 0:
        out = torch.zeros((0, 5 + 1 + c3.shape[1] + 1), device=dev)
        keep = torch.zeros((0,), dtype=torch.int64, device=dev)
        return (out.cpu().numpy(), keep.cpu().numpy()) if as_numpy else (out, keep)
    # :1260-1266  descending score order (numpy's argsort on -score is not stable; ties keep the lower index here)
    sorted_inds = torch.sort(sc_key if f64_in else sc, descending=True, stable=True)[1]
    n_pre = min(int(get('nms_topN_pre', 3000)), A)                                              # :1286-1290
    sorted_inds = sorted_inds[:n_pre]
    if use_diff:                                                                                # :1293-1320
        n = min(max_boxes, n_pre)
        sel = sorted_inds[:n]
        params = ops.make_params(get('nms_thres', 0.4), get('diff_nms_pruning_method', "linear"), get('diff_nms_temperature', 1),
                                 get('diff_nms_valid_box_prob_threshold', 0.3), False, bool(get('diff_nms_group_boxes', True)),
                                 bool(get('diff_nms_mask_group_boxes', True)), get('diff_nms_group_size', 100))
        s_in = sc[sel].contiguous()[None]
        box2d = c2[sel].contiguous()
        iou2d_64 = None
        if f64_in and overlap_in_nms in ("2d", "product"):                                      # :1295 in float64
            b64 = c2_64[sel].contiguous()
            iou2d_64 = ops.overlap2d_f64(b64, b64)
        if overlap_in_nms == "2d" and iou2d_64 is not None:
            st = ops.forward_matrix(s_in, iou2d_64.float()[None], params)
        elif overlap_in_nms == "2d":                                                            # :1297-1298
            st = ops.forward_boxes(s_in, box2d[None], _lib.BOX_2D, params)
        else:
            rec = ops.box3d_records(ops.corners_from_boxes7(c3r[sel][:, :7].contiguous()), mutate_input=False)   # :1303-1311
            if overlap_in_nms == "3d":                                                          # :1314-1315
                st = ops.forward_boxes(s_in, rec[None], _lib.BOX_3D_REC, params, generalized=True, affine=True)
            elif iou2d_64 is not None:                                                          # product :1316-1317, float64 x float32
                ov3 = ops.overlap3d(rec, rec, False, True, generalized=True, affine=True)[1]
                st = ops.forward_matrix(s_in, (iou2d_64 * ov3.double()).float()[None], params)
            else:
                iou2d = ops.overlap2d(box2d, box2d)
                ov = ops.overlap3d(rec, rec, False, True, generalized=True, affine=True, mul2d=iou2d)[1]
                st = ops.forward_matrix(s_in, ov[None], params)
        keep = st.valid_idx[0, :int(st.counts[0, 0])]                                          # the one host sync: the count
        num_boxes = n
    else:                                                                                       # :1334
        dets = torch.cat([c2[sorted_inds], sc[sorted_inds, None]], dim=1).contiguous()
        k, nk = ops.hard_nms(dets, float(get('nms_thres', 0.4)), shift=1.0, cmp=_lib.CMP_GT)
        keep = k[:int(nk.item())].long()
        num_boxes = n_pre
    rows = sorted_inds[:num_boxes][keep]                                                        # :1337-1340
    if f64_in:                                                                                  # float64 detections stay float64 (np.hstack promotes)
        d = torch.float64
        out = torch.cat([c2_64[rows], sc_key[rows, None], _to_dev(cls_pred, dev, d).reshape(-1, 1)[rows], _to_dev(coords_3d, dev, d)[rows],
                         _to_dev(tracker, dev, d).reshape(-1, 1)[rows]], dim=1)
    else:
        out = torch.cat([c2[rows], sc[rows, None], cls[rows], c3[rows], trk[rows]], dim=1)
    return (out.cpu().numpy(), keep.cpu().numpy()) if as_numpy else (out, keep)


def targets_overlaps(rois, gts_val, gts_ign):
    """The overlap part of the reference's compute_targets (lib/rpn_util.py:439-461) on the GPU, in float64 with the call
    site's numpy promotion rules (float32 rois keep a float32 area_a).  rois [M,>=4], gts_val [G,4], gts_ign [Gi,4]
    (numpy or tensors; G, Gi <= 64, either may be empty).
    -> dict with what the reference derives right there: ols [M,G], ols_max = amax(ols, 1), targets = argmax(ols, 1),
       gt_best_rois = argmax(ols, 0), gt_best_ols = amax(ols, 0) (absent when there is no valid ground truth) and
       ols_ign_max = amax(iou_ign(rois, gts_ign), 1) (zeros when there is no ignore region, :443).
    numpy in -> numpy out."""
    as_numpy = not torch.is_tensor(rois)
    dev = rois.device if torch.is_tensor(rois) and rois.is_cuda else device()
    area_f32 = int((rois.dtype == torch.float32) if torch.is_tensor(rois) else (np.asarray(rois).dtype == np.float32))
    r = _to_dev(rois, dev, torch.float64).contiguous()
    M = r.shape[0]
    lib = _lib.load()

    def run(gts, kind, want_ols):
        g = _to_dev(gts, dev, torch.float64).reshape(-1, 4).contiguous()
        G = g.shape[0]
        ols = torch.empty((M, G), dtype=torch.float64, device=dev) if want_ols else None
        rmax = torch.empty((M,), dtype=torch.float64, device=dev); rarg = torch.empty((M,), dtype=torch.int64, device=dev)
        cmax = torch.empty((G,), dtype=torch.float64, device=dev); carg = torch.empty((G,), dtype=torch.int64, device=dev)
        if M and G:
            ws = torch.empty((int(lib.gnms_targets_overlaps_workspace_bytes(M, G)),), dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.gnms_targets_overlaps_f64(ops._p(r), r.stride(0), M, ops._p(g), G, kind, area_f32, ops._p(ols), ops._p(rmax),
                                                         ops._p(rarg), ops._p(cmax), ops._p(carg), ops._p(ws), ops._stream(dev)),
                           "gnms_targets_overlaps_f64")
        return ols, rmax, rarg, cmax, carg

    out = {}
    if (gts_val.shape[0] if hasattr(gts_val, "shape") else len(gts_val)) > 0:
        ols, rmax, rarg, cmax, carg = run(gts_val, 0, True)
        out.update(ols=ols, ols_max=rmax, targets=rarg, gt_best_rois=carg, gt_best_ols=cmax)
    if (gts_ign.shape[0] if hasattr(gts_ign, "shape") else len(gts_ign)) > 0:
        out["ols_ign_max"] = run(gts_ign, 1, False)[1]
    else:
        out["ols_ign_max"] = torch.zeros((M,), dtype=torch.float32, device=dev)
    if as_numpy:
        out = {k: v.cpu().numpy() for k, v in out.items()}
    return out
