"""GrooMeD-NMS branch of the detection loss (reference lib/loss/rpn_3d.py:71-96, 721-883, 1091-1148) on the sm_100a
kernels.

The reference's RPN_3D_loss.forward is a 1400-line method of which only the GrooMeD branch is on the hot path
(SURVEY.md section 8 rows a11/a12); the rest (target assignment on numpy, cls/bbox/iou losses) stays the reference's.
This module exposes that branch as two calls a maintainer drops into the reference method in place of the inlined
code (INTEGRATION.md shows the patch):

    branch = GroomedNMSLossBranch(conf)                                    # same conf keys / defaults as :71-96
    fg_idx, scores_after, best = branch.image(scores_to_nms_img, fg_inds_tensor, boxes7_raw, coords_2d_512_img,
                                              p2, scale_factor, gts_2d, gts_3d)        # replaces :731-825
    loss = branch.after_nms_loss(scores_after_nms, targets_after_nms, bbox_weights)    # replaces :1091-1137

or, for the whole batch in one launch sequence and without the reference's three host syncs per image
(SURVEY.md section 8(f) rank 1):

    scores_after_nms, targets_after_nms = branch.batch(scores_to_nms, fg_mask, boxes7_raw, coords_2d_512, p2s,
                                                       scale_factors, gts_2d_list, gts_3d_list)   # replaces the loop body
"""
import numpy as np
import torch
import torch.nn.functional as F

from ... import _lib, ops
from .aploss import APLoss


class GroomedNMSLossBranch(object):
    def __init__(self, conf):
        def get(key, default):
            return conf[key] if key in conf else default
        # reference lib/loss/rpn_3d.py:71-96 (same keys, same defaults)
        self.use_nms_in_loss = get('use_nms_in_loss', False)
        self.diff_nms_pruning_method = get('diff_nms_pruning_method', "linear")
        self.diff_nms_temperature = get('diff_nms_temperature', 1)
        self.diff_nms_valid_box_prob_threshold = get('diff_nms_valid_box_prob_threshold', 0.3)
        self.diff_nms_boxes_2d = get('diff_nms_boxes_2d', "normal")
        self.diff_nms_group_boxes = get('diff_nms_group_boxes', True)
        self.diff_nms_mask_group_boxes = get('diff_nms_mask_group_boxes', True)
        self.diff_nms_group_size = get('diff_nms_group_size', 100)
        self.after_nms_lambda = get('after_nms_lambda', 1)
        self.after_nms_loss_mode = get('after_nms_loss_mode', "rank")
        self.overlap_in_nms = get('overlap_in_nms', "2d")
        self.rank_boxes_of_all_images_at_once = get('rank_boxes_of_all_images_at_once', False)
        self.best_target_box_beta = get('best_target_box_beta', 0.3)
        self.nms_thres = get('nms_thres', 0.4)
        self.max_boxes = 500                                           # :732 "at max 500 boxes in the NMS"
        self.apLoss = APLoss()

    def params(self):
        return ops.make_params(self.nms_thres, self.diff_nms_pruning_method, self.diff_nms_temperature,
                               self.diff_nms_valid_box_prob_threshold, False, bool(self.diff_nms_group_boxes),
                               bool(self.diff_nms_mask_group_boxes), self.diff_nms_group_size)

    # ---------------------------------------------------------------------------------------------- :731-825
    def image(self, scores_to_nms_img, fg_inds_tensor, boxes7_raw, coords_2d_512_img, p2, scale_factor, gts_2d, gts_3d):
        """One image of the batch.

        scores_to_nms_img [A]   per-anchor score that goes into NMS (requires grad)            (:722-727)
        fg_inds_tensor   [F]    int64 indices of the foreground anchors
        boxes7_raw       [A,7]  (x3d,y3d,z3d,w3d,h3d,l3d,ry3d) raw 3D boxes of all anchors      (:746-752)
        coords_2d_512_img[A,4]  2D boxes x1,y1,x2,y2
        p2 [4,4], scale_factor  projection matrix and image scale (only for diff_nms_boxes_2d == "projected")
        gts_2d [G,4], gts_3d [G,16]  ground truths as the reference lays them out (:804-811)
        Returns (fg_index_for_nms [n<=500] int64, scores_after_nms_img [n] (differentiable wrt the scores),
                 max_indices_nms_img int64 -- the anchors that become positive targets after NMS)."""
        dev = scores_to_nms_img.device
        fg_scores = scores_to_nms_img[fg_inds_tensor]
        _, sorted_index = torch.sort(fg_scores, descending=True)                               # :731
        n = min(self.max_boxes, int(sorted_index.shape[0]))                                     # :732
        fg_index_for_nms = fg_inds_tensor[sorted_index[:n]]                                     # :737
        b7 = boxes7_raw[fg_index_for_nms].detach().float().contiguous()
        corners_b1 = ops.corners_from_boxes7(b7)                                                # :746-752
        box2d = coords_2d_512_img[fg_index_for_nms].detach().float().contiguous()
        if self.diff_nms_boxes_2d == "projected":                                               # :754-768,774
            pts = corners_b1.transpose(1, 2).reshape((-1, 3)).transpose(0, 1).contiguous()
            pr = ops.project_points(p2.to(dev), pts, True)
            pr = pr.transpose(0, 1).reshape((-1, 8, 4)).transpose(1, 2)
            nms_box2d = (torch.stack([pr[:, 0].min(dim=1)[0], pr[:, 1].min(dim=1)[0], pr[:, 0].max(dim=1)[0],
                                      pr[:, 1].max(dim=1)[0]], dim=1) * scale_factor).contiguous()
        else:
            nms_box2d = box2d                                                                   # :772
        scores_in = scores_to_nms_img[fg_index_for_nms]
        params = self.params()
        if self.overlap_in_nms == "2d":                                                         # :776-777
            # fused path: overlaps evaluated on the fly by the tile kernel, nothing n^2 in HBM
            prob = ops.GroomedNMSBoxesFunction.apply(scores_in, nms_box2d, _lib.BOX_2D, params, False, False)[0]
            rec_b1 = ops.box3d_records(corners_b1, mutate_input=False)
        else:
            # the reference's first iou3d_approximate call overwrites Y with Z in corners_3d_b1 (:780, lib/core.py:379-380)
            rec_b1 = ops.box3d_records(corners_b1, mutate_input=True)
            if self.overlap_in_nms == "3d":                                                     # :782-783
                prob = ops.GroomedNMSBoxesFunction.apply(scores_in, rec_b1, _lib.BOX_3D_REC, params, True, True)[0]
            else:                                                                               # "product" :784-786
                iou2d = ops.overlap2d(nms_box2d, nms_box2d)
                _, ov = ops.overlap3d(rec_b1, rec_b1, False, True, generalized=True, affine=True, mul2d=iou2d)
                prob = ops.GroomedNMSFunction.apply(scores_in, ov, params)[0]
            # ... and the GT matching below then sees the mutated corners (:813)
            rec_b1 = ops.box3d_records(corners_b1, mutate_input=False)
        # ---- best box per ground truth (:801-825)
        g3 = gts_3d.to(dev).float()
        g2 = gts_2d.to(dev).float().contiguous()
        gt7 = torch.stack([g3[:, 7], g3[:, 8], g3[:, 9], g3[:, 3], g3[:, 4], g3[:, 5], g3[:, 10]], dim=1).contiguous()
        rec_b2 = ops.box3d_records(ops.corners_from_boxes7(gt7), mutate_input=False)
        iou2d_gt = ops.overlap2d(box2d, g2[:, :4].contiguous())                                 # :814
        _, score_gt = ops.overlap3d(rec_b1, rec_b2, False, True, generalized=True, affine=True, mul2d=iou2d_gt)  # :813,817
        best, max_idx = torch.max(score_gt, dim=0)                                              # :818
        max_idx = max_idx[best > self.best_target_box_beta].flatten()                           # :820-822
        return fg_index_for_nms, prob, fg_index_for_nms[max_idx]

    # ---------------------------------------------------------------------------------------------- :721-828, whole batch
    def batch(self, scores_to_nms, fg_mask, boxes7_raw, coords_2d_512, p2s, scale_factors, gts_2d_list, gts_3d_list):
        """All images of the batch at once (ragged: every image has its own number of foreground anchors and of GTs).

        scores_to_nms [B,A] (requires grad)   fg_mask [B,A] bool   boxes7_raw [B,A,7]   coords_2d_512 [B,A,4]
        p2s [B,4,4] and scale_factors [B] (only read when diff_nms_boxes_2d == "projected")
        gts_2d_list / gts_3d_list: per image [G_b,4] / [G_b,16] tensors or arrays (G_b may be 0)
        Returns (scores_after_nms [B,A] -- zero outside the boxes that went through NMS, differentiable wrt the scores --
                 targets_after_nms [B,A] float, 1 at the best box of every ground truth that clears best_target_box_beta).
        Same arithmetic as image() per image; the top-500 selection, corners, records, NMS forward and backward are single
        launches over the batch (n_per_image carries the ragged counts on the device: no .cpu() / .item() anywhere); the
        top-500 selection is one radix-select launch (gnms_masked_topk_f32) and the best-box assignment of all ground truths
        of the batch one launch (gnms_best_box_per_gt_f32)."""
        dev = scores_to_nms.device
        B, A = scores_to_nms.shape
        K = min(self.max_boxes, A)
        fg_mask = fg_mask.to(dev).bool()
        # :731-737  foreground anchors by descending score, at most 500 (stable: ties keep the lower anchor index): one
        # radix-select launch over the batch instead of a full sort of all A scores per image
        top, n_img = ops.masked_topk(scores_to_nms.detach(), fg_mask, K)                        # [B,K] anchor ids, [B] live counts
        live = torch.arange(K, device=dev)[None, :] < n_img[:, None]
        b7 = torch.gather(boxes7_raw.detach().float(), 1, top[:, :, None].expand(B, K, 7)).contiguous()
        box2d = torch.gather(coords_2d_512.detach().float(), 1, top[:, :, None].expand(B, K, 4)).contiguous()
        corners = ops.corners_from_boxes7(b7.view(B * K, 7))                                     # :746-752
        if self.diff_nms_boxes_2d == "projected":                                               # :754-768,774
            nms_box2d = torch.empty_like(box2d)
            for b in range(B):
                pts = corners[b * K:(b + 1) * K].transpose(1, 2).reshape((-1, 3)).transpose(0, 1).contiguous()
                pr = ops.project_points(p2s[b].to(dev), pts, True).transpose(0, 1).reshape((-1, 8, 4)).transpose(1, 2)
                nms_box2d[b] = torch.stack([pr[:, 0].min(dim=1)[0], pr[:, 1].min(dim=1)[0], pr[:, 0].max(dim=1)[0],
                                            pr[:, 1].max(dim=1)[0]], dim=1) * float(scale_factors[b])
        else:
            nms_box2d = box2d                                                                   # :772
        scores_in = torch.gather(scores_to_nms, 1, top)                                         # keeps the graph
        params = self.params()
        if self.overlap_in_nms == "2d":                                                         # :776-777
            prob = ops.GroomedNMSBatchFunction.apply(scores_in, nms_box2d, _lib.BOX_2D, params, False, False, n_img)[0]
            rec = ops.box3d_records(corners, mutate_input=False)
        else:
            rec = ops.box3d_records(corners, mutate_input=True)                                 # :780 Y <- Z quirk
            if self.overlap_in_nms == "3d":                                                     # :782-783
                prob = ops.GroomedNMSBatchFunction.apply(scores_in, rec.view(B, K, 8), _lib.BOX_3D_REC, params, True, True, n_img)[0]
            else:                                                                               # "product" :784-786
                ov = torch.empty((B, K, K), dtype=torch.float32, device=dev)
                for b in range(B):
                    iou2d = ops.overlap2d(nms_box2d[b], nms_box2d[b])
                    ov[b] = ops.overlap3d(rec[b * K:(b + 1) * K], rec[b * K:(b + 1) * K], False, True, generalized=True, affine=True, mul2d=iou2d)[1]
                prob = ops.GroomedNMSBatchFunction.apply(scores_in, ov, None, params, False, False, n_img)[0]
            rec = ops.box3d_records(corners, mutate_input=False)                                # GT matching sees the mutated corners (:813)
        # :793 (dead slots all point at anchor 0 and add 0)
        scores_after_nms = torch.zeros((B, A), dtype=prob.dtype, device=dev).scatter_add(1, top, prob * live)
        # ---- best box per ground truth (:801-825): every ground truth of the batch in one launch
        targets_after_nms = torch.zeros((B, A), dtype=torch.float32, device=dev)
        g3_all, g2_all, owner = [], [], []
        for b in range(B):
            g3 = np.asarray(gts_3d_list[b].cpu() if torch.is_tensor(gts_3d_list[b]) else gts_3d_list[b], dtype=np.float32).reshape(-1, 16)
            if g3.shape[0] == 0:
                continue
            g2 = np.asarray(gts_2d_list[b].cpu() if torch.is_tensor(gts_2d_list[b]) else gts_2d_list[b], dtype=np.float32).reshape(g3.shape[0], -1)
            g3_all.append(g3); g2_all.append(g2[:, :4]); owner.append(np.full(g3.shape[0], b, np.int32))
        if g3_all:
            g3 = torch.from_numpy(np.concatenate(g3_all)).to(dev, non_blocking=True)
            g2 = torch.from_numpy(np.ascontiguousarray(np.concatenate(g2_all))).to(dev, non_blocking=True)
            gt_image = torch.from_numpy(np.concatenate(owner)).to(dev, non_blocking=True)
            gt7 = torch.stack([g3[:, 7], g3[:, 8], g3[:, 9], g3[:, 3], g3[:, 4], g3[:, 5], g3[:, 10]], dim=1).contiguous()
            rec_gt = ops.box3d_records(ops.corners_from_boxes7(gt7), mutate_input=False)
            ops.best_box_per_gt(rec.view(B, K, 8), box2d, n_img, rec_gt, g2, gt_image, self.best_target_box_beta, top, targets_after_nms)
        return scores_after_nms, targets_after_nms

    # ---------------------------------------------------------------------------------------------- :1091-1137
    def after_nms_loss(self, scores_after_nms, targets_after_nms, bbox_weights):
        """scores_after_nms / targets_after_nms [B,A]; bbox_weights numpy or tensor [B,A] (> 0 = foreground)."""
        dev = scores_after_nms.device
        w = torch.as_tensor(np.asarray(bbox_weights) if not torch.is_tensor(bbox_weights) else bbox_weights, device=dev)
        active = w > 0
        mode = self.after_nms_loss_mode
        if not bool(active.any()):
            loss = torch.zeros((), device=dev)
        elif mode == "rank" and not self.rank_boxes_of_all_images_at_once:                     # :1120-1131
            loss = torch.zeros((), device=dev)
            cnt = 0
            for b in range(scores_after_nms.shape[0]):
                if bool(active[b].any()):
                    cnt += 1
                    loss = loss + self.apLoss(scores_after_nms[b, active[b]], targets_after_nms[b, active[b]].detach()).squeeze()
            if cnt > 0:
                loss = loss / cnt
        elif mode == "rank":                                                                    # :1118-1119
            loss = self.apLoss(scores_after_nms[active], targets_after_nms[active].detach()).squeeze()
        elif mode == "classify":                                                                # :1103-1115
            t = targets_after_nms[active].detach()
            wts = torch.ones_like(t)
            npos, nneg = int((t == 1).sum()), int((t == 0).sum())
            if npos > 0 and nneg > 0:
                wts[t == 0] = float(np.power(npos / nneg, 0.25))
            l = wts * F.binary_cross_entropy(scores_after_nms[active], t, reduction='none')
            loss = l[torch.isfinite(l)].mean()
        elif mode == "regress":                                                                 # :1133-1135
            l = F.l1_loss(scores_after_nms[active], targets_after_nms[active].detach().clone(), reduction='none')
            loss = l[torch.isfinite(l)].mean()
        else:
            raise ValueError("unknown after_nms_loss_mode {}".format(mode))
        return self.after_nms_lambda * loss                                                     # :1137
