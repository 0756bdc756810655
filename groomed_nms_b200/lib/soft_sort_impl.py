"""SoftSort variant (reference lib/groomed_nms.py:131-165, sorting_method="soft").

No shipped config uses it (lowest priority).  The soft permutation is a dense
N x N softmax followed by one plain GEMM; it is expressed with torch composite ops (cuBLAS GEMM) so that autograd
reaches the scores through the permutation, and the grouped rescore + its backward still run in the sm_100a
kernels (the soft-permuted matrix receives its gradient through gnms_backward_f32's grad_iou)."""
import torch

from .. import ops


def soft_sort(scores, full_matrix=None, temperature=0.01):
    hard_sorted_scores, _ = torch.sort(scores, descending=True)
    init = -torch.abs(scores - hard_sorted_scores.unsqueeze(1))
    max_m = torch.max(init, dim=1)[0]
    comb = torch.exp((init - max_m.unsqueeze(-1)) / temperature)          # :149-153 (stable softmax)
    comb = comb / (torch.sum(comb, dim=1) + 1e-3)                         # :154-155 (row vector broadcast, as written)
    soft_scores = torch.matmul(comb, scores)                              # :159
    if full_matrix is None:
        return soft_scores, comb
    return soft_scores, comb, torch.matmul(comb, full_matrix)             # :164 rows only


def differentiable_nms_soft(scores, iou, nms_threshold, pruning_method, temperature, valid_box_prob_threshold,
                            return_sorted_prob, sorting_temperature, group_boxes, mask_group_boxes, group_size):
    _, indices = torch.sort(scores, descending=True, stable=True)         # :41
    s_soft, _, iou_soft = soft_sort(scores, full_matrix=iou, temperature=sorting_temperature)
    params = ops.make_params(nms_threshold, pruning_method, temperature, valid_box_prob_threshold,
                             return_sorted_prob, bool(group_boxes), bool(mask_group_boxes), group_size)
    if not iou_soft.requires_grad and s_soft.requires_grad:
        iou_soft = iou_soft.requires_grad_(True)
    prob, valid, invalid, counts = ops.GroomedNMSFunction.apply(s_soft, iou_soft, params)
    nv, ni = counts.tolist()
    return indices[valid[:nv]], indices[invalid[:ni]], prob
