"""SoftSort variant (reference lib/groomed_nms.py:131-165, sorting_method="soft") on the hand-written kernels of
csrc/softsort.cu: rank by counting, the exp / row-sum / column-wise division the reference performs, the matvec and the
N x N x N product (fp32 FMA-pipe GEMM), and the analytic backward -- one autograd.Function (ops.SoftSortFunction).
No shipped config uses it; it completes the module surface."""
import torch

from .. import ops
from ._util import Origin, to_cuda_f32


def soft_sort(scores, full_matrix=None, temperature=0.01):
    """-> (soft_sorted_scores, convex_comb_matrix[, soft_sorted_matrix]) like the reference (:131-165)."""
    origin = Origin(scores)
    s = to_cuda_f32(scores)
    m = to_cuda_f32(full_matrix) if full_matrix is not None else None
    out = ops.SoftSortFunction.apply(s, m, float(temperature))
    return tuple(origin.back(t) for t in out)


def differentiable_nms_soft(scores, iou, nms_threshold, pruning_method, temperature, valid_box_prob_threshold,
                            return_sorted_prob, sorting_temperature, group_boxes, mask_group_boxes, group_size):
    _, indices = torch.sort(scores, descending=True, stable=True)         # :41
    s_soft, _, iou_soft = ops.SoftSortFunction.apply(scores, iou, float(sorting_temperature))      # :45
    params = ops.make_params(nms_threshold, pruning_method, temperature, valid_box_prob_threshold,
                             return_sorted_prob, bool(group_boxes), bool(mask_group_boxes), group_size)
    prob, valid, invalid, counts = ops.GroomedNMSFunction.apply(s_soft, iou_soft, params)
    nv, ni = counts.tolist()
    return indices[valid[:nv]], indices[invalid[:ni]], prob
