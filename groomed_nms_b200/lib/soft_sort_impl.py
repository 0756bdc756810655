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
    """sorting_method="soft" (reference lib/groomed_nms.py:42-45 and on).  The reference never re-sorts the soft-sorted rows:
    Phi is the lower triangle of the soft matrix in the HARD score order (:72) and the result comes back in that order, while
    get_groups sorts by the SOFT scores once more (:213).  The soft scores of near-tied boxes need not be monotone (rows with
    more near neighbours are divided by a larger sum + 1e-3, :156-157), so the two orders can differ: the grouping kernels work
    in soft-score order, cut Phi by input position (GNMS_MODE_GROUP_MASK_INPUT_TRIL) and the probabilities are returned in
    input order.  The inverse modes (group_boxes=False or mask_group_boxes=False) have no such cut and are exact only
    when the soft scores are monotone; otherwise they raise."""
    _, indices = torch.sort(scores, descending=True, stable=True)         # :41
    s_soft, _, iou_soft = ops.SoftSortFunction.apply(scores, iou, float(sorting_temperature))      # :45
    masked = bool(group_boxes) and bool(mask_group_boxes)
    if not masked and s_soft.numel() > 1 and not bool((s_soft[:-1] >= s_soft[1:]).all()):
        raise NotImplementedError("differentiable_nms(sorting_method='soft') without group masking: the soft-sorted scores are not "
                                  "monotone for this input (near-tied scores at this sorting_temperature); only the "
                                  "group_boxes=True, mask_group_boxes=True form is defined for that case")
    params = ops.make_params(nms_threshold, pruning_method, temperature, valid_box_prob_threshold,
                             return_sorted_prob, bool(group_boxes), bool(mask_group_boxes), group_size, tril_in_input_order=masked)
    prob, valid, invalid, counts = ops.GroomedNMSInputOrderFunction.apply(s_soft, iou_soft, params)
    nv, ni = counts.tolist()
    return indices[valid[:nv]], indices[invalid[:ni]], prob
