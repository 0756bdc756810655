"""Mirror of the reference's lib/groomed_nms.py on the sm_100a kernels.

Same public names and signatures: differentiable_nms, get_groups, pruning_function, indices_copy, soft_sort,
sigmoid_numpy, cast_to_cpu_cuda_tensor.  Reference: abhi1kumar/groomed_nms lib/groomed_nms.py (file:line cited
per function)."""
import numpy as np
import torch

from .. import _lib, ops
from ._util import Origin, to_cuda_f32


def differentiable_nms(scores_unsorted, iou_unsorted, nms_threshold=0.4, pruning_method="linear", temperature=0.01,
                       valid_box_prob_threshold=0.3, return_sorted_prob=False, sorting_method="hard",
                       sorting_temperature=None, group_boxes=True, mask_group_boxes=True, group_size=100, debug=False):
    """GrooMeD-NMS (reference lib/groomed_nms.py:10-129).

    Returns (valid_boxes_index, invalid_boxes_index, non_suppression_prob) exactly as the reference:
    indices are int64 in INPUT index space (valid ordered by rescored probability, descending); the probabilities
    are in SCORE-SORTED order (unthresholded when group_boxes, thresholded otherwise, :124-127; sorted and
    thresholded when return_sorted_prob, :116-119).  numpy inputs are accepted (:34-36) and answered with CPU
    tensors like the reference.  Gradients flow to scores, and to iou if it requires grad, through one
    autograd.Function whose backward is the analytic kernel."""
    origin = Origin(scores_unsorted)
    scores = to_cuda_f32(scores_unsorted)
    iou = to_cuda_f32(iou_unsorted)
    if sorting_method == "soft":
        from .soft_sort_impl import differentiable_nms_soft
        if sorting_temperature is None:
            sorting_temperature = temperature                                           # :43-44
        out = differentiable_nms_soft(scores, iou, nms_threshold, pruning_method, temperature,
                                      valid_box_prob_threshold, return_sorted_prob, sorting_temperature,
                                      group_boxes, mask_group_boxes, group_size)
    else:
        params = ops.make_params(nms_threshold, pruning_method, temperature, valid_box_prob_threshold,
                                 return_sorted_prob, bool(group_boxes), bool(mask_group_boxes), group_size)
        prob, valid, invalid, counts = ops.GroomedNMSFunction.apply(scores, iou, params)
        nv, ni = counts.tolist()               # the reference returns variable-length index tensors: one sync
        out = (valid[:nv], invalid[:ni], prob)
    if origin.numpy:
        return tuple(t.cpu() for t in out)     # the reference converts numpy inputs to CPU tensors (:34-36)
    return tuple(origin.back(t) for t in out)


def get_groups(iou_unsorted, group_threshold, scores_unsorted, group_size=100, return_original_indices=True):
    """Greedy grouping (reference lib/groomed_nms.py:208-270) -> list[LongTensor].

    Leaders are the classical-NMS keep set on the score-sorted matrix (strict `>`), every other box joins the
    first leader that overlaps it, a group keeps its first group_size+1 boxes and overflow boxes belong to no
    group (:247-262)."""
    origin = Origin(scores_unsorted)
    scores = to_cuda_f32(scores_unsorted)
    iou = to_cuda_f32(iou_unsorted)
    n = scores.shape[0]
    gid, grk, ng = ops.get_groups_raw(scores, iou, group_threshold, group_size)
    ng = int(ng.item())
    members = torch.nonzero(gid >= 0).flatten()
    key = gid[members].long() * (n + 1) + grk[members].long()
    members = members[torch.argsort(key)]
    sizes = torch.bincount(gid[members].long(), minlength=ng).tolist()
    if not return_original_indices:
        rank = torch.empty(n, dtype=torch.long, device=scores.device)
        rank[torch.sort(scores, descending=True, stable=True)[1]] = torch.arange(n, device=scores.device)
        members = rank[members]
    members = members.to(origin.device)
    return list(torch.split(members, sizes))


def pruning_function(iou, nms_threshold=0.4, temperature=0.01, pruning_method="linear"):
    """reference lib/groomed_nms.py:167-189 (torch and numpy inputs)."""
    if pruning_method not in _lib.PRUNE:
        raise NotImplementedError("Pruning method not implemented!")
    origin = Origin(iou)
    if pruning_method == "linear":
        return iou
    x = to_cuda_f32(iou)
    return origin.back(ops.prune(x, nms_threshold, temperature, pruning_method), keep_np_dtype=True)


def sigmoid_numpy(x):
    """reference lib/groomed_nms.py:191-199."""
    return pruning_function(np.asarray(x), nms_threshold=0.0, temperature=1.0, pruning_method="sigmoidal")


def cast_to_cpu_cuda_tensor(input, reference_tensor):
    """reference lib/groomed_nms.py:201-206."""
    if reference_tensor.is_cuda and not input.is_cuda:
        input = input.cuda()
    if not reference_tensor.is_cuda and input.is_cuda:
        input = input.cpu()
    return input


def indices_copy(A, B, indA, indB=None, inplace=True):
    """Scatter B into A (reference lib/groomed_nms.py:272-336).

    indA 1-D: A[indA[i], indA[j]] = B[i, j]; indA 2-D: rows of (row, col) pairs; indB: pairs into B (default all of
    B in row-major order).  3-D tensors carry a trailing channel dimension."""
    from .indices_copy_impl import indices_copy as _impl
    return _impl(A, B, indA, indB, inplace)


def soft_sort(scores, full_matrix=None, temperature=0.01):
    """SoftSort (reference lib/groomed_nms.py:131-165)."""
    from .soft_sort_impl import soft_sort as _impl
    return _impl(scores, full_matrix, temperature)
