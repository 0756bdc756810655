"""Device-tensor operators over the C-ABI (libgroomed_b200.so).

PyTorch is plumbing here: it owns device memory and streams; every computation below is a call into the
hand-written sm_100a kernels through ctypes.  All functions take CUDA tensors, enqueue on the current stream of
the tensor's device and never synchronise unless they must return a python int.  There is no CPU fallback.
"""
import ctypes

import torch

from . import _lib
from ._lib import Params, Saved, check

MAX_BOXES = _lib.MAX_BOXES


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("groomed_nms_b200: %s must be a CUDA tensor (no CPU fallback)" % name)


# ------------------------------------------------------------------------------------------------ overlaps
def overlap2d(box_a, box_b, kind=_lib.KIND_IOU):
    """[M,4] x [N,4] -> [M,N] IoU (lib/core.py:480) or intersection area (lib/core.py:178, transposed to [M,N])."""
    _require_cuda(box_a, "box_a")
    a, b = _f32c(box_a), _f32c(box_b)
    M, N = a.shape[0], b.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    if M and N:
        with torch.cuda.device(a.device):
            check(_lib.load().gnms_overlap2d_f32(_p(a), M, _p(b), N, _p(out), N, kind, _stream(a.device)), "gnms_overlap2d_f32")
    return out


def overlap2d_list(box_a, box_b, kind=_lib.KIND_IOU):
    _require_cuda(box_a, "box_a")
    a, b = _f32c(box_a), _f32c(box_b)
    M = a.shape[0]
    if b.shape[0] != M:
        raise ValueError("list mode needs the same number of boxes")
    out = torch.empty((M,), dtype=torch.float32, device=a.device)
    if M:
        with torch.cuda.device(a.device):
            check(_lib.load().gnms_overlap2d_list_f32(_p(a), _p(b), M, _p(out), kind, _stream(a.device)), "gnms_overlap2d_list_f32")
    return out


def overlap2d_f64(box_a, box_b, kind=_lib.KIND_IOU, list_mode=False, area_f32=0):
    """fp64 IoU / intersection of float64 CUDA boxes ([M, >=4] x [N, >=4], unit column stride) -> [M,N] or [M] float64.
    area_f32: bit 0 / 1 = side a / b was float32 before promotion (its areas are rounded to float32)."""
    _require_cuda(box_a, "box_a")
    a = box_a if box_a.stride(-1) == 1 else box_a.contiguous()
    b = box_b if box_b.stride(-1) == 1 else box_b.contiguous()
    M, N = a.shape[0], b.shape[0]
    if list_mode and M != N:
        raise ValueError("list mode needs the same number of boxes")
    out = torch.empty((M,) if list_mode else (M, N), dtype=torch.float64, device=a.device)
    if M and N:
        with torch.cuda.device(a.device):
            check(_lib.load().gnms_overlap2d_f64(_p(a), a.stride(0), M, _p(b), b.stride(0), N, kind, int(bool(list_mode)), int(area_f32), _p(out),
                                                 _stream(a.device)), "gnms_overlap2d_f64")
    return out


class Overlap2dFunction(torch.autograd.Function):
    """iou() with autograd (the reference's iou is a differentiable torch composite; the detection loss
    back-propagates through the list mode, lib/loss/rpn_3d.py:620)."""

    @staticmethod
    def forward(ctx, box_a, box_b, list_mode):
        ctx.list_mode = list_mode
        ctx.save_for_backward(box_a, box_b)
        if list_mode:
            return overlap2d_list(box_a, box_b, _lib.KIND_IOU)
        return overlap2d(box_a, box_b, _lib.KIND_IOU)

    @staticmethod
    def backward(ctx, g):
        box_a, box_b = ctx.saved_tensors
        ga, gb = overlap2d_backward(box_a, box_b, g, ctx.list_mode)
        return ga, gb, None


def overlap2d_backward(box_a, box_b, g, list_mode):
    a, b, g = _f32c(box_a), _f32c(box_b), _f32c(g)
    ga, gb = torch.zeros_like(a), torch.zeros_like(b)
    M, N = a.shape[0], b.shape[0]
    if M and N:
        with torch.cuda.device(a.device):
            check(_lib.load().gnms_iou2d_backward_f32(_p(a), M, _p(b), N, _p(g), int(list_mode), _p(ga), _p(gb),
                                                      _stream(a.device)), "gnms_iou2d_backward_f32")
    return ga, gb


def project_points(p2, pts, pad_ones):
    """lib/math_3d.py:47-72 on device: p2 [4,4], pts [3,n] (pad_ones) or [4,n] -> [4,n]."""
    _require_cuda(pts, "points")
    P, X = _f32c(p2), _f32c(pts)
    n = X.shape[1]
    out = torch.empty((4, n), dtype=torch.float32, device=X.device)
    if n:
        with torch.cuda.device(X.device):
            check(_lib.load().gnms_project_points_f32(_p(P), _p(X), n, int(pad_ones), _p(out), _stream(X.device)), "gnms_project_points_f32")
    return out


def corners_from_boxes7(boxes7, iou_3d_convention=True):
    """[N,7] (x,y,z,w,h,l,ry) -> corners [N,3,8] (lib/math_3d.py:364-435); iou_3d_convention selects the vertex order."""
    _require_cuda(boxes7, "boxes7")
    b = _f32c(boxes7)
    N = b.shape[0]
    out = torch.empty((N, 3, 8), dtype=torch.float32, device=b.device)
    if N:
        with torch.cuda.device(b.device):
            check(_lib.load().gnms_corners_from_boxes7_ex_f32(_p(b), b.stride(0), N, int(bool(iou_3d_convention)), _p(out),
                                                              _stream(b.device)), "gnms_corners_from_boxes7_f32")
    return out


def box3d_records(corners, mutate_input=False):
    """corners [N,3,8] (contiguous fp32) -> records [N,8]; optionally reproduces the reference's Y<-Z write."""
    _require_cuda(corners, "corners")
    if corners.dtype != torch.float32 or not corners.is_contiguous():
        if mutate_input:
            raise RuntimeError("mutate_input needs a contiguous fp32 corners tensor")
        corners = _f32c(corners)
    N = corners.shape[0]
    rec = torch.empty((N, 8), dtype=torch.float32, device=corners.device)
    if N:
        with torch.cuda.device(corners.device):
            check(_lib.load().gnms_box3d_records_f32(_p(corners), N, _p(rec), int(bool(mutate_input)), _stream(corners.device)), "gnms_box3d_records_f32")
    return rec


def overlap3d(rec_a, rec_b, want_bev=True, want_3d=True, generalized=False, affine=False, mul2d=None):
    M, N = rec_a.shape[0], rec_b.shape[0]
    dev = rec_a.device
    bev = torch.empty((M, N), dtype=torch.float32, device=dev) if want_bev else None
    o3d = torch.empty((M, N), dtype=torch.float32, device=dev) if want_3d else None
    if mul2d is not None:
        mul2d = _f32c(mul2d)
    if M and N:
        with torch.cuda.device(dev):
            check(_lib.load().gnms_overlap3d_f32(_p(rec_a), M, _p(rec_b), N, _p(bev), _p(o3d), N, int(generalized),
                                                 int(affine), _p(mul2d), _stream(dev)), "gnms_overlap3d_f32")
    return bev, o3d


def overlap3d_list(rec_a, rec_b, generalized=False, affine=False):
    M = rec_a.shape[0]
    dev = rec_a.device
    bev = torch.empty((M,), dtype=torch.float32, device=dev)
    o3d = torch.empty((M,), dtype=torch.float32, device=dev)
    if M:
        with torch.cuda.device(dev):
            check(_lib.load().gnms_overlap3d_list_f32(_p(rec_a), _p(rec_b), M, _p(bev), _p(o3d), int(generalized),
                                                      int(affine), _stream(dev)), "gnms_overlap3d_list_f32")
    return bev, o3d


class CornersFunction(torch.autograd.Function):
    """get_corners_of_cuboid with its analytic backward (the reference's composite of cos / sin / bmm is differentiable, and
    the acceptance-probability target differentiates it: lib/loss/rpn_3d.py:663-679)."""

    @staticmethod
    def forward(ctx, boxes7, iou_3d_convention):
        ctx.convention = bool(iou_3d_convention)
        b = _f32c(boxes7)
        ctx.save_for_backward(b)
        return corners_from_boxes7(b, ctx.convention)

    @staticmethod
    def backward(ctx, g):
        (b,) = ctx.saved_tensors
        return corners_backward(b, g, ctx.convention), None


def corners_backward(boxes7, grad_corners, iou_3d_convention=True):
    """dL/d(x, y, z, w, h, l, ry) [N,7] from dL/dcorners [N,3,8]."""
    b, g = _f32c(boxes7), _f32c(grad_corners)
    N = b.shape[0]
    gb = torch.empty((N, 7), dtype=torch.float32, device=b.device)
    if N:
        with torch.cuda.device(b.device):
            check(_lib.load().gnms_corners_backward_f32(_p(b), b.stride(0), N, int(bool(iou_3d_convention)), _p(g), _p(gb),
                                                        _stream(b.device)), "gnms_corners_backward_f32")
    return gb


class Iou3dApproxFunction(torch.autograd.Function):
    """iou3d_approximate (lib/core.py:305-421) -> (iou_bev, iou_3d) with the analytic backward wrt both corner sets.
    Works on private copies of the corners: the reference's in-place Y<-Z write (:379-380) is replayed by the caller as an
    ordinary torch op so that autograd records it exactly as it does in the reference."""

    @staticmethod
    def forward(ctx, corners_a, corners_b, list_mode, generalized):
        a, b = _f32c(corners_a.detach()).clone(), _f32c(corners_b.detach()).clone()
        ctx.list_mode, ctx.generalized = bool(list_mode), bool(generalized)
        ctx.save_for_backward(a, b)
        ra, rb = box3d_records(a), box3d_records(b)
        if list_mode:
            bev, o3d = overlap3d_list(ra, rb, generalized=generalized)
        else:
            bev, o3d = overlap3d(ra, rb, True, True, generalized=generalized)
        return bev, o3d

    @staticmethod
    def backward(ctx, g_bev, g_3d):
        a, b = ctx.saved_tensors
        ga, gb = iou3d_approx_backward(a, b, g_bev, g_3d, ctx.list_mode, ctx.generalized)
        return ga, gb, None, None


def iou3d_approx_backward(corners_a, corners_b, g_bev, g_3d, list_mode, generalized):
    """(dL/dcorners_a [M,3,8], dL/dcorners_b [N,3,8]) from dL/diou_bev and / or dL/diou_3d ([M,N], or [M] in list mode)."""
    a, b = _f32c(corners_a), _f32c(corners_b)
    M, N = a.shape[0], b.shape[0]
    ga, gb = torch.zeros_like(a), torch.zeros_like(b)
    if g_bev is None and g_3d is None:
        return ga, gb
    g_bev = _f32c(g_bev) if g_bev is not None else None
    g_3d = _f32c(g_3d) if g_3d is not None else None
    if M and N:
        with torch.cuda.device(a.device):
            check(_lib.load().gnms_iou3d_approx_backward_f32(_p(a), M, _p(b), N, int(bool(list_mode)), int(bool(generalized)),
                                                             _p(g_bev), _p(g_3d), _p(ga), _p(gb), _stream(a.device)),
                  "gnms_iou3d_approx_backward_f32")
    return ga, gb


# ------------------------------------------------------------------------------------------------ GrooMeD-NMS
def make_params(nms_threshold=0.4, pruning_method="linear", temperature=0.01, valid_box_prob_threshold=0.3,
                return_sorted_prob=False, group_boxes=True, mask_group_boxes=True, group_size=100, tril_in_input_order=False):
    if pruning_method not in _lib.PRUNE:
        raise NotImplementedError("Pruning method not implemented!")      # lib/groomed_nms.py:177
    mode = _lib.MODE_GROUP_MASK if (group_boxes and mask_group_boxes) else (
        _lib.MODE_GROUP_NOMASK if group_boxes else _lib.MODE_NOGROUP)
    if tril_in_input_order:
        if mode != _lib.MODE_GROUP_MASK:
            raise NotImplementedError("tril_in_input_order exists for group_boxes=True, mask_group_boxes=True only")
        mode = _lib.MODE_GROUP_MASK_INPUT_TRIL
    p = Params()
    p.nms_threshold = float(nms_threshold)
    p.temperature = float(temperature)
    p.valid_box_prob_threshold = float(valid_box_prob_threshold)
    p.pruning_method = _lib.PRUNE[pruning_method]
    p.mode = mode
    p.group_size = int(min(int(group_size), 2 ** 31 - 2))
    p.thresholded_output = 0 if group_boxes else 1
    p.sorted_output = 1 if return_sorted_prob else 0
    return p


_ws_cache = {}


def workspace(dev, N, batch):
    """Per-(device, stream) scratch, grown on demand; kernels are stream ordered so one buffer per stream is safe."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    need = _lib.load().gnms_workspace_bytes(N, batch)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty((int(need),), dtype=torch.uint8, device=dev)
        _ws_cache[key] = buf
    return buf


class ForwardState(object):
    """Outputs + saved tensors of one forward call (batch of B images, N boxes each)."""
    __slots__ = ("params", "B", "N", "prob", "valid_idx", "invalid_idx", "counts", "order", "sorted_scores", "lead",
                 "pval", "dpval", "pre", "ws", "n_per_image", "iou")

    def saved(self):
        return Saved(_p(self.order), _p(self.sorted_scores), _p(self.lead), _p(self.pval), _p(self.dpval), _p(self.pre))


def _alloc_state(dev, B, N, params, n_per_image, private_ws):
    st = ForwardState()
    st.params, st.B, st.N, st.n_per_image = params, B, N, n_per_image
    st.prob = torch.empty((B, N), dtype=torch.float32, device=dev)
    st.valid_idx = torch.empty((B, N), dtype=torch.int64, device=dev)
    st.invalid_idx = torch.empty((B, N), dtype=torch.int64, device=dev)
    st.counts = torch.empty((B, 2), dtype=torch.int32, device=dev)
    st.order = torch.empty((B, N), dtype=torch.int32, device=dev)
    st.lead = torch.empty((B, N), dtype=torch.int32, device=dev)
    fl = torch.empty((4, B, N), dtype=torch.float32, device=dev)
    st.sorted_scores, st.pval, st.dpval, st.pre = fl[0], fl[1], fl[2], fl[3]
    if private_ws:   # the backward of sorted_output / no-mask modes reads state kept in the workspace
        st.ws = torch.empty((int(_lib.load().gnms_workspace_bytes(N, B)),), dtype=torch.uint8, device=dev)
    else:
        st.ws = workspace(dev, N, B)
    st.iou = None
    return st


def forward_matrix(scores, iou, params, n_per_image=None, private_ws=None, opts=None):
    """scores [B,N] fp32 cuda, iou [B,N,N] fp32 cuda (unit column stride) -> ForwardState."""
    _require_cuda(scores, "scores")
    scores = _f32c(scores)
    if iou.dtype != torch.float32:
        iou = iou.float()
    if iou.stride(-1) != 1 or (iou.dim() == 3 and iou.stride(0) != iou.shape[1] * iou.stride(1)):
        iou = iou.contiguous()
    B, N = scores.shape
    if N > MAX_BOXES:
        raise RuntimeError("groomed_nms_b200: at most %d boxes per image (got %d)" % (MAX_BOXES, N))
    dev = scores.device
    if private_ws is None:
        private_ws = bool(params.sorted_output) or params.mode in (_lib.MODE_GROUP_NOMASK, _lib.MODE_NOGROUP)
    st = _alloc_state(dev, B, N, params, n_per_image, private_ws)
    st.iou = iou
    if B and N:
        ld = iou.stride(-2)
        with torch.cuda.device(dev):
            check(_lib.load().gnms_forward_ex_f32(_p(scores), _p(iou), ld, N, B, _p(n_per_image), ctypes.byref(params),
                                                  _p(st.prob), _p(st.valid_idx), _p(st.invalid_idx), _p(st.counts),
                                                  st.saved(), _p(st.ws), _lib.opts_ref(opts), _stream(dev)), "gnms_forward_f32")
    else:
        st.counts.zero_()
    return st


def forward_boxes(scores, boxes, box_kind, params, generalized=False, affine=False, n_per_image=None, private_ws=None,
                  overlap_out=None, opts=None):
    """Fused forward from boxes [B,N,4] (box_kind BOX_2D) or records [B,N,8] (BOX_3D_REC); overlap_out: optional
    [B,N,N] fp32 tensor that receives the overlap matrix (written once by the tile kernel)."""
    _require_cuda(scores, "scores")
    scores, boxes = _f32c(scores), _f32c(boxes)
    B, N = scores.shape
    if N > MAX_BOXES:
        raise RuntimeError("groomed_nms_b200: at most %d boxes per image (got %d)" % (MAX_BOXES, N))
    dev = scores.device
    if private_ws is None:
        private_ws = bool(params.sorted_output) or params.mode in (_lib.MODE_GROUP_NOMASK, _lib.MODE_NOGROUP)
    st = _alloc_state(dev, B, N, params, n_per_image, private_ws)
    if B and N:
        with torch.cuda.device(dev):
            check(_lib.load().gnms_forward_boxes_ex_f32(_p(scores), _p(boxes), box_kind, int(generalized), int(affine), N, B,
                                                        _p(n_per_image), ctypes.byref(params), _p(overlap_out), _p(st.prob),
                                                        _p(st.valid_idx), _p(st.invalid_idx), _p(st.counts), st.saved(),
                                                        _p(st.ws), _lib.opts_ref(opts), _stream(dev)), "gnms_forward_boxes_f32")
    else:
        st.counts.zero_()
    return st


def backward(st, grad_prob, need_grad_iou=False):
    """-> (grad_scores [B,N], grad_iou [B,N,N] or None)."""
    dev = grad_prob.device
    g = _f32c(grad_prob)
    B, N = st.B, st.N
    gs = torch.empty((B, N), dtype=torch.float32, device=dev)
    gi = torch.zeros((B, N, N), dtype=torch.float32, device=dev) if need_grad_iou else None
    if B and N:
        iou = st.iou
        with torch.cuda.device(dev):
            check(_lib.load().gnms_backward_f32(_p(g), _p(st.prob), _p(iou), iou.stride(-2) if iou is not None else 0, N, B,
                                                _p(st.n_per_image), ctypes.byref(st.params), st.saved(), _p(gs), _p(gi),
                                                N, _p(st.ws), _stream(dev)), "gnms_backward_f32")
    return gs, gi


class GroomedNMSFunction(torch.autograd.Function):
    """The autograd.Function that sits behind differentiable_nms: forward and the analytic
    backward are single C-ABI calls; iou receives a gradient only if it requires one (the training loss detaches
    it, lib/loss/rpn_3d.py:791)."""

    @staticmethod
    def forward(ctx, scores, iou, params):
        st = forward_matrix(scores.detach().unsqueeze(0), iou.detach().unsqueeze(0), params)
        ctx.st = st
        ctx.need_gi = iou.requires_grad
        ctx.mark_non_differentiable(st.valid_idx, st.invalid_idx, st.counts)
        return st.prob[0], st.valid_idx[0], st.invalid_idx[0], st.counts[0]

    @staticmethod
    def backward(ctx, g_prob, g_valid, g_invalid, g_counts):
        gs, gi = backward(ctx.st, g_prob.unsqueeze(0), need_grad_iou=ctx.need_gi)
        return gs[0], (gi[0] if gi is not None else None), None


class GroomedNMSInputOrderFunction(torch.autograd.Function):
    """GroomedNMSFunction whose probabilities come back in INPUT order (prob[i] belongs to input box i) instead of score-rank
    order: the form the reference's sorting_method="soft" path returns (its rows are never re-sorted, lib/groomed_nms.py:42-45).
    With return_sorted_prob the vector is sorted by value and the two coincide."""

    @staticmethod
    def forward(ctx, scores, iou, params):
        st = forward_matrix(scores.detach().unsqueeze(0), iou.detach().unsqueeze(0), params)
        ctx.st = st
        ctx.need_gi = iou.requires_grad
        ctx.mark_non_differentiable(st.valid_idx, st.invalid_idx, st.counts)
        if params.sorted_output:
            return st.prob[0], st.valid_idx[0], st.invalid_idx[0], st.counts[0]
        prob_in = torch.empty_like(st.prob[0])
        prob_in[st.order[0].long()] = st.prob[0]
        return prob_in, st.valid_idx[0], st.invalid_idx[0], st.counts[0]

    @staticmethod
    def backward(ctx, g_prob, g_valid, g_invalid, g_counts):
        st = ctx.st
        g = g_prob if st.params.sorted_output else g_prob[st.order[0].long()]
        gs, gi = backward(st, g.unsqueeze(0), need_grad_iou=ctx.need_gi)
        return gs[0], (gi[0] if gi is not None else None), None


class GroomedNMSBoxesFunction(torch.autograd.Function):
    """Matrix-free variant: overlaps are evaluated on the fly from boxes / 3D records (no gradient to the boxes,
    matching the detached overlap matrix of the training loss)."""

    @staticmethod
    def forward(ctx, scores, boxes, box_kind, params, generalized, affine):
        st = forward_boxes(scores.detach().unsqueeze(0), boxes.detach().unsqueeze(0), box_kind, params, generalized, affine)
        ctx.st = st
        ctx.mark_non_differentiable(st.valid_idx, st.invalid_idx, st.counts)
        return st.prob[0], st.valid_idx[0], st.invalid_idx[0], st.counts[0]

    @staticmethod
    def backward(ctx, g_prob, g_valid, g_invalid, g_counts):
        gs, _ = backward(ctx.st, g_prob.unsqueeze(0), need_grad_iou=False)
        return gs[0], None, None, None, None, None


class GroomedNMSBatchFunction(torch.autograd.Function):
    """Batched, ragged form of the two functions above: scores [B,N] (slots past n_per_image[b] are dead and come back
    as 0 with zero gradient), data = boxes [B,N,4] / records [B,N,8] (box_kind BOX_2D / BOX_3D_REC) or an overlap
    matrix [B,N,N] (box_kind None).  One forward call and one backward call for the whole batch, no host sync."""

    @staticmethod
    def forward(ctx, scores, data, box_kind, params, generalized, affine, n_per_image):
        if box_kind is None:
            st = forward_matrix(scores.detach(), data.detach(), params, n_per_image=n_per_image)
        else:
            st = forward_boxes(scores.detach(), data.detach(), box_kind, params, generalized, affine, n_per_image=n_per_image)
        ctx.st = st
        ctx.mark_non_differentiable(st.valid_idx, st.invalid_idx, st.counts)
        return st.prob, st.valid_idx, st.invalid_idx, st.counts

    @staticmethod
    def backward(ctx, g_prob, g_valid, g_invalid, g_counts):
        gs, _ = backward(ctx.st, g_prob.contiguous(), need_grad_iou=False)
        return gs, None, None, None, None, None, None


class SoftSortFunction(torch.autograd.Function):
    """soft_sort (lib/groomed_nms.py:131-165) on the hand-written kernels: forward (rank, exp / row sums, permutation matrix,
    P s, P M) and the analytic backward are one C-ABI call each.  Returns (soft_scores, P, soft_matrix or None)."""

    @staticmethod
    def forward(ctx, scores, matrix, temperature):
        _require_cuda(scores, "scores")
        s = _f32c(scores.detach())
        m = _f32c(matrix.detach()) if matrix is not None else None
        N = s.shape[0]
        if N > MAX_BOXES:
            raise RuntimeError("groomed_nms_b200: at most %d boxes (got %d)" % (MAX_BOXES, N))
        dev = s.device
        ss = torch.empty((N,), dtype=torch.float32, device=dev)
        P = torch.empty((N, N), dtype=torch.float32, device=dev)
        sm = torch.empty((N, N), dtype=torch.float32, device=dev) if m is not None else None
        perm = torch.empty((N,), dtype=torch.int32, device=dev)
        hs = torch.empty((N,), dtype=torch.float32, device=dev)
        S = torch.empty((N,), dtype=torch.float32, device=dev)
        if N:
            with torch.cuda.device(dev):
                check(_lib.load().gnms_soft_sort_forward_f32(_p(s), N, float(temperature), _p(m), m.stride(0) if m is not None else 0, _p(ss), _p(P),
                                                             _p(sm), _p(perm), _p(hs), _p(S), _stream(dev)), "gnms_soft_sort_forward_f32")
        ctx.save_for_backward(s, m, P, perm, hs, S)
        ctx.temperature = float(temperature)
        ctx.need_m = matrix is not None and matrix.requires_grad
        if sm is None:
            return ss, P
        return ss, P, sm

    @staticmethod
    def backward(ctx, g_ss, g_P, g_sm=None):
        s, m, P, perm, hs, S = ctx.saved_tensors
        N = s.shape[0]
        dev = s.device
        lib = _lib.load()
        grad_s = torch.empty((N,), dtype=torch.float32, device=dev)
        grad_m = torch.empty((N, N), dtype=torch.float32, device=dev) if (m is not None and ctx.need_m) else None
        if N:
            scratch = torch.empty((N, N), dtype=torch.float32, device=dev)
            ws = torch.empty((int(lib.gnms_soft_sort_workspace_bytes(N)),), dtype=torch.uint8, device=dev)
            g_ss = _f32c(g_ss) if g_ss is not None else None
            g_P = _f32c(g_P) if g_P is not None else None
            g_sm = _f32c(g_sm) if g_sm is not None else None
            with torch.cuda.device(dev):
                check(lib.gnms_soft_sort_backward_f32(_p(s), N, ctx.temperature, _p(m), m.stride(0) if m is not None else 0, _p(P), _p(perm), _p(hs),
                                                      _p(S), _p(g_ss), _p(g_P), _p(g_sm), _p(grad_s), _p(grad_m), _p(scratch), _p(ws),
                                                      _stream(dev)), "gnms_soft_sort_backward_f32")
        return grad_s, grad_m, None


def masked_topk(scores, mask, k):
    """scores [B,A], mask [B,A] (bool / uint8) -> (idx int64 [B,k] anchor ids by descending score among the masked ones, stable;
    n int32 [B] = min(#masked, k)).  One launch for the batch (lib/loss/rpn_3d.py:731-737)."""
    _require_cuda(scores, "scores")
    s = _f32c(scores)
    m = mask.to(torch.uint8).contiguous() if mask.dtype != torch.uint8 else mask.contiguous()
    B, A = s.shape
    idx = torch.empty((B, k), dtype=torch.int64, device=s.device)
    n = torch.empty((B,), dtype=torch.int32, device=s.device)
    with torch.cuda.device(s.device):
        check(_lib.load().gnms_masked_topk_f32(_p(s), _p(m), A, B, k, _p(idx), _p(n), _stream(s.device)), "gnms_masked_topk_f32")
    return idx, n


def best_box_per_gt(rec, box2d, n_per_image, gt_rec, gt_2d, gt_image, beta, top, targets):
    """Marks, in targets [B,A] (fp32, modified in place), the best candidate box of every ground truth whose score
    (0.5 (1 + GIoU3D)) * IoU2D exceeds beta (lib/loss/rpn_3d.py:813-825).  rec [B,K,8], box2d [B,K,4], top int64 [B,K],
    gt_rec [G,8], gt_2d [G,4], gt_image int32 [G].  Returns (best_slot int32 [G], best_score fp32 [G])."""
    B, K = rec.shape[0], rec.shape[1]
    G = gt_rec.shape[0]
    slot = torch.empty((G,), dtype=torch.int32, device=rec.device)
    score = torch.empty((G,), dtype=torch.float32, device=rec.device)
    if G and K:
        with torch.cuda.device(rec.device):
            check(_lib.load().gnms_best_box_per_gt_f32(_p(_f32c(rec)), _p(_f32c(box2d)), _p(n_per_image), K, _p(_f32c(gt_rec)), _p(_f32c(gt_2d)),
                                                       _p(gt_image), G, float(beta), _p(top), targets.shape[1], _p(targets), _p(slot),
                                                       _p(score), _stream(rec.device)), "gnms_best_box_per_gt_f32")
    return slot, score


def score_head_forward(x, wb):
    """scores[..] = sigmoid(x[.., :K] . wb[:K] + wb[K]) (K = 64): the per-box linear + sigmoid head of the batched
    configuration (the reference's 1x1-conv acceptance head, models/densenet121_3d_dilate_decomp_alpha.py:112-121,230)."""
    _require_cuda(x, "x")
    x, wb = _f32c(x), _f32c(wb)
    K = x.shape[-1]
    M = x.numel() // K
    out = torch.empty(x.shape[:-1], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().gnms_score_head_forward_f32(_p(x), M, K, _p(wb), _p(out), _stream(x.device)), "gnms_score_head_forward_f32")
    return out


def score_head_backward(x, scores, grad_scores, out=None):
    """-> grad_wb[K+1] = dL/d(weights, bias); `out` may be a view of a gradient bucket."""
    x, scores, g = _f32c(x), _f32c(scores), _f32c(grad_scores)
    K = x.shape[-1]
    M = x.numel() // K
    lib = _lib.load()
    if out is None:
        out = torch.empty((K + 1,), dtype=torch.float32, device=x.device)
    ws = torch.empty((int(lib.gnms_score_head_workspace_bytes(K)),), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.gnms_score_head_backward_f32(_p(x), M, K, _p(scores), _p(g), _p(out), _p(ws), _stream(x.device)), "gnms_score_head_backward_f32")
    return out


class ScoreHeadFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, wb):
        s = score_head_forward(x.detach(), wb.detach())
        ctx.save_for_backward(x, s)
        return s

    @staticmethod
    def backward(ctx, g):
        x, s = ctx.saved_tensors
        return None, score_head_backward(x, s, g.contiguous())


def prune(x, nms_threshold, temperature, pruning_method):
    """Elementwise pruning function (lib/groomed_nms.py:167-189)."""
    _require_cuda(x, "iou")
    xc = _f32c(x)
    out = torch.empty_like(xc)
    n = xc.numel()
    if n:
        with torch.cuda.device(xc.device):
            check(_lib.load().gnms_prune_f32(_p(xc), n, _lib.PRUNE[pruning_method], float(nms_threshold),
                                             float(temperature), _p(out), _stream(xc.device)), "gnms_prune_f32")
    return out


# ------------------------------------------------------------------------------------------------ groups / classical NMS
def get_groups_raw(scores, iou, group_threshold, group_size):
    """-> (group_id[N] int32 by input index, group_rank[N] int32, n_groups int32[1]) on device."""
    _require_cuda(scores, "scores")
    scores = _f32c(scores)
    if iou.dtype != torch.float32:
        iou = iou.float()
    if iou.stride(-1) != 1:
        iou = iou.contiguous()
    N = scores.shape[0]
    dev = scores.device
    gid = torch.empty((N,), dtype=torch.int32, device=dev)
    grk = torch.empty((N,), dtype=torch.int32, device=dev)
    ng = torch.zeros((1,), dtype=torch.int32, device=dev)
    if N:
        ws = workspace(dev, N, 1)
        with torch.cuda.device(dev):
            check(_lib.load().gnms_get_groups_f32(_p(scores), _p(iou), iou.stride(0), N, float(group_threshold),
                                                  int(min(int(group_size), 2 ** 31 - 2)), _p(gid), _p(grk), _p(ng),
                                                  _p(ws), _stream(dev)), "gnms_get_groups_f32")
    return gid, grk, ng


def hard_nms(dets, thresh, shift=1.0, cmp=_lib.CMP_GT):
    """dets [N,5] cuda -> (keep int32[N], n_keep int32[1]) on device; keep[:n_keep] are original indices."""
    _require_cuda(dets, "dets")
    d = _f32c(dets)
    N = d.shape[0]
    dev = d.device
    keep = torch.empty((max(N, 1),), dtype=torch.int32, device=dev)
    nk = torch.zeros((1,), dtype=torch.int32, device=dev)
    if N:
        ws = workspace(dev, N, 1)
        with torch.cuda.device(dev):
            check(_lib.load().gnms_hard_nms_f32(_p(d), N, float(thresh), float(shift), int(cmp), _p(keep), _p(nk), _p(ws),
                                                _stream(dev)), "gnms_hard_nms_f32")
    return keep, nk


def soft_nms(dets64, sigma, Nt, threshold, method, shift):
    """dets [N,5] float64 cuda -> (keep int32[N], scores f64[N], n_keep int32[1])."""
    _require_cuda(dets64, "dets")
    d = dets64.double().contiguous()
    N = d.shape[0]
    dev = d.device
    keep = torch.empty((max(N, 1),), dtype=torch.int32, device=dev)
    ks = torch.empty((max(N, 1),), dtype=torch.float64, device=dev)
    nk = torch.zeros((1,), dtype=torch.int32, device=dev)
    if N:
        ws = torch.empty((int(_lib.load().gnms_soft_nms_workspace_bytes(N)),), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(_lib.load().gnms_soft_nms_f64(_p(d), N, float(sigma), float(Nt), float(threshold), int(method),
                                                float(shift), _p(keep), _p(ks), _p(nk), _p(ws), _stream(dev)), "gnms_soft_nms_f64")
    return keep, ks, nk


def aploss(logits, targets):
    """logits[n], targets[n] cuda -> (loss[1], grad[n]) (lib/loss/aploss.py:16-81)."""
    _require_cuda(logits, "logits")
    x = _f32c(logits.reshape(-1))
    t = _f32c(targets.reshape(-1))
    n = x.shape[0]
    dev = x.device
    loss = torch.zeros((1,), dtype=torch.float32, device=dev)
    grad = torch.zeros((n,), dtype=torch.float32, device=dev)
    if n:
        nb = int(_lib.load().gnms_aploss_workspace_bytes(n))
        ws = torch.empty((nb,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(_lib.load().gnms_aploss_f32(_p(x), _p(t), n, _p(loss), _p(grad), _p(ws), nb, _stream(dev)), "gnms_aploss_f32")
    return loss, grad


def iou3d_exact(corners_a, corners_b, vol=None, list_mode=False):
    """corners_a [M,rows>=3,8], corners_b [N,rows>=3,8] cuda (any float dtype, computed in float64) ->
    (iou_bev, iou_3d), each float64 [M,N] (or [M] in list mode): exact polygon overlap (lib/core.py:246-302).
    vol: None or the per-pair volume sums, same shape as the result."""
    _require_cuda(corners_a, "corners_a")
    _require_cuda(corners_b, "corners_b")
    a = corners_a.double().contiguous()
    b = corners_b.double().contiguous()
    if a.dim() != 3 or b.dim() != 3 or a.shape[1] < 3 or b.shape[1] < 3 or a.shape[2] != 8 or b.shape[2] != 8:
        raise ValueError("corners must be [M, rows >= 3, 8]")
    M, N = a.shape[0], b.shape[0]
    if list_mode and M != N:
        raise ValueError("list mode needs as many boxes in both sets")
    dev = a.device
    shape = (M,) if list_mode else (M, N)
    bev = torch.empty(shape, dtype=torch.float64, device=dev)
    v3d = torch.empty(shape, dtype=torch.float64, device=dev)
    v = None
    if vol is not None:
        v = vol.to(dev).double().reshape(shape).contiguous()
    if M and N:
        with torch.cuda.device(dev):
            check(_lib.load().gnms_iou3d_exact_f64(_p(a), a.shape[1] * 8, M, _p(b), b.shape[1] * 8, N,
                                                   _p(v) if v is not None else None, int(bool(list_mode)), _p(bev), _p(v3d),
                                                   _stream(dev)), "gnms_iou3d_exact_f64")
    return bev, v3d
