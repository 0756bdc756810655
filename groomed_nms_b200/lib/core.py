"""Mirror of the overlap functions of the reference's lib/core.py:178-532 on the sm_100a tile kernels.

Public names and argument meaning follow the reference: intersect, iou, iou3d, iou3d_approximate, get_volume, get_hull,
remove_rotation_in_boxes (plus iou3d_batch, the many-pairs form of iou3d).  Results are bitwise equal to the reference's torch CPU ops (separately rounded fp32)."""
import numpy as np
import torch

from .. import _lib, ops
from ._util import Origin, is_f64, no_autograd, to_cuda_f32, to_cuda_f64
from ._util import device as _device


def _check_type(data_type, x):
    if data_type is None:
        data_type = type(x)
    if data_type not in (np.ndarray, torch.Tensor):
        raise ValueError('type {} is not implemented'.format(data_type))      # lib/core.py:216
    return data_type


def intersect(box_a, box_b, mode='combinations', data_type=None):
    """Intersection areas (reference lib/core.py:178-243).  combinations: box_a[M,4], box_b[N,4] -> [N,M]
    (b-major, like the reference :210-218); list: [M]."""
    _check_type(data_type, box_a)
    no_autograd("lib.core.intersect", box_a, box_b)
    origin = Origin(box_a)
    if mode not in ('combinations', 'list'):
        raise ValueError('unknown mode {}'.format(mode))                       # lib/core.py:243
    if is_f64(box_a) or is_f64(box_b):                                        # float64 on either side: float64 arithmetic (type promotion, as the reference)
        a, b = to_cuda_f64(box_a), to_cuda_f64(box_b)
        if mode == 'combinations':
            return origin.back(ops.overlap2d_f64(b, a, _lib.KIND_INTERSECT))
        return origin.back(ops.overlap2d_f64(a, b, _lib.KIND_INTERSECT, list_mode=True))
    a, b = to_cuda_f32(box_a), to_cuda_f32(box_b)
    if mode == 'combinations':
        out = ops.overlap2d(b[:, :4], a[:, :4], _lib.KIND_INTERSECT)
    elif mode == 'list':
        out = ops.overlap2d_list(a[:, :4], b[:, :4], _lib.KIND_INTERSECT)
    else:
        raise ValueError('unknown mode {}'.format(mode))                       # lib/core.py:243
    return origin.back(out, keep_np_dtype=True)


def iou(box_a, box_b, mode='combinations', data_type=None):
    """IoU (reference lib/core.py:480-532).  combinations: [M,4] x [N,4] -> [M,N]; list: [M].  No +1 convention;
    0/0 (coincident zero-area boxes) is NaN as in the reference.  The reference returns the combinations result
    as a transposed view (:508); this returns the same values contiguous."""
    _check_type(data_type, box_a)
    origin = Origin(box_a)
    if mode not in ('combinations', 'list'):
        raise ValueError('unknown mode {}'.format(mode))                       # lib/core.py:532
    if (is_f64(box_a) or is_f64(box_b)) and not (torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in (box_a, box_b))):
        # float64 in, float64 arithmetic (as the reference: lib/rpn_util.py:1295 thresholds float64 IoUs); the analytic backward
        # exists in fp32 only, so double tensors that require grad take the fp32 route below
        a, b = to_cuda_f64(box_a), to_cuda_f64(box_b)
        af32 = (0 if is_f64(box_a) else 1) | (0 if is_f64(box_b) else 2)     # a float32 side keeps float32 areas (type promotion)
        return origin.back(ops.overlap2d_f64(a, b, _lib.KIND_IOU, list_mode=(mode == 'list'), area_f32=af32))
    a, b = to_cuda_f32(box_a), to_cuda_f32(box_b)
    if mode == 'combinations':
        out = ops.Overlap2dFunction.apply(a[:, :4], b[:, :4], False)
    elif mode == 'list':
        out = ops.Overlap2dFunction.apply(a[:, :4], b[:, :4], True)
    else:
        raise ValueError('unknown mode {}'.format(mode))                       # lib/core.py:532
    return origin.back(out, keep_np_dtype=True)


def get_volume(corners_3d):
    """Axis-aligned volume of [N,3,8] (or [3,8]) corners (reference lib/core.py:434-451)."""
    no_autograd("lib.core.get_volume", corners_3d)
    origin = Origin(corners_3d)
    c = to_cuda_f32(corners_3d)
    if c.dim() == 2:
        c = c.unsqueeze(0)
    rec = ops.box3d_records(c.contiguous(), mutate_input=False)
    vol = rec[:, 6].contiguous()
    if origin.numpy:
        return origin.back(vol, keep_np_dtype=True)[0]                         # numpy branch returns a scalar (:449)
    return origin.back(vol)


def remove_rotation_in_boxes(boxes):
    """[N,4,2] -> [N,4] axis-aligned hull x1,y1,x2,y2 (reference lib/core.py:463-477)."""
    x1 = torch.min(boxes[:, :, 0], dim=1)[0]
    x2 = torch.max(boxes[:, :, 0], dim=1)[0]
    y1 = torch.min(boxes[:, :, 1], dim=1)[0]
    y2 = torch.max(boxes[:, :, 1], dim=1)[0]
    return torch.stack((x1, y1, x2, y2), dim=1)


def get_hull(y_min_b1, y_max_b1, y_min_b2, y_max_b2, mode="list"):
    """Hull extent along one axis (reference lib/core.py:423-432)."""
    if mode == "combinations":
        lo = torch.min(y_min_b1.unsqueeze(1), y_min_b2.unsqueeze(0))
        hi = torch.max(y_max_b1.unsqueeze(1), y_max_b2.unsqueeze(0))
    else:
        lo = torch.min(y_min_b1, y_min_b2)
        hi = torch.max(y_max_b1, y_max_b2)
    return torch.clamp(hi - lo, min=0)


def iou3d(corners_3d_b1, corners_3d_b2, vol=None):
    """Exact (polygon) BEV IoU and 3D IoU of two cuboids given as (3, 8) corner arrays (reference lib/core.py:246-302,
    which intersects two shapely polygons on the host).  Returns (iou_bev, iou_3d) as Python floats like the reference.
    `vol` is the sum of the two volumes (default: the sum of the two get_volume values, :278-279).
    The inputs are not modified (the reference works on copies, :273-274)."""
    c1, c2 = _corners_f64(corners_3d_b1), _corners_f64(corners_3d_b2)
    v = None if vol is None else torch.as_tensor(float(vol), dtype=torch.float64).reshape(1)
    bev, v3d = ops.iou3d_exact(c1[None], c2[None], vol=v, list_mode=True)
    out = torch.stack([bev[0], v3d[0]]).cpu()
    return float(out[0]), float(out[1])


def iou3d_batch(corners_3d_b1, corners_3d_b2, mode="combinations", vol=None):
    """Batched form of `iou3d`: corners [M,3,8] against [N,3,8] -> (iou_bev, iou_3d), each [M,N] ("combinations") or [M]
    ("list"), float64, in the container type of the first argument.  One launch instead of M*N shapely calls (the loops
    at lib/rpn_util.py:1770-1780, test/get_oracle_nms.py:125-131)."""
    if mode not in ("combinations", "list"):
        raise ValueError('unknown mode {}'.format(mode))
    origin = Origin(corners_3d_b1) if isinstance(corners_3d_b1, (np.ndarray, torch.Tensor)) else Origin(np.zeros(0))
    c1, c2 = _corners_f64(corners_3d_b1), _corners_f64(corners_3d_b2)
    if vol is not None and not isinstance(vol, torch.Tensor):
        vol = torch.as_tensor(np.asarray(vol, dtype=np.float64))
    bev, v3d = ops.iou3d_exact(c1, c2, vol=vol, list_mode=(mode == "list"))
    return origin.back(bev), origin.back(v3d)


def _corners_f64(c):
    if isinstance(c, np.ndarray):
        c = torch.from_numpy(np.ascontiguousarray(c, dtype=np.float64))
    elif not isinstance(c, torch.Tensor):
        c = torch.as_tensor(np.asarray(c, dtype=np.float64))
    if not c.is_cuda:
        c = c.to(_device())
    return c.double()


def iou3d_approximate(corners_3d_b1, corners_3d_b2, mode="list", method="normal"):
    """Axis-aligned approximate 3D IoU (reference lib/core.py:305-421) -> (iou_bev, iou_3d).

    corners are [M,3,8] / [N,3,8] (or [3,8]); combinations -> [M,N], list -> [M].  method "generalized" adds the
    GIoU hull term (:390-419).  Like the reference this OVERWRITES the Y row of its inputs with Z (:379-380) when
    they are contiguous fp32 CUDA tensors -- the training loss relies on that quirk (lib/loss/rpn_3d.py:813)."""
    if mode not in ("list", "combinations"):
        raise ValueError('unknown mode {}'.format(mode))
    origin = Origin(corners_3d_b1)
    c1, c2 = to_cuda_f32(corners_3d_b1), to_cuda_f32(corners_3d_b2)
    if c1.dim() == 2:
        c1, c2 = c1.unsqueeze(0), c2.unsqueeze(0)
    gen = method == "generalized"
    if torch.is_grad_enabled() and (c1.requires_grad or c2.requires_grad):
        # differentiable call (lib/loss/rpn_3d.py:663-679): analytic backward from private copies of the corners, then the
        # reference's own in-place Y <- Z write as a torch op, so that autograd sees the mutation exactly as in the reference
        # (and refuses a leaf that requires grad with torch's own error, as the reference does)
        bev, i3d = ops.Iou3dApproxFunction.apply(c1, c2, mode == "list", gen)
        for c in ((corners_3d_b1,) if corners_3d_b1 is corners_3d_b2 else (corners_3d_b1, corners_3d_b2)):
            if isinstance(c, torch.Tensor):
                v = c.unsqueeze(0) if c.dim() == 2 else c
                v[:, 1, :] = v[:, 2, :]
        return origin.back(bev), origin.back(i3d)

    def records(c, orig):
        inplace = isinstance(orig, torch.Tensor) and orig.is_cuda and orig.dtype == torch.float32 and c.is_contiguous() \
            and c.data_ptr() == orig.data_ptr()
        return ops.box3d_records(c if inplace else c.contiguous(), mutate_input=inplace)

    same = isinstance(corners_3d_b1, torch.Tensor) and corners_3d_b1 is corners_3d_b2
    r1 = records(c1, corners_3d_b1)
    r2 = r1 if same else records(c2, corners_3d_b2)
    if mode == "combinations":
        bev, i3d = ops.overlap3d(r1, r2, True, True, generalized=gen, affine=False)
    else:
        bev, i3d = ops.overlap3d_list(r1, r2, generalized=gen, affine=False)
    return origin.back(bev), origin.back(i3d)
