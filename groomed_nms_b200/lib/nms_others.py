"""Mirror of the reference's lib/nms_others.py (comparison NMS variants) on the sm_100a kernels."""
import numpy as np
import torch

from .. import _lib, ops
from ._util import device


def navneeth_soft_nms(boxes, sigma=0.5, Nt=0.4, threshold=0.001, method=0, shift=1):
    """Soft-NMS (reference lib/nms_others.py:6-116): method 0 hard / 1 linear / 2 gaussian decay, boxes whose score
    falls below `threshold` are discarded; returns `keep_orig[:N]` (original indices in selection order).
    Arithmetic is float64.  Unlike the reference the input array is not permuted in place."""
    b = np.ascontiguousarray(boxes, dtype=np.float64)
    if b.shape[0] == 0:
        return np.arange(0)
    keep, _, nk = ops.soft_nms(torch.from_numpy(b).to(device()), sigma, Nt, threshold, method, shift)
    return keep[:int(nk.item())].cpu().numpy().astype(np.int64)


def girshick_nms(dets, thresh, shift=1):
    """Hard NMS with a configurable pixel shift (reference lib/nms_others.py:119-150) -> list of kept indices."""
    d = np.ascontiguousarray(dets, dtype=np.float32)
    if d.shape[0] == 0:
        return []
    keep, nk = ops.hard_nms(torch.from_numpy(d).to(device()), thresh, shift=float(shift), cmp=_lib.CMP_NLE)
    return keep[:int(nk.item())].cpu().numpy().astype(np.int64).tolist()
