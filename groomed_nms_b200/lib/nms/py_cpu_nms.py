"""Mirror of the reference's lib/nms/py_cpu_nms.py:10-38 (survivors are `ovr <= thresh`)."""
import numpy as np
import torch

from ... import _lib, ops
from .._util import device


def py_cpu_nms(dets, thresh):
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.shape[0] == 0:
        return []
    keep, nk = ops.hard_nms(torch.from_numpy(dets).to(device()), thresh, shift=1.0, cmp=_lib.CMP_NLE)
    return keep[:int(nk.item())].cpu().numpy().astype(np.int64).tolist()
