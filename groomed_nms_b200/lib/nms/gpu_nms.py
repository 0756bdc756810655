"""Mirror of the reference's Cython/CUDA lib/nms/gpu_nms.pyx:16-31 (+ nms_kernel.cu): device-resident hard NMS."""
import numpy as np
import torch

from ... import _lib, ops
from .._util import device


def gpu_nms(dets, thresh, device_id=0):
    """dets float32[N,5] numpy (x1,y1,x2,y2,score) -> list of kept ORIGINAL indices in descending score order.
    '+1' pixel convention, suppress when IoU > thresh (lib/nms/nms_kernel.cu:27-30,71)."""
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.shape[0] == 0:
        return []
    dev = torch.device("cuda", int(device_id)) if torch.cuda.is_available() else device()
    d = torch.from_numpy(dets).to(dev)
    keep, nk = ops.hard_nms(d, thresh, shift=1.0, cmp=_lib.CMP_GT)
    n = int(nk.item())
    return keep[:n].cpu().numpy().astype(np.int64).tolist()
