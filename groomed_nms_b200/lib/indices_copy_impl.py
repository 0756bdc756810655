"""indices_copy (reference lib/groomed_nms.py:272-336) on the gnms_indices_copy_f32 scatter kernel."""
import ctypes

import numpy as np
import torch

from .. import _lib
from ..ops import _p, _stream
from ._util import device


class _IndicesCopy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, B, ra, ca, rb, cb, inplace):
        out = A if inplace else A.clone()
        if inplace:
            ctx.mark_dirty(A)
        C = 1 if A.dim() == 2 else A.shape[2]
        npairs = ra.shape[0]
        ctx.save_for_backward(ra, ca if ca is not None else ra.new_empty(0), rb if rb is not None else ra.new_empty(0),
                              cb if cb is not None else ra.new_empty(0))
        ctx.product = ca is None
        ctx.shapeB = B.shape
        if npairs:
            with torch.cuda.device(out.device):
                _lib.check(_lib.load().gnms_indices_copy_f32(_p(out), out.shape[1], _p(B), B.shape[1], C, _p(ra), _p(ca),
                                                             _p(rb), _p(cb), npairs, _stream(out.device)),
                           "gnms_indices_copy_f32")
        return out

    @staticmethod
    def backward(ctx, g):
        ra, ca, rb, cb = ctx.saved_tensors
        gA = g.clone()
        if ctx.product:
            gB = g[ra][:, ra].reshape(ctx.shapeB).clone()
            gA[ra.unsqueeze(1), ra.unsqueeze(0)] = 0
        else:
            gB = torch.zeros(ctx.shapeB, dtype=g.dtype, device=g.device)
            gB[rb, cb] = g[ra, ca]
            gA[ra, ca] = 0
        return gA, gB, None, None, None, None, None


def indices_copy(A, B, indA, indB=None, inplace=True):
    dev_in = A.device
    dev = A.device if A.is_cuda else device()
    shapeA = A.shape
    A_c = A.to(dev).float()
    B_c = B.to(dev).float()
    A_c = A_c if A_c.is_contiguous() else A_c.contiguous()
    B_c = B_c if B_c.is_contiguous() else B_c.contiguous()
    if B_c.dim() == 1:
        B_c = B_c.unsqueeze(1)
    indA = indA.to(dev).long().contiguous()
    if indA.dim() == 1 and indB is None:
        out = _IndicesCopy.apply(A_c, B_c, indA, None, None, None, inplace and A_c.data_ptr() == A.data_ptr())
    else:
        if indA.dim() == 1:                                           # :301-307 all combinations of indA
            g = torch.cartesian_prod(indA, indA)
            ra, ca = g[:, 0].contiguous(), g[:, 1].contiguous()
        else:
            ra, ca = indA[:, 0].contiguous(), indA[:, 1].contiguous()
        if indB is None:                                              # :309-314 all of B, row-major
            rb = torch.arange(B_c.shape[0], device=dev).repeat_interleave(B_c.shape[1])
            cb = torch.arange(B_c.shape[1], device=dev).repeat(B_c.shape[0])
        else:
            indB = indB.to(dev).long()
            rb, cb = indB[:, 0].contiguous(), indB[:, 1].contiguous()
        if rb.shape[0] != ra.shape[0]:
            raise RuntimeError("indices_copy: %d destination pairs but %d source pairs" % (ra.shape[0], rb.shape[0]))
        out = _IndicesCopy.apply(A_c, B_c, ra, ca, rb, cb, inplace and A_c.data_ptr() == A.data_ptr())
    out = out.view(shapeA)
    return out if out.device == dev_in else out.to(dev_in)
