"""Mirror of the reference's lib/loss/aploss.py:14-97 (AP-loss, Chen et al. CVPR 2019) on one sm_100a kernel."""
import torch
import torch.nn as nn

from ... import ops
from .._util import to_cuda_f32


class backpropAPLoss(torch.autograd.Function):
    """forward returns 1 - mean interpolated precision and stores the hand-derived gradient; backward scales it
    (reference lib/loss/aploss.py:16-87).  delta is fixed to 1.0 as in the reference (:18)."""

    @staticmethod
    def forward(ctx, logits, targets, delta=1.0, positive_label=1, negative_label=0):
        dev_in = logits.device
        t = to_cuda_f32(targets.detach())
        if positive_label != 1 or negative_label != 0:
            # the kernel knows positives as 1, negatives as 0 and everything else as "ignored" (:31,36 compare with the labels)
            t = torch.where(t == positive_label, torch.ones_like(t), torch.where(t == negative_label, torch.zeros_like(t), -torch.ones_like(t)))
        loss, grad = ops.aploss(to_cuda_f32(logits.detach()), t)
        ctx.grad = grad.reshape(logits.shape).to(dev_in)
        if bool((targets.max() <= 0).item()) if targets.numel() else True:
            return torch.zeros(1, device=dev_in)                    # reference returns `metric` of shape (1,) (:27-29)
        return loss.reshape(()).to(dev_in)

    @staticmethod
    def backward(ctx, grad_output):
        return ctx.grad * grad_output, None, None, None, None


class APLoss(nn.Module):
    def __init__(self, delta=1.0, positive_label=1, negative_label=0):
        super(APLoss, self).__init__()
        self.delta = delta
        self.positive_label = positive_label
        self.negative_label = negative_label

    def forward(self, logits, targets):
        return backpropAPLoss.apply(logits, targets, self.delta, self.positive_label, self.negative_label)
