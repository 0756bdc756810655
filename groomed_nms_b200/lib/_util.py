"""Input normalisation shared by the mirrored reference functions.

The reference functions accept numpy arrays, CPU tensors and CUDA tensors and answer in kind.  The product has no
CPU compute path: host inputs are copied to the current CUDA device, processed by the sm_100a kernels and copied
back, so the caller still gets the type/device it passed in."""
import numpy as np
import torch


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("groomed_nms_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class Origin(object):
    """Remembers what the caller passed (numpy / cpu tensor / cuda tensor) to answer in kind."""

    def __init__(self, x):
        self.numpy = isinstance(x, np.ndarray)
        self.np_dtype = x.dtype if self.numpy else None
        self.device = x.device if isinstance(x, torch.Tensor) else torch.device("cpu")

    def back(self, t, keep_np_dtype=False):
        if self.numpy:
            out = t.detach().cpu().numpy()
            if keep_np_dtype and out.dtype != self.np_dtype and np.issubdtype(self.np_dtype, np.floating):
                out = out.astype(self.np_dtype)
            return out
        return t if t.device == self.device else t.to(self.device)


def to_cuda_f32(x):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if not x.is_cuda:
        x = x.to(device())
    return x.float() if x.dtype != torch.float32 else x


def is_f64(x):
    """float64 numpy array or tensor: the reference's functions compute in the dtype they are given."""
    return (isinstance(x, np.ndarray) and x.dtype == np.float64) or (isinstance(x, torch.Tensor) and x.dtype == torch.float64)


def to_cuda_f64(x):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if not x.is_cuda:
        x = x.to(device())
    return x.double() if x.dtype != torch.float64 else x


def no_autograd(name, *tensors):
    """The kernel behind `name` has no backward: refuse inputs that require grad instead of silently dropping it."""
    if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        raise RuntimeError("groomed_nms_b200.%s has no autograd backward: detach() its inputs (as the GrooMeD-NMS path does, "
                           "lib/loss/rpn_3d.py:791) or call it through groomed_nms_b200.install(), which routes differentiable "
                           "calls to the reference's own composite" % name)
