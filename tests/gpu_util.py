import numpy as np
import torch

F32 = np.float32


def cuda(x, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(x)).to("cuda").to(dtype)


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=F32)
    b = np.ascontiguousarray(b, dtype=F32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def nbits_diff(a, b):
    a = np.ascontiguousarray(a, dtype=F32)
    b = np.ascontiguousarray(b, dtype=F32)
    return int((a.view(np.uint32) != b.view(np.uint32)).sum())
