import numpy as np
import torch

F32 = np.float32


def cuda(x, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(x)).to("cuda").to(dtype)


def bits_equal(a, b):
    """Bitwise equality of fp32 arrays; NaNs must sit at the same places (their payload/sign is not compared:
    x86 produces 0xFFC00000 for 0/0, the GPU 0x7FFFFFFF)."""
    a = np.ascontiguousarray(a, dtype=F32)
    b = np.ascontiguousarray(b, dtype=F32)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


def nbits_diff(a, b):
    a = np.ascontiguousarray(a, dtype=F32)
    b = np.ascontiguousarray(b, dtype=F32)
    return int((a.view(np.uint32) != b.view(np.uint32)).sum())
