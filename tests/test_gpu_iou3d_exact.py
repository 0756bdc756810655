"""GPU parity: exact cuboid overlap `iou3d` (lib/core.py:246-302) -- the kernel against the
oracle (a different construction of the same area, pinned on closed forms in test_oracle_golden.py) and, bit for bit,
against the host build of the very source it runs."""
import math

import numpy as np
import pytest
import torch

from polygon_cases import cuboid, host_core, random_cuboids

pytestmark = pytest.mark.gpu


def _set():
    return np.concatenate([random_cuboids(70, seed=11, spread=4.0),
                           np.stack([cuboid(0, 12, 4, 2, 0), cuboid(0, 12, 4, 2, 0), cuboid(4, 12, 4, 2, 0),
                                     cuboid(0, 12, 2, 1, 0), cuboid(0, 12, 4, 2, math.pi / 2), cuboid(40, 12, 4, 2, 0.2)])])


def test_kernel_equals_host_build_bitwise_and_oracle(tmp_path):
    from groomed_nms_b200 import ops
    from oracle import polygon_oracle as PO
    c = _set()
    n = c.shape[0]
    t = torch.from_numpy(c).cuda()
    bev, v3 = ops.iou3d_exact(t, t)
    bev, v3 = bev.cpu().numpy(), v3.cpu().numpy()
    hb, h3 = host_core(tmp_path)(c, c)
    assert np.array_equal(bev.view(np.uint64), hb.view(np.uint64))
    assert np.array_equal(v3.view(np.uint64), h3.view(np.uint64))
    want = np.array([[PO.iou3d(c[i], c[j]) for j in range(n)] for i in range(n)])
    assert (bev > 0).sum() > n + 200
    np.testing.assert_allclose(bev, want[:, :, 0], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(v3, want[:, :, 1], rtol=1e-11, atol=1e-13)


def test_list_mode_volume_argument_padding_row_and_empty(tmp_path):
    from groomed_nms_b200 import ops
    from oracle import polygon_oracle as PO
    c = _set()
    n = c.shape[0]
    c4 = np.concatenate([c, np.ones((n, 1, 8))], axis=1)                 # callers hold (4, 8) homogeneous corners
    vol = np.arange(n) + 50.0
    lb, l3 = ops.iou3d_exact(torch.from_numpy(c4).cuda(), torch.from_numpy(c4[::-1].copy()).cuda(),
                             vol=torch.from_numpy(vol), list_mode=True)
    for i in range(n):
        wb, w3 = PO.iou3d(c[i], c[n - 1 - i], vol=vol[i])
        assert abs(float(lb[i]) - wb) < 1e-12 and abs(float(l3[i]) - w3) < 1e-12
    e = torch.zeros((0, 3, 8), dtype=torch.float64, device="cuda")
    b0, _ = ops.iou3d_exact(e, torch.from_numpy(c).cuda())
    assert tuple(b0.shape) == (0, n)
    with pytest.raises(ValueError):
        ops.iou3d_exact(torch.from_numpy(c[:3]).cuda(), torch.from_numpy(c[:4]).cuda(), list_mode=True)


def test_reference_call_shape_scalar_pair_and_batch():
    """`iou3d(corners_3d_b1, corners_3d_b2, vol)` as the evaluation loops call it (lib/rpn_util.py:1776): numpy (3, 8)
    in, two Python floats out, inputs untouched; the batched form agrees with it."""
    from groomed_nms_b200.lib import core
    from oracle import polygon_oracle as PO
    a, b = cuboid(0.3, 15, 4.2, 1.8, 0.4, h=1.6), cuboid(0.9, 15.5, 3.9, 1.7, -0.2, h=1.5)
    a0, b0 = a.copy(), b.copy()
    bev, v3 = core.iou3d(a, b, 1.6 * 4.2 * 1.8 + 1.5 * 3.9 * 1.7)
    assert isinstance(bev, float) and isinstance(v3, float)
    wb, w3 = PO.iou3d(a, b, 1.6 * 4.2 * 1.8 + 1.5 * 3.9 * 1.7)
    assert abs(bev - wb) < 1e-13 and abs(v3 - w3) < 1e-13 and 0.2 < bev < 0.9
    assert np.array_equal(a, a0) and np.array_equal(b, b0)
    bev2, _ = core.iou3d(a.astype(np.float32), b.astype(np.float32))     # float32 corners are widened
    assert abs(bev2 - wb) < 1e-5
    c = _set()[:20]
    mb, m3 = core.iou3d_batch(c, c[:7])
    assert isinstance(mb, np.ndarray) and mb.shape == (20, 7) and mb.dtype == np.float64
    for i, j in ((0, 0), (3, 5), (19, 6)):
        sb, s3 = core.iou3d(c[i], c[j])
        assert sb == mb[i, j] and s3 == m3[i, j]
    lb, _ = core.iou3d_batch(torch.from_numpy(c[:7]), torch.from_numpy(c[:7]), mode="list")
    assert isinstance(lb, torch.Tensor) and not lb.is_cuda and torch.allclose(lb, torch.ones(7, dtype=torch.float64))
    with pytest.raises(ValueError):
        core.iou3d_batch(c, c, mode="pairs")
