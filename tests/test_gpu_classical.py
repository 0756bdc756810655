"""GPU parity: classical hard NMS family, Soft-NMS and AP loss vs the reference's golden vectors / the oracle."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import cuda

pytestmark = pytest.mark.gpu


def test_hard_nms_family_golden():
    from groomed_nms_b200.lib.nms.gpu_nms import gpu_nms
    from groomed_nms_b200.lib.nms.cpu_nms import cpu_nms
    from groomed_nms_b200.lib.nms.py_cpu_nms import py_cpu_nms
    from groomed_nms_b200.lib.nms_others import girshick_nms
    from oracle import groomed_oracle as O
    g = load_golden("classical_nms")
    dets = g["dets"]
    for thr in (0.3, 0.5, 0.7):
        want = list(g["py_cpu_nms_%g" % thr])
        assert py_cpu_nms(dets, thr) == want
        assert gpu_nms(dets, thr, device_id=0) == want          # no NaN / ties here: `>` and `!(<=)` agree
        assert cpu_nms(dets, thr) == O.hard_nms(dets, thr, shift=1, ge=True)
        for shift in (0, 1):
            assert girshick_nms(dets, thr, shift=shift) == list(g["girshick_%g_s%d" % (thr, shift)])
    assert gpu_nms(dets[:0], 0.5) == [] and gpu_nms(dets[:1], 0.5) == [0]


@pytest.mark.parametrize("n", [2, 63, 64, 65, 1000, 3000, 8192])
def test_hard_nms_random_vs_oracle(n):
    from groomed_nms_b200 import synthetic
    from groomed_nms_b200.lib.nms.gpu_nms import gpu_nms
    from oracle import groomed_oracle as O
    boxes, sc, _ = synthetic.clustered_boxes_2d(n, max(1, n // 40), seed=n, canvas=(1200.0, 400.0), size=(70.0, 50.0), jitter=0.2)
    dets = np.concatenate([np.round(boxes), sc[:, None]], 1).astype(np.float32)
    assert gpu_nms(dets, 0.5) == O.hard_nms(dets, 0.5, shift=1)


def test_nms_host_abi_drop_in_for__nms():
    """gnms_nms_host has the reference's `_nms` signature (lib/nms/gpu_nms.hpp:1-2): host pointers, sorted boxes."""
    from groomed_nms_b200 import _lib, synthetic
    from oracle import groomed_oracle as O
    boxes, sc, _ = synthetic.clustered_boxes_2d(500, 10, seed=3, canvas=(800.0, 300.0), size=(70.0, 50.0), jitter=0.2)
    dets = np.concatenate([np.round(boxes), sc[:, None]], 1).astype(np.float32)
    order = dets[:, 4].argsort()[::-1]
    sdets = np.ascontiguousarray(dets[order])
    keep = np.zeros(500, np.int32)
    num = ctypes.c_int(0)
    rc = _lib.load().gnms_nms_host(keep.ctypes.data, ctypes.addressof(num), sdets.ctypes.data, 500, 5, 0.5, 0)
    assert rc == 0
    assert list(order[keep[:num.value]]) == O.hard_nms(dets, 0.5, shift=1)


def test_soft_nms_golden():
    from groomed_nms_b200 import ops
    from groomed_nms_b200.lib.nms_others import navneeth_soft_nms
    g = load_golden("classical_nms")
    dets = g["dets"].astype(np.float64)
    for method in (0, 1, 2):
        for shift in (0, 1):
            thr = 0.05 if method else 0.001
            want = g["soft_m%d_s%d_keep" % (method, shift)]
            keep = navneeth_soft_nms(dets.copy(), sigma=0.5, Nt=0.4, threshold=thr, method=method, shift=shift)
            assert keep.tolist() == want.tolist()
            k, ks, nk = ops.soft_nms(cuda(dets, torch.float64), 0.5, 0.4, thr, method, shift)
            n = int(nk.item())
            assert np.allclose(ks[:n].cpu().numpy(), g["soft_m%d_s%d_scores" % (method, shift)], rtol=1e-12)


def test_aploss_golden_and_autograd():
    from groomed_nms_b200.lib.loss.aploss import APLoss
    g = load_golden("aploss")
    x = cuda(g["logits"]).requires_grad_(True)
    loss = APLoss()(x, cuda(g["labels"]))
    assert np.allclose(loss.item(), g["joint_loss"], rtol=1e-6) and np.allclose(loss.item(), 0.49164116, rtol=1e-6)
    loss.backward()
    assert np.allclose(x.grad.cpu().numpy(), g["joint_grad"], rtol=1e-5, atol=1e-7)
    x.grad.zero_()
    loss = (APLoss()(x[0], cuda(g["labels"])[0]) + APLoss()(x[1], cuda(g["labels"])[1])) / 2
    loss.backward()
    assert np.allclose(loss.item(), g["mean_loss"], rtol=1e-6)
    assert np.allclose(x.grad.cpu().numpy(), g["mean_grad"], rtol=1e-5, atol=1e-7)
    x = cuda(g["rand_logits"]).requires_grad_(True)
    loss = APLoss()(x, cuda(g["rand_targets"]))
    (loss * 1.7).backward()
    assert np.allclose(loss.item(), g["rand_loss"], rtol=1e-5)
    assert np.allclose(x.grad.cpu().numpy(), g["rand_grad_x1p7"], rtol=1e-4, atol=1e-7)
    z = APLoss()(cuda(g["rand_logits"]), torch.zeros(400, device="cuda"))
    assert z.shape == (1,) and z.item() == 0.0


def test_aploss_custom_labels():
    """positive_label / negative_label (lib/loss/aploss.py:31,36): relabelled targets give the same loss and gradient."""
    from groomed_nms_b200.lib.loss.aploss import APLoss
    g = load_golden("aploss")
    t = cuda(g["rand_targets"])
    relabelled = torch.where(t == 1, torch.full_like(t, 5.0), torch.where(t == 0, torch.full_like(t, 2.0), torch.full_like(t, -3.0)))
    x = cuda(g["rand_logits"]).requires_grad_(True)
    loss = APLoss(positive_label=5, negative_label=2)(x, relabelled)
    loss.backward()
    assert np.allclose(loss.item(), g["rand_loss"], rtol=1e-5)
    assert np.allclose(x.grad.cpu().numpy() * 1.7, g["rand_grad_x1p7"], rtol=1e-4, atol=1e-7)
