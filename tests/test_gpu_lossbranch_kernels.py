"""The two batched kernels around the NMS in the loss branch: masked top-K selection
(lib/loss/rpn_3d.py:731-737) and best box per ground truth (:813-825).  Index work: exact against torch's stable sort /
argmax on the same numbers."""
import numpy as np
import pytest
import torch

from gpu_util import cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("A,K", [(69120, 500), (5000, 500), (300, 500), (4096, 1024), (70000, 37)])
def test_masked_topk_matches_stable_sort(A, K):
    from groomed_nms_b200 import ops
    g = torch.Generator("cuda").manual_seed(A + K)
    B = 5
    s = torch.rand(B, A, device="cuda", generator=g)
    s[1] = (s[1] * 50).round() / 50                          # heavy ties: 51 distinct values
    s[2, ::7] = -s[2, ::7]                                   # negative scores and -0 / +0
    s[2, 5] = 0.0; s[2, 6] = -0.0
    mask = torch.rand(B, A, device="cuda", generator=g) < 0.3
    mask[3] = False                                          # an image without foreground
    mask[4] = torch.rand(A, device="cuda", generator=g) < (K * 0.5 / A)     # fewer than K foreground anchors
    idx, n = ops.masked_topk(s, mask, K)
    torch.cuda.synchronize()
    for b in range(B):
        fg = torch.nonzero(mask[b]).flatten()
        order = torch.sort(s[b][fg], descending=True, stable=True)[1]
        want = fg[order][:K]
        assert int(n[b]) == len(want)
        assert torch.equal(idx[b, :len(want)], want), b
        assert not idx[b, len(want):].any()


def test_best_box_per_gt_matches_torch_composite():
    from groomed_nms_b200 import ops
    rng = np.random.default_rng(0)
    B, K, A = 3, 500, 4000
    b7 = np.stack([rng.uniform(-20, 20, (B, K)), 1 + 0.1 * rng.standard_normal((B, K)), rng.uniform(6, 50, (B, K)), 1.6 + 0.1 * rng.standard_normal((B, K)),
                   1.5 + 0.1 * rng.standard_normal((B, K)), 4 + 0.4 * rng.standard_normal((B, K)), rng.uniform(-3, 3, (B, K))], 2).astype(np.float32)
    c = rng.uniform(0, 1200, (B, K, 2)); wh = rng.uniform(20, 200, (B, K, 2))
    box2d = np.concatenate([c - wh / 2, c + wh / 2], 2).astype(np.float32)
    n_img = np.array([500, 137, 0], np.int32)
    top = np.stack([rng.permutation(A)[:K] for _ in range(B)]).astype(np.int64)
    owners = np.array([0, 0, 0, 1, 1, 2], np.int32)
    pick = [(0, 3), (0, 77), (0, 499), (1, 5), (1, 400), (2, 1)]            # ground truths = jittered copies of candidates
    gt7 = np.stack([b7[b, i] for b, i in pick]) + 0.02
    gt2 = np.stack([box2d[b, i] for b, i in pick]) + 1.0
    rec = ops.box3d_records(ops.corners_from_boxes7(cuda(b7).view(B * K, 7))).view(B, K, 8)
    rec_gt = ops.box3d_records(ops.corners_from_boxes7(cuda(gt7)))
    targets = torch.zeros((B, A), device="cuda")
    slot, score = ops.best_box_per_gt(rec, cuda(box2d), cuda(n_img, torch.int32), rec_gt, cuda(gt2), cuda(owners, torch.int32), 0.3,
                                      cuda(top, torch.int64), targets)
    torch.cuda.synchronize()
    want_t = torch.zeros((B, A), device="cuda")
    for g, b in enumerate(owners):
        n = int(n_img[b])
        if n == 0:
            assert int(slot[g]) == -1
            continue
        iou2d = ops.overlap2d(cuda(box2d[b, :n]), cuda(gt2[g:g + 1]))
        _, sc = ops.overlap3d(rec[b, :n].contiguous(), rec_gt[g:g + 1].contiguous(), False, True, generalized=True, affine=True, mul2d=iou2d)
        best, arg = torch.max(sc[:, 0], dim=0)
        assert int(slot[g]) == int(arg) and float(score[g]) == float(best), g
        if float(best) > 0.3:
            want_t[b, top[b, int(arg)]] = 1
    assert torch.equal(targets, want_t) and int(want_t.sum()) >= 3
