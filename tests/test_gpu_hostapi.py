"""GPU parity of the allocation-free runners behind bench.py (groomed_nms_b200.hostapi): the CUDA-graph plan with the
matrix written on its own graph branch, the matrix-free plan, the blocking host-buffer call and the streaming host
pipeline must all give what the plain C-ABI calls give -- and those are checked against the oracle elsewhere."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(B, N, seed=5):
    from groomed_nms_b200 import synthetic
    boxes = np.stack([synthetic.config_c3(seed=seed + 10 * i, n=N, k=max(2, N // 128))[0] for i in range(B)])
    scores = np.stack([synthetic.config_c3(seed=seed + 10 * i, n=N, k=max(2, N // 128))[1] for i in range(B)])
    grads = np.random.default_rng(seed).standard_normal((B, N)).astype(np.float32)
    return boxes, scores, grads


def _reference(boxes, scores, grads, params):
    """Plain calls: corners -> records -> overlap matrix -> forward from the MATRIX -> backward."""
    from groomed_nms_b200 import ops
    B, N = scores.shape
    dev = torch.device("cuda", 0)
    rec = ops.box3d_records(ops.corners_from_boxes7(torch.from_numpy(boxes).to(dev).view(B * N, 7))).view(B, N, 8)
    ov = torch.stack([ops.overlap3d(rec[b], rec[b], False, True, generalized=True, affine=True)[1] for b in range(B)])
    st = ops.forward_matrix(torch.from_numpy(scores).to(dev), ov, params)
    gs, _ = ops.backward(st, torch.from_numpy(grads).to(dev))
    return ov, st, gs


@pytest.mark.parametrize("B,N", [(3, 1024), (2, 1000), (5, 256)])
@pytest.mark.parametrize("tiles_per_cta", [0, 3])
@pytest.mark.parametrize("matrix_kernel", [1, 2])
def test_plans_match_plain_calls(B, N, tiles_per_cta, matrix_kernel):
    from groomed_nms_b200 import _lib, ops
    from groomed_nms_b200.hostapi import Nms3dPlan
    dev = torch.device("cuda", 0)
    params = ops.make_params()
    boxes, scores, grads = _inputs(B, N)
    ov, st, gs = _reference(boxes, scores, grads, params)
    for mat, branch in ((True, True), (True, False), (False, False)):
        pl = Nms3dPlan(B, N, dev, params, materialise=mat, overlap_branch=branch)
        pl.matrix_opts = _lib.launch_opts(matrix_kernel=matrix_kernel, tiles_per_cta=tiles_per_cta)
        pl.forward_opts = _lib.launch_opts(matrix_kernel=matrix_kernel)
        pl.boxes7.copy_(torch.from_numpy(boxes)); pl.scores.copy_(torch.from_numpy(scores)); pl.grad_prob.copy_(torch.from_numpy(grads))
        g = pl.capture()
        for _ in range(3):                                  # replays must be idempotent
            g.replay()
        torch.cuda.synchronize()
        if mat:
            assert torch.equal(pl.overlap.view(torch.int32), ov.view(torch.int32))
        assert torch.equal(pl.prob, st.prob) and torch.equal(pl.counts, st.counts) and torch.equal(pl.lead, st.lead)
        for b in range(B):
            nv = int(st.counts[b, 0])
            assert torch.equal(pl.valid_idx[b, :nv], st.valid_idx[b, :nv])
        assert torch.allclose(pl.grad_scores, gs, rtol=1e-6, atol=1e-7)


def test_host_runner_and_pipeline_match_plain_calls():
    from groomed_nms_b200 import ops
    from groomed_nms_b200.hostapi import HostPipeline, HostRunner
    dev = torch.device("cuda", 0)
    params = ops.make_params()
    B, N = 4, 768
    batches = [_inputs(B, N, seed=20 + k) for k in range(5)]
    want = [_reference(*x, params) for x in batches]
    pinned = [[torch.from_numpy(a).pin_memory() for a in x] for x in batches]
    runner = HostRunner(B, N, dev, params)
    for x, (_, st, gs) in zip(pinned, want):
        prob, grad, valid, counts = runner.run_host(*x)
        assert torch.equal(prob, st.prob.cpu()) and torch.equal(counts, st.counts.cpu())
        assert torch.allclose(grad, gs.cpu(), rtol=1e-6, atol=1e-7)
    pipe = HostPipeline(B, N, dev, params, depth=3)
    tickets = []
    for k, x in enumerate(pinned):
        tickets.append(pipe.submit(*x))
        if k >= 2:                                          # read a result while later calls are in flight
            j = k - 2
            prob, grad, valid, counts = pipe.wait(tickets[j])
            assert torch.equal(prob, want[j][1].prob.cpu()) and torch.equal(counts, want[j][1].counts.cpu())
            assert torch.allclose(grad, want[j][2].cpu(), rtol=1e-6, atol=1e-7)
    pipe.drain()
    for j in (3, 4):
        prob, grad, valid, counts = pipe.wait(tickets[j])
        assert torch.equal(prob, want[j][1].prob.cpu())
        nv = int(counts[0, 0])
        assert valid.dtype == torch.int32 and valid.shape == (B, 512)                    # compact keep lists (keep_cap)
        assert torch.equal(valid[0, :nv].long(), want[j][1].valid_idx[0, :nv].cpu()) and bool((valid[0, nv:] == -1).all())
    # the full int64 lists (keep_cap = 0) and an upstream gradient that stays on the device
    full = HostRunner(B, N, dev, params, keep_cap=0, grad_on_device=True)
    full.set_grad(torch.from_numpy(batches[0][2]).to(dev))
    prob, grad, valid, counts = full.run_host(pinned[0][0], pinned[0][1])
    assert valid.dtype == torch.int64 and valid.shape == (B, N)
    nv = int(counts[1, 0])
    assert torch.equal(valid[1, :nv], want[0][1].valid_idx[1, :nv].cpu()) and torch.allclose(grad, want[0][2].cpu(), rtol=1e-6, atol=1e-7)
    assert full.h2d_bytes == B * N * 32 and runner.h2d_bytes == B * N * 36


def test_benched_shape_parity_64_images_of_4096():
    """Exactly what bench.py times: Nms3dPlan(64, 4096, materialise=True) -- the matrix kernel on its own graph branch with
    CTAs that retire after one 256 x 64 tile, the NMS kernels on a higher-priority branch, one CUDA graph -- replayed; 4 sampled
    images' overlap matrices bit for bit against the general (non-symmetric, non-tiled) overlap kernel, and every image's
    rescored scores, leaders, keep lists and score gradients against per-image plain calls."""
    from groomed_nms_b200 import _lib, ops, synthetic
    from groomed_nms_b200.hostapi import Nms3dPlan
    dev = torch.device("cuda", 0)
    B, N = 64, 4096
    params = ops.make_params(nms_threshold=0.4, pruning_method="linear", temperature=0.01, valid_box_prob_threshold=0.3,
                             group_boxes=True, mask_group_boxes=True, group_size=100)
    boxes = np.stack([synthetic.config_c3(seed=3 + 10 * i)[0] for i in range(B)])
    scores = np.stack([synthetic.config_c3(seed=3 + 10 * i)[1] for i in range(B)])
    grads = np.random.default_rng(1234).standard_normal((B, N)).astype(np.float32)
    pl = Nms3dPlan(B, N, dev, params, materialise=True, overlap_branch=True)
    pl.matrix_opts = _lib.launch_opts(tiles_per_cta=4)                  # bench.py's default
    pl.boxes7.copy_(torch.from_numpy(boxes)); pl.scores.copy_(torch.from_numpy(scores)); pl.grad_prob.copy_(torch.from_numpy(grads))
    g = pl.capture()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    rec = pl.rec.view(B, N, 8)
    for b in (0, 21, 42, 63):
        other = rec[b].clone()                                           # a different pointer: the general M x N kernel, not the tiles
        _, want = ops.overlap3d(rec[b], other, False, True, generalized=True, affine=True)
        assert torch.equal(pl.overlap[b].view(torch.int32), want.view(torch.int32)), b
    up = torch.from_numpy(grads).to(dev)
    for b in range(B):
        st = ops.forward_boxes(pl.scores[b:b + 1], rec[b:b + 1], _lib.BOX_3D_REC, params, generalized=True, affine=True,
                               opts=_lib.launch_opts(election=_lib.ELECT_MASK))     # the all-pairs route, not the direct election
        assert torch.equal(pl.prob[b], st.prob[0]) and torch.equal(pl.lead[b], st.lead[0]) and torch.equal(pl.counts[b], st.counts[0]), b
        nv = int(st.counts[0, 0])
        assert torch.equal(pl.valid_idx[b, :nv], st.valid_idx[0, :nv]), b
        gs, _ = ops.backward(st, up[b:b + 1])
        assert torch.allclose(pl.grad_scores[b], gs[0], rtol=1e-6, atol=1e-7), b
