"""Score head of the batched-image configuration (the reference's 1x1-conv + sigmoid acceptance head,
models/densenet121_3d_dilate_decomp_alpha.py:112-121,230) and the whole batched step against plain torch fp32 autograd
composed with the (separately pinned) NMS function.  Floating point: 1e-5 relative (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from gpu_util import cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m", [1, 31, 4096, 70001])
def test_score_head_forward_backward_vs_torch(m):
    from groomed_nms_b200 import ops
    g = torch.Generator("cuda").manual_seed(m)
    x = torch.randn(m, 64, device="cuda", generator=g)
    wb = (torch.randn(65, device="cuda", generator=g) * 0.3).requires_grad_(True)
    up = torch.randn(m, device="cuda", generator=g)
    s = ops.ScoreHeadFunction.apply(x, wb)
    s.backward(up)
    wb2 = wb.detach().clone().requires_grad_(True)
    s2 = torch.sigmoid(x.double() @ wb2[:64].double() + wb2[64].double())
    s2.backward(up.double())
    assert torch.allclose(s.detach(), s2.detach().float(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(wb.grad, wb2.grad, rtol=1e-4, atol=1e-5 * float(wb2.grad.abs().max()))
    # deterministic: the two-stage reduction has a fixed order
    again = ops.score_head_backward(x, s.detach(), up)
    assert torch.equal(again, wb.grad)


@pytest.mark.parametrize("materialise", [False, True])
def test_batched_head_plan_matches_autograd_composition(materialise):
    from groomed_nms_b200 import _lib, ops, synthetic
    from groomed_nms_b200.hostapi import BatchedHeadPlan
    B, N = 3, 1024
    p = ops.make_params(group_size=100)
    pl = BatchedHeadPlan(B, N, torch.device("cuda", 0), p, materialise=materialise, bucket_pad_elems=1000)
    boxes = np.stack([synthetic.config_c4_image(i, n=N)[0] for i in range(B)])
    rng = np.random.default_rng(0)
    pl.boxes.copy_(cuda(boxes)); pl.x.copy_(cuda(rng.standard_normal((B, N, 64)).astype(np.float32)))
    pl.grad_prob.copy_(cuda(rng.standard_normal((B, N)).astype(np.float32))); pl.wb.copy_(cuda((rng.standard_normal(65) * 0.3).astype(np.float32)))
    g = pl.capture(reduce=True)                       # single process: the all-reduce is a no-op
    g.replay(); g.replay()
    torch.cuda.synchronize()
    wb = pl.wb.clone().requires_grad_(True)
    s = ops.ScoreHeadFunction.apply(pl.x, wb)
    prob = ops.GroomedNMSBatchFunction.apply(s, pl.boxes, _lib.BOX_2D, p, False, False, None)[0]
    (prob * pl.grad_prob).sum().backward()
    assert torch.equal(prob.detach(), pl.prob)
    assert torch.allclose(pl.grad_wb, wb.grad, rtol=1e-6, atol=0)
    assert not pl.bucket.flat[68:].any()
    if materialise:
        want = ops.overlap2d(pl.boxes[1], pl.boxes[1])
        assert torch.equal(pl.overlap[1], want)


def test_head_backward_with_in_kernel_allreduce_single_rank():
    """gnms_score_head_backward_allreduce_f32 with a one-rank exchange (the 2-GPU form is tests/test_gpu_multi.py): the exchange
    buffer life cycle through the C-ABI, step parity over several calls (the counter lives in the buffer), an empty shard, and
    the result -- with one rank the rank-ordered sum is the rank's own gradient, bit for bit the plain backward."""
    import ctypes
    from groomed_nms_b200 import _lib, ops
    lib = _lib.load()
    own = ctypes.c_void_p()
    handle = ctypes.create_string_buffer(64)
    _lib.check(lib.gnms_peer_buffer_create(ctypes.byref(own), handle), "create")
    assert own.value and any(handle.raw)
    peers = _lib.Peers()
    peers.buf[0] = own.value
    peers.rank, peers.world = 0, 1
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws = torch.empty(int(lib.gnms_score_head_workspace_bytes(64)), dtype=torch.uint8, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator("cuda").manual_seed(5)
    for step, m in enumerate([3000, 17, 0, 5000]):
        x = torch.randn(m, 64, device="cuda", generator=g)
        s = torch.rand(m, device="cuda", generator=g)
        up = torch.randn(m, device="cuda", generator=g)
        out = torch.full((65,), 7.0, device="cuda")
        _lib.check(lib.gnms_score_head_backward_allreduce_f32(ops._p(x), m, 64, ops._p(s), ops._p(up), ops._p(out), ops._p(ws),
                                                              ctypes.byref(peers), ops._p(status), st), "allreduce")
        want = ops.score_head_backward(x, s, up) if m else torch.zeros(65, device="cuda")
        assert torch.equal(out, want), step
    assert int(status.item()) == 0
    bad = _lib.Peers()
    bad.rank, bad.world = 0, 1                                                 # no buffer
    assert lib.gnms_score_head_backward_allreduce_f32(ops._p(x), 1, 64, ops._p(s), ops._p(up), ops._p(out), ops._p(ws),
                                                      ctypes.byref(bad), None, st) != 0
    torch.cuda.synchronize()
    assert lib.gnms_peer_buffer_close(own, 1) == 0
