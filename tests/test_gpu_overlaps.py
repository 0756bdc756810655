"""GPU parity: overlap kernels (through the C-ABI) vs the oracle and the reference's golden vectors.
Bar: bitwise equal fp32 for every overlap computed from boxes / corners; 1e-6 relative for the cos/sin + rotation
that produces corners from 7-DoF boxes (backend-defined op order in the reference)."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from gpu_util import bits_equal, cuda, nbits_diff

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from groomed_nms_b200.lib import core, math_3d
    return core, math_3d


def test_iou2d_golden_bitwise(L):
    core, _ = L
    g = load_golden("iou2d")
    a, b = cuda(g["box_a"]), cuda(g["box_b"])
    assert bits_equal(core.iou(a, b).cpu().numpy(), g["iou_comb"])
    assert bits_equal(core.intersect(a, b).cpu().numpy(), g["inter_comb"])
    assert bits_equal(core.iou(a, b[:37], mode="list").cpu().numpy(), g["iou_list"])
    assert bits_equal(core.intersect(a, b[:37], mode="list").cpu().numpy(), g["inter_list"])
    assert bits_equal(core.iou(a, a).cpu().numpy(), g["iou_self"])
    # numpy in -> numpy out, same values
    out = core.iou(g["box_a"], g["box_b"])
    assert isinstance(out, np.ndarray) and bits_equal(out, g["iou_comb"])
    with pytest.raises(ValueError):
        core.iou(a, b, mode="nope")


@pytest.mark.parametrize("M,N", [(1, 1), (3, 5), (17, 513), (130, 1030), (1024, 1024), (777, 777)])
def test_iou2d_random_vs_oracle(L, M, N):
    from oracle import groomed_oracle as O
    core, _ = L
    rng = np.random.default_rng(M * 1000 + N)
    a = rng.uniform(0, 200, (M, 2)); a = np.concatenate([a, a + rng.uniform(0, 80, (M, 2))], 1).astype(np.float32)
    b = rng.uniform(0, 200, (N, 2)); b = np.concatenate([b, b + rng.uniform(0, 80, (N, 2))], 1).astype(np.float32)
    assert bits_equal(core.iou(cuda(a), cuda(b)).cpu().numpy(), O.iou(a, b))
    assert bits_equal(core.intersect(cuda(a), cuda(b)).cpu().numpy(), O.intersect(a, b))


def test_iou2d_empty_and_unaligned(L):
    core, _ = L
    a = torch.rand(0, 4, device="cuda"); b = torch.rand(5, 4, device="cuda")
    assert core.iou(a, b).shape == (0, 5) and core.iou(b, a).shape == (5, 0)
    from oracle import groomed_oracle as O
    x = np.random.default_rng(0).uniform(0, 50, (9, 5)).astype(np.float32)
    x[:, 2:4] += x[:, 0:2]
    got = core.iou(cuda(x)[:, :4], cuda(x)[:, :4])          # non-contiguous slice of a [N,5] dets array
    assert bits_equal(got.cpu().numpy(), O.iou(x[:, :4], x[:, :4]))


def test_iou2d_autograd_matches_torch_composite(L):
    core, _ = L
    rng = np.random.default_rng(3)
    a = rng.uniform(0, 100, (40, 2)); a = np.concatenate([a, a + rng.uniform(5, 60, (40, 2))], 1).astype(np.float32)
    b = rng.uniform(0, 100, (40, 2)); b = np.concatenate([b, b + rng.uniform(5, 60, (40, 2))], 1).astype(np.float32)

    def torch_iou(A, B, mode):
        if mode == "list":
            mx = torch.min(A[:, 2:], B[:, 2:]); mn = torch.max(A[:, :2], B[:, :2])
            inter = torch.clamp(mx - mn, 0); inter = inter[:, 0] * inter[:, 1]
            aa = (A[:, 2] - A[:, 0]) * (A[:, 3] - A[:, 1]); ab = (B[:, 2] - B[:, 0]) * (B[:, 3] - B[:, 1])
            return inter / (aa + ab - inter)
        mx = torch.min(A[:, 2:4], B[:, 2:4].unsqueeze(1)); mn = torch.max(A[:, 0:2], B[:, 0:2].unsqueeze(1))
        inter = torch.clamp(mx - mn, 0); inter = inter[:, :, 0] * inter[:, :, 1]
        aa = (A[:, 2] - A[:, 0]) * (A[:, 3] - A[:, 1]); ab = (B[:, 2] - B[:, 0]) * (B[:, 3] - B[:, 1])
        return (inter / (aa.unsqueeze(0) + ab.unsqueeze(1) - inter)).permute(1, 0)

    for mode in ("list", "combinations"):
        A1, B1 = cuda(a).requires_grad_(True), cuda(b).requires_grad_(True)
        A2, B2 = cuda(a).requires_grad_(True), cuda(b).requires_grad_(True)
        o1 = core.iou(A1, B1, mode=mode); o2 = torch_iou(A2, B2, mode)
        w = torch.randn_like(o2)
        (o1 * w).sum().backward(); (o2 * w).sum().backward()
        assert torch.allclose(A1.grad, A2.grad, rtol=1e-4, atol=1e-6)
        assert torch.allclose(B1.grad, B2.grad, rtol=1e-4, atol=1e-6)


def test_corners_and_projection(L):
    _, m3 = L
    g = load_golden("corners")
    a = g["args64"].astype(np.float32)
    c = m3.get_corners_of_cuboid(*[cuda(a[:, i]) for i in range(7)])
    assert c.shape == (5, 3, 8)
    assert np.allclose(c.cpu().numpy(), g["corners_torch"], rtol=1e-6, atol=1e-5)
    cn = m3.get_corners_of_cuboid(*[g["args64"][:, i] for i in range(7)])
    assert isinstance(cn, np.ndarray) and np.allclose(cn, g["corners_np"], rtol=1e-5, atol=1e-4)
    pts = cuda(g["corners_torch"]).transpose(1, 2).reshape((-1, 3)).transpose(0, 1)
    pr = m3.project_3d_points_in_4D_format(cuda(g["p2"]), pts, pad_ones=True)
    assert np.allclose(pr.cpu().numpy(), g["projected"], rtol=1e-5, atol=1e-3)


def test_iou3d_golden_bitwise_and_input_mutation(L):
    core, _ = L
    g = load_golden("iou3d")
    for method in ("normal", "generalized"):
        ca, cb = cuda(g["corners_a"]), cuda(g["corners_b"])
        bev, i3d = core.iou3d_approximate(ca, cb, mode="combinations", method=method)
        assert bits_equal(bev.cpu().numpy(), g["bev_comb_" + method])
        assert bits_equal(i3d.cpu().numpy(), g["i3d_comb_" + method])
        ca, cb = cuda(g["corners_a"][:64]), cuda(g["corners_b"])
        bev, i3d = core.iou3d_approximate(ca, cb, mode="list", method=method)
        assert bits_equal(bev.cpu().numpy(), g["bev_list_" + method])
        assert bits_equal(i3d.cpu().numpy(), g["i3d_list_" + method])
    cs = cuda(g["corners_a"])
    _, s = core.iou3d_approximate(cs, cs, mode="combinations", method="generalized")
    assert bits_equal(s.cpu().numpy(), g["i3d_self_generalized"])
    assert bits_equal(cs.cpu().numpy(), g["corners_after_self_call"])      # the reference's in-place Y<-Z quirk
    assert bits_equal(core.get_volume(cuda(g["corners_a"])).cpu().numpy(), g["volume_a"])


def test_iou3d_c3_full_size_bitwise_vs_oracle(L):
    """BASELINE config 3 at full size: N=4096 7-DoF boxes; parity defined from the corners onward."""
    from groomed_nms_b200 import synthetic, ops
    from oracle import groomed_oracle as O
    b7, _ = synthetic.config_c3()
    corners = ops.corners_from_boxes7(cuda(b7))
    cn = corners.cpu().numpy()
    ref_c = O.get_corners_of_cuboid(*[b7[:, i] for i in range(7)])
    assert np.allclose(cn, ref_c, rtol=1e-6, atol=2e-5)
    rec = ops.box3d_records(corners)
    _, got = ops.overlap3d(rec, rec, want_bev=False, want_3d=True, generalized=True, affine=True)
    _, want = O.iou3d_approximate(cn, cn, "combinations", "generalized")
    want = (np.float32(0.5) * (np.float32(1) + want)).astype(np.float32)
    got = got.cpu().numpy()
    assert nbits_diff(got, want) == 0
    assert bits_equal(got, got.T)


MATRIX_VARIANTS = {"auto": (0, 0), "direct": (1, 0), "direct_chunked": (1, 4), "tma": (2, 0), "tma_chunked": (2, 4), "direct_scalar": (1, 0)}


@pytest.mark.parametrize("variant", sorted(MATRIX_VARIANTS))
@pytest.mark.parametrize("kind,n,batch", [("2d", 2051, 1), ("3d", 2052, 1), ("2d", 700, 3), ("3d", 1028, 2), ("3d", 4096, 2), ("2d", 8192, 1),
                                          ("3d", 300, 7), ("3d", 2051, 1)])
def test_batched_self_overlap_tall_tiles_bitwise(kind, n, batch, variant):
    """The batched self-overlap entry points run the matrix-only tile kernels (256 x 64 symmetric tiles, packed fp32x2
    arithmetic for 3D) -- register-direct mirrored stores or shared-memory staging + TMA tensor stores, persistent or
    chunked launches, selected per call through gnms_launch_opts: odd sizes (partial tiles clipped by the tensor map, scalar
    store path when N % 4 != 0), several images per launch (a tall tile must not run into the next image), degenerate
    boxes (exact-division fallback) -- bitwise equal to the oracle, and symmetric."""
    from groomed_nms_b200 import _lib, ops
    from oracle import groomed_oracle as O
    lib = _lib.load()
    mk, tpc = MATRIX_VARIANTS[variant]
    o = _lib.launch_opts(matrix_kernel=mk, tiles_per_cta=tpc, flags=_lib.OPT_SCALAR_MATH if variant.endswith("scalar") else 0)
    rng = np.random.default_rng(n + batch)
    outs, wants = [], []
    guard = 64                                                                       # canary rows after the last image
    if kind == "2d":
        c = rng.uniform(0, 1500, (batch, n, 2)); wh = rng.uniform(5, 200, (batch, n, 2))
        boxes = np.concatenate([c - wh / 2, c + wh / 2], 2).astype(np.float32)
        boxes[:, 7, 2:] = boxes[:, 7, :2]; boxes[:, 11] = boxes[:, 7]          # zero-area twins: 0/0 = NaN
        d = cuda(boxes)
        buf = torch.full((batch * n * n + guard * n,), -3.0, dtype=torch.float32, device="cuda")
        out = buf[:batch * n * n].view(batch, n, n)
        _lib.check(lib.gnms_overlap2d_batched_ex_f32(ops._p(d), n, batch, ops._p(out), _lib.opts_ref(o), ops._stream(d.device)), "overlap2d_batched")
        for b in range(batch if n <= 2100 else 1):
            wants.append(O.iou(boxes[b], boxes[b])); outs.append(out[b].cpu().numpy())
    else:
        b7 = np.stack([rng.uniform(-30, 30, (batch, n)), 1.6 + 0.2 * rng.standard_normal((batch, n)), rng.uniform(5, 70, (batch, n)),
                       1.6 + 0.2 * rng.standard_normal((batch, n)), 1.5 + 0.1 * rng.standard_normal((batch, n)),
                       4 + 0.5 * rng.standard_normal((batch, n)), rng.uniform(-np.pi, np.pi, (batch, n))], 2).astype(np.float32)
        corners = ops.corners_from_boxes7(cuda(b7).view(batch * n, 7))
        rec = ops.box3d_records(corners).view(batch, n, 8).contiguous()
        buf = torch.full((batch * n * n + guard * n,), -3.0, dtype=torch.float32, device="cuda")
        out = buf[:batch * n * n].view(batch, n, n)
        _lib.check(lib.gnms_overlap3d_batched_ex_f32(ops._p(rec), n, batch, ops._p(out), 1, 1, _lib.opts_ref(o), ops._stream(rec.device)), "overlap3d_batched")
        cn = corners.view(batch, n, 3, 8).cpu().numpy()
        for b in range(batch if n <= 2100 else 1):
            g3 = O.iou3d_approximate(cn[b], cn[b], "combinations", "generalized")[1]
            wants.append((np.float32(0.5) * (np.float32(1) + g3)).astype(np.float32)); outs.append(out[b].cpu().numpy())
    torch.cuda.synchronize()
    for got, want in zip(outs, wants):
        assert bits_equal(got, want)
    for b in range(batch):                                                            # every image is its own mirror image
        img = out[b]
        assert torch.equal(torch.nan_to_num(img, nan=-7.0), torch.nan_to_num(img.t(), nan=-7.0))
    assert bool((buf[batch * n * n:] == -3.0).all())                                  # nothing written past the last image


def test_launch_opts_are_validated_and_versioned():
    """gnms_launch_opts: unknown enumerators are rejected, a shorter (older) struct is accepted (struct_size versioning)."""
    import ctypes
    from groomed_nms_b200 import _lib, ops
    lib = _lib.load()
    rec = torch.zeros((1, 64, 8), device="cuda"); out = torch.empty((1, 64, 64), device="cuda")
    bad = _lib.launch_opts(matrix_kernel=7)
    assert lib.gnms_overlap3d_batched_ex_f32(ops._p(rec), 64, 1, ops._p(out), 1, 1, ctypes.byref(bad), None) == -1
    short = _lib.launch_opts(matrix_kernel=1, tiles_per_cta=4)
    short.struct_size = 12                                                            # only matrix_kernel and tiles_per_cta are covered
    short.rank_method = 99                                                            # beyond struct_size: must be ignored
    assert lib.gnms_overlap3d_batched_ex_f32(ops._p(rec), 64, 1, ops._p(out), 1, 1, ctypes.byref(short), None) == 0
    torch.cuda.synchronize()


def test_corners_kitti_vertex_order_gpu():
    """get_corners_of_cuboid(..., iou_3d_convention=False) (lib/math_3d.py:405-426): the reference's own output for N = 4
    (golden), the oracle for any N, torch and numpy containers."""
    from conftest import load_golden
    from groomed_nms_b200.lib import math_3d as M
    from oracle import groomed_oracle as O
    g = load_golden("corners_kitti_order")
    b = g["boxes7"]
    got = M.get_corners_of_cuboid(*[cuda(b[:, i]) for i in range(7)], iou_3d_convention=False)
    assert np.allclose(got.cpu().numpy(), g["corners"], rtol=1e-6, atol=1e-5)
    rng = np.random.default_rng(5)
    b = np.stack([rng.uniform(-20, 20, 333), rng.uniform(0.5, 2, 333), rng.uniform(5, 60, 333), rng.uniform(1.4, 1.9, 333),
                  rng.uniform(1.3, 1.8, 333), rng.uniform(3, 5, 333), rng.uniform(-np.pi, np.pi, 333)], 1).astype(np.float32)
    want = O.get_corners_of_cuboid(*[b[:, i] for i in range(7)], iou_3d_convention=False)
    got_np = M.get_corners_of_cuboid(*[b[:, i] for i in range(7)], iou_3d_convention=False)
    assert isinstance(got_np, np.ndarray) and np.allclose(got_np, want, rtol=1e-6, atol=1e-5)


# ------------------------------------------------------------------------------------------------ autograd of the 3D feeders
def _rel(a, b):
    return np.abs(a - b).max() / max(1e-12, np.abs(b).max())


def test_corners_autograd_matches_reference(L):
    """get_corners_of_cuboid backward (gnms_corners_backward_f32) vs the reference's autograd (golden grad3d.npz)."""
    _, m3d = L
    g = load_golden("grad3d")
    t = [cuda(g["corners_boxes7"][:, i].copy()).requires_grad_(True) for i in range(7)]
    c = m3d.get_corners_of_cuboid(*t)
    assert c.requires_grad and np.allclose(c.detach().cpu().numpy(), g["corners_out"], rtol=1e-6, atol=1e-5)
    (c * cuda(g["corners_up"])).sum().backward()
    got = np.stack([v.grad.cpu().numpy() for v in t], 1)
    assert _rel(got, g["corners_grad"]) < 1e-6
    # the other vertex order: same kernel, other sign table -- against the oracle's VJP
    from oracle import groomed_oracle as O
    t = [cuda(g["corners_boxes7"][:, i].copy()).requires_grad_(True) for i in range(7)]
    (m3d.get_corners_of_cuboid(*t, iou_3d_convention=False) * cuda(g["corners_up"])).sum().backward()
    want = O.get_corners_of_cuboid_backward(g["corners_boxes7"], g["corners_up"], iou_3d_convention=False)
    assert _rel(np.stack([v.grad.cpu().numpy() for v in t], 1), want) < 1e-6


@pytest.mark.parametrize("mode", ["list", "combinations"])
@pytest.mark.parametrize("method", ["normal", "generalized"])
def test_iou3d_approximate_autograd_matches_reference(L, mode, method):
    """iou3d_approximate backward (gnms_iou3d_approx_backward_f32) wrt both corner sets vs the reference's autograd."""
    core, _ = L
    g = load_golden("grad3d")
    k = "pairs_%s_%s_" % (mode, method)
    la = cuda(g["pairs_a"]).requires_grad_(True)
    lb = cuda(g["pairs_b"][:64 if mode == "combinations" else 96]).requires_grad_(True)
    with pytest.raises(RuntimeError):                                       # a leaf that requires grad: torch refuses the in-place
        core.iou3d_approximate(la, lb, mode=mode, method=method)            # Y <- Z write, exactly as under the reference
    ca, cb = la * 1.0, lb * 1.0
    bev, i3d = core.iou3d_approximate(ca, cb, mode=mode, method=method)
    assert bits_equal(bev.detach().cpu().numpy(), g[k + "bev"]) and bits_equal(i3d.detach().cpu().numpy(), g[k + "3d"])
    assert torch.equal(ca[:, 1], ca[:, 2]) and torch.equal(cb[:, 1], cb[:, 2])            # the reference's mutation (lib/core.py:379-380)
    ((bev * cuda(g[k + "up_bev"])).sum() + (i3d * cuda(g[k + "up_3d"])).sum()).backward()
    assert _rel(la.grad.cpu().numpy(), g[k + "grad_a"]) < 1e-5 and _rel(lb.grad.cpu().numpy(), g[k + "grad_b"]) < 1e-5


def test_acceptance_target_chain_autograd_matches_reference(L):
    """parameters -> corners -> iou3d_approximate(list) as lib/loss/rpn_3d.py:663-679 differentiates it, real cuboids (every
    per-box min / max is a tie), and the NMS-style self call whose mutated corners are used afterwards."""
    core, m3d = L
    g = load_golden("grad3d")
    ta = [cuda(g["chain_a"][:, i].copy()).requires_grad_(True) for i in range(7)]
    tb = [cuda(g["chain_b"][:, i].copy()).requires_grad_(True) for i in range(7)]
    _, i3d = core.iou3d_approximate(m3d.get_corners_of_cuboid(*ta), m3d.get_corners_of_cuboid(*tb))
    assert np.allclose(i3d.detach().cpu().numpy(), g["chain_3d"], rtol=1e-4, atol=1e-6)
    (i3d * cuda(g["chain_up"])).sum().backward()
    assert _rel(np.stack([v.grad.cpu().numpy() for v in ta], 1), g["chain_grad_a"]) < 1e-4
    assert _rel(np.stack([v.grad.cpu().numpy() for v in tb], 1), g["chain_grad_b"]) < 1e-4
    ts = [cuda(g["corners_boxes7"][:, i].copy()).requires_grad_(True) for i in range(7)]
    cs = m3d.get_corners_of_cuboid(*ts)
    bev, i3d = core.iou3d_approximate(cs, cs, mode="combinations", method="generalized")
    assert np.allclose(cs.detach().cpu().numpy(), g["self_corners_after"], rtol=1e-6, atol=1e-5)
    ((bev * cuda(g["self_up_bev"])).sum() + (i3d * cuda(g["self_up_3d"])).sum() + (cs * cuda(g["corners_up"])).sum()).backward()
    assert _rel(np.stack([v.grad.cpu().numpy() for v in ts], 1), g["self_grad"]) < 1e-4


def test_iou2d_float64_and_mixed_dtypes_match_reference(L):
    """float64 (and mixed float32 / float64) inputs are computed in float64 like the reference's type promotion does
    (lib/rpn_util.py:448, :1295): bit-exact against golden iou2d_f64.npz, numpy and torch containers."""
    core, _ = L
    g = load_golden("iou2d_f64")

    def same(got, want):
        got = got.cpu().numpy() if isinstance(got, torch.Tensor) else got
        return got.dtype == np.float64 and got.shape == want.shape and np.array_equal(np.isnan(got), np.isnan(want)) \
            and np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])
    for ka, kb in (("a64", "b64"), ("a32", "b64"), ("a64", "b32")):
        for mode in ("combinations", "list"):
            assert same(core.iou(g[ka], g[kb], mode=mode), g["iou_%s_%s_%s" % (ka, kb, mode)]), (ka, kb, mode)
            assert same(core.intersect(g[ka], g[kb], mode=mode), g["inter_%s_%s_%s" % (ka, kb, mode)]), (ka, kb, mode)
            ta, tb = torch.from_numpy(g[ka]).cuda(), torch.from_numpy(g[kb]).cuda()
            assert same(core.iou(ta, tb, mode=mode), g["iou_%s_%s_%s" % (ka, kb, mode)])
    # a [N,5] detections array sliced like lib/rpn_util.py:1295 (non-contiguous rows)
    dets = np.concatenate([g["a64"], np.ones((57, 1))], 1)
    assert same(core.iou(dets[:, 0:4], dets[:, 0:4]), core.iou(g["a64"], g["a64"]))
