"""GPU parity: GrooMeD-NMS forward + analytic backward (through the reference-named API and the C-ABI) against
the reference's golden vectors and the oracle.  Bar (BASELINE.json north_star): group / keep indices bit-exact,
rescored scores and gradients within 1e-5 relative fp32."""
import ast

import numpy as np
import pytest
import torch

from conftest import load_golden, unflatten_groups
from gpu_util import bits_equal, cuda

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(scope="module")
def G():
    from groomed_nms_b200.lib import groomed_nms
    return groomed_nms


def check_against(o, prob, valid, invalid, cond=1.0):
    """o: oracle forward dict; prob/valid/invalid: numpy from the CUDA path."""
    assert np.allclose(prob, o["prob"], rtol=RTOL, atol=1e-6 * cond), np.abs(prob - o["prob"]).max()
    thr = o["cfg"]["valid_box_prob_threshold"]
    vals = o["r_thr"][o["r_thr"] >= thr]
    if len(np.unique(vals)) == len(vals) and cond == 1.0:
        assert list(valid) == list(o["valid"])
    else:
        assert set(valid) == set(o["valid"])
    assert set(invalid) == set(o["invalid"])
    assert len(valid) + len(invalid) == len(o["prob"])


def test_kats_of_the_reference_tests(G):
    g = load_golden("kat_forward")
    for n in ("4", "5"):
        v, i, p = G.differentiable_nms(cuda(g["s" + n]), cuda(g["iou" + n]), nms_threshold=0.4, temperature=0.1,
                                       valid_box_prob_threshold=0.3, pruning_method="linear", sorting_method="hard",
                                       return_sorted_prob=False, group_boxes="True", debug=False)
        assert np.allclose(p.cpu().numpy(), g["printed" + n], atol=5e-4)
        assert np.array_equal(p.cpu().numpy(), g["prob" + n])
        assert v.dtype == torch.int64 and list(v.cpu().numpy()) == list(g["valid" + n])
        assert set(i.cpu().numpy()) == set(g["invalid" + n])


def test_seeded_5box_keep_matches_every_nms(G):
    from groomed_nms_b200.lib import core
    from groomed_nms_b200.lib.nms.gpu_nms import gpu_nms
    from groomed_nms_b200.lib.nms_others import girshick_nms, navneeth_soft_nms
    g = load_golden("seeded_5box")
    ab = g["aboxes"]
    iou = core.iou(cuda(ab[:, :4]), cuda(ab[:, :4]))
    assert bits_equal(iou.cpu().numpy(), g["iou"])
    v, i, p = G.differentiable_nms(cuda(ab[:, 4]), iou, nms_threshold=0.4, temperature=0.1, valid_box_prob_threshold=0.3)
    assert list(v.cpu().numpy()) == [1, 0, 4] == list(g["keep_ours"])
    assert np.array_equal(p.cpu().numpy(), g["prob_ours"])
    assert gpu_nms(ab.astype(np.float32), 0.4, device_id=0) == [1, 0, 4]
    assert girshick_nms(ab, 0.4, shift=1) == list(g["keep_girshick"])
    assert navneeth_soft_nms(ab.copy(), Nt=0.4, shift=1).tolist() == list(g["keep_soft"])
    # numpy in -> CPU tensors out, like the reference (lib/groomed_nms.py:34-36)
    v2, _, p2 = G.differentiable_nms(ab[:, 4], g["iou"], nms_threshold=0.4, temperature=0.1)
    assert not p2.is_cuda and list(v2.numpy()) == [1, 0, 4]


def test_get_groups_reference_case_and_golden(G):
    g = load_golden("get_groups_10")
    got = G.get_groups(scores_unsorted=cuda(g["scores"]), iou_unsorted=cuda(g["iou"]), group_threshold=0.4)
    assert [x.tolist() for x in got] == [[0, 1, 4, 5, 6, 7, 9], [2, 3, 8]]
    assert all(x.dtype == torch.int64 for x in got)
    gg = load_golden("groups")
    inp = load_golden("dnms_inputs")
    for tag, (sc, io) in dict(g240=(inp["scores240"], inp["iou240"]), g96=(inp["scores96"], inp["iou96"]),
                              gns=(inp["s_ns"], inp["iou_ns"])).items():
        for gs in (1, 7, 100):
            want = unflatten_groups(gg["%s_gs%d_flat" % (tag, gs)], gg["%s_gs%d_len" % (tag, gs)])
            got = G.get_groups(cuda(io), 0.4, cuda(sc), group_size=gs)
            assert [x.tolist() for x in got] == [list(x) for x in want]


def _cases():
    g = load_golden("dnms_cases")
    return sorted({k.split("__")[0] for k in g.files})


@pytest.mark.parametrize("tag", _cases())
def test_golden_forward_backward(G, tag):
    """Every differentiable_nms case the real reference produced (modes A/B/C x pruning x group sizes)."""
    from oracle import groomed_oracle as O
    g = load_golden("dnms_cases")
    inp = load_golden("dnms_inputs")
    cfg = dict(ast.literal_eval(str(g[tag + "__cfg"])))
    if "240" in tag:
        sc, io, up = inp["scores240"], inp["iou240"], inp["g240"]
    elif "120ns" in tag:
        sc, io, up = inp["s_ns"], inp["iou_ns"], inp["g_ns"]
    else:
        sc, io, up = inp["scores96"], inp["iou96"], inp["g96"]
    need_gi = (tag + "__grad_iou") in g.files
    s = cuda(sc).requires_grad_(True)
    m = cuda(io).requires_grad_(need_gi)
    v, i, p = G.differentiable_nms(s, m, **cfg)
    o = O.differentiable_nms(sc, io, **cfg)
    cond = 1.0 if o["T"] is None else max(1.0, float(np.abs(o["T"]).max()))
    want = g[tag + "__prob"]
    assert np.allclose(p.detach().cpu().numpy(), want, rtol=RTOL, atol=1e-6 * cond)
    check_against(o, p.detach().cpu().numpy(), v.cpu().numpy(), i.cpu().numpy(), cond)
    assert set(v.cpu().numpy()) == set(g[tag + "__valid"]) and set(i.cpu().numpy()) == set(g[tag + "__invalid"])
    p.backward(cuda(up))
    wgs = g[tag + "__grad_scores"]
    sc_ = max(1.0, np.abs(wgs).max())
    assert np.allclose(s.grad.cpu().numpy(), wgs, rtol=RTOL, atol=2e-6 * sc_ * cond), np.abs(s.grad.cpu().numpy() - wgs).max()
    if need_gi:
        wgi = g[tag + "__grad_iou"]
        sc_ = max(1.0, np.abs(wgi).max())
        assert np.allclose(m.grad.cpu().numpy(), wgi, rtol=RTOL, atol=2e-6 * sc_ * cond)


def _rand_case(seed, n, k, nonsym=False):
    from groomed_nms_b200 import synthetic
    from oracle import groomed_oracle as O
    rng = np.random.default_rng(seed)
    if nonsym:
        iou = (rng.uniform(0, 1, (n, n)) ** 4).astype(np.float32)
        np.fill_diagonal(iou, 1.0)
        sc = synthetic.distinct_scores(rng, n, 0.05, 1.0)
    else:
        boxes, sc, _ = synthetic.clustered_boxes_2d(n, k, seed=seed, jitter=0.08)
        iou = O.iou(boxes, boxes)
    return sc, iou, rng.standard_normal(n).astype(np.float32)


@pytest.mark.parametrize("n,k,nonsym", [(1, 1, False), (2, 1, False), (33, 2, False), (200, 4, False), (500, 9, False),
                                        (257, 1, True), (1000, 20, False), (64, 64, False)])
@pytest.mark.parametrize("pm,temp", [("linear", 0.1), ("sigmoidal", 0.1), ("soft_nms", 0.5)])
@pytest.mark.parametrize("gs", [1, 20, 100, 10 ** 9])
def test_mode_a_random_vs_oracle(G, n, k, nonsym, pm, temp, gs):
    from oracle import groomed_oracle as O
    sc, iou, up = _rand_case(1000 + n, n, k, nonsym)
    cfg = dict(nms_threshold=0.4, pruning_method=pm, temperature=temp, valid_box_prob_threshold=0.3, group_size=gs)
    o = O.differentiable_nms(sc, iou, dense=False, **cfg)
    s = cuda(sc).requires_grad_(True)
    m = cuda(iou).requires_grad_(n <= 257)
    v, i, p = G.differentiable_nms(s, m, **cfg)
    check_against(o, p.detach().cpu().numpy(), v.cpu().numpy(), i.cpu().numpy())
    p.backward(cuda(up))
    gs_, gi_ = O.differentiable_nms_backward(o, up, need_grad_iou=(n <= 257))
    assert np.allclose(s.grad.cpu().numpy(), gs_, rtol=RTOL, atol=2e-6 * max(1.0, np.abs(gs_).max()))
    if n <= 257:
        assert np.allclose(m.grad.cpu().numpy(), gi_, rtol=RTOL, atol=2e-6 * max(1.0, np.abs(gi_).max()))


def test_saved_state_groups_are_bit_exact():
    """lead[] (group of every box) and order[] from the C-ABI state vs the oracle, incl. the group_size cap."""
    from groomed_nms_b200 import ops
    from oracle import groomed_oracle as O
    for seed, n, k, gs in [(5, 700, 6, 100), (6, 700, 6, 15), (7, 1500, 3, 100), (8, 300, 300, 5)]:
        sc, iou, _ = _rand_case(seed, n, k)
        o = O.differentiable_nms(sc, iou, group_size=gs, dense=False)
        st = ops.forward_matrix(cuda(sc)[None], cuda(iou)[None], ops.make_params(group_size=gs))
        assert np.array_equal(st.order[0].cpu().numpy(), o["order"])
        assert np.array_equal(st.lead[0].cpu().numpy(), o["lead"])
        assert np.array_equal(st.pre[0].cpu().numpy(), o["pre"])


def test_nan_overlap_leaves_the_pool_silently(G):
    """NaN compares false on both `>` and `<=` (lib/groomed_nms.py:249-250): the box joins no group, r = 0."""
    from oracle import groomed_oracle as O
    sc, iou, _ = _rand_case(77, 120, 3)
    order = np.argsort(-sc)
    iou = iou.copy()
    iou[order[5], order[0]] = np.nan
    iou[order[40], order[1]] = np.nan
    o = O.differentiable_nms(sc, iou, dense=False)
    v, i, p = G.differentiable_nms(cuda(sc), cuda(iou))
    check_against(o, p.cpu().numpy(), v.cpu().numpy(), i.cpu().numpy())
    assert p[5].item() == 0.0


@pytest.mark.parametrize("box_kind", ["2d", "3d"])
def test_matrix_free_forward_equals_matrix_path(box_kind):
    """gnms_forward_boxes_f32 (overlaps on the fly, nothing N^2 in HBM) == overlap kernel + gnms_forward_f32."""
    from groomed_nms_b200 import synthetic, ops, _lib
    if box_kind == "2d":
        boxes, sc, _ = synthetic.clustered_boxes_2d(1500, 7, seed=9, jitter=0.06)
        bx = cuda(boxes)
        iou = ops.overlap2d(bx, bx)
        p = ops.make_params(group_size=50)
        ov_out = torch.empty(1, 1500, 1500, device="cuda")
        st2 = ops.forward_boxes(cuda(sc)[None], bx[None], _lib.BOX_2D, p, overlap_out=ov_out)
    else:
        b7, sc = synthetic.config_c3(seed=4, n=2000, k=16)
        rec = ops.box3d_records(ops.corners_from_boxes7(cuda(b7)))
        _, iou = ops.overlap3d(rec, rec, False, True, generalized=True, affine=True)
        p = ops.make_params(group_size=50)
        ov_out = torch.empty(1, 2000, 2000, device="cuda")
        st2 = ops.forward_boxes(cuda(sc)[None], rec[None], _lib.BOX_3D_REC, p, generalized=True, affine=True,
                                overlap_out=ov_out)
    st1 = ops.forward_matrix(cuda(sc)[None], iou[None], p)
    torch.cuda.synchronize()
    assert torch.equal(ov_out[0], iou)          # the tile kernel's matrix output is bitwise the overlap kernel's
    for f in ("prob", "order", "lead", "pre", "pval", "counts"):
        assert torch.equal(getattr(st1, f), getattr(st2, f)), f
    nv = int(st1.counts[0, 0])
    assert torch.equal(st1.valid_idx[0, :nv], st2.valid_idx[0, :nv])
    g = torch.randn(1, sc.shape[0], device="cuda")
    g1, _ = ops.backward(st1, g)
    g2, _ = ops.backward(st2, g)
    assert torch.allclose(g1, g2, rtol=1e-6, atol=1e-7)


def test_batched_ragged_equals_per_image():
    """[B,N] batch with n_per_image (padding) == B independent calls."""
    from groomed_nms_b200 import ops
    B, N = 5, 640
    ns = [640, 1, 333, 0, 500]
    sc = np.zeros((B, N), np.float32); iou = np.zeros((B, N, N), np.float32)
    singles = []
    for b, n in enumerate(ns):
        if n:
            s, m, _ = _rand_case(300 + b, n, 5)
            sc[b, :n] = s; iou[b, :n, :n] = m
        singles.append(n)
    p = ops.make_params(group_size=30)
    npi = torch.tensor(ns, dtype=torch.int32, device="cuda")
    st = ops.forward_matrix(cuda(sc), cuda(iou), p, n_per_image=npi)
    g = torch.randn(B, N, device="cuda")
    gs, _ = ops.backward(st, g)
    for b, n in enumerate(ns):
        assert st.counts[b].sum().item() == n
        if n == 0:
            continue
        s1 = ops.forward_matrix(cuda(sc[b:b + 1, :n]), cuda(iou[b:b + 1, :n, :n]), p)
        assert torch.equal(s1.prob[0], st.prob[b, :n])
        nv = int(s1.counts[0, 0])
        assert torch.equal(s1.valid_idx[0, :nv], st.valid_idx[b, :nv])
        g1, _ = ops.backward(s1, g[b:b + 1, :n].contiguous())
        assert torch.allclose(g1[0], gs[b, :n], rtol=1e-6, atol=1e-7)
        assert (gs[b, n:] == 0).all() and (st.prob[b, n:] == 0).all()


@pytest.mark.parametrize("N,ns", [(640, [640, 1, 333, 0, 500]), (4096, [4096, 4000]), (1100, [1100, 1024, 1025]), (8192, [8192])])
@pytest.mark.parametrize("src", ["matrix", "boxes2d"])
def test_rank_by_radix_sort_equals_rank_by_counting(N, ns, src):
    """The per-image radix sort (chosen for large batches) and the counting kernel give the same stable order, also with
    tied scores, ragged batches and the padded tail: every output of the forward is identical."""
    from groomed_nms_b200 import _lib, ops, synthetic
    if src == "matrix" and N > 2048:
        pytest.skip("matrix inputs of that size are covered by the box path")
    lib = _lib.load()
    B = len(ns)
    rng = np.random.default_rng(N)
    sc = np.zeros((B, N), np.float32)
    boxes = np.zeros((B, N, 4), np.float32)
    for b in range(B):
        bx, s, _ = synthetic.clustered_boxes_2d(N, 7, seed=N + b, jitter=0.08)
        s[rng.integers(0, N, N // 3)] = s[rng.integers(0, N, N // 3)]          # many exact ties
        s[rng.integers(0, N, 5)] = -0.0
        sc[b], boxes[b] = s, bx
    npi = torch.tensor(ns, dtype=torch.int32, device="cuda")
    p = ops.make_params(group_size=40)
    outs = []
    for method in (_lib.RANK_COUNT, _lib.RANK_SORT):
        o = _lib.launch_opts(rank_method=method)
        if src == "matrix":
            iou = torch.stack([ops.overlap2d(cuda(boxes[b]), cuda(boxes[b])) for b in range(B)])
            st = ops.forward_matrix(cuda(sc), iou, p, n_per_image=npi, opts=o)
        else:
            st = ops.forward_boxes(cuda(sc), cuda(boxes), _lib.BOX_2D, p, n_per_image=npi, opts=o)
        torch.cuda.synchronize()
        outs.append(st)
    a, c = outs
    assert torch.equal(a.order, c.order) and torch.equal(a.counts, c.counts) and torch.equal(a.lead, c.lead)
    assert torch.equal(a.prob, c.prob) and torch.equal(a.sorted_scores, c.sorted_scores)
    for b, n in enumerate(ns):
        nv = int(a.counts[b, 0])
        assert torch.equal(a.valid_idx[b, :nv], c.valid_idx[b, :nv])
        # stable: ties keep the lower input index first
        o = a.order[b, :n].cpu().numpy()
        want = np.argsort(-sc[b, :n], kind="stable")
        assert np.array_equal(o, want)


def test_config_c1_one_group(G):
    from groomed_nms_b200 import synthetic
    from groomed_nms_b200.lib import core
    from oracle import groomed_oracle as O
    boxes, sc = synthetic.config_c1()
    iou = core.iou(cuda(boxes), cuda(boxes))
    assert len(O.get_groups(iou.cpu().numpy(), 0.4, sc)) == 1
    v, i, p = G.differentiable_nms(cuda(sc), iou, nms_threshold=0.4, temperature=0.1, valid_box_prob_threshold=0.3)
    o = O.differentiable_nms(sc, iou.cpu().numpy(), temperature=0.1)
    check_against(o, p.cpu().numpy(), v.cpu().numpy(), i.cpu().numpy())
    assert len(G.get_groups(iou, 0.4, cuda(sc))) == 1


def test_config_c2_four_groups_fwd_bwd(G):
    from groomed_nms_b200 import synthetic
    from groomed_nms_b200.lib import core
    from oracle import groomed_oracle as O
    boxes, sc = synthetic.config_c2()
    iou = core.iou(cuda(boxes), cuda(boxes))
    iou_np = iou.cpu().numpy()
    assert bits_equal(iou_np, O.iou(boxes, boxes))
    assert len(O.get_groups(iou_np, 0.4, sc)) == 4
    up = np.random.default_rng(2).standard_normal(1024).astype(np.float32)
    for gs in (100, 10 ** 9):
        s = cuda(sc).requires_grad_(True)
        v, i, p = G.differentiable_nms(s, iou.clone().detach(), group_size=gs)
        o = O.differentiable_nms(sc, iou_np, group_size=gs, dense=False)
        check_against(o, p.detach().cpu().numpy(), v.cpu().numpy(), i.cpu().numpy())
        assert int((o["lead"] >= 0).sum()) == (404 if gs == 100 else 1024)
        p.backward(cuda(up))
        want, _ = O.differentiable_nms_backward(o, up, need_grad_iou=False)
        assert np.allclose(s.grad.cpu().numpy(), want, rtol=RTOL, atol=2e-6 * np.abs(want).max())


def test_config_c3_full_size_fwd_bwd():
    """BASELINE headline config: N=4096 7-DoF boxes, overlap = 0.5*(1+giou3d); matrix path and matrix-free path
    against the oracle (fed with the CUDA corners: parity is defined from the corners onward)."""
    from groomed_nms_b200 import synthetic, ops, _lib
    from oracle import groomed_oracle as O
    b7, sc = synthetic.config_c3()
    corners = ops.corners_from_boxes7(cuda(b7))
    rec = ops.box3d_records(corners)
    _, ov = ops.overlap3d(rec, rec, False, True, generalized=True, affine=True)
    cn = corners.cpu().numpy()
    _, want_ov = O.iou3d_approximate(cn, cn, "combinations", "generalized")
    want_ov = (np.float32(0.5) * (np.float32(1) + want_ov)).astype(np.float32)
    assert bits_equal(ov.cpu().numpy(), want_ov)
    o = O.differentiable_nms(sc, want_ov, dense=False)
    up = np.random.default_rng(5).standard_normal(4096).astype(np.float32)
    want_g, _ = O.differentiable_nms_backward(o, up, need_grad_iou=False)
    p = ops.make_params()
    for st in (ops.forward_matrix(cuda(sc)[None], ov[None], p),
               ops.forward_boxes(cuda(sc)[None], rec[None], _lib.BOX_3D_REC, p, generalized=True, affine=True)):
        nv, ni = st.counts[0].tolist()
        check_against(o, st.prob[0].cpu().numpy(), st.valid_idx[0, :nv].cpu().numpy(), st.invalid_idx[0, :ni].cpu().numpy())
        assert np.array_equal(st.lead[0].cpu().numpy(), o["lead"])
        gs, _ = ops.backward(st, cuda(up)[None])
        assert np.allclose(gs[0].cpu().numpy(), want_g, rtol=RTOL, atol=2e-6 * np.abs(want_g).max())
        # size-independent properties: leaders keep their score, members never gain, counts add up
        lead = st.lead[0].cpu().numpy(); r = st.prob[0].cpu().numpy(); ss = st.sorted_scores[0].cpu().numpy()
        L = lead == np.arange(4096)
        assert np.array_equal(r[L], np.clip(ss[L], 0, 1)) and (r <= ss + 1e-7).all() and (r[lead < 0] == 0).all()
        assert nv + ni == 4096


def test_idempotent_and_deterministic_forward():
    from groomed_nms_b200 import ops
    sc, iou, _ = _rand_case(11, 2048, 16)
    p = ops.make_params()
    a = ops.forward_matrix(cuda(sc)[None], cuda(iou)[None], p, private_ws=True)
    b = ops.forward_matrix(cuda(sc)[None], cuda(iou)[None], p, private_ws=True)
    for f in ("prob", "valid_idx", "counts", "lead", "pre"):
        assert torch.equal(getattr(a, f)[..., :int(a.counts[0, 0])] if f == "valid_idx" else getattr(a, f),
                           getattr(b, f)[..., :int(a.counts[0, 0])] if f == "valid_idx" else getattr(b, f))


def test_errors(G):
    s, m = torch.rand(4, device="cuda"), torch.eye(4, device="cuda")
    with pytest.raises(NotImplementedError):
        G.differentiable_nms(s, m, pruning_method="bogus")
    with pytest.raises(NotImplementedError):
        G.pruning_function(m, pruning_method="bogus")
    with pytest.raises(RuntimeError):
        G.differentiable_nms(torch.rand(9000, device="cuda"), torch.zeros(1, 1, device="cuda").expand(9000, 9000))


def test_pruning_function_and_indices_copy(G):
    from oracle import groomed_oracle as O
    x = np.random.default_rng(1).uniform(0, 1, (50, 50)).astype(np.float32)
    for pm, t in (("linear", 0.1), ("sigmoidal", 0.1), ("soft_nms", 0.5)):
        got = G.pruning_function(cuda(x), 0.4, t, pm).cpu().numpy()
        assert np.allclose(got, O.pruning_function(x, 0.4, t, pm), rtol=1e-5, atol=1e-6)
        assert isinstance(G.pruning_function(x, 0.4, t, pm), np.ndarray)
    A = torch.zeros(5, 5, device="cuda"); B = torch.rand(3, 3, device="cuda")
    ind = torch.tensor([1, 2, 4], device="cuda")
    out = G.indices_copy(A, B, ind)
    want = torch.zeros(5, 5, device="cuda"); want[ind.unsqueeze(1), ind.unsqueeze(0)] = B
    assert torch.equal(out, want) and torch.equal(A, want)
    A2 = torch.zeros(3, 4, device="cuda")
    out2 = G.indices_copy(A2, B, torch.tensor([[0, 0], [0, 1], [1, 1]], device="cuda"),
                          torch.tensor([[1, 1], [2, 1], [2, 2]], device="cuda"), inplace=False)
    assert out2[0, 0] == B[1, 1] and out2[0, 1] == B[2, 1] and out2[1, 1] == B[2, 2] and A2.abs().sum() == 0


@pytest.mark.parametrize("mode", ["B", "C"])
@pytest.mark.parametrize("n,k,gs", [(150, 3, 20), (400, 6, 100), (64, 64, 5), (333, 1, 10 ** 9)])
def test_inverse_modes_random_vs_oracle(G, mode, n, k, gs):
    """group_boxes/mask_group_boxes variants that need (I+Phi)^-1: forward substitution vs the oracle's fp64 inverse."""
    from oracle import groomed_oracle as O
    sc, iou, up = _rand_case(4000 + n, n, k)
    cfg = dict(nms_threshold=0.4, pruning_method="linear", temperature=0.1, valid_box_prob_threshold=0.3, group_size=gs,
               group_boxes=(mode == "B"), mask_group_boxes=False)
    o = O.differentiable_nms(sc, iou, **cfg)
    cond = max(1.0, float(np.abs(o["T"]).max()))
    s = cuda(sc).requires_grad_(True)
    m = cuda(iou).requires_grad_(True)
    v, i, p = G.differentiable_nms(s, m, **cfg)
    check_against(o, p.detach().cpu().numpy(), v.cpu().numpy(), i.cpu().numpy(), cond if cond > 1.5 else 1.0 + 1e-9)
    p.backward(cuda(up))
    gs_, gi_ = O.differentiable_nms_backward(o, up, need_grad_iou=True)
    assert np.allclose(s.grad.cpu().numpy(), gs_, rtol=RTOL, atol=2e-6 * cond * max(1.0, np.abs(gs_).max()))
    assert np.allclose(m.grad.cpu().numpy(), gi_, rtol=RTOL, atol=2e-6 * cond * max(1.0, np.abs(gi_).max()))


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("n,k", [(700, 5), (3000, 40)])
def test_inverse_modes_from_boxes_equal_matrix_path(mode, n, k):
    """Modes GROUP_NOMASK / NOGROUP from boxes: the grouping of mode B comes from the direct leader election (or, forced, from
    the all-pairs bitmask) and the triangular solves gather their overlaps on the fly -- all must equal the matrix path; with
    and without the overlap matrix as an extra output."""
    from groomed_nms_b200 import synthetic, ops, _lib
    boxes, sc, _ = synthetic.clustered_boxes_2d(n, k, seed=19, jitter=0.06)
    bx = cuda(boxes)
    iou = ops.overlap2d(bx, bx)
    p = ops.make_params(group_size=30, group_boxes=(mode == 1), mask_group_boxes=False)
    st1 = ops.forward_matrix(cuda(sc)[None], iou[None], p)
    g = torch.randn(1, n, device="cuda")
    g1, _ = ops.backward(st1, g)
    for election, flags in ((_lib.ELECT_BATCHED, 0), (_lib.ELECT_DIRECT, 0), (_lib.ELECT_MASK, 0), (_lib.ELECT_MASK, _lib.OPT_ONE_PASS),
                            (_lib.ELECT_MASK, _lib.OPT_ONE_PASS | _lib.OPT_INLINE_HITS)):
        for want_matrix in (False, True):
            ov = torch.empty((1, n, n), device="cuda") if want_matrix else None
            st2 = ops.forward_boxes(cuda(sc)[None], bx[None], _lib.BOX_2D, p, overlap_out=ov, opts=_lib.launch_opts(election=election, flags=flags))
            for f in ("prob", "lead", "pre", "counts"):
                assert torch.equal(getattr(st1, f), getattr(st2, f)), (f, election, want_matrix)
            if want_matrix:
                assert torch.equal(ov[0].view(torch.int32), iou.view(torch.int32))
            g2, _ = ops.backward(st2, g)
            assert torch.equal(g1, g2)


def test_soft_sort_path_runs_and_matches_torch_composite(G):
    """sorting_method="soft" (lib/groomed_nms.py:131-165): the soft permutation is applied to the ROWS of the overlap
    matrix only (:164), then the usual prune / group / rescore runs on (soft scores, row-permuted matrix).  The
    composite is checked against the oracle fed with exactly those soft-sorted inputs."""
    from oracle import groomed_oracle as O
    sc, iou, up = _rand_case(91, 60, 3)
    # (inputs already in score order: with unsorted inputs the row-only permutation leaves a diagonal that is not a
    #  self-overlap and the reference's own grouping loop never terminates, lib/groomed_nms.py:247-262)
    pre_order = np.argsort(-sc, kind="stable")
    sc, iou = sc[pre_order].copy(), iou[pre_order][:, pre_order].copy()
    s = cuda(sc).requires_grad_(True)
    v, i, p = G.differentiable_nms(s, cuda(iou), temperature=0.1, sorting_method="soft", sorting_temperature=1e-4)
    ss, perm, ms = G.soft_sort(cuda(sc), full_matrix=cuda(iou), temperature=1e-4)
    # (the reference divides by the row sums broadcast along the LAST axis, :155, so rows are not renormalised:
    #  soft scores are only approximately the sorted scores -- reproduced as written)
    assert torch.allclose(ss, torch.sort(cuda(sc), descending=True)[0], rtol=2e-2) and perm.shape == (60, 60)
    o = O.differentiable_nms(ss.cpu().numpy(), ms.cpu().numpy(), temperature=0.1)
    assert np.allclose(p.detach().cpu().numpy(), o["prob"], rtol=1e-5, atol=1e-6)
    hard_order = np.argsort(-sc, kind="stable")
    assert list(v.cpu().numpy()) == list(hard_order[o["valid"]])     # indices map back through the HARD sort (:118)
    p.sum().backward()
    assert torch.isfinite(s.grad).all() and s.grad.abs().sum() > 0


@pytest.mark.parametrize("kind,n,seed", [("2d", 3000, 1), ("2d", 777, 2), ("3d", 2500, 3), ("3d", 4096, 4), ("3d", 100, 5)])
def test_culled_fused_pass_equals_matrix_path_on_unclustered_boxes(kind, n, seed):
    """The matrix-free pass skips tile pairs that provably hold no overlap above the threshold (spatial culling).
    Uniformly scattered boxes with many near-threshold neighbours, duplicates and zero-size boxes: the group
    structure and probabilities must still equal the materialise-everything path bit for bit."""
    from groomed_nms_b200 import ops, _lib
    rng = np.random.default_rng(seed)
    sc = (rng.uniform(0.05, 1.0, n) + np.arange(n) * 1e-7).astype(np.float32)
    if kind == "2d":
        c = rng.uniform(0, 600, (n, 2)); wh = rng.uniform(10, 60, (n, 2))
        boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        boxes[5] = boxes[9]                                   # exact duplicates
        boxes[7, 2:] = boxes[7, :2]                           # zero-area boxes, two of them coincident (0/0 = NaN)
        boxes[11] = boxes[7]
        bx = cuda(boxes)
        iou = ops.overlap2d(bx, bx)
        p = ops.make_params(group_size=20)
        st2 = ops.forward_boxes(cuda(sc)[None], bx[None], _lib.BOX_2D, p)
    else:
        b7 = np.stack([rng.uniform(-25, 25, n), 1.6 + 0.2 * rng.standard_normal(n), rng.uniform(5, 60, n),
                       1.6 + 0.2 * rng.standard_normal(n), 1.5 + 0.1 * rng.standard_normal(n),
                       4 + 0.5 * rng.standard_normal(n), rng.uniform(-np.pi, np.pi, n)], 1).astype(np.float32)
        b7[3] = b7[8]
        rec = ops.box3d_records(ops.corners_from_boxes7(cuda(b7)))
        _, iou = ops.overlap3d(rec, rec, False, True, generalized=True, affine=True)
        p = ops.make_params(group_size=20)
        st2 = ops.forward_boxes(cuda(sc)[None], rec[None], _lib.BOX_3D_REC, p, generalized=True, affine=True)
    st1 = ops.forward_matrix(cuda(sc)[None], iou[None], p)
    torch.cuda.synchronize()
    for f in ("order", "lead", "prob", "pre", "counts"):
        assert torch.equal(getattr(st1, f), getattr(st2, f)), f
    nv = int(st1.counts[0, 0])
    assert torch.equal(st1.valid_idx[0, :nv], st2.valid_idx[0, :nv])


@pytest.mark.parametrize("kind", ["2d", "3d"])
def test_direct_election_mixed_batch_and_fallback(kind):
    """Matrix-free path: elect_kernel finds the leaders directly for images with few groups and gives up (flag) for
    images that need too many steps, which then go through the spatial tile list / mask route INSIDE THE SAME LAUNCH
    SEQUENCE.  A ragged batch mixing both kinds of image -- clustered detector-like boxes, uniformly scattered boxes with
    well over 384 leaders, an empty image, a single box -- must equal the matrix path bit for bit, with the direct
    election switched on and off."""
    from groomed_nms_b200 import _lib, ops, synthetic
    lib = _lib.load()
    N = 2048
    rng = np.random.default_rng(77)
    ns = [2048, 2048, 0, 1, 1500, 2048]
    sc = np.zeros((len(ns), N), np.float32)
    if kind == "2d":
        data = np.zeros((len(ns), N, 4), np.float32)
        for b, n in enumerate(ns):
            if b in (0, 4):                                              # clustered: a handful of groups
                bx, s, _ = synthetic.clustered_boxes_2d(N, 9, seed=50 + b, jitter=0.06)
            else:                                                        # scattered: thousands of leaders -> fallback
                c = rng.uniform(0, 2000, (N, 2)); wh = rng.uniform(10, 40, (N, 2))
                bx = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
                s = (rng.uniform(0.05, 1.0, N) + np.arange(N) * 1e-7).astype(np.float32)
            sc[b], data[b] = s, bx
        dev_data = cuda(data)
        iou = torch.stack([ops.overlap2d(dev_data[b], dev_data[b]) for b in range(len(ns))])
        kw = dict(box_kind=_lib.BOX_2D)
    else:
        recs = []
        for b, n in enumerate(ns):
            if b in (0, 4):
                b7, s = synthetic.config_c3(seed=60 + b, n=N, k=12)
            else:
                b7 = np.stack([rng.uniform(-60, 60, N), 1.6 + 0.2 * rng.standard_normal(N), rng.uniform(5, 120, N),
                               1.6 + 0.2 * rng.standard_normal(N), 1.5 + 0.1 * rng.standard_normal(N),
                               4 + 0.5 * rng.standard_normal(N), rng.uniform(-np.pi, np.pi, N)], 1).astype(np.float32)
                s = (rng.uniform(0.05, 1.0, N) + np.arange(N) * 1e-7).astype(np.float32)
            sc[b] = s
            recs.append(ops.box3d_records(ops.corners_from_boxes7(cuda(b7))))
        dev_data = torch.stack(recs)
        iou = torch.stack([ops.overlap3d(r, r, False, True, generalized=True, affine=True)[1] for r in recs])
        kw = dict(box_kind=_lib.BOX_3D_REC, generalized=True, affine=True)
    npi = torch.tensor(ns, dtype=torch.int32, device="cuda")
    p = ops.make_params(group_size=25)
    ref = ops.forward_matrix(cuda(sc), iou, p, n_per_image=npi)
    n_lead = [(int((ref.lead[b, :n] == torch.arange(n, device="cuda")).sum())) for b, n in enumerate(ns)]
    assert n_lead[0] < 100 and n_lead[1] > 384 and n_lead[5] > 384      # both routes are really exercised
    for direct, flags in ((_lib.ELECT_BATCHED, 0), (_lib.ELECT_BATCHED, _lib.OPT_SPLIT_CHAIN), (_lib.ELECT_DIRECT, 0), (_lib.ELECT_MASK, 0)):
        st = ops.forward_boxes(cuda(sc), dev_data, kw["box_kind"], p, kw.get("generalized", False), kw.get("affine", False),
                               n_per_image=npi, opts=_lib.launch_opts(election=direct, flags=flags))
        torch.cuda.synchronize()
        for f in ("order", "lead", "prob", "pre", "counts"):
            assert torch.equal(getattr(ref, f), getattr(st, f)), (direct, f)
        for b in range(len(ns)):
            nv = int(ref.counts[b, 0])
            assert torch.equal(ref.valid_idx[b, :nv], st.valid_idx[b, :nv])
        up = torch.randn(len(ns), N, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
        g1, _ = ops.backward(ref, up)
        g2, _ = ops.backward(st, up)
        assert torch.allclose(g1, g2, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("kind", ["2d", "3d"])
def test_batched_election_queue_overflow_and_odd_sizes(kind):
    """elect2_kernel (up to 32 leaders per step) on inputs built to break its bookkeeping: boxes piled on one spot so that
    nearly every (box, leader) pair passes the gap bound but few exceed the threshold -- the pair queue overflows and steps are
    redone with fewer leaders -- at sizes that are not multiples of 32 / 1024, ragged, against the matrix path bit for bit."""
    from groomed_nms_b200 import _lib, ops
    rng = np.random.default_rng(5)
    N = 4096
    ns = [4096, 3001, 1025, 33, 4095]
    sc = np.zeros((len(ns), N), np.float32)
    if kind == "2d":
        data = np.zeros((len(ns), N, 4), np.float32)
        for b in range(len(ns)):
            c = rng.uniform(0, 260 + 60 * b, (N, 2)); wh = rng.uniform(60, 140, (N, 2))
            data[b] = np.concatenate([c - wh / 2, c + wh / 2], 1)
            sc[b] = (rng.uniform(0.05, 1.0, N) + np.arange(N) * 1e-7).astype(np.float32)
        dev = cuda(data)
        iou = torch.stack([ops.overlap2d(dev[b], dev[b]) for b in range(len(ns))])
        kw = dict(box_kind=_lib.BOX_2D, generalized=False, affine=False)
    else:
        recs = []
        for b in range(len(ns)):
            b7 = np.stack([rng.uniform(-4, 4 + b, N), 1.6 + 0.3 * rng.standard_normal(N), rng.uniform(10, 18 + b, N),
                           1.6 + 0.2 * rng.standard_normal(N), 1.5 + 0.1 * rng.standard_normal(N), 4 + 0.5 * rng.standard_normal(N),
                           rng.uniform(-np.pi, np.pi, N)], 1).astype(np.float32)
            sc[b] = (rng.uniform(0.05, 1.0, N) + np.arange(N) * 1e-7).astype(np.float32)
            recs.append(ops.box3d_records(ops.corners_from_boxes7(cuda(b7))))
        dev = torch.stack(recs)
        iou = torch.stack([ops.overlap3d(r, r, False, True, generalized=True, affine=True)[1] for r in recs])
        kw = dict(box_kind=_lib.BOX_3D_REC, generalized=True, affine=True)
    npi = torch.tensor(ns, dtype=torch.int32, device="cuda")
    for thr, gs in ((0.4, 100), (0.7, 3)):
        p = ops.make_params(nms_threshold=thr, group_size=gs)
        ref = ops.forward_matrix(cuda(sc), iou, p, n_per_image=npi)
        st = ops.forward_boxes(cuda(sc), dev, kw["box_kind"], p, kw["generalized"], kw["affine"], n_per_image=npi,
                               opts=_lib.launch_opts(election=_lib.ELECT_BATCHED))
        torch.cuda.synchronize()
        for f in ("order", "lead", "prob", "pre", "counts"):
            assert torch.equal(getattr(ref, f), getattr(st, f)), (thr, f)
        for b in range(len(ns)):
            nv = int(ref.counts[b, 0])
            assert torch.equal(ref.valid_idx[b, :nv], st.valid_idx[b, :nv])


def test_config_c4_batched_2d_images_vs_oracle(G):
    """BASELINE config 4: a batch of independent N=2048 2D images (16 clusters each), one ragged launch sequence.  Every
    image equals its own single-image call; three of them are checked against the oracle (forward and backward)."""
    from groomed_nms_b200 import _lib, ops, synthetic
    from oracle import groomed_oracle as O
    B, N = 32, 2048
    data = [synthetic.config_c4_image(i) for i in range(B)]
    boxes = np.stack([d[0] for d in data]); sc = np.stack([d[1] for d in data])
    p = ops.make_params(nms_threshold=0.4, temperature=0.1, group_size=100)
    st = ops.forward_boxes(cuda(sc), cuda(boxes), _lib.BOX_2D, p)
    up = np.random.default_rng(4).standard_normal((B, N)).astype(np.float32)
    gs, _ = ops.backward(st, cuda(up))
    torch.cuda.synchronize()
    assert int(st.counts.sum()) == B * N
    for b in (0, 13, 31):
        iou = O.iou(boxes[b], boxes[b])
        o = O.differentiable_nms(sc[b], iou, nms_threshold=0.4, temperature=0.1, group_size=100, dense=False)
        nv = int(st.counts[b, 0])
        check_against(o, st.prob[b].cpu().numpy(), st.valid_idx[b, :nv].cpu().numpy(), st.invalid_idx[b, :N - nv].cpu().numpy())
        assert np.array_equal(st.lead[b].cpu().numpy(), o["lead"])
        want, _ = O.differentiable_nms_backward(o, up[b], need_grad_iou=False)
        assert np.allclose(gs[b].cpu().numpy(), want, rtol=1e-5, atol=2e-6 * np.abs(want).max())
    for b in (1, 7, 20):
        s1 = ops.forward_boxes(cuda(sc[b:b + 1]), cuda(boxes[b:b + 1]), _lib.BOX_2D, p)
        assert torch.equal(s1.prob[0], st.prob[b]) and torch.equal(s1.lead[0], st.lead[b]) and torch.equal(s1.counts[0], st.counts[b])


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_soft_sort_kernels_match_reference(case, G):
    """soft_sort (lib/groomed_nms.py:131-165) on the hand-written kernels (rank by counting, exp / row sums / the column-wise
    division the reference performs, fp32 FMA GEMM for the N x N x N product) and its analytic backward, against the
    reference's own forward and autograd (tests/golden/soft_sort.npz, oracle/gen_golden_misc.py).  1e-5 relative."""
    from conftest import load_golden
    g = load_golden("soft_sort")
    sc, iou, temp = g[case + "_scores"], g[case + "_iou"], float(g[case + "_temp"][0])
    n = len(sc)
    rng = np.random.default_rng(n)
    gs, gp, gm = rng.standard_normal(n).astype(np.float32), rng.standard_normal((n, n)).astype(np.float32), rng.standard_normal((n, n)).astype(np.float32)
    s = cuda(sc).requires_grad_(True)
    m = cuda(iou).requires_grad_(True)
    ss, P, sm = G.soft_sort(s, full_matrix=m, temperature=temp)
    tol = dict(rtol=1e-5, atol=1e-6)
    assert np.allclose(ss.detach().cpu().numpy(), g[case + "_soft_scores"], **tol)
    assert np.allclose(P.detach().cpu().numpy(), g[case + "_P"], **tol)
    assert np.allclose(sm.detach().cpu().numpy(), g[case + "_soft_matrix"], **tol)
    ((ss * cuda(gs)).sum() + (P * cuda(gp)).sum() + (sm * cuda(gm)).sum()).backward()
    want_s, want_m = g[case + "_grad_s"], g[case + "_grad_m"]
    assert np.allclose(m.grad.cpu().numpy(), want_m, rtol=1e-5, atol=1e-5 * np.abs(want_m).max())
    assert np.allclose(s.grad.cpu().numpy(), want_s, rtol=1e-4, atol=1e-5 * np.abs(want_s).max())
    # without a matrix: (soft scores, permutation) only
    ss2, P2 = G.soft_sort(cuda(sc), temperature=temp)
    assert torch.equal(ss2, ss.detach()) and torch.equal(P2, P.detach())
    # the soft path of differentiable_nms on pre-sorted inputs (the regime in which the reference's loop terminates)
    o = g[case + "_dnms_order"]
    s2 = cuda(sc[o]).requires_grad_(True)
    v, i, p = G.differentiable_nms(s2, cuda(iou[o][:, o]), nms_threshold=0.4, temperature=0.1, sorting_method="soft", sorting_temperature=1e-4,
                                   group_size=20)
    assert v.cpu().tolist() == g[case + "_dnms_valid"].tolist() and sorted(i.cpu().tolist()) == sorted(g[case + "_dnms_invalid"].tolist())
    # (cases a and b hold near-tied scores whose soft scores are NOT monotone: the probabilities come back in input order and
    #  Phi is cut by input position, as the reference does -- see lib/soft_sort_impl.py)
    assert np.allclose(p.detach().cpu().numpy(), g[case + "_dnms_prob"], rtol=1e-5, atol=2e-6)
    p.backward(cuda(_dnms_up(n)))
    want = g[case + "_dnms_grad_s"]
    assert np.allclose(s2.grad.cpu().numpy(), want, rtol=1e-4, atol=1e-5 * max(1e-9, np.abs(want).max()))


def _dnms_up(n):
    """The upstream gradient oracle/gen_golden_misc.py drew for the dnms-soft case: the 4th draw of rng(n)."""
    rng = np.random.default_rng(n)
    rng.standard_normal(n); rng.standard_normal((n, n)); rng.standard_normal((n, n))
    return rng.standard_normal(n).astype(np.float32)


@pytest.mark.parametrize("kind", ["2d", "3d"])
def test_batched_election_with_degenerate_boxes(kind):
    """Boxes the straight-line division and the gap bound must not be trusted on -- zero-area, inverted, denormal-size, huge,
    infinite and NaN coordinates -- take the exact path inside elect2_kernel and are never culled: everything still equals the
    matrix path (where a NaN overlap makes the box leave the pool without a group, lib/groomed_nms.py:249-250)."""
    from groomed_nms_b200 import _lib, ops, synthetic
    rng = np.random.default_rng(11)
    N = 1500
    if kind == "2d":
        boxes, sc, _ = synthetic.clustered_boxes_2d(N, 6, seed=31, jitter=0.07)
        bad = rng.choice(N, 60, replace=False)
        boxes[bad[0:10], 2] = boxes[bad[0:10], 0]                               # zero width
        boxes[bad[10:20], 2:] = boxes[bad[10:20], :2] - 5.0                     # inverted
        boxes[bad[20:30], 2:] = boxes[bad[20:30], :2] + 1e-30                   # denormal-size
        boxes[bad[30:40]] *= 1e30                                               # huge
        boxes[bad[40:50], 3] = np.inf
        boxes[bad[50:60], 1] = np.nan
        dev = cuda(boxes)
        iou = ops.overlap2d(dev, dev)
        kw = dict(box_kind=_lib.BOX_2D, generalized=False, affine=False)
        data = dev
    else:
        b7, sc = synthetic.config_c3(seed=8, n=N, k=10)
        bad = rng.choice(N, 50, replace=False)
        b7[bad[0:10], 3:6] = 0.0                                                # zero volume
        b7[bad[10:20], 3:6] = 1e-25
        b7[bad[20:30], :3] *= 1e12
        b7[bad[30:40], 2] = np.inf
        b7[bad[40:50], 0] = np.nan
        rec = ops.box3d_records(ops.corners_from_boxes7(cuda(b7)))
        iou = ops.overlap3d(rec, rec, False, True, generalized=True, affine=True)[1]
        kw = dict(box_kind=_lib.BOX_3D_REC, generalized=True, affine=True)
        data = rec
    assert bool(torch.isnan(iou).any())
    for thr in (0.4, 0.6):
        p = ops.make_params(nms_threshold=thr, group_size=30)
        ref = ops.forward_matrix(cuda(sc)[None], iou[None], p)
        for election in (_lib.ELECT_BATCHED, _lib.ELECT_DIRECT, _lib.ELECT_MASK):
            st = ops.forward_boxes(cuda(sc)[None], data[None], kw["box_kind"], p, kw["generalized"], kw["affine"],
                                   opts=_lib.launch_opts(election=election))
            torch.cuda.synchronize()
            for f in ("order", "lead", "counts"):
                assert torch.equal(getattr(ref, f), getattr(st, f)), (thr, election, f)
            assert torch.equal(ref.prob.view(torch.int32), st.prob.view(torch.int32)), (thr, election)


@pytest.mark.parametrize("seed", list(range(40)))
def test_batched_election_fuzz_against_the_bitmask_route(seed):
    """Randomised differential test: ragged batches of random size, cluster structure, threshold and group size through the batched
    election (with the chain at its end, and as two launches) and through the all-pairs bitmask route -- every saved tensor and
    list must agree bit for bit, forward and backward."""
    from groomed_nms_b200 import _lib, ops, synthetic
    rng = np.random.default_rng(1000 + seed)
    N = int(rng.choice([1, 31, 32, 33, 500, 1024, 1500, 2048, 3000, 4096]))
    B = int(rng.integers(1, 5))
    thr = float(rng.choice([0.1, 0.4, 0.55, 0.7]))
    gs = int(rng.choice([0, 1, 5, 30, 100, 5000]))
    kind = "3d" if seed % 2 else "2d"
    ns = [int(rng.integers(0, N + 1)) for _ in range(B)]
    ns[0] = N
    sc = np.zeros((B, N), np.float32)
    if kind == "2d":
        data = np.zeros((B, N, 4), np.float32)
        for b in range(B):
            k = int(rng.integers(1, max(2, N // 20 + 1)))
            bx, s, _ = synthetic.clustered_boxes_2d(N, k, seed=seed * 10 + b, jitter=float(rng.choice([0.03, 0.1, 0.3])))
            data[b], sc[b] = bx, s
        dev = cuda(data)
        kw = dict(box_kind=_lib.BOX_2D, generalized=False, affine=False)
    else:
        recs = []
        for b in range(B):
            b7, s = synthetic.config_c3(seed=seed * 10 + b, n=N, k=int(rng.integers(1, max(2, N // 16 + 1))))
            sc[b] = s
            recs.append(ops.box3d_records(ops.corners_from_boxes7(cuda(b7))))
        dev = torch.stack(recs)
        kw = dict(box_kind=_lib.BOX_3D_REC, generalized=True, affine=True)
    npi = torch.tensor(ns, dtype=torch.int32, device="cuda")
    p = ops.make_params(nms_threshold=thr, group_size=gs, pruning_method=str(rng.choice(["linear", "sigmoidal", "soft_nms"])), temperature=0.1)
    up = torch.randn(B, N, device="cuda", generator=torch.Generator("cuda").manual_seed(seed))
    outs = []
    for election, flags in ((_lib.ELECT_MASK, 0), (_lib.ELECT_BATCHED, 0), (_lib.ELECT_BATCHED, _lib.OPT_SPLIT_CHAIN)):
        st = ops.forward_boxes(cuda(sc), dev, kw["box_kind"], p, kw["generalized"], kw["affine"], n_per_image=npi,
                               opts=_lib.launch_opts(election=election, flags=flags))
        g, _ = ops.backward(st, up)
        torch.cuda.synchronize()
        outs.append((st, g))
    ref, gref = outs[0]
    for st, g in outs[1:]:
        for f in ("order", "lead", "counts", "pval", "dpval"):
            assert torch.equal(getattr(ref, f), getattr(st, f)), (f, N, ns, thr, gs, kind)
        assert torch.equal(ref.prob.view(torch.int32), st.prob.view(torch.int32)) and torch.equal(ref.pre.view(torch.int32), st.pre.view(torch.int32))
        for b in range(B):
            nv = int(ref.counts[b, 0])
            assert torch.equal(ref.valid_idx[b, :nv], st.valid_idx[b, :nv])
            assert torch.equal(ref.invalid_idx[b, :ns[b] - nv], st.invalid_idx[b, :ns[b] - nv])
        assert torch.equal(gref, g)
