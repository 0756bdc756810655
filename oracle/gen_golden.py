"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported under oracle/ref_shim.py) on CPU in the build container.

    python oracle/gen_golden.py            # rewrites tests/golden/

The fixtures hold inputs AND the reference's outputs (forward values, index lists, autograd gradients), so
the oracle restatement and the CUDA path can be checked anywhere, including on the GPU box where
/root/reference does not exist.  Deterministic: all seeds fixed; torch CPU, fp32.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from groomed_nms_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote %-40s %7.1f KiB" % (name + ".npz", os.path.getsize(path) / 1024.0))


def run_dnms(ref, scores, iou, g=None, need_grad_iou=True, **kw):
    """Reference differentiable_nms forward (+ autograd backward with upstream gradient g)."""
    s = T(scores).clone().requires_grad_(True)
    m = T(iou).clone().requires_grad_(need_grad_iou)
    valid, invalid, prob = ref.groomed_nms.differentiable_nms(s, m, **kw)
    out = dict(valid=valid.numpy().astype(np.int64), invalid=invalid.numpy().astype(np.int64),
               prob=prob.detach().numpy().astype(np.float32))
    if g is not None:
        prob.backward(T(g))
        out["grad_scores"] = s.grad.numpy().astype(np.float32) if s.grad is not None else np.zeros_like(scores)
        if need_grad_iou:
            out["grad_iou"] = m.grad.numpy().astype(np.float32) if m.grad is not None else np.zeros_like(iou)
    return out


def main():
    ref = ref_shim.load()
    torch.set_num_threads(4)
    dn = ref.groomed_nms.differentiable_nms

    # ------------------------------------------------------------------ KATs of the reference's own tests
    # test/test_differentiable_nms_forward.py:127-140 (printed expectations at :129 and :137)
    iou4 = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0.9, 0.9, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
    s4 = np.array([0.99, 0.98, 0.8, 0.7], dtype=np.float32)
    iou5 = np.array([[1, 0, 0, 0, 0], [0, 1, 0, 0, 0], [0.9, 0.9, 1, 0, 0], [0.9, 0.9, 0, 1, 0], [0, 0, 0.9, 0.9, 1]],
                    dtype=np.float32)
    s5 = np.array([0.99, 0.98, 0.8, 0.7, 0.6], dtype=np.float32)
    kw = dict(nms_threshold=0.4, temperature=0.1, valid_box_prob_threshold=0.3, pruning_method="linear",
              sorting_method="hard", return_sorted_prob=False, group_boxes="True")
    o4 = run_dnms(ref, s4, iou4, **kw)
    o5 = run_dnms(ref, s5, iou5, **kw)
    save("kat_forward", iou4=iou4, s4=s4, prob4=o4["prob"], valid4=o4["valid"], invalid4=o4["invalid"],
         printed4=np.array([0.990, 0.980, 0.000, 0.700], dtype=np.float32),
         iou5=iou5, s5=s5, prob5=o5["prob"], valid5=o5["valid"], invalid5=o5["invalid"],
         printed5=np.array([0.990, 0.980, 0.000, 0.000, 0.600], dtype=np.float32))

    # seeded 5-box scenario, test/test_differentiable_nms_forward.py:49-122 (same RNG call sequence)
    torch.manual_seed(0)
    scores = torch.FloatTensor(5).uniform_(0.4, 1.0)           # get_scores_iou :18
    _ = torch.rand(5, 5)                                       # :22 (consumed, unused)
    w = torch.rand(5) * 10                                     # :95
    aboxes = torch.zeros((5, 5))
    aboxes[:, 2] = w
    aboxes[:, 3] = w
    aboxes[:, 4] = scores
    iou_ov = ref.core.iou(aboxes[:, :4], aboxes[:, :4], mode="combinations")
    keep_ours, inv_ours, prob_ours = dn(aboxes[:, 4], iou_ov, nms_threshold=0.4, temperature=0.1,
                                        valid_box_prob_threshold=0.3, pruning_method="linear",
                                        sorting_method="hard", group_boxes="True")
    keep_soft = ref.nms_others.navneeth_soft_nms(aboxes.clone().numpy(), Nt=0.4, shift=1)
    keep_girshick = ref.nms_others.girshick_nms(aboxes.clone().numpy(), 0.4, shift=1)
    keep_py = ref.py_cpu_nms.py_cpu_nms(aboxes.clone().numpy(), 0.4)
    save("seeded_5box", aboxes=aboxes.numpy(), iou=iou_ov.contiguous().numpy(), keep_ours=keep_ours.numpy(),
         invalid_ours=inv_ours.numpy(), prob_ours=prob_ours.numpy(), keep_soft=np.asarray(keep_soft),
         keep_girshick=np.asarray(keep_girshick), keep_py=np.asarray(keep_py))

    # test/test_get_groups.py (prints [[0,1,4,5,6,7,9],[2,3,8]])
    torch.manual_seed(0)
    iou10 = torch.rand(10, 10)
    for i in range(10):
        iou10[i, i] = 1
    iou10 = 0.5 * (iou10.transpose(1, 0) + iou10)
    sc10 = torch.sort(torch.rand(10), descending=True)[0]
    groups = ref.groomed_nms.get_groups(scores_unsorted=sc10, iou_unsorted=iou10, group_threshold=0.4)
    save("get_groups_10", iou=iou10.numpy(), scores=sc10.numpy(),
         groups_flat=np.concatenate([g.numpy() for g in groups]),
         groups_len=np.array([len(g) for g in groups]))

    # ------------------------------------------------------------------ overlaps
    rng = np.random.default_rng(11)
    a = rng.uniform(0, 100, (37, 2))
    box_a = np.concatenate([a, a + rng.uniform(1, 60, (37, 2))], axis=1).astype(np.float32)
    b = rng.uniform(0, 100, (53, 2))
    box_b = np.concatenate([b, b + rng.uniform(1, 60, (53, 2))], axis=1).astype(np.float32)
    box_a[5] = box_b[7]                                        # an exact duplicate pair
    box_b[9, 2:] = box_b[9, :2]                                # a zero-area box
    box_a[6] = box_b[9]                                        # ... coincident with a zero-area box -> 0/0 = NaN
    with np.errstate(all="ignore"):
        save("iou2d",
             box_a=box_a, box_b=box_b,
             iou_comb=ref.core.iou(T(box_a), T(box_b), mode="combinations").contiguous().numpy(),
             inter_comb=ref.core.intersect(T(box_a), T(box_b), mode="combinations").contiguous().numpy(),
             iou_list=ref.core.iou(T(box_a), T(box_b[:37]), mode="list").numpy(),
             inter_list=ref.core.intersect(T(box_a), T(box_b[:37]), mode="list").numpy(),
             iou_self=ref.core.iou(T(box_a), T(box_a), mode="combinations").contiguous().numpy(),
             iou_comb_np=ref.core.iou(box_a, box_b, mode="combinations"))

    b7, _ = synthetic.config_c3(seed=5, n=96, k=6)
    b7b, _ = synthetic.config_c3(seed=5, n=64, k=6)
    b7b = b7b[::-1].copy()
    gc = ref.math_3d.get_corners_of_cuboid

    def corners_of(x):
        return gc(*[T(x[:, i].copy()) for i in range(7)])

    c1 = corners_of(b7)
    c2 = corners_of(b7b)
    c1n = c1.numpy().copy()
    c2n = c2.numpy().copy()
    out = {}
    for method in ("normal", "generalized"):
        bev, i3d = ref.core.iou3d_approximate(c1.clone(), c2.clone(), mode="combinations", method=method)
        out["bev_comb_" + method] = bev.contiguous().numpy()
        out["i3d_comb_" + method] = i3d.contiguous().numpy()
        bev, i3d = ref.core.iou3d_approximate(c1[:64].clone(), c2.clone(), mode="list", method=method)
        out["bev_list_" + method] = bev.contiguous().numpy()
        out["i3d_list_" + method] = i3d.contiguous().numpy()
    cs = c1.clone()
    bev, i3d = ref.core.iou3d_approximate(cs, cs, mode="combinations", method="generalized")
    out["i3d_self_generalized"] = i3d.contiguous().numpy()
    out["corners_after_self_call"] = cs.numpy().copy()        # documents the in-place Y<-Z mutation
    save("iou3d", boxes7_a=b7, boxes7_b=b7b, corners_a=c1n, corners_b=c2n,
         volume_a=ref.core.get_volume(T(c1n)).numpy(), **out)

    # test/test_get_corners_of_cuboid_numpy.py inputs
    np.random.seed(0)
    m = 5
    x3d = 30 * np.random.uniform(size=(m))
    y3d = 10 * np.random.uniform(size=(m))
    z3d = 15 * np.random.uniform(size=(m))
    l3d = 4 * np.random.uniform(size=(m))
    w3d = 5 * np.random.uniform(size=(m))
    h3d = 6 * np.random.uniform(size=(m))
    r3d = np.random.uniform(low=-1.57, high=1.57, size=(m))
    co_np = gc(x3d, y3d, z3d, w3d, h3d, l3d, r3d)
    co_t = gc(*[torch.from_numpy(v).float() for v in (x3d, y3d, z3d, w3d, h3d, l3d, r3d)]).numpy()
    p2 = np.array([[721.5377, 0, 609.5593, 44.85728], [0, 721.5377, 172.854, 0.2163791],
                   [0, 0, 1, 0.002745884], [0, 0, 0, 1]], dtype=np.float32)
    pts = torch.from_numpy(co_t).transpose(1, 2).reshape((-1, 3)).transpose(0, 1)
    proj = ref.math_3d.project_3d_points_in_4D_format(T(p2), pts, pad_ones=True).numpy()
    save("corners", args64=np.stack([x3d, y3d, z3d, w3d, h3d, l3d, r3d], axis=1), corners_np=co_np,
         corners_torch=co_t, p2=p2, projected=proj)

    # ------------------------------------------------------------------ differentiable_nms fwd + autograd bwd
    cases = {}
    boxes, scores, _ = synthetic.clustered_boxes_2d(240, 5, seed=21, jitter=0.08)
    iou240 = ref.core.iou(T(boxes), T(boxes), mode="combinations").contiguous().numpy()
    g240 = np.random.default_rng(22).standard_normal(240).astype(np.float32)
    boxes96, scores96, _ = synthetic.clustered_boxes_2d(96, 3, seed=23, jitter=0.10)
    iou96 = ref.core.iou(T(boxes96), T(boxes96), mode="combinations").contiguous().numpy()
    g96 = np.random.default_rng(24).standard_normal(96).astype(np.float32)
    # a non-symmetric random matrix (the reference only ever reads the lower triangle of the sorted matrix)
    rr = np.random.default_rng(25)
    iou_ns = rr.uniform(0, 1, (120, 120)).astype(np.float32) ** 3
    np.fill_diagonal(iou_ns, 1.0)
    s_ns = synthetic.distinct_scores(rr, 120, 0.05, 1.0)
    g_ns = rr.standard_normal(120).astype(np.float32)
    save("dnms_inputs", boxes240=boxes, scores240=scores, iou240=iou240, g240=g240,
         boxes96=boxes96, scores96=scores96, iou96=iou96, g96=g96, iou_ns=iou_ns, s_ns=s_ns, g_ns=g_ns)

    def add(tag, scores, iou, g, need_grad_iou=True, **kw):
        full = dict(nms_threshold=0.4, pruning_method="linear", temperature=0.1, valid_box_prob_threshold=0.3,
                    return_sorted_prob=False, group_boxes=True, mask_group_boxes=True, group_size=100)
        full.update(kw)
        o = run_dnms(ref, scores, iou, g=g, need_grad_iou=need_grad_iou, **full)
        for k, v in o.items():
            cases[tag + "__" + k] = v
        cases[tag + "__cfg"] = np.array(repr(sorted(full.items())))

    for pm, temp in (("linear", 0.1), ("sigmoidal", 0.1), ("soft_nms", 0.5)):
        add("A240_%s" % pm, scores, iou240, g240, pruning_method=pm, temperature=temp, need_grad_iou=False)
        add("A96_%s" % pm, scores96, iou96, g96, pruning_method=pm, temperature=temp)
        add("B96_%s" % pm, scores96, iou96, g96, pruning_method=pm, temperature=temp, mask_group_boxes=False,
            group_size=20)
        add("C96_%s" % pm, scores96, iou96, g96, pruning_method=pm, temperature=temp, group_boxes=False)
    for gs in (1, 15, 20, 100, 10 ** 9):
        add("A240_gs%d" % gs, scores, iou240, g240, group_size=gs, need_grad_iou=False)
        add("B240_gs%d" % gs, scores, iou240, g240, group_size=min(gs, 300), mask_group_boxes=False,
            need_grad_iou=False)
    add("A120ns", s_ns, iou_ns, g_ns, group_size=7)
    add("B120ns", s_ns, iou_ns, g_ns, group_size=7, mask_group_boxes=False)
    add("C120ns", s_ns, iou_ns, g_ns, group_boxes=False)
    add("A96_sortedprob", scores96, iou96, g96, return_sorted_prob=True)
    add("C96_sortedprob", scores96, iou96, g96, return_sorted_prob=True, group_boxes=False)
    add("A96_thr", scores96, iou96, g96, nms_threshold=0.6, valid_box_prob_threshold=0.5)
    save("dnms_cases", **cases)

    # groups on the same inputs
    gr = {}
    for tag, (sc, io) in dict(g240=(scores, iou240), g96=(scores96, iou96), gns=(s_ns, iou_ns)).items():
        for gs in (1, 7, 100):
            groups = ref.groomed_nms.get_groups(T(io), 0.4, T(sc), group_size=gs)
            gr["%s_gs%d_flat" % (tag, gs)] = np.concatenate([x.numpy() for x in groups])
            gr["%s_gs%d_len" % (tag, gs)] = np.array([len(x) for x in groups])
    save("groups", **gr)

    # ------------------------------------------------------------------ classical NMS family
    nm = {}
    boxes_n, scores_n, _ = synthetic.clustered_boxes_2d(300, 12, seed=31, canvas=(600.0, 300.0), size=(60.0, 45.0),
                                                        jitter=0.15)
    dets = np.concatenate([np.round(boxes_n), scores_n[:, None]], axis=1).astype(np.float32)
    nm["dets"] = dets
    for thr in (0.3, 0.5, 0.7):
        nm["py_cpu_nms_%g" % thr] = np.asarray(ref.py_cpu_nms.py_cpu_nms(dets.copy(), thr))
        for shift in (0, 1):
            nm["girshick_%g_s%d" % (thr, shift)] = np.asarray(ref.nms_others.girshick_nms(dets.copy(), thr, shift=shift))
    for method in (0, 1, 2):
        for shift in (0, 1):
            d64 = dets.astype(np.float64).copy()
            keep = ref.nms_others.navneeth_soft_nms(d64, sigma=0.5, Nt=0.4, threshold=0.05 if method else 0.001,
                                                    method=method, shift=shift)
            nm["soft_m%d_s%d_keep" % (method, shift)] = np.asarray(keep)
            nm["soft_m%d_s%d_scores" % (method, shift)] = d64[:len(keep), 4]
    save("classical_nms", **nm)

    # ------------------------------------------------------------------ AP loss (test/test_aploss.py:200-212)
    ap = {}
    logits = np.array([[1.00, 0.00, 0.0069, 0.0064, 0.0108], [0.0059, 0.8, 0.0025, 0.0023, 0.0064]], dtype=np.float32)
    labels = np.array([[1, 0, 0, 0, 0], [0, 0, 1, 0, 0]], dtype=np.float32)
    x = T(logits).clone().requires_grad_(True)
    loss = ref.aploss.APLoss()(x, T(labels))
    loss.backward()
    ap["logits"], ap["labels"], ap["joint_loss"], ap["joint_grad"] = logits, labels, loss.detach().numpy(), x.grad.numpy().copy()
    x.grad.zero_()
    loss = (ref.aploss.APLoss()(x[0], T(labels)[0]) + ref.aploss.APLoss()(x[1], T(labels)[1])) / 2
    loss.backward()
    ap["mean_loss"], ap["mean_grad"] = loss.detach().numpy(), x.grad.numpy().copy()
    rr = np.random.default_rng(41)
    lg = rr.uniform(0, 1, 400).astype(np.float32)
    tg = (rr.uniform(0, 1, 400) < 0.04).astype(np.float32)
    tg[rr.integers(0, 400, 20)] = -1
    x = T(lg).clone().requires_grad_(True)
    loss = ref.aploss.APLoss()(x, T(tg))
    (loss * 1.7).backward()
    ap["rand_logits"], ap["rand_targets"], ap["rand_loss"], ap["rand_grad_x1p7"] = lg, tg, loss.detach().numpy(), x.grad.numpy().copy()
    zero = ref.aploss.APLoss()(T(lg), T(np.zeros_like(tg)))
    ap["allbg_loss"] = zero.numpy()
    save("aploss", **ap)


if __name__ == "__main__":
    main()
