"""TEST INFRASTRUCTURE ONLY -- small golden vectors from the UNMODIFIED reference that the other generators do not cover.

    python oracle/gen_golden_misc.py          # rewrites tests/golden/corners_kitti_order.npz, soft_sort.npz and grad3d.npz

corners_kitti_order: get_corners_of_cuboid(..., iou_3d_convention=False) (lib/math_3d.py:405-426) for N = 4 boxes -- the one
batch size besides 1 for which the reference's torch branch broadcasts (`corners[:, 0, [1,2,3,4]] = l3d`, :422)."""
import os
import signal
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402


def soft_sort_cases(ref):
    """soft_sort (lib/groomed_nms.py:131-165) forward + autograd backward, and differentiable_nms(sorting_method="soft")."""
    from groomed_nms_b200 import synthetic
    g = {}
    for name, n, temp in (("a", 64, 0.05), ("b", 120, 0.01), ("c", 33, 0.5)):
        boxes, sc, _ = synthetic.clustered_boxes_2d(n, 5, seed=40 + n, jitter=0.08)
        iou = ref.core.iou(torch.from_numpy(boxes), torch.from_numpy(boxes)).contiguous().numpy()
        rng = np.random.default_rng(n)
        gs, gp, gm = rng.standard_normal(n).astype(np.float32), rng.standard_normal((n, n)).astype(np.float32), rng.standard_normal((n, n)).astype(np.float32)
        s = torch.from_numpy(sc).clone().requires_grad_(True)
        m = torch.from_numpy(iou).clone().requires_grad_(True)
        ss, P, sm = ref.groomed_nms.soft_sort(s, full_matrix=m, temperature=temp)
        ((ss * torch.from_numpy(gs)).sum() + (P * torch.from_numpy(gp)).sum() + (sm * torch.from_numpy(gm)).sum()).backward()
        g.update({name + "_scores": sc, name + "_iou": iou, name + "_temp": np.array([temp], np.float32),       # upstream gradients: rng(n), see the tests
                  name + "_soft_scores": ss.detach().numpy(), name + "_P": P.detach().numpy(), name + "_soft_matrix": sm.detach().numpy(),
                  name + "_grad_s": s.grad.numpy(), name + "_grad_m": m.grad.numpy()})
        # the whole soft path of differentiable_nms.  soft_sort permutes the ROWS of the matrix only (:164), so unless the boxes
        # already come in descending score order the "diagonal" the grouping loop relies on is an arbitrary overlap, the leader
        # never leaves the pool and the reference loops forever (lib/groomed_nms.py:247-262).  The only inputs on which the
        # reference's soft path terminates are (nearly) pre-sorted ones: the fixture sorts the boxes by score first and uses a
        # sharp sorting temperature.  (The alarm is the safety net.)
        o = np.argsort(-sc, kind="stable")
        sc_s, iou_s = sc[o].copy(), iou[o][:, o].copy()
        s2 = torch.from_numpy(sc_s).clone().requires_grad_(True)
        up = rng.standard_normal(n).astype(np.float32)
        signal.alarm(120)
        valid, invalid, prob = ref.groomed_nms.differentiable_nms(s2, torch.from_numpy(iou_s).clone(), nms_threshold=0.4, temperature=0.1,
                                                                  sorting_method="soft", sorting_temperature=1e-4, group_size=20)
        signal.alarm(0)
        g.update({name + "_dnms_order": o.astype(np.int64)})
        prob.backward(torch.from_numpy(up))
        g.update({name + "_dnms_up": up, name + "_dnms_prob": prob.detach().numpy(), name + "_dnms_valid": valid.numpy().astype(np.int64),
                  name + "_dnms_invalid": invalid.numpy().astype(np.int64), name + "_dnms_grad_s": s2.grad.numpy()})
    path = os.path.join(ROOT, "tests", "golden", "soft_sort.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, "%.1f KiB" % (os.path.getsize(path) / 1024.0))


def grad3d_cases(ref):
    """Autograd gradients of the reference's get_corners_of_cuboid (lib/math_3d.py:364-435) and iou3d_approximate
    (lib/core.py:305-421), the two composites the acceptance-probability target differentiates (lib/loss/rpn_3d.py:663-679).
      corners_*    : d(sum(g * corners)) / d(x, y, z, w, h, l, ry)
      pairs_<mode>_<method>_* : gradients wrt both corner sets; the corners are cuboids plus a small random offset per corner so
                     that no two corners of a box tie in a min / max (torch sends a tied reduction's gradient to an
                     implementation-defined corner)
      chain_*      : parameters -> corners -> iou3d_approximate(list) -> sum(g * iou_3d), real cuboids (ties and all), as the
                     loss does it; gradients wrt the six regressed parameters (ry is detached there, :669)
      self_*       : iou3d_approximate(c, c, combinations, generalized) -- one tensor on both sides, as the NMS call sites do"""
    from groomed_nms_b200 import synthetic
    core, m3d = ref.core, ref.math_3d
    g = {}
    b7, _ = synthetic.config_c3(seed=11, n=96, k=6)
    rng = np.random.default_rng(77)
    b7b = b7.copy()
    b7b[:, :3] += rng.normal(0, 0.25, (96, 3)).astype(np.float32)
    b7b[:, 3:6] *= (1 + rng.normal(0, 0.05, (96, 3))).astype(np.float32)
    b7b[:, 6] += rng.normal(0, 0.1, 96).astype(np.float32)

    def corners_of(b, grad=True, convention=True):
        t = [torch.from_numpy(b[:, i].copy()).requires_grad_(grad) for i in range(7)]
        return t, m3d.get_corners_of_cuboid(*t, iou_3d_convention=convention)

    up = rng.standard_normal((96, 3, 8)).astype(np.float32)
    t, c = corners_of(b7)
    (c * torch.from_numpy(up)).sum().backward()
    g.update(corners_boxes7=b7, corners_up=up, corners_out=c.detach().numpy(), corners_grad=np.stack([v.grad.numpy() for v in t], 1))

    with torch.no_grad():
        ca = corners_of(b7, False)[1].numpy() + rng.normal(0, 0.02, (96, 3, 8)).astype(np.float32)
        cb = corners_of(b7b, False)[1].numpy() + rng.normal(0, 0.02, (96, 3, 8)).astype(np.float32)
    g.update(pairs_a=ca, pairs_b=cb)
    for mode in ("list", "combinations"):
        for method in ("normal", "generalized"):
            la, lb = torch.from_numpy(ca.copy()).requires_grad_(True), torch.from_numpy(cb[:64 if mode == "combinations" else 96].copy()).requires_grad_(True)
            bev, i3d = core.iou3d_approximate(la * 1.0, lb * 1.0, mode=mode, method=method)        # * 1.0: the in-place Y <- Z write refuses leaves
            ub, u3 = rng.standard_normal(tuple(bev.shape)).astype(np.float32), rng.standard_normal(tuple(i3d.shape)).astype(np.float32)
            ((bev * torch.from_numpy(ub)).sum() + (i3d * torch.from_numpy(u3)).sum()).backward()
            k = "pairs_%s_%s_" % (mode, method)
            g.update({k + "up_bev": ub, k + "up_3d": u3, k + "bev": bev.detach().numpy(), k + "3d": i3d.detach().numpy(),
                      k + "grad_a": la.grad.numpy(), k + "grad_b": lb.grad.numpy()})

    ta, c1 = corners_of(b7)
    tb, c2 = corners_of(b7b)
    _, i3d = core.iou3d_approximate(c1, c2)
    u = rng.standard_normal(96).astype(np.float32)
    (i3d * torch.from_numpy(u)).sum().backward()
    g.update(chain_a=b7, chain_b=b7b, chain_up=u, chain_3d=i3d.detach().numpy(), chain_grad_a=np.stack([v.grad.numpy() for v in ta], 1),
             chain_grad_b=np.stack([v.grad.numpy() for v in tb], 1))

    ts, cs = corners_of(b7)
    bev, i3d = core.iou3d_approximate(cs, cs, mode="combinations", method="generalized")
    ub, u3 = rng.standard_normal((96, 96)).astype(np.float32), rng.standard_normal((96, 96)).astype(np.float32)
    ((bev * torch.from_numpy(ub)).sum() + (i3d * torch.from_numpy(u3)).sum() + (cs * torch.from_numpy(up)).sum()).backward()   # cs AFTER the Y <- Z write
    g.update(self_up_bev=ub, self_up_3d=u3, self_grad=np.stack([v.grad.numpy() for v in ts], 1), self_corners_after=cs.detach().numpy())
    path = os.path.join(ROOT, "tests", "golden", "grad3d.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, "%.1f KiB" % (os.path.getsize(path) / 1024.0))


def main():
    ref = ref_shim.load()
    if "--grad3d-only" in sys.argv:
        return grad3d_cases(ref)
    grad3d_cases(ref)
    soft_sort_cases(ref)
    rng = np.random.default_rng(21)
    b7 = np.stack([rng.uniform(-20, 20, 4), rng.uniform(0.5, 2, 4), rng.uniform(5, 60, 4), rng.uniform(1.4, 1.9, 4),
                   rng.uniform(1.3, 1.8, 4), rng.uniform(3, 5, 4), rng.uniform(-np.pi, np.pi, 4)], 1).astype(np.float32)
    t = [torch.from_numpy(b7[:, i].copy()) for i in range(7)]
    # NOTE the reference's `= l3d` (no unsqueeze) assigns l3d[j] to CORNER j of every box when N == 4, not box i's length to its
    # corners; with identical dimensions for the 4 boxes both readings coincide, so the fixture uses one (w, h, l) for all four
    for i in (3, 4, 5):
        t[i][:] = t[i][0]
        b7[:, i] = b7[0, i]
    out = ref.math_3d.get_corners_of_cuboid(*t, iou_3d_convention=False).numpy()
    path = os.path.join(ROOT, "tests", "golden", "corners_kitti_order.npz")
    np.savez_compressed(path, boxes7=b7, corners=out.astype(np.float32))
    print("wrote", path)


if __name__ == "__main__":
    main()
