"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the GrooMeD-NMS hot path (numpy, fp32, separately rounded ops).

This is a *restatement* of the reference algorithm (abhi1kumar/groomed_nms @ ad10dbb), not a copy: every
function cites the reference file:line it follows (paths relative to the reference root).  It is pinned against
the real reference by tests/golden/*.npz, which oracle/gen_golden.py produced by importing the unmodified
reference under oracle/ref_shim.py in the build container (tests/test_oracle_golden.py re-checks every vector).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product path (groomed_nms_b200/) must never import it: it fails loudly if the CUDA library is missing.

Conventions shared with the CUDA path (where the reference leaves behaviour implementation-defined):
  * score sort is descending and STABLE (ties: lower input index first); reference uses an unstable
    torch.sort (lib/groomed_nms.py:41), so tests use distinct scores.
  * `invalid_boxes_index` is ordered by ascending score-sorted position (what a stable descending sort of
    the thresholded vector yields); reference order among the exact ties at 0 is implementation-defined
    (lib/groomed_nms.py:121-123) and tests compare it as a set as well.
"""
import numpy as np

F32 = np.float32


def _f32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=F32))


# ---------------------------------------------------------------------------------------------------------
# 2D overlaps                                                                 reference: lib/core.py:178-532
# ---------------------------------------------------------------------------------------------------------
def intersect(box_a, box_b, mode="combinations"):
    """lib/core.py:178-243.  combinations: box_a[M,4], box_b[N,4] -> [N,M] (b-major); list: [M]."""
    a, b = _f32(box_a), _f32(box_b)
    if mode == "combinations":
        max_xy = np.minimum(a[None, :, 2:4], b[:, None, 2:4])          # :210 / :203
        min_xy = np.maximum(a[None, :, 0:2], b[:, None, 0:2])          # :211 / :204
        inter = np.maximum(max_xy - min_xy, F32(0))                    # :212 clamp(.,0)
        return inter[:, :, 0] * inter[:, :, 1]                         # :218
    elif mode == "list":
        max_xy = np.minimum(a[:, 2:4], b[:, 2:4])                      # :226
        min_xy = np.maximum(a[:, 0:2], b[:, 0:2])
        inter = np.maximum(max_xy - min_xy, F32(0))
        return inter[:, 0] * inter[:, 1]                               # :240
    raise ValueError("unknown mode {}".format(mode))                   # :243


def iou(box_a, box_b, mode="combinations"):
    """lib/core.py:480-532.  combinations -> logical [M,N]; list -> [M].  No +1 pixel convention."""
    a, b = _f32(box_a), _f32(box_b)
    if mode == "combinations":
        inter = intersect(a, b, "combinations")                        # [N,M]           :497
        area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])             # :498-499
        area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])             # :500-501
        union = (area_a[None, :] + area_b[:, None]) - inter            # :505 (left-to-right)
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.ascontiguousarray((inter / union).T)             # :506 permute(1,0)
    elif mode == "list":
        inter = intersect(a, b, "list")                                # :520
        area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
        area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
        union = (area_a + area_b) - inter                              # :523
        with np.errstate(divide="ignore", invalid="ignore"):
            return inter / union
    raise ValueError("unknown mode {}".format(mode))


# ---------------------------------------------------------------------------------------------------------
# 3D overlaps (axis-aligned approximation)                                    reference: lib/core.py:305-477
# ---------------------------------------------------------------------------------------------------------
BEV_CORNERS = [2, 3, 6, 7]                                             # lib/core.py:383-384


def overlap2d_f64(box_a, box_b, mode="combinations", kind="iou"):
    """iou / intersect on float64 or mixed float32 / float64 inputs (lib/core.py:196-206, 498-515, numpy branch): everything in
    float64 except the box areas of a float32 side, which numpy / torch type promotion keep in float32.  Returns [M,N] for iou
    and [N,M] for intersect in combinations mode, like the reference; [M] in list mode."""
    a, b = np.asarray(box_a), np.asarray(box_b)
    A, B = a.astype(np.float64), b.astype(np.float64)
    if mode == "combinations":
        A_, B_ = A[:, None, :], B[None, :, :]
    elif mode == "list":
        A_, B_ = A, B
    else:
        raise ValueError("unknown mode {}".format(mode))
    iw = np.maximum(np.minimum(A_[..., 2], B_[..., 2]) - np.maximum(A_[..., 0], B_[..., 0]), 0.0)
    ih = np.maximum(np.minimum(A_[..., 3], B_[..., 3]) - np.maximum(A_[..., 1], B_[..., 1]), 0.0)
    inter = iw * ih
    if kind == "intersect":
        return inter.T if mode == "combinations" else inter
    area = lambda x: ((x[:, 2] - x[:, 0]) * (x[:, 3] - x[:, 1])).astype(np.float64)      # in x's own dtype, then promoted
    aa, ab = area(a), area(b)
    with np.errstate(divide="ignore", invalid="ignore"):
        if mode == "combinations":
            return inter / ((aa[:, None] + ab[None, :]) - inter)
        return inter / ((aa + ab) - inter)


def get_volume(corners_3d):
    """lib/core.py:434-451: prod over (x,y,z) of max-min over the 8 corners, ((dx*dy)*dz)."""
    c = _f32(corners_3d)
    if c.ndim == 2:
        c = c[None]
    d = c.max(axis=2) - c.min(axis=2)                                  # :444-446
    return (d[:, 0] * d[:, 1]) * d[:, 2]                               # :448 torch.prod order


def box_records_from_corners(corners_3d):
    """Per-box quantities iou3d_approximate derives from [N,3,8] corners (lib/core.py:354-388).

    Returns dict of fp32 [N] arrays: ymin,ymax (all 8 corners, :365-368), bx1,bx2,bz1,bz2 (min/max of x and z
    over corners [2,3,6,7] only, :383-388,463-477), vol (:354-355), area_bev."""
    c = _f32(corners_3d)
    rec = {}
    rec["ymin"] = c[:, 1, :].min(axis=1)
    rec["ymax"] = c[:, 1, :].max(axis=1)
    bev_x = c[:, 0, :][:, BEV_CORNERS]
    bev_z = c[:, 2, :][:, BEV_CORNERS]
    rec["bx1"], rec["bx2"] = bev_x.min(axis=1), bev_x.max(axis=1)
    rec["bz1"], rec["bz2"] = bev_z.min(axis=1), bev_z.max(axis=1)
    rec["vol"] = get_volume(c)
    rec["area_bev"] = (rec["bx2"] - rec["bx1"]) * (rec["bz2"] - rec["bz1"])
    return rec


def iou3d_approximate(corners_3d_b1, corners_3d_b2, mode="list", method="normal"):
    """lib/core.py:305-421 -> (iou_bev, iou_3d); combinations: [M,N], list: [M].

    Unlike the reference this does not mutate its inputs (the reference writes Z into Y of the caller's
    storage through a transposed view at :379-380; the product mirrors that quirk, the oracle returns the
    mutated corners separately via `mutated_corners`)."""
    r1 = box_records_from_corners(corners_3d_b1)
    r2 = box_records_from_corners(corners_3d_b2)
    zero = F32(0)
    if mode == "combinations":
        A = lambda v: v[:, None]
        B = lambda v: v[None, :]
    elif mode == "list":
        A = lambda v: v
        B = lambda v: v
    else:
        raise ValueError("unknown mode {}".format(mode))
    vol = A(r1["vol"]) + B(r2["vol"])                                                    # :356-359
    y_int = np.maximum(zero, np.minimum(A(r1["ymax"]), B(r2["ymax"]))
                       - np.maximum(A(r1["ymin"]), B(r2["ymin"])))                       # :370-376
    iw = np.maximum(np.minimum(A(r1["bx2"]), B(r2["bx2"])) - np.maximum(A(r1["bx1"]), B(r2["bx1"])), zero)
    ih = np.maximum(np.minimum(A(r1["bz2"]), B(r2["bz2"])) - np.maximum(A(r1["bz1"]), B(r2["bz1"])), zero)
    inter_bev = iw * ih                                                                  # :410 intersect()
    union_bev = (A(r1["area_bev"]) + B(r2["area_bev"])) - inter_bev
    with np.errstate(divide="ignore", invalid="ignore"):
        iou_bev = inter_bev / union_bev                                                  # :408 iou()
        inter_3d = inter_bev * y_int                                                     # :415
        union_3d = vol - inter_3d                                                        # :416
        iou_3d = inter_3d / union_3d                                                     # :417
        if method == "generalized":
            x_h = np.maximum(zero, np.maximum(A(r1["bx2"]), B(r2["bx2"])) - np.minimum(A(r1["bx1"]), B(r2["bx1"])))
            y_h = np.maximum(zero, np.maximum(A(r1["ymax"]), B(r2["ymax"])) - np.minimum(A(r1["ymin"]), B(r2["ymin"])))
            z_h = np.maximum(zero, np.maximum(A(r1["bz2"]), B(r2["bz2"])) - np.minimum(A(r1["bz1"]), B(r2["bz1"])))
            vol_hull = (x_h * y_h) * z_h                                                 # :406
            iou_3d = iou_3d - ((vol_hull - union_3d) / vol_hull)                         # :419
    return iou_bev.astype(F32), iou_3d.astype(F32)


def mutated_corners(corners_3d):
    """The state the reference leaves its *input* in after iou3d_approximate (lib/core.py:379-380): Y <- Z."""
    c = _f32(corners_3d).copy()
    c[:, 1, :] = c[:, 2, :]
    return c


# ---------------------------------------------------------------------------------------------------------
# 7-DoF -> corners, projection                                      reference: lib/math_3d.py:47-72,364-490
# ---------------------------------------------------------------------------------------------------------
def get_corners_of_cuboid(x3d, y3d, z3d, w3d, h3d, l3d, ry3d, iou_3d_convention=True):
    """lib/math_3d.py:364-435 -> [N,3,8] fp32; iou_3d_convention=False is the vertex order of :405-426 (the reference's torch
    branch there only broadcasts for N in {1, 4}; tests/golden/corners_kitti_order.npz holds its N = 4 output).

    bmm row products are evaluated as (cos*cx + 0*cy) + sin*cz etc. in fp32; the reference's bmm/einsum
    accumulation order/FMA use is backend-defined, so corner parity is a 1e-6-relative check, and overlap
    parity is defined from the corners onward (hard part 1)."""
    x3d, y3d, z3d, w3d, h3d, l3d, ry3d = [_f32(v) for v in (x3d, y3d, z3d, w3d, h3d, l3d, ry3d)]
    n = x3d.shape[0]
    c = np.zeros((n, 3, 8), dtype=F32)
    c[:, 0, [1, 3, 5, 6] if iou_3d_convention else [1, 2, 3, 4]] = l3d[:, None]   # :401 / :422
    c[:, 1, [2, 3, 6, 7]] = h3d[:, None]                                          # :402 / :423
    c[:, 2, [4, 5, 6, 7] if iou_3d_convention else [3, 4, 5, 6]] = w3d[:, None]   # :403 / :424
    c[:, 0] -= (l3d / F32(2))[:, None]                                 # :424-426
    c[:, 1] -= (h3d / F32(2))[:, None]
    c[:, 2] -= (w3d / F32(2))[:, None]
    cs, sn = np.cos(ry3d).astype(F32), np.sin(ry3d).astype(F32)        # :370-374
    out = np.empty_like(c)
    out[:, 0] = cs[:, None] * c[:, 0] + sn[:, None] * c[:, 2]          # R row 0 = [cos, 0, sin]
    out[:, 1] = c[:, 1]                                                # R row 1 = [0, 1, 0]
    out[:, 2] = (-sn)[:, None] * c[:, 0] + cs[:, None] * c[:, 2]       # R row 2 = [-sin, 0, cos]
    out[:, 0] += x3d[:, None]                                          # :433-435
    out[:, 1] += y3d[:, None]
    out[:, 2] += z3d[:, None]
    return out


def get_corners_of_cuboid_backward(boxes7, grad_corners, iou_3d_convention=True):
    """Vector-Jacobian product of get_corners_of_cuboid (lib/math_3d.py:364-435): dL/dcorners [N,3,8] -> dL/d(x,y,z,w,h,l,ry)
    [N,7].  Pinned by tests/golden/grad3d.npz (the reference's autograd)."""
    b, g = _f32(boxes7).astype(np.float64), _f32(grad_corners).astype(np.float64)
    hx = np.full(8, -0.5); hx[[1, 3, 5, 6] if iou_3d_convention else [1, 2, 3, 4]] = 0.5
    hy = np.full(8, -0.5); hy[[2, 3, 6, 7]] = 0.5
    hz = np.full(8, -0.5); hz[[4, 5, 6, 7] if iou_3d_convention else [3, 4, 5, 6]] = 0.5
    w, h, l, ry = b[:, 3], b[:, 4], b[:, 5], b[:, 6]
    cs, sn = np.cos(ry)[:, None], np.sin(ry)[:, None]
    px, pz = hx[None] * l[:, None], hz[None] * w[:, None]
    ga, gb_, gc = g[:, 0], g[:, 1], g[:, 2]
    out = np.empty((b.shape[0], 7))
    out[:, 0], out[:, 1], out[:, 2] = ga.sum(1), gb_.sum(1), gc.sum(1)
    out[:, 3] = (hz[None] * (sn * ga + cs * gc)).sum(1)
    out[:, 4] = (hy[None] * gb_).sum(1)
    out[:, 5] = (hx[None] * (cs * ga - sn * gc)).sum(1)
    out[:, 6] = (ga * (-sn * px + cs * pz) + gc * (-cs * px - sn * pz)).sum(1)
    return out.astype(F32)


def iou3d_approximate_backward(corners_3d_b1, corners_3d_b2, g_bev, g_3d, mode="list", method="normal"):
    """Vector-Jacobian product of iou3d_approximate (lib/core.py:305-421) wrt both (unmutated) corner sets, with torch's
    sub-gradients: clamp(x, 0) of intersect (:210-218) passes on x >= 0; binary min / max and max(zeros, x) (:376, :430) split
    ties evenly; the min / max over the corners of a box go to the first extremal corner.  Pinned by tests/golden/grad3d.npz."""
    ca, cb = _f32(corners_3d_b1).astype(np.float64), _f32(corners_3d_b2).astype(np.float64)
    comb = mode == "combinations"
    M, N = ca.shape[0], cb.shape[0]
    shape = (M, N) if comb else (M,)
    gbev = np.zeros(shape) if g_bev is None else np.asarray(g_bev, np.float64)
    g3 = np.zeros(shape) if g_3d is None else np.asarray(g_3d, np.float64)

    def ext(c):
        e = {}
        for name, row, idx in (("x8", 0, slice(None)), ("y", 1, slice(None)), ("z8", 2, slice(None)), ("bx", 0, BEV_CORNERS), ("bz", 2, BEV_CORNERS)):
            v = c[:, row, :][:, idx]
            ids = np.arange(8)[idx]
            e[name + "lo"], e[name + "hi"] = v.min(1), v.max(1)
            e["i" + name + "lo"], e["i" + name + "hi"] = (row, ids[v.argmin(1)]), (row, ids[v.argmax(1)])
        return e
    ea, eb = ext(ca), ext(cb)
    A = (lambda v: v[:, None]) if comb else (lambda v: v)
    B = (lambda v: v[None, :]) if comb else (lambda v: v)
    names = ["x8lo", "x8hi", "ylo", "yhi", "z8lo", "z8hi", "bxlo", "bxhi", "bzlo", "bzhi"]
    Ga, Gb = {k: np.zeros(shape) for k in names}, {k: np.zeros(shape) for k in names}

    def d_min(ka, kb, g):
        a, b = np.broadcast_arrays(A(ea[ka]), B(eb[kb]))
        Ga[ka] += g * np.where(a < b, 1.0, np.where(a == b, 0.5, 0.0)); Gb[kb] += g * np.where(b < a, 1.0, np.where(a == b, 0.5, 0.0))

    def d_max(ka, kb, g):
        a, b = np.broadcast_arrays(A(ea[ka]), B(eb[kb]))
        Ga[ka] += g * np.where(a > b, 1.0, np.where(a == b, 0.5, 0.0)); Gb[kb] += g * np.where(b > a, 1.0, np.where(a == b, 0.5, 0.0))
    relu_b = lambda x: np.where(x > 0, 1.0, np.where(x == 0, 0.5, 0.0))
    relu_c = lambda x: (x >= 0).astype(np.float64)
    dw = np.minimum(A(ea["bxhi"]), B(eb["bxhi"])) - np.maximum(A(ea["bxlo"]), B(eb["bxlo"]))
    dh = np.minimum(A(ea["bzhi"]), B(eb["bzhi"])) - np.maximum(A(ea["bzlo"]), B(eb["bzlo"]))
    iw, ih = np.maximum(dw, 0), np.maximum(dh, 0)
    ibev = iw * ih
    area = lambda e: (e["bxhi"] - e["bxlo"]) * (e["bzhi"] - e["bzlo"])
    U2 = A(area(ea)) + B(area(eb)) - ibev
    dy = np.minimum(A(ea["yhi"]), B(eb["yhi"])) - np.maximum(A(ea["ylo"]), B(eb["ylo"]))
    yint = np.maximum(dy, 0)
    vol = lambda e: (e["x8hi"] - e["x8lo"]) * (e["yhi"] - e["ylo"]) * (e["z8hi"] - e["z8lo"])
    V = A(vol(ea)) + B(vol(eb))
    i3 = ibev * yint
    un = V - i3
    with np.errstate(divide="ignore", invalid="ignore"):
        g_un, g_i3 = -g3 * i3 / (un * un), g3 / un
        if method == "generalized":
            hx = np.maximum(A(ea["bxhi"]), B(eb["bxhi"])) - np.minimum(A(ea["bxlo"]), B(eb["bxlo"]))
            hy = np.maximum(A(ea["yhi"]), B(eb["yhi"])) - np.minimum(A(ea["ylo"]), B(eb["ylo"]))
            hz = np.maximum(A(ea["bzhi"]), B(eb["bzhi"])) - np.minimum(A(ea["bzlo"]), B(eb["bzlo"]))
            xh, yh, zh = np.maximum(hx, 0), np.maximum(hy, 0), np.maximum(hz, 0)
            vh = xh * yh * zh
            g_vh = -g3 * un / (vh * vh)
            g_un = g_un + g3 / vh
        g_V = g_un
        g_i3 = g_i3 - g_un
        g_ibev = g_i3 * yint + gbev / U2
        g_U2 = -gbev * ibev / (U2 * U2)
    g_ibev = g_ibev - g_U2
    g_dw, g_dh = g_ibev * ih * relu_c(dw), g_ibev * iw * relu_c(dh)
    d_min("bxhi", "bxhi", g_dw); d_max("bxlo", "bxlo", -g_dw); d_min("bzhi", "bzhi", g_dh); d_max("bzlo", "bzlo", -g_dh)
    for G, e, S in ((Ga, ea, A), (Gb, eb, B)):
        G["bxhi"] += g_U2 * S(e["bzhi"] - e["bzlo"]); G["bxlo"] -= g_U2 * S(e["bzhi"] - e["bzlo"])
        G["bzhi"] += g_U2 * S(e["bxhi"] - e["bxlo"]); G["bzlo"] -= g_U2 * S(e["bxhi"] - e["bxlo"])
        dx_, dy_, dz_ = S(e["x8hi"] - e["x8lo"]), S(e["yhi"] - e["ylo"]), S(e["z8hi"] - e["z8lo"])
        G["x8hi"] += g_V * dy_ * dz_; G["x8lo"] -= g_V * dy_ * dz_
        G["yhi"] += g_V * dx_ * dz_; G["ylo"] -= g_V * dx_ * dz_
        G["z8hi"] += g_V * dx_ * dy_; G["z8lo"] -= g_V * dx_ * dy_
    g_dy = g_i3 * ibev * relu_b(dy)
    d_min("yhi", "yhi", g_dy); d_max("ylo", "ylo", -g_dy)
    if method == "generalized":
        gx, gy, gz = g_vh * yh * zh * relu_b(hx), g_vh * xh * zh * relu_b(hy), g_vh * xh * yh * relu_b(hz)
        d_max("bxhi", "bxhi", gx); d_min("bxlo", "bxlo", -gx); d_max("yhi", "yhi", gy); d_min("ylo", "ylo", -gy)
        d_max("bzhi", "bzhi", gz); d_min("bzlo", "bzlo", -gz)
    out_a, out_b = np.zeros_like(ca), np.zeros_like(cb)
    for out, G, e, axis in ((out_a, Ga, ea, 1), (out_b, Gb, eb, 0)):
        for k in names:
            row, idx = e["i" + k]
            tot = G[k].sum(axis=axis) if comb else G[k]
            np.add.at(out, (np.arange(out.shape[0]), row, idx), tot)
    return out_a.astype(F32), out_b.astype(F32)


def project_3d_points_in_4D_format(p2, points, pad_ones=False):
    """lib/math_3d.py:47-72: p2[4,4] @ [pts;1], x,y divided by z where |z| > 1e-2."""
    p2 = _f32(p2)
    pts = _f32(points)
    if pad_ones:
        pts = np.vstack((pts, np.ones((1, pts.shape[1]), dtype=F32)))
    out = (p2.astype(np.float64) @ pts.astype(np.float64)).astype(F32)  # fp32 matmul, order backend-defined
    ind = np.abs(out[2]) > F32(1e-2)
    out[:2, ind] /= out[2, ind]
    return out


def projected_boxes_from_corners(p2, corners_3d, scale_factor=1.0):
    """lib/loss/rpn_3d.py:754-768: min/max of the projected corners -> [N,4] (x1,y1,x2,y2) * scale_factor."""
    c = _f32(corners_3d)
    n = c.shape[0]
    pts = c.transpose(0, 2, 1).reshape(-1, 3).T
    pr = project_3d_points_in_4D_format(p2, pts, pad_ones=True)
    pr = pr.T.reshape(n, 8, 4).transpose(0, 2, 1)
    box = np.stack([pr[:, 0].min(axis=1), pr[:, 1].min(axis=1), pr[:, 0].max(axis=1), pr[:, 1].max(axis=1)], axis=1)
    return (box * F32(scale_factor)).astype(F32)


# ---------------------------------------------------------------------------------------------------------
# GrooMeD-NMS                                                             reference: lib/groomed_nms.py
# ---------------------------------------------------------------------------------------------------------
def stable_sort_desc(scores):
    """Descending, ties by lower index (see module docstring)."""
    s = _f32(scores)
    return np.argsort(-s.astype(np.float64), kind="stable")


def pruning_function(iou_m, nms_threshold=0.4, temperature=0.01, pruning_method="linear"):
    """lib/groomed_nms.py:167-189 (fp32)."""
    x = _f32(iou_m)
    if pruning_method == "linear":
        return x
    if pruning_method == "sigmoidal":
        z = (x - F32(nms_threshold)) / F32(temperature)
        with np.errstate(over="ignore"):
            return (F32(1) / (F32(1) + np.exp(-z))).astype(F32)
    if pruning_method == "soft_nms":
        with np.errstate(over="ignore"):
            return (F32(1) - np.exp(-(x * x) / F32(temperature))).astype(F32)
    raise NotImplementedError("Pruning method not implemented!")       # :177,187


def pruning_derivative(iou_m, nms_threshold, temperature, pruning_method):
    """d p(iou) / d iou, used by the analytic backward."""
    x = _f32(iou_m)
    if pruning_method == "linear":
        return np.ones_like(x)
    if pruning_method == "sigmoidal":
        p = pruning_function(x, nms_threshold, temperature, "sigmoidal")
        return (p * (F32(1) - p) / F32(temperature)).astype(F32)
    if pruning_method == "soft_nms":
        return ((F32(2) * x / F32(temperature)) * np.exp(-(x * x) / F32(temperature))).astype(F32)
    raise NotImplementedError("Pruning method not implemented!")


def get_groups_sorted(iou_sorted, group_threshold, group_size=100):
    """The shrinking-pool loop of lib/groomed_nms.py:242-262 on an already score-sorted matrix.

    Returns a list of int64 arrays of sorted positions.  Uses column 0 of the shrinking matrix, i.e.
    iou_sorted[j, leader] (the LOWER triangle); `>` joins the group, `<=` stays in the pool, NaN does neither
    and silently leaves the pool (:249-250); only the first group_size+1 boxes of a group are kept (:254-255)."""
    m = _f32(iou_sorted)
    pool = np.arange(m.shape[0], dtype=np.int64)
    groups = []
    thr = F32(group_threshold)
    guard = 0
    while pool.size > 0:
        col = m[pool, pool[0]]
        with np.errstate(invalid="ignore"):
            high = col > thr
            low = col <= thr
        g = pool[high][: min(int(high.sum()), group_size + 1)]
        groups.append(g.astype(np.int64))
        if low.sum() == 0:                                             # :258-260
            break
        new_pool = pool[low]
        if new_pool.size == pool.size:
            # reference would loop forever here (leader never leaves the pool: diag <= thr)
            raise RuntimeError("get_groups precondition violated: iou[l,l] must be > group_threshold")
        pool = new_pool
        guard += 1
    return groups


def get_groups(iou_unsorted, group_threshold, scores_unsorted, group_size=100, return_original_indices=True):
    """lib/groomed_nms.py:208-270."""
    idx = stable_sort_desc(scores_unsorted)                            # :213
    m = _f32(iou_unsorted)[idx][:, idx]                                # :214
    groups = get_groups_sorted(m, group_threshold, group_size)
    if return_original_indices:
        groups = [idx[g] for g in groups]                              # :266-268
    return groups


def differentiable_nms(scores_unsorted, iou_unsorted, nms_threshold=0.4, pruning_method="linear",
                       temperature=0.01, valid_box_prob_threshold=0.3, return_sorted_prob=False,
                       group_boxes=True, mask_group_boxes=True, group_size=100, dense=True):
    """lib/groomed_nms.py:10-129 with sorting_method="hard".

    Returns a dict: valid, invalid (int64, input index space), prob (what the reference returns as
    non_suppression_prob), plus the internals the analytic backward and the CUDA parity tests need:
    order, scores_sorted, groups (sorted positions), lead[N] (sorted position of the group leader, -1 if the
    box is in no group), pre (pre-clamp rescored vector), r (clamped, unthresholded), T (dense transform).
    dense=True follows the reference's dense N x N construction (:56-111); dense=False uses the closed forms
    of SURVEY.md section 0.2 for mode A (used for large N where N x N fp32 temporaries are wasteful)."""
    s_u = _f32(scores_unsorted)
    iou_u = _f32(iou_unsorted)
    n = s_u.shape[0]
    order = stable_sort_desc(s_u)                                      # :41
    s = s_u[order]                                                     # :47
    m = iou_u[order][:, order]                                         # :48
    P = np.tril(pruning_function(m, nms_threshold, temperature, pruning_method), -1).astype(F32)  # :71-73
    lead = np.full(n, -1, dtype=np.int64)
    groups = None
    T = None
    if group_boxes:
        groups = get_groups_sorted(m, nms_threshold, group_size)       # :85 (raw iou, not pruned)
        for g in groups:
            if g.size:
                lead[g] = g[0]
        if mask_group_boxes and not dense:
            pre = np.zeros(n, dtype=F32)
            for g in groups:
                if g.size == 0:
                    continue
                l = g[0]
                pre[g] = s[g] - P[g, l] * s[l]                          # row g of (I - Phi) @ s, two non-zeros
                pre[l] = s[l]
        else:
            T = np.zeros((n, n), dtype=F32)                            # :65
            if mask_group_boxes:
                M = np.zeros((n, n), dtype=F32)                        # :95-99
                for g in groups:
                    if g.size:
                        M[g, g[0]] = 1
                Phi = P * M                                            # :100
            else:
                Phi = P
            eye = np.eye(n, dtype=F32)
            for g in groups:                                           # :103-108
                if g.size == 0:
                    continue
                sub = np.ix_(g, g)
                if mask_group_boxes:
                    T[sub] = eye[sub] - Phi[sub]                       # :105
                else:
                    T[sub] = np.linalg.inv((eye[sub] + Phi[sub]).astype(np.float64)).astype(F32)  # :107 (fp64 inverse, see note)
            pre = (T.astype(np.float64) @ s.astype(np.float64)).astype(F32) if not mask_group_boxes else (T @ s).astype(F32)  # :111
    else:
        T = np.linalg.inv((np.eye(n, dtype=F32) + P).astype(np.float64)).astype(F32)  # :110 (fp64 inverse, see note)
        pre = (T.astype(np.float64) @ s.astype(np.float64)).astype(F32)
    r = np.clip(pre, F32(0), F32(1)).astype(F32)                       # :111 clamp
    thr = F32(valid_box_prob_threshold)
    r_thr = r.copy()
    r_thr[r_thr < thr] = 0                                             # :115
    sidx = np.argsort(-r_thr.astype(np.float64), kind="stable")        # :117/:121
    r_sorted = r_thr[sidx]
    valid = order[sidx[r_sorted >= thr]]                               # :118/:122
    invalid = order[sidx[r_sorted < thr]]                              # :119/:123
    if return_sorted_prob:
        prob = r_sorted                                                # :117
    elif group_boxes:
        prob = r                                                       # :125 (the unthresholded clone)
    else:
        prob = r_thr                                                   # :127
    return dict(valid=valid.astype(np.int64), invalid=invalid.astype(np.int64), prob=prob.astype(F32),
                order=order.astype(np.int64), scores_sorted=s, iou_sorted=m, P=P, groups=groups, lead=lead,
                pre=pre, r=r, r_thr=r_thr, sidx=sidx, T=T,
                cfg=dict(nms_threshold=nms_threshold, pruning_method=pruning_method, temperature=temperature,
                         valid_box_prob_threshold=valid_box_prob_threshold,
                         return_sorted_prob=return_sorted_prob, group_boxes=group_boxes,
                         mask_group_boxes=mask_group_boxes, group_size=group_size))


def differentiable_nms_backward(fwd, grad_prob, need_grad_iou=True):
    """Analytic vector-Jacobian product of differentiable_nms (what autograd computes through
    lib/groomed_nms.py:41-127), derived in SURVEY.md section 8(a) and pinned by tests/golden (reference autograd).

    grad_prob is dL/d(prob) for the `prob` returned by the forward.  Returns (grad_scores[N] in input order,
    grad_iou[N,N] in input order or None)."""
    cfg = fwd["cfg"]
    n = fwd["order"].shape[0]
    g = _f32(grad_prob).copy()
    thr = F32(cfg["valid_box_prob_threshold"])
    if cfg["return_sorted_prob"]:
        gs = np.zeros(n, dtype=F32)
        gs[fwd["sidx"]] = g                                            # undo the second sort (:117)
        g = gs * (fwd["r"] >= thr)                                     # in-place zeroing kills those grads (:115)
    elif not cfg["group_boxes"]:
        g = g * (fwd["r"] >= thr)                                      # :115 acts on the returned tensor (:127)
    pre, s, P = fwd["pre"], fwd["scores_sorted"], fwd["P"]
    a = ((pre >= 0) & (pre <= 1)).astype(F32)                          # clamp passes grad on the closed interval
    gt = (g * a).astype(F32)
    ds = np.zeros(n, dtype=F32)
    dPhi = np.zeros((n, n), dtype=F32) if need_grad_iou else None
    if cfg["group_boxes"] and cfg["mask_group_boxes"]:
        lead = fwd["lead"]
        ingrp = lead >= 0
        gt = gt * ingrp
        ds += gt
        mem = np.nonzero(ingrp & (lead != np.arange(n)))[0]
        np.subtract.at(ds, lead[mem], (P[mem, lead[mem]] * gt[mem]).astype(F32))
        if need_grad_iou:
            dPhi[mem, lead[mem]] = -(s[lead[mem]] * gt[mem])
    elif cfg["group_boxes"]:
        for gidx in fwd["groups"]:
            if gidx.size == 0:
                continue
            sub = np.ix_(gidx, gidx)
            Tg = fwd["T"][sub].astype(np.float64)
            dsg = Tg.T @ gt[gidx].astype(np.float64)
            ds[gidx] = dsg.astype(F32)
            if need_grad_iou:
                dPhi[sub] = np.tril(-np.outer(dsg, pre[gidx].astype(np.float64)), -1).astype(F32)
    else:
        T = fwd["T"].astype(np.float64)
        dsf = T.T @ gt.astype(np.float64)
        ds = dsf.astype(F32)
        if need_grad_iou:
            dPhi = np.tril(-np.outer(dsf, pre.astype(np.float64)), -1).astype(F32)
    order = fwd["order"]
    grad_scores = np.zeros(n, dtype=F32)
    grad_scores[order] = ds
    grad_iou = None
    if need_grad_iou:
        dp = pruning_derivative(fwd["iou_sorted"], cfg["nms_threshold"], cfg["temperature"], cfg["pruning_method"])
        d_sorted = (dPhi * np.tril(dp, -1)).astype(F32)
        grad_iou = np.zeros((n, n), dtype=F32)
        grad_iou[np.ix_(order, order)] = d_sorted
    return grad_scores, grad_iou


# ---------------------------------------------------------------------------------------------------------
# Classical NMS family                                    reference: lib/nms/*, lib/nms_others.py
# ---------------------------------------------------------------------------------------------------------
def hard_nms(dets, thresh, shift=1.0, ge=False):
    """Greedy hard NMS on dets[N,5]=(x1,y1,x2,y2,score) -> kept original indices in score order.

    shift=1, ge=False : lib/nms/py_cpu_nms.py:10-38 and the CUDA path lib/nms/nms_kernel.cu:24-78,127-139
                        (suppress when IoU > thresh);
    shift=1, ge=True  : lib/nms/cpu_nms.pyx:17-68 (suppress when IoU >= thresh, :65);
    any shift         : lib/nms_others.py:119-150 girshick_nms.
    Score order is numpy's `scores.argsort()[::-1]` (quick-sort, reversed) as in the reference; tests use
    distinct scores.  IoU in fp32 with the op order of py_cpu_nms.py:21-33."""
    d = _f32(dets)
    x1, y1, x2, y2, sc = d[:, 0], d[:, 1], d[:, 2], d[:, 3], d[:, 4]
    sh = F32(shift)
    areas = (x2 - x1 + sh) * (y2 - y1 + sh)
    order = sc.argsort()[::-1]
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        rest = order[1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(F32(0), xx2 - xx1 + sh)
        h = np.maximum(F32(0), yy2 - yy1 + sh)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
            survive = (ovr < F32(thresh)) if ge else (ovr <= F32(thresh))
        order = rest[survive]
    return keep


def soft_nms(boxes, sigma=0.5, Nt=0.4, threshold=0.001, method=0, shift=1):
    """lib/nms_others.py:6-116 navneeth_soft_nms, restated on index/score arrays instead of in-place row swaps.

    The reference does a selection sort with swaps and removes a box (swap with the last live one) once its
    decayed score drops below `threshold`.  Equivalent formulation: keep a `live` list whose physical order
    evolves exactly as the reference's rows do; arithmetic is float64 as in the reference (numpy default)."""
    b = np.array(boxes, dtype=np.float64, copy=True)
    n = b.shape[0]
    keep = np.arange(n)
    N = n
    i = 0
    while i < N:
        maxpos = i + int(np.argmax(b[i:N, 4])) if N > i else i          # first maximum, as the `<` scan at :31-35
        b[[i, maxpos]] = b[[maxpos, i]]
        keep[[i, maxpos]] = keep[[maxpos, i]]
        tx1, ty1, tx2, ty2 = b[i, 0], b[i, 1], b[i, 2], b[i, 3]
        pos = i + 1
        while pos < N:
            x1, y1, x2, y2 = b[pos, 0], b[pos, 1], b[pos, 2], b[pos, 3]
            area = (x2 - x1 + shift) * (y2 - y1 + shift)
            iw = min(tx2, x2) - max(tx1, x1) + shift
            if iw > 0:
                ih = min(ty2, y2) - max(ty1, y1) + shift
                if ih > 0:
                    ua = float((tx2 - tx1 + shift) * (ty2 - ty1 + shift) + area - iw * ih)
                    ov = iw * ih / ua
                    if method == 1:
                        weight = 1 - ov if ov > Nt else 1
                    elif method == 2:
                        weight = np.exp(-(ov * ov) / sigma)
                    else:
                        weight = 0 if ov > Nt else 1
                    b[pos, 4] = weight * b[pos, 4]
                    if b[pos, 4] < threshold:
                        b[[pos, N - 1]] = b[[N - 1, pos]]
                        keep[[pos, N - 1]] = keep[[N - 1, pos]]
                        N -= 1
                        pos -= 1
            pos += 1
        i += 1
    return keep[:N], b[:N, 4]


# ---------------------------------------------------------------------------------------------------------
# AP loss                                                                reference: lib/loss/aploss.py:14-97
# ---------------------------------------------------------------------------------------------------------
def aploss(logits, targets, delta=1.0):
    """lib/loss/aploss.py:16-81 -> (loss scalar fp32, grad[n] fp32).  delta is forced to 1.0 (:18)."""
    delta = 1.0
    x = _f32(logits).reshape(-1)
    t = np.asarray(targets).reshape(-1)
    grad = np.zeros_like(x)
    if t.size == 0 or t.max() <= 0:                                    # :27-29
        return F32(0), grad
    pos = t == 1
    fg = x[pos]
    thr = fg.min() - F32(delta)                                        # :33
    valid_n = (t == 0) & (x >= thr)                                    # :36
    bg = x[valid_n]
    bg_grad = np.zeros_like(bg)
    nfg = fg.shape[0]
    prec = np.zeros(nfg, dtype=F32)
    order = np.argsort(fg, kind="stable")                              # :47
    max_prec = F32(0)
    for ii in order:                                                   # :50-68
        tmp1 = np.clip((fg - fg[ii]) / F32(2 * delta) + F32(0.5), 0, 1).astype(F32)
        tmp2 = np.clip((bg - fg[ii]) / F32(2 * delta) + F32(0.5), 0, 1).astype(F32)
        a = F32(tmp1.sum(dtype=F32) + F32(0.5))
        b = F32(tmp2.sum(dtype=F32))
        tmp2 = tmp2 / (a + b)
        cur = F32(a / (a + b))
        if max_prec <= cur:
            max_prec = cur
        else:
            tmp2 = tmp2 * ((F32(1) - max_prec) / (F32(1) - cur))
        bg_grad += tmp2.astype(F32)
        prec[ii] = max_prec
    grad[valid_n] = bg_grad
    grad[pos] = -(F32(1) - prec)
    nfg = max(nfg, 1)
    grad /= F32(nfg)
    metric = prec.sum(dtype=F32) / F32(nfg)
    return F32(1) - metric, grad


# ---------------------------------------------------------------------------------------------------------
# Target-assignment overlaps                              reference: lib/rpn_util.py:439-461 (compute_targets)
# ---------------------------------------------------------------------------------------------------------
def targets_overlaps(rois, gts_val, gts_ign):
    """numpy restatement of the overlap part of compute_targets with the call site's dtypes: rois float32 [M,>=4],
    ground truths float64.  lib/core.py:205-207,218 (intersect), :497-513 (iou), :556-564 (iou_ign): min / max / clip and
    the union are evaluated after numpy's promotion to float64, but area_a = (x2-x1)*(y2-y1) of the rois stays float32.
    -> dict(ols [M,G] f64, ols_max, targets, gt_best_rois, gt_best_ols, ols_ign_max)"""
    a32 = np.asarray(rois)[:, :4]
    a = a32.astype(np.float64)

    def inter(b):
        hi = np.minimum(a[None, :, 2:4], b[:, None, 2:4])
        lo = np.maximum(a[None, :, 0:2], b[:, None, 0:2])
        d = np.clip(hi - lo, 0, None)
        return d[:, :, 0] * d[:, :, 1]                                                   # [G,M]

    area_a = ((a32[:, 2] - a32[:, 0]) * (a32[:, 3] - a32[:, 1])).astype(np.float64)      # float32 arithmetic when rois are float32
    out = {}
    g = np.asarray(gts_val, dtype=np.float64).reshape(-1, 4)
    if g.shape[0]:
        it = inter(g)
        area_b = (g[:, 2] - g[:, 0]) * (g[:, 3] - g[:, 1])
        with np.errstate(divide="ignore", invalid="ignore"):
            ols = (it / (area_a[None, :] + area_b[:, None] - it)).T
        out.update(ols=ols, ols_max=np.amax(ols, axis=1), targets=np.argmax(ols, axis=1),
                   gt_best_rois=np.argmax(ols, axis=0), gt_best_ols=np.amax(ols, axis=0))
    gi = np.asarray(gts_ign, dtype=np.float64).reshape(-1, 4)
    if gi.shape[0]:
        it = inter(gi)
        area_b = (gi[:, 2] - gi[:, 0]) * (gi[:, 3] - gi[:, 1])
        with np.errstate(divide="ignore", invalid="ignore"):
            ols_ign = (it / (area_a[None, :] + area_b[:, None] * 0 - it * 0)).T
        out["ols_ign_max"] = np.amax(ols_ign, axis=1)
    else:
        out["ols_ign_max"] = np.zeros(a.shape[0], dtype=np.float32)                     # lib/rpn_util.py:443
    return out
