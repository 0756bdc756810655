"""Seeded synthetic box workloads (numpy, host side) for the BASELINE.json configs.

C1: 64 2D boxes, one cluster -> one group.            C2: N=1024 2D boxes, 4 well separated clusters.
C3: N=4096 7-DoF boxes, 32 objects x 128 proposals.   C4: 32 images x N=2048 2D boxes (16 clusters each).
Scores are made pairwise distinct so that the score sort is unique (the reference's torch.sort is unstable,
lib/groomed_nms.py:41).
"""
import numpy as np

F32 = np.float32


def distinct_scores(rng, n, lo=0.4, hi=1.0):
    """U(lo,hi) scores, made distinct in fp32 by rejection."""
    s = rng.uniform(lo, hi, size=n).astype(F32)
    for _ in range(100):
        u, first = np.unique(s, return_index=True)
        if u.size == n:
            return s
        dup = np.setdiff1d(np.arange(n), first)
        s[dup] = rng.uniform(lo, hi, size=dup.size).astype(F32)
    raise RuntimeError("could not draw distinct scores")


def clustered_boxes_2d(n, k, seed, canvas=(1760.0, 512.0), size=(100.0, 80.0), jitter=0.05, shuffle=True):
    """n boxes (x1,y1,x2,y2) in k clusters laid out on a grid over the canvas; centre jitter = jitter*size
    (sigma), size jitter = jitter (sigma, relative).  Returns (boxes[n,4] f32, scores[n] f32, cluster_id[n])."""
    rng = np.random.default_rng(seed)
    cols = int(np.ceil(np.sqrt(k * canvas[0] / canvas[1])))
    rows = int(np.ceil(k / cols))
    cx = (np.arange(cols) + 0.5) * canvas[0] / cols
    cy = (np.arange(rows) + 0.5) * canvas[1] / rows
    centres = np.array([(cx[i % cols], cy[i // cols]) for i in range(k)])
    cid = np.arange(n) % k
    if shuffle:
        cid = rng.permutation(cid)
    w = size[0] * (1.0 + jitter * rng.standard_normal(n))
    h = size[1] * (1.0 + jitter * rng.standard_normal(n))
    x = centres[cid, 0] + jitter * size[0] * rng.standard_normal(n)
    y = centres[cid, 1] + jitter * size[1] * rng.standard_normal(n)
    boxes = np.stack([x - w / 2, y - h / 2, x + w / 2, y + h / 2], axis=1).astype(F32)
    return boxes, distinct_scores(rng, n), cid


def config_c1(seed=0):
    """64 2D boxes, one cluster at (500,200), 100x80, 5% jitter -> 1 group."""
    rng = np.random.default_rng(seed)
    n = 64
    w = 100.0 * (1 + 0.05 * rng.standard_normal(n))
    h = 80.0 * (1 + 0.05 * rng.standard_normal(n))
    x = 500.0 + 5.0 * rng.standard_normal(n)
    y = 200.0 + 4.0 * rng.standard_normal(n)
    boxes = np.stack([x - w / 2, y - h / 2, x + w / 2, y + h / 2], axis=1).astype(F32)
    return boxes, distinct_scores(rng, n)


def config_c2(seed=1, n=1024, k=4, jitter=0.03):
    """N=1024 2D boxes in 4 well separated clusters (256 each)."""
    boxes, scores, _ = clustered_boxes_2d(n, k, seed, jitter=jitter)
    return boxes, scores


def config_c3(seed=3, n=4096, k=32):
    """N=4096 7-DoF boxes (x,y,z,w,h,l,ry): 32 KITTI-car-like objects x 128 proposals, 'KITTI-like anchor
    spread' (centre jitter 0.15 m, dim jitter 3 %, yaw jitter 0.05 rad).  Returns (boxes7[n,7], scores[n])."""
    rng = np.random.default_rng(seed)
    ox = rng.uniform(-30, 30, k)
    oz = rng.uniform(5, 70, k)
    oy = 1.65 + 0.1 * rng.standard_normal(k)
    ow = 1.63 + 0.1 * rng.standard_normal(k)
    oh = 1.53 + 0.1 * rng.standard_normal(k)
    ol = 3.9 + 0.4 * rng.standard_normal(k)
    ory = rng.uniform(-np.pi, np.pi, k)
    cid = rng.permutation(np.arange(n) % k)
    b = np.empty((n, 7), dtype=np.float64)
    b[:, 0] = ox[cid] + 0.15 * rng.standard_normal(n)
    b[:, 1] = oy[cid] + 0.15 * rng.standard_normal(n)
    b[:, 2] = oz[cid] + 0.15 * rng.standard_normal(n)
    b[:, 3] = ow[cid] * (1 + 0.03 * rng.standard_normal(n))
    b[:, 4] = oh[cid] * (1 + 0.03 * rng.standard_normal(n))
    b[:, 5] = ol[cid] * (1 + 0.03 * rng.standard_normal(n))
    b[:, 6] = ory[cid] + 0.05 * rng.standard_normal(n)
    return b.astype(F32), distinct_scores(rng, n)


def config_c4_image(i, n=2048, k=16):
    """Image i of the batched config: an independent C2-style draw with 16 clusters, seed 100+i."""
    boxes, scores, _ = clustered_boxes_2d(n, k, 100 + i, jitter=0.04)
    return boxes, scores


# ------------------------------------------------------------------------------------------------ C5 (in-model)
KITTI_P2 = np.array([[721.5377, 0.0, 609.5593, 44.85728], [0.0, 721.5377, 172.854, 0.2163791],
                     [0.0, 0.0, 1.0, 0.002745884], [0.0, 0.0, 0.0, 1.0]], dtype=np.float64)


def c5_anchors(feat_stride=16, test_scale=512, percent_anc_h=(0.0625, 0.75), ratios=(0.5, 1.0, 1.5), n_scales=12):
    """36 x 11 anchor table laid out like the reference's (lib/rpn_util.py:24-216: x1,y1,x2,y2 centred on the stride
    cell, then the 3D priors z, w3d, h3d, l3d, alpha, alpha_sin, alpha_cos).  The reference derives the 3D priors from
    dataset statistics; here z is the depth at which a 1.5 m tall object has the anchor's pixel height."""
    lo, hi = test_scale * percent_anc_h[0], test_scale * percent_anc_h[1]
    base = (hi / lo) ** (1.0 / (n_scales - 1))
    rows = []
    for i in range(n_scales):
        for r in ratios:
            h, w = lo * base ** i, lo * base ** i * r
            c = (feat_stride - 1) / 2.0
            z = min(KITTI_P2[0, 0] * 1.5 / h, 60.0)
            rows.append([-w / 2 + c, -h / 2 + c, w / 2 + c, h / 2 + c, z, 1.63, 1.53, 3.9, 0.0, 0.0, -np.pi / 2])
    return np.asarray(rows, dtype=np.float32)


def c5_rois(anchors, feat_size, stride=16):
    """[(a*H + h)*W + w] -> (x1,y1,x2,y2,anchor id): the unrolling of the reference's locate_anchors(convert_tensor=True)
    (lib/rpn_util.py:965-1034)."""
    H, W = feat_size
    a = np.arange(anchors.shape[0])[:, None, None]
    sy = (np.arange(H) * float(stride))[None, :, None]
    sx = (np.arange(W) * float(stride))[None, None, :]
    z = np.zeros((anchors.shape[0], H, W))
    out = np.stack([anchors[a, 0] + sx + z, anchors[a, 1] + sy + z, anchors[a, 2] + sx + z, anchors[a, 3] + sy + z, a + z], -1)
    return out.reshape(-1, 5)


def _corners_np(b7):
    """[n,7] (x,y,z,w,h,l,ry) -> [n,3,8] corners, the reference's iou_3d_convention (lib/math_3d.py:379-426)."""
    x, y, z, w, h, l, ry = [b7[:, i] for i in range(7)]
    n = b7.shape[0]
    sx = np.array([-1, 1, -1, 1, -1, 1, 1, -1]) * 0.5
    sy = np.array([-1, -1, 1, 1, -1, -1, 1, 1]) * 0.5
    sz = np.array([-1, -1, -1, -1, 1, 1, 1, 1]) * 0.5
    cx, cy, cz = l[:, None] * sx, h[:, None] * sy, w[:, None] * sz
    c, s = np.cos(ry)[:, None], np.sin(ry)[:, None]
    out = np.empty((n, 3, 8))
    out[:, 0] = c * cx + s * cz + x[:, None]
    out[:, 1] = cy + y[:, None]
    out[:, 2] = -s * cx + c * cz + z[:, None]
    return out


def _wrap_pi(a):
    return (a + np.pi) % (2 * np.pi) - np.pi


def c5_scene(seed=0, batch=2, feat_size=(24, 80), n_gt=(3, 8), near_iou=0.3, noise=0.15, stride=16):
    """Synthetic inputs of one training step of the reference's RPN_3D_loss.forward (lib/loss/rpn_3d.py:162) for config
    C5: network outputs for every anchor of a feat_size feature map and KITTI-like car ground
    truths.  Anchors that overlap a ground truth predict its target transform plus noise (so the foreground boxes of an
    object pile up on it, which is what NMS is for); all others predict noise.  Pure numpy, deterministic per seed.

    Returns a dict of float32 arrays: cls [B,T,4] logits, bbox_2d [B,T,4], bbox_3d [B,T,10], acc_logit [B,T] (acceptance
    probability before the sigmoid), anchors [36,11], bbox_means / bbox_stds [1,13], rois [T,5], p2 [4,4] (float64),
    gts = per image list of dicts (cls, ign, visibility, bbox_full [x,y,w,h], bbox_3d [16]), feat_size, scale_factor."""
    rng = np.random.default_rng(seed)
    H, W = feat_size
    anchors = c5_anchors(stride)
    rois = c5_rois(anchors, feat_size, stride)
    T = rois.shape[0]
    aid = rois[:, 4].astype(np.int64)
    means = np.zeros((1, 13), dtype=np.float32)
    stds = np.array([[0.14, 0.13, 0.25, 0.25, 0.12, 0.12, 4.0, 0.07, 0.07, 0.1, 1.7, 0.8, 0.8]], dtype=np.float32)
    p2 = KITTI_P2.copy()
    f, cxp, cyp, ph = p2[0, 0], p2[0, 2], p2[1, 2], p2[2, 3]
    img_w, img_h = W * stride, H * stride
    rw, rh = rois[:, 2] - rois[:, 0] + 1.0, rois[:, 3] - rois[:, 1] + 1.0
    rcx, rcy = rois[:, 0] + 0.5 * rw, rois[:, 1] + 0.5 * rh
    out = dict(cls=np.empty((batch, T, 4), F32), bbox_2d=np.empty((batch, T, 4), F32), bbox_3d=np.empty((batch, T, 10), F32),
               acc_logit=np.empty((batch, T), F32), anchors=anchors, bbox_means=means, bbox_stds=stds,
               rois=rois.astype(F32), p2=p2, gts=[], feat_size=[H, W], scale_factor=1.0)
    for b in range(batch):
        g = int(rng.integers(n_gt[0], n_gt[1] + 1))
        z = rng.uniform(6.0, 30.0, g)
        x = rng.uniform(-0.42, 0.42, g) * z * img_w / f
        y = 1.0 + 0.1 * rng.standard_normal(g)
        dims = np.stack([1.63 + 0.1 * rng.standard_normal(g), 1.53 + 0.1 * rng.standard_normal(g), 3.9 + 0.4 * rng.standard_normal(g)], 1)
        rot_y = rng.uniform(-np.pi, np.pi, g)
        b7 = np.concatenate([x[:, None], y[:, None], z[:, None], dims, rot_y[:, None]], 1)
        cn = _corners_np(b7)
        hom = np.concatenate([cn, np.ones((g, 1, 8))], 1)
        pr = np.einsum('ij,gjk->gik', p2, hom)
        u, v = pr[:, 0] / pr[:, 2], pr[:, 1] / pr[:, 2]
        x1, y1, x2, y2 = np.clip(u.min(1), 0, img_w - 1), np.clip(v.min(1), 0, img_h - 1), np.clip(u.max(1), 0, img_w - 1), np.clip(v.max(1), 0, img_h - 1)
        ctr = p2 @ np.stack([x, y, z, np.ones(g)])
        alpha = _wrap_pi(rot_y - np.arctan2(-z, x) - 0.5 * np.pi)
        a_sin = alpha.copy()
        a_sin[a_sin > np.pi / 2] -= np.pi
        a_sin[a_sin <= -np.pi / 2] += np.pi
        a_cos = alpha.copy()
        a_cos[a_cos > 0] -= np.pi
        axis = (np.abs(np.sin(alpha)) < np.abs(np.cos(alpha))).astype(np.float64)
        base = np.where(axis == 1, a_sin, a_cos)
        head = (np.abs(_wrap_pi(base - alpha)) > 1e-6).astype(np.float64)
        gts = []
        names = ['Car'] * g
        if g > 3:
            names[-1] = 'Van'                                       # an ignored class (conf.ilbls) exercises iou_ign
        for i in range(g):
            b3 = [ctr[0, i] / ctr[2, i], ctr[1, i] / ctr[2, i], z[i] + ph, dims[i, 0], dims[i, 1], dims[i, 2], alpha[i], x[i], y[i], z[i],
                  rot_y[i], 0.0, a_sin[i], a_cos[i], axis[i], head[i]]
            gts.append(dict(cls=names[i], ign=False, visibility=1.0, bbox_full=np.array([x1[i], y1[i], x2[i] - x1[i] + 1, y2[i] - y1[i] + 1]),
                            bbox_3d=[float(t) for t in b3]))
        out['gts'].append(gts)
        # overlap of every anchor with every ground truth (pixel boxes, no +1): who predicts what
        gx1, gy1, gx2, gy2 = x1[None], y1[None], x2[None], y2[None]
        iw = np.clip(np.minimum(rois[:, 2:3], gx2) - np.maximum(rois[:, 0:1], gx1), 0, None)
        ih = np.clip(np.minimum(rois[:, 3:4], gy2) - np.maximum(rois[:, 1:2], gy1), 0, None)
        inter = iw * ih
        ov = inter / ((rois[:, 2:3] - rois[:, 0:1]) * (rois[:, 3:4] - rois[:, 1:2]) + (gx2 - gx1) * (gy2 - gy1) - inter)
        best, who = ov.max(1), ov.argmax(1)
        near = best >= near_iou
        gw, gh = x2 - x1 + 1.0, y2 - y1 + 1.0
        gcx, gcy = x1 + 0.5 * gw, y1 + 0.5 * gh
        t2 = np.stack([(gcx[who] - rcx) / rw, (gcy[who] - rcy) / rh, np.log(gw[who] / rw), np.log(gh[who] / rh)], 1)
        A = anchors[aid].astype(np.float64)
        t3 = np.stack([(ctr[0, who] / ctr[2, who] - rcx) / rw, (ctr[1, who] / ctr[2, who] - rcy) / rh, z[who] + ph - A[:, 4],
                       np.log(dims[who, 0] / A[:, 5]), np.log(dims[who, 1] / A[:, 6]), np.log(dims[who, 2] / A[:, 7]),
                       a_sin[who] - A[:, 9], a_cos[who] - A[:, 10]], 1)
        t2 = (t2 - means[:, 0:4]) / stds[:, 0:4]
        t3 = (t3 - means[:, [4, 5, 6, 7, 8, 9, 11, 12]]) / stds[:, [4, 5, 6, 7, 8, 9, 11, 12]]
        n2 = rng.standard_normal((T, 4))
        n3 = rng.standard_normal((T, 8))
        out['bbox_2d'][b] = np.where(near[:, None], t2 + noise * n2, 0.5 * n2)
        out['bbox_3d'][b, :, :8] = np.where(near[:, None], t3 + noise * n3, 0.5 * n3)
        lab = np.stack([axis[who], head[who]], 1)
        out['bbox_3d'][b, :, 8:] = np.clip(np.where(near[:, None], lab, 0.5) + 0.2 * rng.standard_normal((T, 2)), 0.02, 0.98)
        cls = rng.standard_normal((T, 4))
        cls[:, 0] += np.where(near, -1.0, 2.5)
        cls[:, 1] += np.where(near, 3.0 * best + 1.0, 0.0)
        out['cls'][b] = cls
        out['acc_logit'][b] = np.where(near, 6.0 * best - 2.0, -3.0) + 0.8 * rng.standard_normal(T) + 1e-6 * np.arange(T)
    return out
