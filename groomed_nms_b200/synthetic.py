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
