"""CPU oracle for the reference's exact cuboid overlap `iou3d` (lib/core.py:246-302).  TEST INFRASTRUCTURE ONLY: imported
by tests/ (and nothing in the product path).

The reference delegates the geometry to a third-party dependency that is absent from /root/reference and from this image:
shapely (`from shapely.geometry import Polygon`, lib/core.py:26; the repo pins no version, any 1.x/2.x gives the same
areas) -- `Polygon([...]).intersection(other).area` on the two bottom faces.  PARITY UNPINNED against shapely itself: the
functions below restate the published semantics (area of the intersection of two simple convex polygons) with a method
that differs on purpose from the product's clipping kernel -- candidate vertices = corners of one face inside the other +
proper edge crossings, ordered by angle, shoelace -- and are pinned in tests/test_oracle_golden.py on closed-form areas
(axis-aligned rectangles, a square against its 45-degree rotation, crossed rectangles, containment, disjoint/touching
faces) and on a raster count of random rotated faces.  Everything around the polygon call follows the reference line by
line: y overlap (:281-285), volume sum (:278-279), iou_bev (:298), iou_3d (:299).
"""
import math

import numpy as np

POLYGON_ORDER = (7, 2, 3, 6)           # lib/core.py:290 (the closing vertex 7 is implicit)


def _area2(P):
    s = 0.0
    for i in range(len(P)):
        x0, z0 = P[i]
        x1, z1 = P[(i + 1) % len(P)]
        s += x0 * z1 - x1 * z0
    return s


def polygon_area(P):
    return abs(_area2(P)) * 0.5


def _inside(pt, P, orient, eps):
    for i in range(len(P)):
        x0, z0 = P[i]
        x1, z1 = P[(i + 1) % len(P)]
        if orient * ((x1 - x0) * (pt[1] - z0) - (z1 - z0) * (pt[0] - x0)) < -eps:
            return False
    return True


def convex_intersection_area(P, Q):
    """Area of P n Q for convex polygons given as lists of (x, z)."""
    aP, aQ = _area2(P), _area2(Q)
    if aP == 0.0 or aQ == 0.0:
        return 0.0
    scale = max(max(abs(c) for pt in P + Q for c in pt), 1.0)
    eps = 1e-12 * scale * scale
    pts = [p for p in P if _inside(p, Q, math.copysign(1.0, aQ), eps)]
    pts += [q for q in Q if _inside(q, P, math.copysign(1.0, aP), eps)]
    for i in range(len(P)):
        p0, p1 = P[i], P[(i + 1) % len(P)]
        r = (p1[0] - p0[0], p1[1] - p0[1])
        for j in range(len(Q)):
            q0, q1 = Q[j], Q[(j + 1) % len(Q)]
            s = (q1[0] - q0[0], q1[1] - q0[1])
            den = r[0] * s[1] - r[1] * s[0]
            if abs(den) <= 1e-300:
                continue                                  # parallel edges contribute no proper crossing
            w = (q0[0] - p0[0], q0[1] - p0[1])
            t = (w[0] * s[1] - w[1] * s[0]) / den
            u = (w[0] * r[1] - w[1] * r[0]) / den
            if 0.0 <= t <= 1.0 and 0.0 <= u <= 1.0:
                pts.append((p0[0] + t * r[0], p0[1] + t * r[1]))
    if len(pts) < 3:
        return 0.0
    cx = sum(p[0] for p in pts) / len(pts)
    cz = sum(p[1] for p in pts) / len(pts)
    pts.sort(key=lambda p: math.atan2(p[1] - cz, p[0] - cx))
    return polygon_area(pts)


def get_volume(corners_3d):
    """lib/core.py:453-458 (numpy branch)."""
    c = np.asarray(corners_3d, dtype=np.float64)
    return float(np.prod(c.max(axis=1) - c.min(axis=1)))


def face(corners_3d):
    c = np.asarray(corners_3d, dtype=np.float64)
    return [(float(c[0, i]), float(c[2, i])) for i in POLYGON_ORDER]


def iou3d(corners_3d_b1, corners_3d_b2, vol=None):
    """(iou_bev, iou_3d) of two (3, 8) corner arrays, lib/core.py:246-302."""
    b1 = np.asarray(corners_3d_b1, dtype=np.float64)[:3]
    b2 = np.asarray(corners_3d_b2, dtype=np.float64)[:3]
    if vol is None:
        vol = get_volume(b1) + get_volume(b2)                                    # :278-279
    y_int = max(0.0, min(b1[1].max(), b2[1].max()) - max(b1[1].min(), b2[1].min()))   # :281-285
    f1, f2 = face(b1), face(b2)
    inter = convex_intersection_area(f2, f1)                                     # :295
    i3d = y_int * inter                                                          # :296
    a1, a2 = polygon_area(f1), polygon_area(f2)
    den_bev, den_3d = (a2 + a1) - inter, vol - i3d
    iou_bev = inter / den_bev if den_bev != 0.0 else float("nan")                # :298 (ZeroDivisionError in the reference)
    iou_3d = i3d / den_3d if den_3d != 0.0 else float("nan")                     # :299
    return iou_bev, iou_3d


def raster_intersection_area(P, Q, res=1500):
    """Independent pin: fraction of a res x res grid of sample points inside both convex polygons."""
    P, Q = np.asarray(P, dtype=np.float64), np.asarray(Q, dtype=np.float64)
    lo = np.minimum(P.min(0), Q.min(0))
    hi = np.maximum(P.max(0), Q.max(0))
    xs = lo[0] + (np.arange(res) + 0.5) * (hi[0] - lo[0]) / res
    zs = lo[1] + (np.arange(res) + 0.5) * (hi[1] - lo[1]) / res
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    ok = np.ones_like(X, dtype=bool)
    for poly in (P, Q):
        orient = 1.0 if _area2([tuple(p) for p in poly]) > 0 else -1.0
        for i in range(len(poly)):
            x0, z0 = poly[i]
            x1, z1 = poly[(i + 1) % len(poly)]
            ok &= orient * ((x1 - x0) * (Z - z0) - (z1 - z0) * (X - x0)) >= 0
    return float(ok.mean() * (hi[0] - lo[0]) * (hi[1] - lo[1]))
