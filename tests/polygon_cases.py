"""Shared inputs of the exact cuboid overlap tests (CPU and GPU): rotated cuboids as (3, 8) float64 corner arrays in the
reference's iou_3d_convention corner order (lib/math_3d.py:364-435), and the host build of the kernel's core."""
import ctypes
import math
import os
import subprocess

import numpy as np

from conftest import ROOT


def cuboid(cx, cz, l, w, ry, y=1.0, h=1.5):
    """Corners of one cuboid: x row, y row, z row; bottom face = corners 7, 2, 3, 6."""
    c = np.zeros((3, 8))
    c[0, [1, 3, 5, 6]] = l
    c[1, [2, 3, 6, 7]] = h
    c[2, [4, 5, 6, 7]] = w
    c -= np.array([l, h, w])[:, None] / 2
    cs, sn = math.cos(ry), math.sin(ry)
    out = np.stack([cs * c[0] + sn * c[2], c[1], -sn * c[0] + cs * c[2]])
    return out + np.array([cx, y, cz])[:, None]


def random_cuboids(n, seed, spread=6.0):
    rng = np.random.default_rng(seed)
    return np.stack([cuboid(rng.uniform(-spread, spread), rng.uniform(10, 10 + 2 * spread), rng.uniform(3, 5),
                            rng.uniform(1.5, 2.5), rng.uniform(-math.pi, math.pi), rng.uniform(0.8, 1.8),
                            rng.uniform(1.4, 1.9)) for _ in range(n)])


def host_core(tmp_dir):
    """Compile tests/native/polygon_host.cpp (the kernel's host+device core, built for the host) and wrap it."""
    so = os.path.join(str(tmp_dir), "polygon_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "native", "polygon_host.cpp")])
    lib = ctypes.CDLL(so)
    vp = ctypes.c_void_p

    def run(a, b, vol=None, list_mode=False):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        M, N = a.shape[0], b.shape[0]
        shape = (M,) if list_mode else (M, N)
        ob, o3 = np.empty(shape), np.empty(shape)
        v = None if vol is None else np.ascontiguousarray(vol, dtype=np.float64)
        lib.polygon_host_iou3d(vp(a.ctypes.data), ctypes.c_long(a.shape[1] * 8), M, vp(b.ctypes.data),
                               ctypes.c_long(b.shape[1] * 8), N, None if v is None else vp(v.ctypes.data),
                               int(list_mode), vp(ob.ctypes.data), vp(o3.ctypes.data))
        return ob, o3
    return run
