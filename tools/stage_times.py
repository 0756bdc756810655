#!/usr/bin/env python
"""Per-kernel device times of the forward on the bench workload (run on the GPU box):
    python tools/stage_times.py [--images 16] [--n 4096]
Every stage of gnms_forward_boxes_f32 is launched back to back on its own (debug stage mask of the library), for the
matrix-producing and the matrix-free pipeline, and the two pipelines are checked against each other."""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from groomed_nms_b200 import _lib, ops, synthetic          # noqa: E402
from groomed_nms_b200.hostapi import Nms3dPlan             # noqa: E402

STAGES = {1: "rank", 32: "elect", 3: "rank+spatial", 4: "tile", 8: "has_earlier", 16: "chain", 0xff: "forward (all)"}


def time_call(fn, stream, iters=20, warm=3):
    s = ctypes.c_void_p(stream.cuda_stream)
    for _ in range(warm):
        fn(s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream.synchronize()
    e0.record(stream)
    for _ in range(iters):
        fn(s)
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--rank-by-sort", type=int, default=-1, help="-1 library heuristic, 0 counting kernel, 1 per-image radix sort")
    ap.add_argument("--tile-queue", type=int, default=1, help="1: hits go through a shared-memory queue drained by the whole CTA; 0: inline loop")
    ap.add_argument("--election", type=int, default=0, help="0 library heuristic, 1 direct leader election, 2 suppression-bit route")
    args = ap.parse_args()
    lib = _lib.load()
    rank_method = {-1: _lib.RANK_AUTO, 0: _lib.RANK_COUNT, 1: _lib.RANK_SORT}[args.rank_by_sort]
    base_flags = 0 if args.tile_queue else _lib.OPT_INLINE_HITS

    def opts(mask=0):
        return _lib.launch_opts(rank_method=rank_method, election=args.election, stage_mask=mask, flags=base_flags)
    dev = torch.device("cuda", 0)
    B, N = args.images, args.n
    params = ops.make_params()
    boxes = np.stack([synthetic.config_c3(seed=3 + 10 * i, n=N)[0] for i in range(B)])
    scores = np.stack([synthetic.config_c3(seed=3 + 10 * i, n=N)[1] for i in range(B)])
    st = torch.cuda.current_stream(dev)
    s = ctypes.c_void_p(st.cuda_stream)
    res = {}
    for mat in (True, False):
        pl = Nms3dPlan(B, N, dev, params, materialise=mat)
        pl.forward_opts = opts()
        pl.boxes7.copy_(torch.from_numpy(boxes)); pl.scores.copy_(torch.from_numpy(scores))
        pl.grad_prob.normal_()
        pl.stage_corners(s); pl.stage_records(s); pl.stage_forward(s)
        torch.cuda.synchronize()
        res[mat] = dict(prob=pl.prob.clone(), counts=pl.counts.clone(), lead=pl.lead.clone())
        if not mat:
            nl = (pl.lead == torch.arange(N, device=dev)[None, :]).sum(dim=1)
            print("leaders per image: min %d mean %.1f max %d" % (int(nl.min()), float(nl.float().mean()), int(nl.max())))
        for m, nm in STAGES.items():
            if m in (3, 32) and mat:
                continue
            pl.forward_opts = opts(m)
            us = time_call(pl.stage_forward, st)
            extra = "  -> %.0f GB/s of matrix written" % (B * 4.0 * N * N / us / 1e3) if (m == 4 and mat) else ""
            print("matrix=%d  %-14s %9.1f us%s" % (mat, nm, us, extra))
        pl.forward_opts = opts()
        del pl
    print("matrix vs matrix-free:", " ".join("%s=%s" % (k, torch.equal(res[True][k], res[False][k])) for k in ("prob", "counts", "lead")))


if __name__ == "__main__":
    main()
