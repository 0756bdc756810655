#!/usr/bin/env python
"""bench.py -- GrooMeD-NMS forward+backward throughput (boxes/s) on BASELINE.json's headline configuration:
N=4096 7-DoF boxes per image (32 objects x 128 proposals, overlap = 0.5*(1+GIoU3D approx), group+mask, linear
pruning), fp32.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--images B] [--path materialised|fused]

One step = one forward+backward pass of the hot path over a batch of B independent images per GPU:
7-DoF boxes -> corners -> per-box records -> [N x N overlap matrix ->] sort -> suppression bitmask -> greedy
grouping + rescore + keep lists -> analytic backward (grad wrt scores).  Inputs are resident in HBM for `value`;
`e2e` repeats the measurement through groomed_nms_b200.hostapi.HostRunner.run_host with pinned HOST buffers
(H2D of boxes/scores/upstream grad and D2H of rescored scores / score grads / keep lists inside the timed region).

Multi-GPU (torchrun, one rank per GPU): images are independent (NMS is per image), so every
rank processes its own B images with no data-path collective: weak scaling, value = all ranks' boxes / max time.

`--impl reference` times the reference's own CPU PyTorch path on the host cores (the unmodified sources staged in the
git-ignored baseline/_ref by tools/stage_reference.py, imported under oracle/ref_shim.py), one image per worker process;
if nothing is staged it falls back to the numpy oracle port and says so (`cpu_baseline.kind`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BOXES = 4096
BOX_DOF = 7
METRIC = "groomed_nms_fwd_bwd_boxes_per_s_N4096"


def algorithmic_bytes(n, d):
    """SURVEY.md section 8(d) / BASELINE.md section 3: API-compatible bytes per image of n boxes."""
    return 8 * n * n + (4 * d + 24) * n


def ncu_traffic(images):
    """DRAM bytes (read + write) of one tile-kernel launch from the committed `ncu --set full` capture of this command
    (profiles/r2_tile_kernel_traffic.json), or None if no capture exists for this batch size."""
    path = os.path.join(ROOT, "profiles", "r2_tile_kernel_traffic.json")
    try:
        d = json.load(open(path))
        if int(d["images_per_launch"]) == int(images):
            return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except Exception:
        pass
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------ CPU arm
WORKLOAD = ("C3: N=4096 7-DoF boxes/image, 32 objects x 128 proposals, overlap 0.5*(1+GIoU3D approx), "
            "group+mask, linear pruning, group_size 100, fwd+bwd (grad wrt scores)")
_REF = None


def _ref_init():
    """Worker initialiser: one torch thread per worker process, the UNMODIFIED reference imported under the shim."""
    global _REF
    os.environ["OMP_NUM_THREADS"] = "1"
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import ref_shim
    import torch
    torch.set_num_threads(1)
    _REF = ref_shim.load()


def _ref_one_image(seed):
    """One C3 image through the reference's own CPU PyTorch path, as its loss strings the calls together
    (lib/loss/rpn_3d.py:746-752 corners, :780-781 overlaps, :791 differentiable_nms) + autograd backward."""
    import numpy as np
    import torch
    from groomed_nms_b200 import synthetic
    if _REF is None:
        _ref_init()
    b7, sc = synthetic.config_c3(seed=seed)
    up = torch.from_numpy(np.random.default_rng(seed + 1).standard_normal(N_BOXES).astype(np.float32))
    t = torch.from_numpy(b7)
    s = torch.from_numpy(sc).requires_grad_(True)
    t0 = time.perf_counter()
    corners = _REF.math_3d.get_corners_of_cuboid(x3d=t[:, 0], y3d=t[:, 1], z3d=t[:, 2], w3d=t[:, 3], h3d=t[:, 4], l3d=t[:, 5], ry3d=t[:, 6])
    _, ov = _REF.core.iou3d_approximate(corners, corners, mode="combinations", method="generalized")
    ov = 0.5 * (1 + ov)
    _, _, prob = _REF.groomed_nms.differentiable_nms(scores_unsorted=s, iou_unsorted=ov.clone().detach(), nms_threshold=0.4,
                                                     pruning_method="linear", temperature=0.01, valid_box_prob_threshold=0.3,
                                                     return_sorted_prob=False, group_boxes=True, mask_group_boxes=True, group_size=100)
    prob.backward(up)
    return time.perf_counter() - t0, float(s.grad.sum()) + float(prob.detach().sum())


def _port_init():
    os.environ["OMP_NUM_THREADS"] = "1"
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)


def _port_one_image(seed):
    import numpy as np
    from groomed_nms_b200 import synthetic
    from oracle import groomed_oracle as O
    b7, sc = synthetic.config_c3(seed=seed)
    up = np.random.default_rng(seed + 1).standard_normal(N_BOXES).astype(np.float32)
    t0 = time.perf_counter()
    corners = O.get_corners_of_cuboid(*[b7[:, i] for i in range(7)])
    _, ov = O.iou3d_approximate(corners, corners, "combinations", "generalized")
    ov = (np.float32(0.5) * (np.float32(1) + ov)).astype(np.float32)
    fwd = O.differentiable_nms(sc, ov, dense=False)
    gs, _ = O.differentiable_nms_backward(fwd, up, need_grad_iou=False)
    return time.perf_counter() - t0, float(gs.sum()) + float(fwd["prob"].sum())


def reference_staged():
    from oracle import ref_shim
    return ref_shim.reference_available()


def host_workers():
    """Worker processes for the CPU arm: one per host core, bounded by memory (the reference's N=4096 path holds a few
    dozen N x N fp32 temporaries, about 3 GB per worker)."""
    cores = os.cpu_count() or 1
    try:
        avail_kb = [int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        cores = max(1, min(cores, int(avail_kb / (3.5 * 2 ** 20))))
    except Exception:
        pass
    return max(1, min(cores, 32))


class CpuArm(object):
    """A pool of worker processes timing the CPU implementation of the path on C3 images, one image per worker call.
    kind = "reference": the reference's own torch code (baseline/_ref or /root/reference under oracle/ref_shim.py);
    kind = "port": the numpy oracle restatement (closed-form mode A, about 10x faster than the reference's code)."""

    def __init__(self, kind, workers):
        import multiprocessing as mp
        self.kind, self.workers = kind, workers
        self.fn = _ref_one_image if kind == "reference" else _port_one_image
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        # spawn, not fork: the GPU arm has used CUDA and autograd threads in this process by the time the CPU leg starts
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(workers, initializer=_ref_init if kind == "reference" else _port_init)
        self.next_seed = 3

    def step(self, n_images):
        seeds = [self.next_seed + 10 * i for i in range(n_images)]
        self.next_seed += 10 * n_images
        t0 = time.perf_counter()
        self.pool.map(self.fn, seeds, chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_throughput(kind, n_images, workers, warm_images=0):
    arm = CpuArm(kind, workers)
    if warm_images:
        arm.step(warm_images)
    dt = arm.step(n_images)
    arm.close()
    return n_images * N_BOXES / dt, dt


def run_reference(args, rank):
    """The reference arm: the reference's own CPU implementation on the box's host cores, same workload / metric / unit
    as the GPU arm.  One step = a bounded sample of the GPU arm's step (one image per worker process instead of
    --images); K and W are honoured unless K steps would run past ~4 minutes."""
    if rank != 0:
        return
    kind = "reference" if reference_staged() else "port"
    workers = host_workers()
    per_step = workers
    arm = CpuArm(kind, workers)
    warm = max(1, args.warmup)
    t_w = [arm.step(per_step) for _ in range(warm)]
    steps = max(1, min(args.steps, int(240.0 / max(t_w[-1], 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        arm.step(per_step)
    dt = time.perf_counter() - t0
    arm.close()
    val = steps * per_step * N_BOXES / dt
    what = ("the reference's own lib.math_3d.get_corners_of_cuboid + lib.core.iou3d_approximate + lib.groomed_nms.differentiable_nms "
            "+ autograd backward (torch CPU, unmodified sources under oracle/ref_shim.py)") if kind == "reference" else "numpy oracle port"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "boxes/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_step": per_step,
                   "note": "each step is a bounded sample of the GPU arm's step: one image per host worker process"},
        "cpu_baseline": {"value": val, "unit": "boxes/s", "cores": workers, "kind": kind,
                         "sample": "%d steps x %d images of N=4096, one image per worker process, 1 torch thread each: %s" % (steps, per_step, what)},
        "e2e": {"value": val, "unit": "boxes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
class ClockSampler(object):
    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def time_stage(torch, fn, stream, iters, warm=3):
    """Average device time (ms) of one enqueue of fn(stream_ptr) over `iters` back-to-back launches."""
    import ctypes
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
    return e0.elapsed_time(e1) / iters



def measure_latency(torch, dev, params):
    """Single-call latency (B = 1), microseconds: what one `differentiable_nms` call of the reference's per-image loop costs.
    `c_abi_*`: Nms3dPlan(1, 4096) -- 7-DoF boxes -> records -> forward -> backward as one CUDA graph (device time per replay
    from CUDA events over back-to-back replays; wall time of replay + stream sync); the 2D sizes (N = 1024: config C2, N = 500: what
    the reference's loss passes per image) the same way from boxes and from a materialised IoU matrix.  `python_api_*`: the reference-facing
    lib.groomed_nms.differentiable_nms(scores, iou) on a materialised matrix + autograd backward, wall clock with the one
    sync the variable-length index results need; overlaps through lib.core.iou / ops.overlap3d are timed separately."""
    import numpy as np
    from groomed_nms_b200 import ops, synthetic
    from groomed_nms_b200.hostapi import Nms3dPlan
    from groomed_nms_b200.lib import core as C
    from groomed_nms_b200.lib import groomed_nms as G
    out = {}

    def wall(fn, n=40, warm=8):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        return 1e6 * ts[len(ts) // 2]

    b7, sc = synthetic.config_c3(seed=3)
    for name, mat in (("c_abi_graph_N4096_3d_matrix_free", False), ("c_abi_graph_N4096_3d_with_matrix", True)):
        pl = Nms3dPlan(1, N_BOXES, dev, params, materialise=mat)
        pl.boxes7.copy_(torch.from_numpy(b7)[None]); pl.scores.copy_(torch.from_numpy(sc)[None]); pl.grad_prob.normal_()
        g = pl.capture()
        for _ in range(10):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(100):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out[name] = {"device_us_per_call": 1e3 * e0.elapsed_time(e1) / 100, "wall_us_per_call": wall(g.replay)}
        del pl, g
    rec = ops.box3d_records(ops.corners_from_boxes7(torch.from_numpy(b7).to(dev)))
    bx2, sc2 = synthetic.config_c2()
    bx5, sc5, _ = synthetic.clustered_boxes_2d(500, 6, seed=11, jitter=0.05)
    # the 2D sizes through the C-ABI as one CUDA graph: forward from the boxes (fused) or from a materialised matrix (the
    # reference's call form), + backward
    from groomed_nms_b200 import _lib as L_
    for tag, bx, scv in (("N1024_2d", bx2, sc2), ("N500_2d", bx5, sc5)):
        for form in ("from_boxes", "from_matrix"):
            try:
                tb, ts_ = torch.from_numpy(bx).to(dev)[None].contiguous(), torch.from_numpy(scv).to(dev)[None].contiguous()
                up2 = torch.randn_like(ts_)
                iou2 = ops.overlap2d(tb[0], tb[0])[None].contiguous() if form == "from_matrix" else None

                def step2():
                    st2 = ops.forward_matrix(ts_, iou2, params) if iou2 is not None else ops.forward_boxes(ts_, tb, L_.BOX_2D, params)
                    ops.backward(st2, up2)
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for _ in range(3):
                        step2()
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize()
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2):
                    step2()
                for _ in range(10):
                    g2.replay()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(100):
                    g2.replay()
                e1.record()
                torch.cuda.synchronize()
                out["c_abi_graph_%s_%s" % (tag, form)] = {"device_us_per_call": 1e3 * e0.elapsed_time(e1) / 100, "wall_us_per_call": wall(g2.replay)}
                del g2
            except Exception as e:                                    # (never let an extra take the bench line down)
                out["c_abi_graph_%s_%s" % (tag, form)] = {"error": "%s: %s" % (type(e).__name__, e)}
    cases = (("python_api_N4096_3d", torch.from_numpy(sc).to(dev), lambda: ops.overlap3d(rec, rec, False, True, generalized=True, affine=True)[1]),
             ("python_api_N1024_2d", torch.from_numpy(sc2).to(dev), lambda b=torch.from_numpy(bx2).to(dev): C.iou(b, b)),
             ("python_api_N500_2d", torch.from_numpy(sc5).to(dev), lambda b=torch.from_numpy(bx5).to(dev): C.iou(b, b)))
    for name, scores, make_iou in cases:
        iou = make_iou()
        up = torch.randn_like(scores)

        def call():
            s = scores.clone().requires_grad_(True)
            valid, invalid, prob = G.differentiable_nms(s, iou, nms_threshold=0.4, pruning_method="linear", temperature=0.01,
                                                       valid_box_prob_threshold=0.3, group_boxes=True, mask_group_boxes=True, group_size=100)
            prob.backward(up)
        out[name] = {"wall_us_fwd_bwd": wall(call), "wall_us_overlap_matrix": wall(make_iou)}
    return out


def measure_sustained(torch, graph, boxes_per_step, seconds=2.5):
    """The step's CUDA graph replayed back to back for `seconds` (the headline `value` is a ~20 ms burst): boxes/s over the
    whole interval, with the SM clock seen meanwhile."""
    for _ in range(5):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    n = 0
    while time.perf_counter() - t0 < seconds:
        for _ in range(50):
            graph.replay()
        n += 50
        torch.cuda.current_stream().synchronize() if n % 500 == 0 else None
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"steps": n, "seconds": ms * 1e-3, "ms_per_step": ms / n, "value": boxes_per_step * n / (ms * 1e-3)}


def run_ours(args, rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    from groomed_nms_b200 import _lib, ops, synthetic
    from groomed_nms_b200.hostapi import HostRunner, Nms3dPlan
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU fallback (use --impl reference for the CPU arm)")
    lib = _lib.load()
    matrix_kernel = {"auto": _lib.MATRIX_KERNEL_AUTO, "direct": _lib.MATRIX_KERNEL_DIRECT, "tma": _lib.MATRIX_KERNEL_TMA}[args.matrix_kernel]
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from groomed_nms_b200.hostapi import bind_host_to_gpu
    orig_affinity = os.sched_getaffinity(0)
    host_binding = bind_host_to_gpu(local)          # before any pinned allocation: first touch puts the buffers on the GPU's node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N = args.images, N_BOXES
    params = ops.make_params(nms_threshold=0.4, pruning_method="linear", temperature=0.01, valid_box_prob_threshold=0.3,
                             group_boxes=True, mask_group_boxes=True, group_size=100)
    # synthetic inputs: B independent C3 draws per rank (distinct seeds per rank/image)
    boxes = np.stack([synthetic.config_c3(seed=3 + 10 * (rank * B + i))[0] for i in range(B)])
    scores = np.stack([synthetic.config_c3(seed=3 + 10 * (rank * B + i))[1] for i in range(B)])
    grads = np.random.default_rng(1234 + rank).standard_normal((B, N)).astype(np.float32)

    plans = {}
    for name, mat in (("materialised", True), ("fused", False)):
        pl = Nms3dPlan(B, N, dev, params, materialise=mat, overlap_branch=(mat and not args.no_overlap_branch))
        pl.matrix_opts = _lib.launch_opts(matrix_kernel=matrix_kernel, tiles_per_cta=args.tiles_per_cta if pl.overlap_branch else 0,
                                          flags=0 if args.packed else _lib.OPT_SCALAR_MATH)
        pl.forward_opts = _lib.launch_opts(matrix_kernel=matrix_kernel, flags=0 if args.packed else _lib.OPT_SCALAR_MATH)
        pl.boxes7.copy_(torch.from_numpy(boxes)); pl.scores.copy_(torch.from_numpy(scores)); pl.grad_prob.copy_(torch.from_numpy(grads))
        plans[name] = pl
    torch.cuda.synchronize()

    def timed_steps(plan, steps, warmup):
        graph = plan.capture()
        for _ in range(warmup):
            graph.replay()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    if args.splits > 1:
        from groomed_nms_b200.hostapi import SplitPlan
        for name, mat in (("materialised", True), ("fused", False)):
            sp = SplitPlan(B, N, dev, params, materialise=mat, splits=args.splits)
            sp.load(torch.from_numpy(boxes).to(dev), torch.from_numpy(scores).to(dev), torch.from_numpy(grads).to(dev))
            plans[name + "_split"] = sp
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sfx = "_split" if args.splits > 1 else ""
    head = plans[args.path + sfx]
    ms_total = timed_steps(head, args.steps, args.warmup)
    other_name = "fused" if args.path == "materialised" else "materialised"
    ms_other = timed_steps(plans[other_name + sfx], args.steps, args.warmup)

    # e2e: same step through the host-buffer API (pinned host inputs/outputs, every step copies all of its inputs in
    # and all of its results out inside the timed region).  Two variants: one blocking call per step (run_host), and
    # the streaming pipeline (HostPipeline, 2 slots) in which the copies of one call overlap the kernels of the next.
    from groomed_nms_b200.hostapi import HostPipeline
    # run_host returns no matrix (fused matrix-free pipeline).  The upstream gradient dL/dprob stays on the device, where a
    # training step's loss kernels produce it; keep lists come back as int32 [B, 512] (+ the true counts)
    runner = HostRunner(B, N, dev, params, materialise=False, keep_cap=512, grad_on_device=True)
    runner.set_grad(torch.from_numpy(grads).to(dev))
    hb = torch.from_numpy(boxes).pin_memory(); hs = torch.from_numpy(scores).pin_memory(); hg = torch.from_numpy(grads).pin_memory()
    e2e_steps = max(args.steps, 200)          # ~40 ms of wall clock: short runs are at the mercy of host scheduling noise

    def time_host(fn, finish):
        for _ in range(max(3, args.warmup)):
            fn()
        finish()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        finish()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    e2e_blocking_s = time_host(lambda: runner.run_host(hb, hs), lambda: None)
    pipe = HostPipeline(B, N, dev, params, depth=args.e2e_depth, keep_cap=512, grad_on_device=True)
    pipe.set_grad(torch.from_numpy(grads).to(dev))
    # host-side timing of a copy-bound pipeline is at the mercy of whatever else the box's host does: two takes, both reported
    # (e2e.samples), the faster one is the value
    e2e_takes = [time_host(lambda: pipe.submit(hb, hs), pipe.drain) for _ in range(2)]
    e2e_s = min(e2e_takes)
    # the same calls with the kernels left out: what the PCIe link of this box allows for this call pattern
    copies_s = time_host(lambda: pipe.submit(hb, hs, copies_only=True), pipe.drain)
    # round-1 call pattern for comparison: the upstream gradient crosses PCIe too and keep lists are int64 [B, N]
    pipe_r1 = HostPipeline(B, N, dev, params, depth=args.e2e_depth, keep_cap=0, grad_on_device=False)
    e2e_r1_s = time_host(lambda: pipe_r1.submit(hb, hs, hg), pipe_r1.drain)
    r1_bytes = (pipe_r1.h2d_bytes, pipe_r1.d2h_bytes)
    del pipe_r1
    clocks = sampler.stop() if rank == 0 else None
    os.sched_setaffinity(0, orig_affinity)          # the CPU legs below use every core the process was given

    # per-kernel device times (rank 0): each stage launched back to back on its own; the working set of the N^2
    # stages (B x 64 MiB) exceeds the 126 MB L2 for B >= 2, so these are HBM-resident timings
    stages = {}
    if rank == 0:
        st = torch.cuda.current_stream(dev)
        pl = plans["materialised"]
        it = max(10, args.steps)
        stages["records_from_boxes7(corners in registers)"] = time_stage(torch, pl.stage_front, st, it)
        stages["forward_boxes+matrix_out(all kernels, one stream)"] = time_stage(torch, pl.stage_forward, st, it)
        # the kernels of that forward one by one (debug stage mask of the library: the same launches, in isolation)
        keep = pl.forward_opts
        for bit, name in ((_lib.STAGE_RANK, "sort_kernel+rank_kernel"), (_lib.STAGE_ELECT, "elect2_kernel(batched leader election)"),
                          (_lib.STAGE_TILES, "tile_kernel(matrix only)"), (_lib.STAGE_CHAIN, "chain_kernel(on its own; the step runs it at the end of elect2_kernel)")):
            # the same launches one at a time: a per-call stage mask (gnms_launch_opts), tiles_per_cta as in the step
            pl.forward_opts = _lib.launch_opts(matrix_kernel=matrix_kernel, stage_mask=bit, flags=keep.flags,
                                               tiles_per_cta=args.tiles_per_cta if pl.overlap_branch else 0)
            stages[name] = time_stage(torch, pl.stage_forward, st, it)
        pl.forward_opts = keep
        stages["backward"] = time_stage(torch, pl.stage_backward, st, it)
        stages["forward_boxes_no_matrix(all kernels)"] = time_stage(torch, plans["fused"].stage_forward, st, it)
        tk = Nms3dPlan(B, N, dev, params, materialise=True, two_kernel=True)
        tk.boxes7.copy_(pl.boxes7); tk.scores.copy_(pl.scores); tk.grad_prob.copy_(pl.grad_prob)
        tk.stage_corners(__import__("ctypes").c_void_p(st.cuda_stream)); tk.stage_records(__import__("ctypes").c_void_p(st.cuda_stream))
        stages["two_kernel:overlap3d_matrix"] = time_stage(torch, tk.stage_overlap, st, it)
        stages["two_kernel:forward_from_matrix(rank+mask+chain)"] = time_stage(torch, tk.stage_forward, st, it)
        del tk
    latency, sustained = None, None
    if rank == 0 and not args.no_extras:
        latency = measure_latency(torch, dev, params)
        smp = ClockSampler(local)
        smp.start()
        sustained = measure_sustained(torch, head.capture(), B * N)
        sustained["clocks"] = smp.stop()

    if world > 1:
        dist.barrier()
    if rank == 0:
        peak, peak_src = measured_peaks()
        ms_step = ms_total / args.steps
        boxes_per_step = B * N * world
        value = boxes_per_step / (ms_step * 1e-3)
        step_bytes = algorithmic_bytes(N, BOX_DOF) * B
        # dominant kernel of the materialised path: the N x N overlap tile kernel (writes 4 N^2 per image) or the
        # matrix -> bitmask stream (reads 4 N^2 per image); algorithmic bytes per launch stated in DESIGN.md
        k_ms = stages["tile_kernel(matrix only)"]
        k_bytes = B * (4 * N * N + 32 * N)      # matrix written once + box records read
        ach = k_bytes / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "boxes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "images_per_step_per_gpu": B, "path": args.path, "cuda_graph": True, "graph_branches": (2 if (args.path == "materialised" and not args.no_overlap_branch) else 1) * args.splits,
                       "l2": "no flush: per-step working set %.0f MiB of overlap matrices vs 126 MB L2" % (B * N * N * 4 / 2 ** 20)
                             if args.path == "materialised" else "matrix-free path: no N^2 HBM traffic; inputs are O(N)",
                       "parallelism": "per-image shard, %d rank(s), no data-path collective" % world},
            "e2e": {"value": boxes_per_step * e2e_steps / e2e_s, "unit": "boxes/s", "h2d_bytes_per_step": runner.h2d_bytes,
                    "d2h_bytes_per_step": runner.d2h_bytes, "samples": [boxes_per_step * e2e_steps / t for t in e2e_takes], "api": "groomed_nms_b200.hostapi.HostPipeline.submit/drain (pinned host buffers, %d slots: copies of one call overlap the kernels of the next; fused matrix-free pipeline: boxes and scores in; probabilities, score gradients, int32 keep lists [B,512] and counts out; the upstream gradient dL/dprob is device-resident)" % args.e2e_depth,
                    "h2d_GBps_per_rank": runner.h2d_bytes * e2e_steps / e2e_s / 1e9,
                    "with_host_gradient_and_int64_keep_lists": {"value": boxes_per_step * e2e_steps / e2e_r1_s, "h2d_bytes_per_step": r1_bytes[0], "d2h_bytes_per_step": r1_bytes[1]},
                    "steps_timed": e2e_steps,
                    "copies_only_value": boxes_per_step * e2e_steps / copies_s,
                    "copies_only_note": "the same pipeline with the kernels left out (copies only): the ceiling the PCIe link of this box sets for e2e (%.1f GB/s H2D)" % (runner.h2d_bytes * e2e_steps / copies_s / 1e9),
                    "blocking_call_value": boxes_per_step * e2e_steps / e2e_blocking_s,
                    "blocking_call_api": "groomed_nms_b200.hostapi.HostRunner.run_host (one synchronous call per step)"},
            "gpu_launches": head.launches_per_step * args.steps,
            "clocks": clocks,
            "host_binding": host_binding,
            "roofline": {"bound": "hbm", "kernel": ("gnms::tile_tall_kernel<3D records, generalized, affine, packed fp32x2> (symmetric 256x64 overlap tiles "
                                                    "stored direct + mirrored straight from registers, %d images per launch)" if args.matrix_kernel == "direct" else
                                                    "gnms::tile_tma_kernel<3D records, generalized, affine, packed fp32x2> (symmetric 256x64 overlap tiles, each 32x32 block "
                                                    "and its mirror image staged in swizzled shared memory and written by TMA tensor stores, %d images per launch)") % B,
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": k_bytes, "avg_launch_ms": k_ms, "traffic": ncu_traffic(B),
                         "timing": "CUDA events around %d back-to-back launches of this kernel alone on the launching stream" % max(10, args.steps),
                         "note": "fp32 instruction issue limits the kernel (41 warp instructions per 32 pairs, 30 % of them FMNMX on the half-rate "
                                 "ALU pipe, 4 warps per sub-partition), not HBM: see DESIGN.md section 5 and profiles/r2_ncu_tile_tma_kernel_b64.txt"},
            "step_roofline": {"algorithmic_bytes_per_step": step_bytes, "effective_GBps": step_bytes / (ms_step * 1e-3) / 1e9,
                              "frac_of_peak": step_bytes / (ms_step * 1e-3) / 1e9 / peak,
                              "note": "section 8(d) byte model 8N^2+(4D+24)N per image over the whole fwd+bwd step"},
            "other_path": {"path": other_name, "ms_per_step": ms_other / args.steps,
                           "value": boxes_per_step / (ms_other / args.steps * 1e-3),
                           "effective_frac_of_peak": step_bytes / (ms_other / args.steps * 1e-3) / 1e9 / peak},
            "stage_ms": stages,
            "extra": {"latency_us": latency, "sustained": sustained},
        }
        if not args.no_cpu and world == 1:                      # (the CPU baseline is an N = 1 figure: the other ranks' processes share these cores)
            workers = host_workers()
            kind = "reference" if reference_staged() else "port"
            v, dt = cpu_throughput(kind, workers, workers, warm_images=workers)       # one warm image + one timed image per worker
            line["cpu_baseline"] = {"value": v, "unit": "boxes/s", "cores": workers, "kind": kind,
                                    "sample": "%d images of N=4096 of the same workload, one per worker process (1 torch thread each), after one warm-up image each: "
                                              "%s, %.1f s" % (workers, "the reference's own torch CPU code (corners + iou3d_approximate + differentiable_nms + autograd backward)"
                                                              if kind == "reference" else "numpy oracle port", dt)}
            if kind == "reference":
                vp, dtp = cpu_throughput("port", 4 * workers, workers)
                line["cpu_baseline_port"] = {"value": vp, "unit": "boxes/s", "cores": workers, "kind": "port",
                                             "sample": "%d images, numpy oracle restatement (closed-form mode A), %.1f s" % (4 * workers, dtp)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()



# ------------------------------------------------------------------------------------------------ C4: batched images + NCCL all-reduce
def c4_inputs(images, n, np, synthetic):
    """Seeded inputs of the listed global image ids: 2D boxes (16 clusters), 64 features per box, upstream gradient."""
    boxes = np.stack([synthetic.config_c4_image(i, n=n)[0] for i in images])
    feats = np.stack([np.random.default_rng(7000 + i).standard_normal((n, 64)).astype(np.float32) for i in images])
    grads = np.stack([np.random.default_rng(9000 + i).standard_normal(n).astype(np.float32) for i in images])
    wb = (np.random.default_rng(42).standard_normal(65) * 0.3).astype(np.float32)
    return boxes, feats, grads, wb


def run_c4(args, rank, world):
    """BASELINE.json configs[3]: batch = 32 images x N = 2048 2D boxes, images sharded over the ranks (shard_range), scores from
    a shared 64 -> 1 head, GrooMeD-NMS forward + backward, ONE NCCL all-reduce of the flat gradient bucket per step.
    value = strong scaling (the 32 images divided over the ranks); extra.weak = 4 images per rank whatever the rank count."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from groomed_nms_b200 import _lib, ops, synthetic
    from groomed_nms_b200.hostapi import BatchedHeadPlan
    from groomed_nms_b200.sharding import shard_range
    _lib.load()
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, TOTAL = 2048, 32
    params = ops.make_params(nms_threshold=0.4, pruning_method="linear", temperature=0.01, valid_box_prob_threshold=0.3,
                             group_boxes=True, mask_group_boxes=True, group_size=100)
    pad = args.bucket_mb * (1 << 20) // 4

    def make_plan(images):
        boxes, feats, grads, wb = c4_inputs(images, N, np, synthetic)
        pl = BatchedHeadPlan(len(images), N, dev, params, materialise=not args.c4_matrix_free, bucket_pad_elems=pad, collective=args.c4_collective)
        pl.boxes.copy_(torch.from_numpy(boxes)); pl.x.copy_(torch.from_numpy(feats)); pl.grad_prob.copy_(torch.from_numpy(grads))
        pl.wb.copy_(torch.from_numpy(wb))
        return pl

    def timed(graph, steps, warmup):
        for _ in range(warmup):
            graph.replay()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    res = {}
    for label, images in (("strong", list(shard_range(TOTAL, world, rank))), ("weak", [4 * rank + i for i in range(4)])):
        pl = make_plan(images)
        try:
            g_full = pl.capture(reduce=True)
            graphed = True
        except Exception as e:                               # NCCL in a graph unavailable: time eager launches instead
            graphed = False

            class Eager(object):
                def __init__(self, f): self.f = f
                def replay(self): self.f()
            g_full = Eager(lambda pl=pl: pl.step(None, True))
        ms_full = timed(g_full, args.steps, args.warmup)
        ms_compute = timed(pl.capture(reduce=False), args.steps, args.warmup)

        class ReduceOnly(object):
            def __init__(self, pl): self.pl = pl
            def replay(self): self.pl.bucket.all_reduce()
        ms_reduce = timed(ReduceOnly(pl), args.steps, args.warmup) if world > 1 else 0.0
        res[label] = dict(images_per_rank=len(images), ms_per_step=ms_full, ms_compute_only=ms_compute, ms_nccl_allreduce_alone=ms_reduce,
                          collective=pl.collective, peer_exchange_ok=(int(pl.exchange.status.item()) == 0) if pl.exchange is not None else None,
                          graphed=graphed, boxes_per_step=(TOTAL if label == "strong" else 4 * world) * N, bucket_bytes=pl.bucket.nbytes,
                          launches=pl.launches_per_step)
        # the reduced gradient must not depend on how the images were sharded (checked against rank 0's own full computation)
        if label == "strong":
            pl.step(None, True)
            torch.cuda.synchronize()
            got = pl.grad_wb.clone()
            if world > 1:
                full = make_plan(list(range(TOTAL)))
                full.step(None, False)
                torch.cuda.synchronize()
                err = float((got - full.grad_wb).abs().max() / full.grad_wb.abs().max().clamp_min(1e-30))
                res[label]["allreduced_grad_rel_err_vs_single_process"] = err
                del full
        pl.close()
        del pl
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        peak, peak_src = measured_peaks()
        st = res["strong"]
        value = st["boxes_per_step"] / (st["ms_per_step"] * 1e-3)
        bytes_step = TOTAL * algorithmic_bytes(N, 4)
        line = {
            "metric": "groomed_nms_fwd_bwd_boxes_per_s_batched_N2048", "value": value, "unit": "boxes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": st["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C4: batch 32 images x N=2048 2D boxes (16 clusters each), shared 64->1 score head, group+mask linear, fwd+bwd, "
                                   "one all-reduce of the gradient bucket per step", "images_total": TOTAL,
                       "images_per_rank": st["images_per_rank"], "matrix": "materialised" if not args.c4_matrix_free else "matrix-free",
                       "bucket_bytes": st["bucket_bytes"],
                       "collective": ("all-reduce inside the head-gradient kernel over NVLink peer memory (gnms_score_head_backward_allreduce_f32: CUDA IPC buffers, "
                                      "release / acquire flags, rank-ordered sum), captured in the step's CUDA graph; ncclAllReduce on the same bucket timed beside it "
                                      "(extra.*.ms_nccl_allreduce_alone)") if st["collective"] == "peer" else
                                     ("ncclAllReduce(sum, fp32) via torch.distributed (NCCL), captured in the step's CUDA graph"
                                      if st["graphed"] else "ncclAllReduce(sum, fp32) via torch.distributed (NCCL), eager"),
                       "parallelism": "per-image shard over %d rank(s)" % world},
            "gpu_launches": st["launches"] * args.steps, "clocks": clocks,
            "step_roofline": {"algorithmic_bytes_per_step": bytes_step, "effective_GBps": bytes_step / (st["ms_per_step"] * 1e-3) / 1e9 / 1.0,
                              "frac_of_peak_per_gpu": bytes_step / world / (st["ms_per_step"] * 1e-3) / 1e9 / peak, "peak": peak, "peak_source": peak_src},
            "extra": {"strong": st, "weak": dict(res["weak"], value=res["weak"]["boxes_per_step"] / (res["weak"]["ms_per_step"] * 1e-3))},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # CUDA graphs that captured NCCL work are still alive here; tearing the process group down underneath them hung
        # (observed at 2 ranks, NCCL 2.28.9 / torch 2.11): leave together and skip the destructors
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=64, help="images (of N=4096 boxes) per GPU per step")
    ap.add_argument("--e2e-depth", type=int, default=3, help="slots of the host pipeline (copies of one call overlap the kernels of the others)")
    ap.add_argument("--path", default="materialised", choices=["materialised", "fused"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the single-call latency and sustained-throughput legs")
    ap.add_argument("--no-overlap-branch", action="store_true", help="matrix kernel and NMS kernels on one stream instead of two graph branches")
    ap.add_argument("--tiles-per-cta", type=int, default=4, help="matrix-only tile kernel on its branch: tiles per CTA (0 = persistent)")
    ap.add_argument("--packed", type=int, default=1, help="packed fp32x2 arithmetic in the matrix-only kernel (0 = scalar)")
    ap.add_argument("--matrix-kernel", default="auto", choices=["auto", "direct", "tma"],
                    help="how the overlap matrix leaves the SM: register-direct STG or shared-memory staging + TMA tensor stores")
    ap.add_argument("--config", default="c3", choices=["c3", "c4"], help="c3 = the headline N=4096 7-DoF workload; c4 = batch 32 x N=2048 2D "
                                                                          "images sharded over the ranks with one NCCL gradient all-reduce per step")
    ap.add_argument("--bucket-mb", type=int, default=0, help="c4: pad the gradient bucket to this many MiB (48 = the reference model's size)")
    ap.add_argument("--c4-matrix-free", action="store_true", help="c4: do not materialise the [N,N] IoU matrices")
    ap.add_argument("--c4-collective", default="auto", choices=["auto", "peer", "nccl"],
                    help="c4: all-reduce inside the head-gradient kernel over NVLink peer memory (auto: when the bucket is not padded) or NCCL")
    ap.add_argument("--splits", type=int, default=1, help="issue the batch as this many sub-batches on parallel graph branches")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
    elif args.config == "c4":
        run_c4(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
