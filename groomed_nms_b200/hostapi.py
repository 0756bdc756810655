"""Allocation-free runner for the GrooMeD-NMS hot path on a batch of images + the host-buffer entry point.

`Nms3dPlan` owns every device buffer a forward+backward step needs for a fixed (batch, N) and enqueues the
sm_100a kernels through the C-ABI with raw pointers, so a whole step (7-DoF boxes -> corners -> records ->
[overlap matrix ->] GrooMeD-NMS forward -> analytic backward) is 5-7 stream-ordered launches with no host sync
and can be captured in a CUDA graph.  `run_host` is the call a user with HOST (pinned) buffers makes: it copies
boxes/scores/upstream-gradient in, runs the step, and copies the rescored scores, the score gradients and the
keep lists back."""
import ctypes

import torch

from . import _lib
from ._lib import Saved, check


def _vp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class Nms3dPlan(object):
    """Fixed-shape plan: `batch` images of `n` 7-DoF boxes each on `device`.

    materialise=True  : the [n,n] overlap matrix 0.5*(1+GIoU3D) of every image is an OUTPUT (the API-visible product
                        of iou3d_approximate): the fused tile kernel streams it to HBM while the same register-
                        resident values feed the grouping stage, so it is written once and never read back.
    materialise=False : same kernels, matrix not requested (nothing n^2 touches HBM).
    two_kernel=True   : the reference's call sequence instead: overlap kernel writes the matrix, then
                        gnms_forward_f32 streams it back (8*n^2 bytes of matrix traffic); kept for comparison."""

    def __init__(self, batch, n, device, params, materialise=True, box_dof=7, two_kernel=False, overlap_branch=None):
        self.lib = _lib.load()
        self.B, self.N, self.dev, self.params, self.materialise = batch, n, device, params, materialise
        self.two_kernel = two_kernel and materialise
        # overlap_branch: the matrix does not depend on the NMS half of the step (leaders are elected directly from the
        # boxes), so it is written by the matrix-only tile kernel on its own stream / graph branch while the NMS kernels
        # run on a higher-priority one.  matrix_opts.tiles_per_cta > 0 makes the matrix kernel's CTAs retire continuously so
        # that the NMS kernels find SM slots.
        if overlap_branch is None:
            overlap_branch = materialise and not two_kernel
        self.overlap_branch = bool(overlap_branch) and materialise and not self.two_kernel
        # per-call launch options of the matrix kernel / the forward (gnms_launch_opts: nothing process-wide is touched)
        self.matrix_opts = _lib.launch_opts(tiles_per_cta=4 if self.overlap_branch else 0)
        self.forward_opts = None
        if self.overlap_branch:
            self.side = torch.cuda.Stream(device, priority=0)       # matrix branch (default priority)
            self.hi = torch.cuda.Stream(device, priority=-1)        # NMS branch
        f = dict(dtype=torch.float32, device=device)
        self.boxes7 = torch.zeros((batch, n, box_dof), **f)
        self.scores = torch.zeros((batch, n), **f)
        self.grad_prob = torch.zeros((batch, n), **f)
        self.corners = torch.empty((batch * n, 3, 8), **f)
        self.rec = torch.empty((batch * n, 8), **f)
        self.overlap = torch.empty((batch, n, n), **f) if materialise else None
        self.prob = torch.empty((batch, n), **f)
        self.grad_scores = torch.empty((batch, n), **f)
        self.valid_idx = torch.empty((batch, n), dtype=torch.int64, device=device)
        self.invalid_idx = torch.empty((batch, n), dtype=torch.int64, device=device)
        self.counts = torch.empty((batch, 2), dtype=torch.int32, device=device)
        self.order = torch.empty((batch, n), dtype=torch.int32, device=device)
        self.lead = torch.empty((batch, n), dtype=torch.int32, device=device)
        self.fl = torch.empty((4, batch, n), **f)
        self.ws = torch.empty((int(self.lib.gnms_workspace_bytes(n, batch)),), dtype=torch.uint8, device=device)
        self.saved = Saved(_vp(self.order), _vp(self.fl[0]), _vp(self.lead), _vp(self.fl[1]), _vp(self.fl[2]), _vp(self.fl[3]))
        # matrix-free:     records (from the 7-DoF boxes), sort, rank, elect2 (batched leader election with the chain at its end), backward
        # matrix produced: the same plus the matrix-only tile kernel
        # (batches below ~10 images of N = 4096 rank by counting: one launch fewer; two_kernel: overlap, rank, mask, chain, backward)
        by_sort = float(batch) * n * n >= 10.0 * 4096 * 4096
        self.launches_per_step = (6 if by_sort else 5) if two_kernel else (4 if by_sort else 3) + 1 + (1 if materialise else 0)

    # -- individual stages (each is one C-ABI call = one kernel launch unless noted)
    def stage_corners(self, s):
        check(self.lib.gnms_corners_from_boxes7_f32(_vp(self.boxes7), self.boxes7.stride(1), self.B * self.N, _vp(self.corners), s), "corners")

    def stage_records(self, s):
        check(self.lib.gnms_box3d_records_f32(_vp(self.corners), self.B * self.N, _vp(self.rec), 0, s), "records")

    def stage_overlap(self, s):          # one launch over all images
        check(self.lib.gnms_overlap3d_batched_ex_f32(_vp(self.rec), self.N, self.B, _vp(self.overlap), 1, 1,
                                                     _lib.opts_ref(self.matrix_opts), s), "overlap3d_batched")

    def stage_forward_no_matrix(self, s):
        p = ctypes.byref(self.params)
        check(self.lib.gnms_forward_boxes_ex_f32(_vp(self.scores), _vp(self.rec), _lib.BOX_3D_REC, 1, 1, self.N, self.B, None, p,
                                                 None, _vp(self.prob), _vp(self.valid_idx), _vp(self.invalid_idx),
                                                 _vp(self.counts), self.saved, _vp(self.ws), _lib.opts_ref(self.forward_opts), s),
              "forward_boxes")

    def stage_forward(self, s):          # sort + mask + chain: 3 launches
        p = ctypes.byref(self.params)
        if self.two_kernel:
            check(self.lib.gnms_forward_ex_f32(_vp(self.scores), _vp(self.overlap), self.N, self.N, self.B, None, p, _vp(self.prob),
                                               _vp(self.valid_idx), _vp(self.invalid_idx), _vp(self.counts), self.saved,
                                               _vp(self.ws), _lib.opts_ref(self.forward_opts), s), "forward")
        else:
            check(self.lib.gnms_forward_boxes_ex_f32(_vp(self.scores), _vp(self.rec), _lib.BOX_3D_REC, 1, 1, self.N, self.B, None, p,
                                                     _vp(self.overlap), _vp(self.prob), _vp(self.valid_idx), _vp(self.invalid_idx),
                                                     _vp(self.counts), self.saved, _vp(self.ws), _lib.opts_ref(self.forward_opts), s),
                  "forward_boxes")

    def stage_backward(self, s):
        check(self.lib.gnms_backward_f32(_vp(self.grad_prob), _vp(self.prob), _vp(self.overlap), self.N, self.N, self.B, None,
                                         ctypes.byref(self.params), self.saved, _vp(self.grad_scores), None, self.N,
                                         _vp(self.ws), s), "backward")

    def stage_front(self, s):            # 7-DoF boxes -> records in one launch (the corners stay in registers)
        check(self.lib.gnms_box3d_records_from_boxes7_f32(_vp(self.boxes7), self.boxes7.stride(1), self.B * self.N, _vp(self.rec),
                                                          None, s), "records_from_boxes7")

    def step(self, stream=None):
        """Enqueue one forward+backward pass over the batch on `stream` (default: torch's current stream)."""
        st = stream if stream is not None else torch.cuda.current_stream(self.dev)
        s = ctypes.c_void_p(st.cuda_stream)
        if self.two_kernel:
            self.stage_corners(s)
            self.stage_records(s)
        else:
            self.stage_front(s)
        if self.overlap_branch:
            # fork: matrix on the side stream, NMS forward + backward on the high-priority stream, join on `st`
            ev = torch.cuda.Event()
            ev.record(st)
            self.side.wait_event(ev)
            self.hi.wait_event(ev)
            self.stage_overlap(ctypes.c_void_p(self.side.cuda_stream))
            sh = ctypes.c_void_p(self.hi.cuda_stream)
            self.stage_forward_no_matrix(sh)
            self.stage_backward(sh)
            e1, e2 = torch.cuda.Event(), torch.cuda.Event()
            e1.record(self.side); e2.record(self.hi)
            st.wait_event(e1); st.wait_event(e2)
            return
        if self.two_kernel:
            self.stage_overlap(s)
        self.stage_forward(s)
        self.stage_backward(s)

    def capture(self):
        """Capture step() into a CUDA graph and return it (replay with .replay())."""
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            self.step(side)                      # warm-up outside capture (one-time function attributes)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.step(torch.cuda.current_stream(self.dev))
        return g


class BatchedHeadPlan(object):
    """Allocation-free training step of the batched-image configuration (BASELINE.json configs[3], SURVEY.md section 8(d)
    C4): this rank's `batch` images of `n` 2D boxes each.

        scores = sigmoid(x . w + b)          shared 64 -> 1 head over per-box features (gnms_score_head_forward_f32)
        prob   = GrooMeD-NMS(scores, boxes)  fused forward from the boxes (+ the [n,n] IoU matrices if materialise)
        dL/dscores                           analytic NMS backward for the upstream gradient grad_prob
        dL/d(w, b)                           head backward, written straight into the gradient bucket
        all-reduce(bucket)                   ONE all-reduce (sum) per step over all ranks: inside the head-gradient kernel over
                                             NVLink peer memory (collective="peer": gnms_score_head_backward_allreduce_f32), or
                                             NCCL on the flat bucket (collective="nccl"; what a padded bucket needs)

    Everything is enqueued on one stream; `capture()` records it (collective included) in a CUDA graph."""

    FEAT = 64

    def __init__(self, batch, n, device, params, materialise=True, bucket_pad_elems=0, group=None, collective="auto"):
        from .sharding import GradBucket, PeerExchange
        self.lib = _lib.load()
        self.B, self.N, self.dev, self.params, self.materialise, self.group = batch, n, device, params, materialise, group
        f = dict(dtype=torch.float32, device=device)
        self.x = torch.zeros((batch, n, self.FEAT), **f)
        self.boxes = torch.zeros((batch, n, 4), **f)
        self.wb = torch.zeros((self.FEAT + 1,), **f)
        self.grad_prob = torch.zeros((batch, n), **f)
        self.scores = torch.empty((batch, n), **f)
        self.overlap = torch.empty((batch, n, n), **f) if materialise else None
        self.prob = torch.empty((batch, n), **f)
        self.grad_scores = torch.empty((batch, n), **f)
        self.valid_idx = torch.empty((batch, n), dtype=torch.int64, device=device)
        self.invalid_idx = torch.empty((batch, n), dtype=torch.int64, device=device)
        self.counts = torch.empty((batch, 2), dtype=torch.int32, device=device)
        self.order = torch.empty((batch, n), dtype=torch.int32, device=device)
        self.lead = torch.empty((batch, n), dtype=torch.int32, device=device)
        self.fl = torch.empty((4, batch, n), **f)
        self.ws = torch.empty((int(self.lib.gnms_workspace_bytes(n, batch)),), dtype=torch.uint8, device=device)
        self.head_ws = torch.empty((int(self.lib.gnms_score_head_workspace_bytes(self.FEAT)),), dtype=torch.uint8, device=device)
        self.saved = Saved(_vp(self.order), _vp(self.fl[0]), _vp(self.lead), _vp(self.fl[1]), _vp(self.fl[2]), _vp(self.fl[3]))
        self.bucket = GradBucket({"head_wb": self.FEAT + 1}, device, pad_elems=bucket_pad_elems)
        self.grad_wb = self.bucket.view("head_wb")
        self.forward_opts = None
        import torch.distributed as dist
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if collective not in ("auto", "peer", "nccl"):
            raise ValueError("collective must be auto, peer or nccl")
        # the gradient of this head alone is 260 bytes: it travels inside the kernel; a padded bucket (a whole model) is NCCL's job
        self.exchange = PeerExchange(device, group) if multi and (collective == "peer" or (collective == "auto" and bucket_pad_elems == 0)) else None
        self.collective = "peer" if self.exchange is not None else ("nccl" if multi else "none")
        # head fwd, [sort,] rank, elect2 (+ chain), [matrix], NMS backward, head backward (2 launches; the second one carries the
        # all-reduce in peer mode, otherwise NCCL's kernel follows, not counted)
        by_sort = float(batch) * n * n >= 10.0 * 4096 * 4096
        self.launches_per_step = 6 + (1 if by_sort else 0) + (1 if materialise else 0)

    def compute(self, s, head_backward=True):
        """Forward + backward of this rank's shard (no collective)."""
        p = ctypes.byref(self.params)
        M = self.B * self.N
        check(self.lib.gnms_score_head_forward_f32(_vp(self.x), M, self.FEAT, _vp(self.wb), _vp(self.scores), s), "score_head_forward")
        check(self.lib.gnms_forward_boxes_ex_f32(_vp(self.scores), _vp(self.boxes), _lib.BOX_2D, 0, 0, self.N, self.B, None, p,
                                                 _vp(self.overlap), _vp(self.prob), _vp(self.valid_idx), _vp(self.invalid_idx),
                                                 _vp(self.counts), self.saved, _vp(self.ws), _lib.opts_ref(self.forward_opts), s),
              "forward_boxes")
        check(self.lib.gnms_backward_f32(_vp(self.grad_prob), _vp(self.prob), None, 0, self.N, self.B, None, p, self.saved,
                                         _vp(self.grad_scores), None, self.N, _vp(self.ws), s), "backward")
        if head_backward:
            check(self.lib.gnms_score_head_backward_f32(_vp(self.x), M, self.FEAT, _vp(self.scores), _vp(self.grad_scores),
                                                        _vp(self.grad_wb), _vp(self.head_ws), s), "score_head_backward")

    def step(self, stream=None, reduce=True):
        """One training step.  reduce=True on EVERY rank of the group (it is a collective), or on none."""
        st = stream if stream is not None else torch.cuda.current_stream(self.dev)
        s = ctypes.c_void_p(st.cuda_stream)
        if reduce and self.exchange is not None:
            self.compute(s, head_backward=False)
            check(self.lib.gnms_score_head_backward_allreduce_f32(_vp(self.x), self.B * self.N, self.FEAT, _vp(self.scores), _vp(self.grad_scores),
                                                                  _vp(self.grad_wb), _vp(self.head_ws), ctypes.byref(self.exchange.peers),
                                                                  _vp(self.exchange.status), s), "score_head_backward_allreduce")
            return
        self.compute(s)
        if reduce:
            with torch.cuda.stream(st):
                self.bucket.all_reduce(self.group)

    def close(self):
        """Release the peer exchange buffers (a collective: every rank calls it, after its last step)."""
        if self.exchange is not None:
            import torch.distributed as dist
            torch.cuda.synchronize(self.dev)
            dist.barrier(self.group)                 # nobody is still reading this rank's buffer
            self.exchange.close()
            self.exchange = None

    def capture(self, reduce=True):
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            self.step(side, reduce)                  # warm-up outside capture (function attributes, NCCL communicator)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.step(torch.cuda.current_stream(self.dev), reduce)
        return g


class SplitPlan(object):
    """The same step over `batch` images, issued as `splits` independent sub-batches on parallel graph branches.
    Images are independent, and every stage except the tile kernel is a small-grid, latency-bound launch (one CTA
    per image): with two or more branches the grouping / backward kernels of one sub-batch run under the tile kernel
    of another instead of leaving the GPU idle."""

    def __init__(self, batch, n, device, params, materialise=True, splits=2):
        assert batch % splits == 0
        self.dev = device
        self.parts = [Nms3dPlan(batch // splits, n, device, params, materialise=materialise) for _ in range(splits)]
        self.B, self.N = batch, n
        self.launches_per_step = sum(p.launches_per_step for p in self.parts)
        self.streams = [torch.cuda.Stream(device) for _ in range(splits)]

    def load(self, boxes7, scores, grad_prob):
        k = self.B // len(self.parts)
        for i, p in enumerate(self.parts):
            p.boxes7.copy_(boxes7[i * k:(i + 1) * k]); p.scores.copy_(scores[i * k:(i + 1) * k]); p.grad_prob.copy_(grad_prob[i * k:(i + 1) * k])

    def step(self, stream=None):
        main = stream if stream is not None else torch.cuda.current_stream(self.dev)
        ev = torch.cuda.Event()
        ev.record(main)
        done = []
        for p, st in zip(self.parts, self.streams):
            st.wait_event(ev)
            p.step(st)
            e = torch.cuda.Event()
            e.record(st)
            done.append(e)
        for e in done:
            main.wait_event(e)

    def capture(self):
        for p, st in zip(self.parts, self.streams):
            st.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(st):
                p.step(st)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.step(torch.cuda.current_stream(self.dev))
        return g


def bind_host_to_gpu(device_index):
    """Pin this process (and the pinned buffers it allocates afterwards: first touch) to the CPUs of the NUMA node the GPU hangs
    off, read from sysfs (/sys/bus/pci/devices/<bdf>/numa_node, local_cpulist).  Returns a small report; does nothing -- and says
    so -- where the platform exposes no locality (numa_node -1, as in most VMs) or the files are missing."""
    import os
    rep = {"device": int(device_index), "bound": False}
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read().strip())
        cpulist = open(base + "/local_cpulist").read().strip()
        rep.update(pci=bdf, numa_node=node, local_cpulist=cpulist)
        if node < 0 or not cpulist:
            rep["why"] = "the platform reports no NUMA locality for this GPU"
            return rep
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            rep["why"] = "already confined to (or not allowed on) the GPU's local CPUs"
            return rep
        os.sched_setaffinity(0, cpus)
        rep["bound"] = True
        rep["cpus"] = len(cpus)
    except Exception as e:                                   # no sysfs, no permission: leave the process where it is
        rep["why"] = "%s: %s" % (type(e).__name__, e)
    return rep


class HostRunner(object):
    """End-to-end entry for HOST buffers: pinned staging + an Nms3dPlan.  run_host(...) returns host tensors.

    keep_cap: the keep lists go back as int32 [batch, keep_cap] (entries past an image's count are -1; counts[b, 0] is the
              true length -- a keep list is a few dozen entries, the [batch, n] int64 array it lives in 8 n bytes per image).
              keep_cap = 0 returns the full int64 [batch, n] array instead.
    grad_on_device: the upstream gradient dL/dprob stays where training produces it -- on the device (plan.grad_prob, set by
              the caller's loss kernels or once with set_grad) -- instead of crossing PCIe with every call."""

    def __init__(self, batch, n, device, params, materialise=False, keep_cap=512, grad_on_device=False):
        self.plan = Nms3dPlan(batch, n, device, params, materialise=materialise)
        self.keep_cap, self.grad_on_device = int(keep_cap), bool(grad_on_device)
        pin = dict(pin_memory=True)
        self.h_prob = torch.empty((batch, n), dtype=torch.float32, **pin)
        self.h_grad = torch.empty((batch, n), dtype=torch.float32, **pin)
        self.h_counts = torch.empty((batch, 2), dtype=torch.int32, **pin)
        if self.keep_cap:
            self.d_keep = torch.empty((batch, self.keep_cap), dtype=torch.int32, device=device)
            self.h_valid = torch.empty((batch, self.keep_cap), dtype=torch.int32, **pin)
        else:
            self.d_keep = None
            self.h_valid = torch.empty((batch, n), dtype=torch.int64, **pin)
        self.h2d_bytes = batch * n * (7 + 1 + (0 if self.grad_on_device else 1)) * 4
        self.d2h_bytes = batch * n * (4 + 4) + self.h_valid.numel() * self.h_valid.element_size() + batch * 8
        self.graph = None                        # the kernel sequence as a CUDA graph, captured on first use

    def set_grad(self, grad_prob):
        self.plan.grad_prob.copy_(grad_prob, non_blocking=True)

    def enqueue_step(self, stream):
        """The kernels of one call (+ the keep-list packing) on `stream`; capturable."""
        p = self.plan
        p.step(stream)
        if self.keep_cap:
            check(p.lib.gnms_pack_keep_i32(_vp(p.valid_idx), _vp(p.counts), p.N, p.B, self.keep_cap, _vp(self.d_keep),
                                           ctypes.c_void_p(stream.cuda_stream)), "pack_keep")

    def capture(self, stream=None):
        p = self.plan
        st = stream if stream is not None else torch.cuda.Stream(p.dev)
        st.wait_stream(torch.cuda.current_stream(p.dev))
        with torch.cuda.stream(st):
            self.enqueue_step(st)                # warm-up outside capture
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            self.enqueue_step(st)
        return g

    def copy_in(self, boxes7_host, scores_host, grad_prob_host):
        p = self.plan
        p.boxes7.copy_(boxes7_host, non_blocking=True)
        p.scores.copy_(scores_host, non_blocking=True)
        if not self.grad_on_device:
            p.grad_prob.copy_(grad_prob_host, non_blocking=True)

    def copy_out(self):
        p = self.plan
        self.h_prob.copy_(p.prob, non_blocking=True)
        self.h_grad.copy_(p.grad_scores, non_blocking=True)
        self.h_valid.copy_(self.d_keep if self.keep_cap else p.valid_idx, non_blocking=True)
        self.h_counts.copy_(p.counts, non_blocking=True)

    def run_host(self, boxes7_host, scores_host, grad_prob_host=None):
        """boxes7 [B,N,7], scores [B,N] (and dL/dprob [B,N] unless grad_on_device), pinned host fp32 ->
        (prob, grad_scores, keep lists, counts) on the host.  One H2D per input, the kernels (one graph launch), one D2H per
        output, one stream sync."""
        p = self.plan
        if self.graph is None:
            self.graph = self.capture()
        self.copy_in(boxes7_host, scores_host, grad_prob_host)
        self.graph.replay()
        self.copy_out()
        torch.cuda.current_stream(p.dev).synchronize()
        return self.h_prob, self.h_grad, self.h_valid, self.h_counts


class HostPipeline(object):
    """Streaming version of HostRunner for back-to-back calls with HOST buffers: `depth` slots, each with its own
    device plan, pinned result buffers, stream and CUDA graph of the kernel sequence.  submit() enqueues H2D ->
    graph -> D2H for one batch on the next slot and returns immediately, so the PCIe copies of one call overlap the
    kernels of its neighbours; wait(ticket) (or drain()) makes a call's results readable.  Every call still moves
    all of its inputs and outputs across PCIe."""

    def __init__(self, batch, n, device, params, depth=2, keep_cap=512, grad_on_device=False):
        self.dev = device
        self.slots = []
        for _ in range(depth):
            r = HostRunner(batch, n, device, params, materialise=False, keep_cap=keep_cap, grad_on_device=grad_on_device)
            st = torch.cuda.Stream(device)
            g = r.capture(st)
            self.slots.append(dict(r=r, st=st, g=g, ev=torch.cuda.Event(), busy=False))
        self.k = 0
        self.h2d_bytes, self.d2h_bytes = self.slots[0]["r"].h2d_bytes, self.slots[0]["r"].d2h_bytes

    def set_grad(self, grad_prob):
        for s in self.slots:
            with torch.cuda.stream(s["st"]):
                s["r"].set_grad(grad_prob)

    def submit(self, boxes7_host, scores_host, grad_prob_host=None, copies_only=False):
        """copies_only: skip the kernels (measures what the PCIe link alone allows for this call pattern)."""
        s = self.slots[self.k % len(self.slots)]
        if s["busy"]:
            s["ev"].synchronize()                 # the slot's previous results must have been produced (and are now overwritten)
        r = s["r"]
        with torch.cuda.stream(s["st"]):
            r.copy_in(boxes7_host, scores_host, grad_prob_host)
            if not copies_only:
                s["g"].replay()
            r.copy_out()
            s["ev"].record(s["st"])
        s["busy"] = True
        self.k += 1
        return self.k - 1

    def wait(self, ticket):
        s = self.slots[ticket % len(self.slots)]
        s["ev"].synchronize()
        r = s["r"]
        return r.h_prob, r.h_grad, r.h_valid, r.h_counts

    def drain(self):
        for s in self.slots:
            if s["busy"]:
                s["ev"].synchronize()
