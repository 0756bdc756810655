"""Per-image sharding of a batch across ranks: NMS is independent per image
(lib/loss/rpn_3d.py:375,721-793), so rank r of W takes a contiguous block of images and the NMS path itself needs no
collective.  What a training step does exchange is the gradient of the shared parameters: ONE all-reduce (sum) of a flat
fp32 bucket per step -- the multi-process replacement of the reference's single-process `nn.DataParallel`
(lib/core.py:68), whose backward reduces the replicas' gradients onto device 0.  `GradBucket` is that bucket."""
import torch


def shard_range(num_images, world_size, rank):
    """Contiguous, balanced block of image indices for `rank` (first `num_images % world_size` ranks get one more)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(num_images, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def gather_per_image(local_values, num_images, group=None):
    """All-gather python objects computed per local image into a list indexed by global image id."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, list(local_values), group=group)
    out = []
    for r in range(world):
        assert len(parts[r]) == len(shard_range(num_images, world, r))
        out.extend(parts[r])
    return out


class GradBucket(object):
    """One flat fp32 buffer holding the gradients of all shared parameters of a step (and, optionally, padding that
    stands for the rest of a model: the reference's DenseNet-121 RPN has about 12 M parameters = 48 MB), reduced with a
    single collective call per step.  Kernels write their gradients straight into views of the buffer (no copy, no
    per-parameter collective)."""

    def __init__(self, sizes, device, pad_elems=0):
        """sizes: dict name -> number of fp32 elements, in bucket order."""
        self.offsets = {}
        off = 0
        for name, n in sizes.items():
            self.offsets[name] = (off, int(n))
            off += (int(n) + 3) // 4 * 4                               # 16-byte aligned views
        self.used = off
        self.flat = torch.zeros((off + int(pad_elems),), dtype=torch.float32, device=device)

    def view(self, name):
        off, n = self.offsets[name]
        return self.flat[off:off + n]

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def all_reduce(self, group=None, average=False):
        """Sum (or mean) over the ranks of `group`, in place, on the current stream: one ncclAllReduce on a NCCL group.
        No-op for a single process."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return self.flat
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.div_(dist.get_world_size(group))
        return self.flat


class PeerExchange(object):
    """The exchange buffers behind gnms_score_head_backward_allreduce_f32: this rank allocates a small device buffer, the ranks
    of `group` swap its CUDA IPC handle (one all_gather_object on the host, once) and map each other's buffers; the kernel then
    moves the gradient over NVLink itself.  torch.distributed is only the channel the 64-byte handles travel through."""

    def __init__(self, device, group=None):
        import ctypes
        import torch.distributed as dist
        from . import _lib
        self.lib = _lib.load()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 16:
            raise ValueError("at most 16 ranks per exchange")
        with torch.cuda.device(device):
            own = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            _lib.check(self.lib.gnms_peer_buffer_create(ctypes.byref(own), handle), "gnms_peer_buffer_create")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            self.ptrs = []
            for r in range(self.world):
                if r == self.rank:
                    self.ptrs.append(own.value)
                else:
                    p = ctypes.c_void_p()
                    _lib.check(self.lib.gnms_peer_buffer_open(handles[r], ctypes.byref(p)), "gnms_peer_buffer_open")
                    self.ptrs.append(p.value)
        self.device = device
        self.peers = _lib.Peers()
        for r in range(self.world):
            self.peers.buf[r] = self.ptrs[r]
        self.peers.rank, self.peers.world = self.rank, self.world
        self.status = torch.zeros((1,), dtype=torch.int32, device=device)
        dist.barrier(group)                                   # every buffer is mapped everywhere before the first step

    def close(self):
        if getattr(self, "ptrs", None):
            with torch.cuda.device(self.device):
                for r, p in enumerate(self.ptrs):
                    self.lib.gnms_peer_buffer_close(p, 1 if r == self.rank else 0)
            self.ptrs = None
