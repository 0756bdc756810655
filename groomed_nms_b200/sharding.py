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
