"""Per-image sharding of a batch across ranks: NMS is independent per image
(lib/loss/rpn_3d.py:375,721-793), so rank r of W takes a contiguous block of images and no data-path collective is
needed; only per-image results (keep counts, losses) are gathered by the caller if it wants them on one rank."""


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
