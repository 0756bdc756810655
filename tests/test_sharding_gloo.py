"""CPU, world_size 2 (gloo): the host-side logic of the N>1 path -- the rank -> image partition (shard_range), the gather
of per-image results, and the product's gradient bucket (GradBucket: kernels write their gradients into views of one flat
fp32 buffer which is all-reduced with ONE collective per step).  The per-image numbers here are stand-ins computed on the
host (there is no GPU in this tier and the product has no CPU compute path); the same code runs over NCCL on the GPU box
(tests/test_gpu_multi.py, bench.py --config c4)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from groomed_nms_b200 import sharding


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 32, 33):
        for w in (1, 2, 3, 8):
            got = [i for r in range(w) for i in sharding.shard_range(n, w, r)]
            assert got == list(range(n))
            sizes = [len(sharding.shard_range(n, w, r)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_grad_bucket_views_alias_the_flat_buffer():
    b = sharding.GradBucket({"head_wb": 65, "other": 3}, torch.device("cpu"), pad_elems=100)
    assert b.flat.numel() == 68 + 4 + 100 and b.view("other").data_ptr() == b.flat[68:].data_ptr()      # 16-byte aligned views
    b.view("head_wb").fill_(2.0)
    b.view("other").fill_(5.0)
    assert float(b.flat.sum()) == 65 * 2 + 3 * 5
    assert b.all_reduce() is b.flat                                                                      # single process: no-op


def _image_grad(i):
    """Stand-in for one image's contribution to the shared-parameter gradient (deterministic, exactly representable)."""
    rng = np.random.default_rng(100 + i)
    return torch.from_numpy(rng.integers(-8, 9, 65).astype(np.float32))


def _worker(rank, world, port, num_images, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_range(num_images, world, rank)
    bucket = sharding.GradBucket({"head_wb": 65}, torch.device("cpu"), pad_elems=32)
    g = bucket.view("head_wb")
    for i in mine:                                              # what the head-backward kernel does for the local shard
        g += _image_grad(i)
    bucket.all_reduce()                                         # ONE collective for the whole bucket
    per_image = sharding.gather_per_image([float(_image_grad(i).sum()) for i in mine], num_images)
    if rank == 0:
        q.put((bucket.view("head_wb").clone().numpy(), float(bucket.flat[68:].abs().sum()), per_image))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bucket_allreduce_equals_single_process():
    num_images = 5
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_images, q)) for r in range(2)]
    for p in procs:
        p.start()
    grad, pad_sum, per_image = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    want = sum(_image_grad(i) for i in range(num_images)).numpy()
    assert np.array_equal(grad, want) and pad_sum == 0.0
    assert per_image == [float(_image_grad(i).sum()) for i in range(num_images)]
