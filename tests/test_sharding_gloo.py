"""CPU, world_size 2 (gloo): the per-image shard of the batched config reproduces the single-process result.
The per-image work is done by the numpy oracle here (no GPU in this tier); on the GPU box bench.py uses the same
partition with one rank per GPU."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from groomed_nms_b200 import sharding, synthetic


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 32, 33):
        for w in (1, 2, 3, 8):
            got = [i for r in range(w) for i in sharding.shard_range(n, w, r)]
            assert got == list(range(n))
            sizes = [len(sharding.shard_range(n, w, r)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _per_image(i):
    from oracle import groomed_oracle as O
    boxes, sc = synthetic.config_c4_image(i, n=192, k=6)
    o = O.differentiable_nms(sc, O.iou(boxes, boxes), dense=False)
    return (int(len(o["valid"])), float(o["prob"].sum()))


def _worker(rank, world, port, num_images, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = [_per_image(i) for i in sharding.shard_range(num_images, world, rank)]
    allv = sharding.gather_per_image(local, num_images)
    if rank == 0:
        q.put(allv)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_equals_single_process():
    num_images = 5
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_images, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    want = [_per_image(i) for i in range(num_images)]
    assert [g[0] for g in got] == [w[0] for w in want]
    assert np.allclose([g[1] for g in got], [w[1] for w in want], rtol=0, atol=0)
