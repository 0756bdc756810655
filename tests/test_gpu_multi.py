"""N > 1 on real GPUs (skipped on a 1-GPU box): the batched-image step (bench.py --config c4) under torchrun with one rank
per GPU over NCCL.  The all-reduced gradient of the shared head must equal what ONE process computes over all 32 images
(fp32 sums in a different order: 1e-5 relative), and the collective must be part of the step."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_c4_two_ranks_nccl_allreduce_matches_single_process():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "bench.py"), "--config", "c4", "--gpus", "2", "--steps", "5", "--warmup", "3"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and line["extra"]["strong"]["images_per_rank"] == 16
    assert line["extra"]["strong"]["allreduced_grad_rel_err_vs_single_process"] < 1e-5
    assert "ncclAllReduce" in line["config"]["collective"]
