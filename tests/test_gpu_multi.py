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


def _run_c4(extra, port):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--config", "c4", "--gpus", "2", "--steps", "5", "--warmup", "3"] + extra
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_c4_two_ranks_nccl_allreduce_matches_single_process():
    line = _run_c4(["--c4-collective", "nccl"], 29731)
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and line["extra"]["strong"]["images_per_rank"] == 16
    assert line["extra"]["strong"]["allreduced_grad_rel_err_vs_single_process"] < 1e-5
    assert "ncclAllReduce" in line["config"]["collective"] and line["extra"]["strong"]["collective"] == "nccl"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_c4_two_ranks_in_kernel_peer_allreduce_matches_single_process():
    """The default for the unpadded bucket: the all-reduce happens inside the head-gradient kernel over NVLink peer memory
    (gnms_score_head_backward_allreduce_f32); same bar as NCCL, and no rank may have timed out waiting for its peer."""
    line = _run_c4([], 29732)
    assert line["n_gpus"] == 2 and line["extra"]["strong"]["collective"] == "peer" and line["extra"]["weak"]["collective"] == "peer"
    assert line["extra"]["strong"]["peer_exchange_ok"] is True and line["extra"]["weak"]["peer_exchange_ok"] is True
    assert line["extra"]["strong"]["allreduced_grad_rel_err_vs_single_process"] < 1e-5
    assert "peer memory" in line["config"]["collective"]
