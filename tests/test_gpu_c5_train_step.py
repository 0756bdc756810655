"""Config C5 (BASELINE.json configs[4]): the reference's training step -- its densenet121_3d_dilate_decomp_alpha RPN, its
RPN_3D_loss with the GrooMeD-NMS branch, its loss_backprop (scripts/train_rpn_3d.py:131-142) -- on synthetic 384 x 1280
images, run stock and with `groomed_nms_b200.install()` active (oracle/ref_harness.py --model, one process per arm; the
reference sources are the unmodified files staged in baseline/_ref).  "Drops into scripts/train_rpn_3d.py unchanged" means:
same losses, same gradients, same parameter trajectory."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

STAGED = os.path.join(ROOT, "baseline", "_ref", "models", "densenet121_3d_dilate_decomp_alpha.py")


def _run(arm, out, iters=3):
    cmd = [sys.executable, "-m", "oracle.ref_harness", "--arm", arm, "--out", out, "--model", "--iters", str(iters), "--batch", "2",
           "--feat", "24x80", "--seed", "0"]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1200)
    assert r.returncode == 0, "ref_harness %s failed:\n%s" % (arm, r.stdout[-4000:])
    print(r.stdout.strip().splitlines()[-1])
    d = np.load(out)
    return {k: d[k] for k in d.files}


@pytest.mark.skipif(not os.path.isfile(STAGED), reason="reference not staged (python tools/stage_reference.py in the build container)")
def test_reference_train_step_stock_vs_installed(tmp_path):
    stock = _run("stock", str(tmp_path / "stock.npz"))
    ours = _run("installed", str(tmp_path / "installed.npz"))
    assert np.array_equal(stock["nms_sizes"], ours["nms_sizes"]) and int(stock["n_nms"][0]) == 6        # 2 images x 3 iterations
    assert np.isclose(stock["losses"][0], ours["losses"][0], rtol=1e-5)
    assert np.allclose(stock["losses"], ours["losses"], rtol=2e-4)                                     # after 1 and 2 SGD steps
    assert list(stock["stat_names"]) == list(ours["stat_names"])
    assert np.allclose(stock["stat_vals"], ours["stat_vals"], rtol=1e-4, atol=1e-6)
    for k in stock:
        if k.startswith("grad_"):
            a, b = stock[k], ours[k]
            assert np.allclose(a, b, rtol=1e-3, atol=1e-5 * max(1e-12, np.abs(a).max())), k
    assert np.abs(stock["grad_acceptance_prob.layer_0.weight"]).max() > 0
    print("C5 ms per iteration: stock %s | installed %s ; loss alone: stock %s | installed %s" % (
        np.round(stock["iter_ms"], 1).tolist(), np.round(ours["iter_ms"], 1).tolist(), np.round(stock["loss_ms"], 1).tolist(),
        np.round(ours["loss_ms"], 1).tolist()))
