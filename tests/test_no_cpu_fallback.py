"""CPU: the product path has no CPU fallback and never reaches into oracle/ -- it fails loudly when the CUDA library or a
CUDA device is missing (the oracle is test infrastructure: only tests/, smoke() and bench.py's CPU legs may import it)."""
import ast
import glob
import os

import numpy as np
import pytest
import torch

from conftest import ROOT


def _imports(path):
    tree = ast.parse(open(path).read(), path)
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                yield a.name
        elif isinstance(node, ast.ImportFrom):
            yield ("." * node.level) + (node.module or "")


def test_package_never_imports_the_oracle():
    files = glob.glob(os.path.join(ROOT, "groomed_nms_b200", "**", "*.py"), recursive=True)
    assert len(files) > 10
    for f in files:
        for mod in _imports(f):
            assert mod.split(".")[0] != "oracle", "%s imports %s" % (f, mod)


def test_missing_library_raises(monkeypatch):
    from groomed_nms_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "groomed_nms_b200", "no_such_library.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
    import groomed_nms_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        groomed_nms_b200.install()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a host without a CUDA device")
def test_mirrored_calls_refuse_to_run_without_a_gpu():
    from groomed_nms_b200.lib import core, groomed_nms
    from groomed_nms_b200.lib.nms.gpu_nms import gpu_nms
    rng = np.random.default_rng(0)
    xy = rng.uniform(0, 100, (6, 2)).astype(np.float32)
    boxes = np.concatenate([xy, xy + 20], 1)
    scores = rng.uniform(0, 1, 6).astype(np.float32)
    with pytest.raises(RuntimeError, match="CUDA device"):
        core.iou(boxes, boxes)
    with pytest.raises(RuntimeError, match="CUDA device"):
        groomed_nms.differentiable_nms(scores, np.eye(6, dtype=np.float32))
    with pytest.raises(RuntimeError):
        gpu_nms(np.concatenate([boxes, scores[:, None]], 1), 0.5)
    with pytest.raises(RuntimeError, match="CUDA device"):
        core.iou3d(np.zeros((3, 8)), np.zeros((3, 8)))
