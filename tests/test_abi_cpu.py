"""CPU: the C-ABI library builds, loads and exports every symbol include/groomed_nms_b200.h declares; argument
validation works without a GPU (no compute calls here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _header_symbols():
    hdr = open(os.path.join(ROOT, "include", "groomed_nms_b200.h")).read()
    return sorted(set(re.findall(r"\b(gnms_[a-z0-9_]+)\s*\(", hdr)))


@pytest.fixture(scope="module")
def lib():
    from groomed_nms_b200 import build, _lib
    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from groomed_nms_b200 import _lib
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_strings(lib):
    assert lib.gnms_version() == 100
    assert lib.gnms_error_string(0) == b"success"
    assert b"bad argument" in lib.gnms_error_string(-1)
    assert b"GNMS_MAX_BOXES" in lib.gnms_error_string(-2)


def test_workspace_sizes_monotone(lib):
    w = [lib.gnms_workspace_bytes(n, 1) for n in (1, 64, 500, 1024, 4096, 8192)]
    assert all(b > a for a, b in zip(w, w[1:]))
    assert lib.gnms_workspace_bytes(2048, 4) > 4 * lib.gnms_workspace_bytes(2048, 1) - 8 * 4096 * 8
    assert lib.gnms_workspace_bytes(4096, 1) >= 4096 * 4096 // 8          # the suppression bitmask


def test_argument_errors_without_a_gpu(lib):
    from groomed_nms_b200._lib import Params, Saved
    z = ctypes.c_void_p(0)
    assert lib.gnms_overlap2d_f32(z, -1, z, 4, z, 4, 0, z) == -1
    assert lib.gnms_overlap2d_f32(z, 4, z, 4, z, 2, 0, z) == -1            # ld < N
    assert lib.gnms_overlap2d_f32(z, 0, z, 4, z, 4, 0, z) == 0             # empty is a no-op
    p = Params(0.4, 0.1, 0.3, 0, 0, 100, 0, 0)
    sv = Saved(z, z, z, z, z, z)
    assert lib.gnms_forward_f32(z, z, 9000, 9000, 1, z, ctypes.byref(p), z, z, z, z, sv, z, z) == -2   # too large
    assert lib.gnms_forward_f32(z, z, 0, 0, 1, z, ctypes.byref(p), z, z, z, z, sv, z, z) == 0
    p.pruning_method = 7
    assert lib.gnms_forward_f32(z, z, 8, 8, 1, z, ctypes.byref(p), z, z, z, z, sv, z, z) == -1
    assert lib.gnms_hard_nms_f32(z, 5, 0.5, 1.0, 9, z, z, z, z) == -1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "groomed_nms_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
