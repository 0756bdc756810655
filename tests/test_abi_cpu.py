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


def test_header_is_plain_c_and_links_from_a_c_program(lib, tmp_path):
    """include/groomed_nms_b200.h compiles as strict C99 and a C program linked against the library reaches the entry
    points (examples/c_abi_check.c: version, workspace sizes, argument validation -- nothing that needs a GPU)."""
    import subprocess
    libdir = os.path.join(ROOT, "groomed_nms_b200")
    exe = os.path.join(str(tmp_path), "c_abi_check")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_check.c"), "-L", libdir, "-lgroomed_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    assert "c abi ok: version 100" in out.stdout


def test_exact_overlap_argument_errors_without_a_gpu(lib):
    z = ctypes.c_void_p(0)
    assert lib.gnms_iou3d_exact_f64(z, 24, 0, z, 24, 3, z, 0, z, z, z) == 0
    assert lib.gnms_iou3d_exact_f64(z, 23, 2, z, 24, 3, z, 0, z, z, z) == -1          # fewer than 3 rows of 8
    assert lib.gnms_iou3d_exact_f64(z, 24, 2, z, 24, 3, z, 1, z, z, z) == -1          # list mode needs M == N
    assert lib.gnms_iou3d_exact_f64(z, 24, 2, z, 24, 2, z, 1, z, z, z) == -1          # null pointers
