"""Builds libgroomed_b200.so in-tree with nvcc for sm_100a (the only target).

    python -m groomed_nms_b200.build [--force]

-fmad=false: overlaps must be evaluated with separately rounded fp32 ops to be bitwise equal to the reference's
torch ops (hard part 1).  -lineinfo keeps ncu's source page usable.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgroomed_b200.so")
SOURCES = ["overlap.cu", "gnms.cu", "misc.cu", "head.cu", "lossbranch.cu", "softsort.cu", "grad3d.cu"]
DEPS = ["common.cuh", "solve.cuh", os.path.join("..", "..", "include", "groomed_nms_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    files = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    files += [os.path.join(CSRC, d) for d in DEPS]
    return any(os.path.getmtime(f) > t for f in files)


def build(force=False, verbose=False):
    """Compile every CUDA source into groomed_nms_b200/libgroomed_b200.so.  Returns the library path."""
    if not force and not needs_build():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    extra = os.environ.get("NVCC_EXTRA", "").split()                   # e.g. -DGNMS_DEBUG (phase clock of chain_kernel)
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stdout))
    if verbose:
        print(res.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
