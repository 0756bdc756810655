#!/usr/bin/env python
"""Stage the UNMODIFIED reference (abhi1kumar/groomed_nms) into the git-ignored `baseline/_ref/`.

    python tools/stage_reference.py [--src /root/reference] [--check]

`/root/reference` exists only in the build container; the GPU box gets a snapshot of this repository, and
`baseline/_ref/` travels with it exactly like the built `.so` files do (git-ignored, not gpurun-ignored).  With the
reference staged there,
  * `bench.py --impl reference` times the reference's OWN CPU PyTorch path (lib.core.iou3d_approximate +
    lib.groomed_nms.differentiable_nms + autograd backward) on the GPU box's host cores,
  * `tests/test_gpu_reference_dropin.py` runs the reference's RPN_3D_loss.forward stock and with
    `groomed_nms_b200.install()` active on the same inputs, and
  * `tools/c5_train_step.py` runs the reference's model + loss + optimiser step (config C5).

Only the Python sources the hot path and its callers need are staged, byte for byte (the manifest below records a
sha256 per file so a test can prove nothing was edited).  Nothing staged here is part of the product or of git
history; the product never imports it.
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
# directories whose *.py (and the lib/nms native sources, for reading) are staged
SUBTREES = ["lib", "models", "scripts", "test"]
KEEP_EXT = (".py", ".pyx", ".hpp", ".cu")
SKIP_DIRS = {"roi_align", "__pycache__"}          # dead code in the reference


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def iter_sources(src):
    for sub in SUBTREES:
        top = os.path.join(src, sub)
        for d, dirs, files in os.walk(top):
            dirs[:] = sorted(x for x in dirs if x not in SKIP_DIRS)
            for fn in sorted(files):
                if fn.endswith(KEEP_EXT):
                    p = os.path.join(d, fn)
                    yield os.path.relpath(p, src), p


def stage(src="/root/reference", dest=DEST, quiet=False):
    """Copy the sources; returns the manifest {relpath: sha256}.  No-op (returns None) when `src` is absent."""
    if not os.path.isfile(os.path.join(src, "lib", "groomed_nms.py")):
        return None
    manifest = {}
    for rel, p in iter_sources(src):
        out = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        if not (os.path.exists(out) and _sha(out) == _sha(p)):
            shutil.copyfile(p, out)
        manifest[rel] = _sha(out)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": "abhi1kumar/groomed_nms (unmodified)", "files": manifest}, f, indent=1, sort_keys=True)
    if not quiet:
        print("staged %d reference files into %s" % (len(manifest), dest))
    return manifest


def check(dest=DEST):
    """True when every staged file still has the sha256 recorded at staging time."""
    mf = os.path.join(dest, "MANIFEST.json")
    if not os.path.exists(mf):
        return False
    files = json.load(open(mf))["files"]
    return all(os.path.exists(os.path.join(dest, r)) and _sha(os.path.join(dest, r)) == h for r, h in files.items())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    if a.check:
        ok = check()
        print("baseline/_ref %s" % ("intact" if ok else "missing or modified"))
        sys.exit(0 if ok else 1)
    if stage(a.src) is None:
        print("no reference tree at %s: nothing staged" % a.src)
        sys.exit(1)
