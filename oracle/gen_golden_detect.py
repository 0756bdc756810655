"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/inference_site_ref.npz by running the UNMODIFIED reference's
im_detect_3d (lib/rpn_util.py:1052-1357) on CPU in the build container (oracle/ref_harness.py --detect, stock arm,
--fake-cuda), with a stand-in network that returns the synthetic scene's outputs.

    python oracle/gen_golden_detect.py

Pins SURVEY.md section 8(f) rank 2 (the NMS block :1258-1341): per case the detections as they enter the block (score
sorted, truncated to nms_topN_pre) and what the reference returns (kept rows in its column layout, keep indices)."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "inference_site_ref.npz")
CASES = {
    "groomed_2d": (dict(seed=11), dict(nms_topN_pre=1200)),
    "groomed_3d": (dict(seed=12), dict(nms_topN_pre=1200, overlap_in_nms="3d")),
    "groomed_product": (dict(seed=13), dict(nms_topN_pre=1200, overlap_in_nms="product")),
    "groomed_2d_nomask": (dict(seed=14), dict(nms_topN_pre=1200, diff_nms_mask_group_boxes=False, diff_nms_group_size=30)),
    "classical": (dict(seed=15), dict(nms_topN_pre=1200, use_nms_in_loss=False)),
}


def main():
    g = {}
    for name, (sa, ov) in CASES.items():
        with tempfile.TemporaryDirectory() as td:
            out = os.path.join(td, "run.npz")
            cmd = [sys.executable, "-m", "oracle.ref_harness", "--arm", "stock", "--fake-cuda", "--detect", "--out", out, "--seed", str(sa["seed"]),
                   "--batch", "1", "--feat", "16x56"]
            for k, v in ov.items():
                cmd += ["--set", "%s=%s" % (k, json.dumps(v))]
            subprocess.run(cmd, check=True, cwd=ROOT, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            d = np.load(out)
            for k in d.files:
                g["%s/%s" % (name, k)] = d[k]
            g[name + "/overrides"] = np.frombuffer(json.dumps(ov).encode(), dtype=np.uint8)
            print("%-20s %4d into NMS -> %3d kept (%s)" % (name, len(d["pre_aboxes"]), len(d["aboxes_out"]), bytes(d["nms_fn"]).decode()))
    np.savez_compressed(OUT, **g)
    print("wrote %s %.1f KiB" % (OUT, os.path.getsize(OUT) / 1024.0))


if __name__ == "__main__":
    main()
