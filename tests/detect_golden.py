"""Reader of tests/golden/inference_site_ref.npz (oracle/gen_golden_detect.py: the unmodified reference's im_detect_3d)."""
import json

from conftest import load_golden


def cases():
    g = load_golden("inference_site_ref")
    out = {}
    for name in sorted({k.split("/")[0] for k in g.files}):
        c = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith(name + "/")}
        c["overrides"] = json.loads(bytes(c["overrides"]).decode())
        conf = dict(use_nms_in_loss=True, nms_thres=0.4, nms_topN_pre=3000, diff_nms_temperature=0.1, diff_nms_pruning_method="linear",
                    diff_nms_group_boxes=True, diff_nms_mask_group_boxes=True, diff_nms_group_size=100)        # scripts/config/groumd_nms.py
        conf.update(c["overrides"])
        c["conf"] = conf
        out[name] = c
    return out
