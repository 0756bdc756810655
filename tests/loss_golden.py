"""Reader of tests/golden/loss_branch_ref.npz (written by oracle/gen_golden_loss.py from the unmodified reference's
RPN_3D_loss.forward).  Shared by the CPU oracle test and the GPU parity tests."""
import json

import numpy as np

from conftest import load_golden


class Case(object):
    def __init__(self, g, name):
        p = name + "/"
        self.name = name
        self.overrides = json.loads(bytes(g[p + "overrides"]).decode())
        self.loss = float(g[p + "loss"][0])
        self.p2 = g[p + "p2"]
        self.images = []
        for i in range(int(g[p + "n_images"][0])):
            q = "%simg%d_" % (p, i)
            self.images.append({k[len(q):]: g[k] for k in g.files if k.startswith(q)})

    def conf(self):
        """The keys of scripts/config/groumd_nms.py the GrooMeD branch reads, with this case's overrides."""
        c = dict(use_nms_in_loss=True, diff_nms_temperature=0.1, diff_nms_pruning_method="linear", diff_nms_boxes_2d="normal",
                 diff_nms_group_boxes=True, diff_nms_mask_group_boxes=True, diff_nms_group_size=100, after_nms_lambda=0.05,
                 after_nms_loss_mode="rank", best_target_box_beta=0.3, nms_thres=0.4, rank_boxes_of_all_images_at_once=False)
        c.update(self.overrides)
        return c


def cases():
    g = load_golden("loss_branch_ref")
    names = sorted({k.split("/")[0] for k in g.files})
    return [Case(g, n) for n in names]


def case_names():
    return [c.name for c in cases()]
