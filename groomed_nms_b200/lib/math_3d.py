"""Mirror of the corner / projection feeders of the 3D overlap (reference lib/math_3d.py:47-72,364-490)."""
import numpy as np
import torch

from .. import ops
from ._util import Origin, no_autograd, to_cuda_f32


def get_corners_of_cuboid(x3d, y3d, z3d, w3d, h3d, l3d, ry3d, iou_3d_convention=True):
    """7-DoF boxes -> [N,3,8] corners (reference lib/math_3d.py:364-490).  iou_3d_convention=True: corner numbering of the
    docstring at :379-398 (what the GrooMeD-NMS path uses); False: the other vertex order of the reference (:405-426:
    length on x for corners {1,2,3,4}, height on y for {2,3,6,7}, width on z for {3,4,5,6}).  The reference's own False
    branch only broadcasts for N in {1, 4} in torch (`corners[:, 0, [1,2,3,4]] = l3d`, :422-424) and is missing from its
    numpy twin (:450-477); this gives the vertex order that branch documents, for any N and both container types.
    torch in -> torch out on the same device, numpy in -> numpy out.  Differentiable like the reference's composite."""
    origin = Origin(x3d)
    cols = [to_cuda_f32(v).reshape(-1) for v in (x3d, y3d, z3d, w3d, h3d, l3d, ry3d)]
    boxes7 = torch.stack(cols, dim=1)
    if torch.is_grad_enabled() and boxes7.requires_grad:
        # analytic backward wrt all seven parameters (the acceptance-probability target differentiates this call,
        # lib/loss/rpn_3d.py:663-679)
        out = ops.CornersFunction.apply(boxes7, bool(iou_3d_convention))
    else:
        out = ops.corners_from_boxes7(boxes7, iou_3d_convention=bool(iou_3d_convention))
    return origin.back(out, keep_np_dtype=True)


def project_3d_points_in_4D_format(p2, points_4d, pad_ones=False):
    """p2[4,4] @ [pts;1] with the x,y rows divided by z where |z| > 1e-2 (reference lib/math_3d.py:47-72)."""
    no_autograd("lib.math_3d.project_3d_points_in_4D_format", p2, points_4d)
    origin = Origin(points_4d)
    pts = to_cuda_f32(points_4d)
    P = to_cuda_f32(p2)
    out = ops.project_points(P, pts, bool(pad_ones))
    return origin.back(out, keep_np_dtype=True)
