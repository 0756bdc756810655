"""Drop-in installation: alias the mirrored modules into `sys.modules['lib.*']` so that reference code such as
`from lib.groomed_nms import differentiable_nms` (lib/loss/rpn_3d.py:14), `from lib.nms.gpu_nms import gpu_nms`
(lib/rpn_util.py:17) or `from lib.nms_others import *` picks up the sm_100a implementation unchanged."""
import functools
import importlib
import sys
import types

_MODULES = ["groomed_nms", "nms", "nms.gpu_nms", "nms.cpu_nms", "nms.py_cpu_nms", "nms_others", "loss.aploss"]
_CORE_SYMBOLS = ["iou", "intersect", "iou3d", "iou3d_approximate", "get_volume", "get_hull", "remove_rotation_in_boxes"]
_MATH3D_SYMBOLS = ["get_corners_of_cuboid", "project_3d_points_in_4D_format"]


def _needs_autograd(args, kwargs):
    import torch
    if not torch.is_grad_enabled():
        return False
    return any(isinstance(a, torch.Tensor) and a.requires_grad for a in list(args) + list(kwargs.values()))


def _grad_aware(mine, original):
    """`iou`, `iou3d_approximate`, `get_corners_of_cuboid` and `differentiable_nms` carry an analytic backward of their own (the
    loss differentiates the first three when the acceptance-probability target is not detached, lib/loss/rpn_3d.py:620,
    :663-679 with :1060).  The remaining kernel-backed mirrors -- project_3d_points_in_4D_format / intersect / get_volume, which
    nothing in the reference differentiates directly -- return plain tensors (no grad_fn): a call of those whose inputs require
    grad goes to the reference's own composite (same results, autograd intact), every other call goes to the kernels."""
    @functools.wraps(original)
    def dispatch(*args, **kwargs):
        if _needs_autograd(args, kwargs):
            return original(*args, **kwargs)
        return mine(*args, **kwargs)
    dispatch.__wrapped_kernel__ = mine
    dispatch.__wrapped_reference__ = original
    return dispatch


_HAS_BACKWARD = {"iou", "iou3d_approximate", "get_corners_of_cuboid"}   # mirrors that carry their own analytic backward


def install(patch_core=True):
    """Register lib.groomed_nms, lib.nms.*, lib.nms_others, lib.loss.aploss.  If the reference's own `lib.core` /
    `lib.math_3d` are importable and patch_core is set, their overlap/corner functions are rebound in place (the
    rest of those modules -- config, checkpoints, LR policy -- is outside the hot path and stays the reference's).
    Rebound functions without an analytic backward keep the reference's autograd: see _grad_aware."""
    from . import _lib
    _lib.load()                                   # fail loudly right here if the CUDA library is missing
    if "lib" not in sys.modules:
        try:
            importlib.import_module("lib")
        except ImportError:
            pkg = types.ModuleType("lib")
            pkg.__path__ = []
            sys.modules["lib"] = pkg
    for name in _MODULES:
        mod = importlib.import_module("groomed_nms_b200.lib." + name)
        sys.modules["lib." + name] = mod
        parent, _, leaf = ("lib." + name).rpartition(".")
        if parent in sys.modules:
            setattr(sys.modules[parent], leaf, mod)
    if patch_core:
        from .lib import core as my_core, math_3d as my_m3d
        for modname, mine, symbols in (("lib.core", my_core, _CORE_SYMBOLS), ("lib.math_3d", my_m3d, _MATH3D_SYMBOLS)):
            try:
                ref = importlib.import_module(modname)
            except Exception:
                sys.modules[modname] = mine
                continue
            for sym in symbols:
                orig = getattr(ref, sym, None)
                orig = getattr(orig, "__wrapped_reference__", orig)          # install() twice: keep the true original
                new = getattr(mine, sym)
                setattr(ref, sym, new if (sym in _HAS_BACKWARD or orig is None) else _grad_aware(new, orig))
