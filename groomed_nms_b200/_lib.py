"""ctypes binding of libgroomed_b200.so (the C-ABI declared in include/groomed_nms_b200.h).

There is no CPU fallback: if the library is missing or a call fails the caller gets an exception."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgroomed_b200.so")

c_f = ctypes.POINTER(ctypes.c_float)
vp = ctypes.c_void_p
i32 = ctypes.c_int
i64 = ctypes.c_int64
f32 = ctypes.c_float
f64 = ctypes.c_double
sz = ctypes.c_size_t

MAX_BOXES = 8192
PRUNE = {"linear": 0, "sigmoidal": 1, "soft_nms": 2}
MODE_GROUP_MASK, MODE_GROUP_NOMASK, MODE_NOGROUP, MODE_GROUP_MASK_INPUT_TRIL = 0, 1, 2, 3
KIND_IOU, KIND_INTERSECT = 0, 1
BOX_2D, BOX_3D_REC = 0, 1
CMP_GT, CMP_GE, CMP_NLE = 0, 1, 2


class Params(ctypes.Structure):
    _fields_ = [("nms_threshold", f32), ("temperature", f32), ("valid_box_prob_threshold", f32),
                ("pruning_method", ctypes.c_int32), ("mode", ctypes.c_int32), ("group_size", ctypes.c_int32),
                ("thresholded_output", ctypes.c_int32), ("sorted_output", ctypes.c_int32)]


class LaunchOpts(ctypes.Structure):
    """gnms_launch_opts (include/groomed_nms_b200.h): per-call tuning / profiling knobs; 0 = library default everywhere."""
    _fields_ = [("struct_size", ctypes.c_uint32), ("matrix_kernel", ctypes.c_int32), ("tiles_per_cta", ctypes.c_int32),
                ("rank_method", ctypes.c_int32), ("election", ctypes.c_int32), ("stage_mask", ctypes.c_uint32),
                ("flags", ctypes.c_uint32)]


MATRIX_KERNEL_AUTO, MATRIX_KERNEL_DIRECT, MATRIX_KERNEL_TMA = 0, 1, 2
RANK_AUTO, RANK_COUNT, RANK_SORT = 0, 1, 2
ELECT_AUTO, ELECT_DIRECT, ELECT_MASK, ELECT_BATCHED = 0, 1, 2, 3
STAGE_RANK, STAGE_SPATIAL, STAGE_TILES, STAGE_EARLIER, STAGE_CHAIN, STAGE_ELECT = 1, 2, 4, 8, 16, 32
OPT_SCALAR_MATH, OPT_INLINE_HITS, OPT_ONE_PASS, OPT_SPLIT_CHAIN = 1, 2, 4, 8


def launch_opts(matrix_kernel=0, tiles_per_cta=0, rank_method=0, election=0, stage_mask=0, flags=0):
    o = LaunchOpts()
    o.struct_size = ctypes.sizeof(LaunchOpts)
    o.matrix_kernel, o.tiles_per_cta, o.rank_method, o.election = matrix_kernel, tiles_per_cta, rank_method, election
    o.stage_mask, o.flags = stage_mask, flags
    return o


def opts_ref(o):
    """ctypes argument for an optional LaunchOpts (None -> NULL)."""
    return ctypes.byref(o) if o is not None else None


class Peers(ctypes.Structure):
    """gnms_peers: the exchange buffers of all ranks (device pointers, this rank's own included)."""
    _fields_ = [("buf", vp * 16), ("rank", ctypes.c_int32), ("world", ctypes.c_int32)]


class Saved(ctypes.Structure):
    _fields_ = [("order", vp), ("sorted_scores", vp), ("lead", vp), ("pval", vp), ("dpval", vp), ("pre", vp)]


# name -> (restype, argtypes); every symbol include/groomed_nms_b200.h declares
SIGNATURES = {
    "gnms_version": (i32, []),
    "gnms_error_string": (ctypes.c_char_p, [i32]),
    "gnms_overlap2d_f32": (i32, [vp, i32, vp, i32, vp, i64, i32, vp]),
    "gnms_overlap2d_list_f32": (i32, [vp, vp, i32, vp, i32, vp]),
    "gnms_overlap2d_f64": (i32, [vp, i64, i32, vp, i64, i32, i32, i32, i32, vp, vp]),
    "gnms_iou2d_backward_f32": (i32, [vp, i32, vp, i32, vp, i32, vp, vp, vp]),
    "gnms_corners_from_boxes7_f32": (i32, [vp, i64, i32, vp, vp]),
    "gnms_corners_from_boxes7_ex_f32": (i32, [vp, i64, i32, i32, vp, vp]),
    "gnms_corners_backward_f32": (i32, [vp, i64, i32, i32, vp, vp, vp]),
    "gnms_iou3d_approx_backward_f32": (i32, [vp, i32, vp, i32, i32, i32, vp, vp, vp, vp, vp]),
    "gnms_project_points_f32": (i32, [vp, vp, i64, i32, vp, vp]),
    "gnms_box3d_records_f32": (i32, [vp, i32, vp, i32, vp]),
    "gnms_box3d_records_from_boxes7_f32": (i32, [vp, i64, i32, vp, vp, vp]),
    "gnms_overlap3d_f32": (i32, [vp, i32, vp, i32, vp, vp, i64, i32, i32, vp, vp]),
    "gnms_overlap2d_batched_f32": (i32, [vp, i32, i32, vp, vp]),
    "gnms_overlap3d_batched_f32": (i32, [vp, i32, i32, vp, i32, i32, vp]),
    "gnms_overlap3d_list_f32": (i32, [vp, vp, i32, vp, vp, i32, i32, vp]),
    "gnms_overlap2d_batched_ex_f32": (i32, [vp, i32, i32, vp, ctypes.POINTER(LaunchOpts), vp]),
    "gnms_overlap3d_batched_ex_f32": (i32, [vp, i32, i32, vp, i32, i32, ctypes.POINTER(LaunchOpts), vp]),
    "gnms_workspace_bytes": (sz, [i32, i32]),
    "gnms_forward_f32": (i32, [vp, vp, i64, i32, i32, vp, ctypes.POINTER(Params), vp, vp, vp, vp, Saved, vp, vp]),
    "gnms_forward_boxes_f32": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, ctypes.POINTER(Params), vp, vp, vp, vp, vp,
                                     Saved, vp, vp]),
    "gnms_forward_ex_f32": (i32, [vp, vp, i64, i32, i32, vp, ctypes.POINTER(Params), vp, vp, vp, vp, Saved, vp,
                                  ctypes.POINTER(LaunchOpts), vp]),
    "gnms_forward_boxes_ex_f32": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, ctypes.POINTER(Params), vp, vp, vp, vp, vp,
                                        Saved, vp, ctypes.POINTER(LaunchOpts), vp]),
    "gnms_backward_f32": (i32, [vp, vp, vp, i64, i32, i32, vp, ctypes.POINTER(Params), Saved, vp, vp, i64, vp, vp]),
    "gnms_pack_keep_i32": (i32, [vp, vp, i32, i32, i32, vp, vp]),
    "gnms_get_groups_f32": (i32, [vp, vp, i64, i32, f32, i32, vp, vp, vp, vp, vp]),
    "gnms_prune_f32": (i32, [vp, i64, i32, f32, f32, vp, vp]),
    "gnms_indices_copy_f32": (i32, [vp, i64, vp, i64, i64, vp, vp, vp, vp, i64, vp]),
    "gnms_soft_sort_workspace_bytes": (sz, [i32]),
    "gnms_soft_sort_forward_f32": (i32, [vp, i32, f32, vp, i64, vp, vp, vp, vp, vp, vp, vp]),
    "gnms_soft_sort_backward_f32": (i32, [vp, i32, f32, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "gnms_hard_nms_f32": (i32, [vp, i32, f32, f32, i32, vp, vp, vp, vp]),
    "gnms_nms_host": (i32, [vp, vp, vp, i32, i32, f32, i32]),
    "gnms_soft_nms_f64": (i32, [vp, i32, f64, f64, f64, i32, f64, vp, vp, vp, vp, vp]),
    "gnms_soft_nms_workspace_bytes": (sz, [i32]),
    "gnms_aploss_f32": (i32, [vp, vp, i32, vp, vp, vp, sz, vp]),
    "gnms_aploss_workspace_bytes": (sz, [i32]),
    "gnms_masked_topk_f32": (i32, [vp, vp, i32, i32, i32, vp, vp, vp]),
    "gnms_best_box_per_gt_f32": (i32, [vp, vp, vp, i32, vp, vp, vp, i32, f32, vp, i32, vp, vp, vp, vp]),
    "gnms_score_head_workspace_bytes": (sz, [i32]),
    "gnms_score_head_forward_f32": (i32, [vp, i64, i32, vp, vp, vp]),
    "gnms_score_head_backward_f32": (i32, [vp, i64, i32, vp, vp, vp, vp, vp]),
    "gnms_peer_buffer_create": (i32, [ctypes.POINTER(vp), ctypes.c_char_p]),
    "gnms_peer_buffer_open": (i32, [ctypes.c_char_p, ctypes.POINTER(vp)]),
    "gnms_peer_buffer_close": (i32, [vp, i32]),
    "gnms_score_head_backward_allreduce_f32": (i32, [vp, i64, i32, vp, vp, vp, vp, ctypes.POINTER(Peers), vp, vp]),
    "gnms_targets_overlaps_workspace_bytes": (sz, [i32, i32]),
    "gnms_targets_overlaps_f64": (i32, [vp, i64, i32, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
    "gnms_iou3d_exact_f64": (i32, [vp, i64, i32, vp, i64, i32, vp, i32, vp, vp, vp]),
}

_lib = None


def load():
    """Load (once) and return the shared library with argtypes set.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "groomed_nms_b200: %s is missing -- build it with `python -m groomed_nms_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class GnmsError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = load().gnms_error_string(rc)
        raise GnmsError("%s failed: rc=%d (%s)" % (what, rc, msg.decode() if msg else "?"))
