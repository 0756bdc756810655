"""Debug: per-phase time of chain_kernel (SM clock stamps at every barrier) for one C3 image.
Needs a library built with -DGNMS_DEBUG (the phase clock is compiled out of the shipped build), e.g. into
tools/exp/libgroomed_b200_debug.so (the default looked up here; override with $GNMS_DEBUG_LIB):
    (cd groomed_nms_b200/csrc && nvcc <flags of build.py> -DGNMS_DEBUG -o ../../tools/exp/libgroomed_b200_debug.so *.cu)
    python tools/chain_phases.py [election: 0 auto, 1 direct, 2 suppression bits]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from groomed_nms_b200 import _lib, ops, synthetic
_lib.LIB_PATH = os.environ.get("GNMS_DEBUG_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "exp", "libgroomed_b200_debug.so"))
lib = ctypes.CDLL(_lib.LIB_PATH)
_lib.load()
OPTS = _lib.launch_opts(election=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
b7, sc = synthetic.config_c3()
dev = torch.device("cuda", 0)
rec = ops.box3d_records(ops.corners_from_boxes7(torch.from_numpy(b7).to(dev)))
p = ops.make_params()
s = torch.from_numpy(sc).to(dev)[None]
for _ in range(3):
    st = ops.forward_boxes(s, rec[None], _lib.BOX_3D_REC, p, generalized=True, affine=True, opts=OPTS)
torch.cuda.synchronize()
lib.gnms_debug_chain_clock(1)
st = ops.forward_boxes(s, rec[None], _lib.BOX_3D_REC, p, generalized=True, affine=True, opts=OPTS)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
lib.gnms_debug_chain_clock_read(buf)
lib.gnms_debug_chain_clock(0)
v = np.array(list(buf), dtype=np.int64)
order = [(i, v[i]) for i in range(64) if v[i] > 0]
order.sort(key=lambda t: t[1])
t0 = order[0][1]
prev = t0
for i, c in order:
    print("phase %2d  +%7.2f us  (cum %7.2f us)" % (i, (c - prev) / 1965.0, (c - t0) / 1965.0))
    prev = c
