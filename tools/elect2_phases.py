"""Debug: cycles per phase of elect2_kernel (image 0) on one C3 image or a batch; needs the -DGNMS_DEBUG library (see
tools/chain_phases.py).   python tools/elect2_phases.py [images]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from groomed_nms_b200 import _lib, ops, synthetic
_lib.LIB_PATH = os.environ.get("GNMS_DEBUG_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "exp", "libgroomed_b200_debug.so"))
lib = ctypes.CDLL(_lib.LIB_PATH)
_lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
recs, scs = [], []
for i in range(B):
    b7, sc = synthetic.config_c3(seed=3 + 10 * i)
    recs.append(ops.box3d_records(ops.corners_from_boxes7(torch.from_numpy(b7).to(dev))))
    scs.append(torch.from_numpy(sc).to(dev))
rec, s = torch.stack(recs), torch.stack(scs)
p = ops.make_params()
buf = (ctypes.c_longlong * 16)()
for _ in range(3):
    st = ops.forward_boxes(s, rec, _lib.BOX_3D_REC, p, generalized=True, affine=True)
torch.cuda.synchronize()
lib.gnms_debug_elect2_clock_read(buf, 1)
st = ops.forward_boxes(s, rec, _lib.BOX_3D_REC, p, generalized=True, affine=True)
torch.cuda.synchronize()
lib.gnms_debug_elect2_clock_read(buf, 1)
v = list(buf)
names = ["load", "A list", "B pairs", "C resolve", "D sweep", "E evaluate", "F pool", "store"]
for k, nm in enumerate(names):
    print("%-10s %8.2f us" % (nm, v[k] / 1965.0))
print("total %.2f us; steps %d, leaders %d, queued pairs %d, redone steps %d" % (sum(v[:8]) / 1965.0, v[8], v[9], v[10], v[11]))
