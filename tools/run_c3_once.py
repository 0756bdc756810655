"""One C3 image (or a batch) through the matrix-free forward + backward a few times: the command profiled by ncu for the per-image
kernels (rank / elect2 / chain / backward).   python tools/run_c3_once.py [images] [iterations]"""
import sys
import torch
sys.path.insert(0, ".")
from groomed_nms_b200 import _lib, ops, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
recs, scs = [], []
for i in range(B):
    b7, sc = synthetic.config_c3(seed=3 + 10 * i)
    recs.append(ops.box3d_records(ops.corners_from_boxes7(torch.from_numpy(b7).to(dev))))
    scs.append(torch.from_numpy(sc).to(dev))
rec, s = torch.stack(recs), torch.stack(scs)
p = ops.make_params()
g = torch.randn(B, 4096, device=dev)
for _ in range(iters):
    st = ops.forward_boxes(s, rec, _lib.BOX_3D_REC, p, generalized=True, affine=True)
    ops.backward(st, g)
torch.cuda.synchronize()
print("leaders", int((st.lead[0] == torch.arange(4096, device=dev)).sum()), "valid", int(st.counts[0, 0]))
