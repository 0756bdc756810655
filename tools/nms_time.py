import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from groomed_nms_b200 import ops, _lib, synthetic
from groomed_nms_b200.lib.nms.gpu_nms import gpu_nms
for n in (500, 3000, 8000):
    bx, sc, _ = synthetic.clustered_boxes_2d(n, 40, seed=1, jitter=0.08)
    dets = np.concatenate([bx, sc[:, None]], 1).astype(np.float32)
    d = torch.from_numpy(dets).cuda()
    for _ in range(3): k, nk = ops.hard_nms(d, 0.4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): k, nk = ops.hard_nms(d, 0.4)
    e1.record(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): keep = gpu_nms(dets, 0.4)
    dt = (time.perf_counter() - t0) / 20
    print("N=%d  device %.1f us per call, gpu_nms() host call %.1f us, kept %d" % (n, e0.elapsed_time(e1) / 20 * 1e3, dt * 1e6, len(keep)))
