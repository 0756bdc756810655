"""Where does a host-buffer call spend its time?  Device time of the kernel graph of HostRunner variants (CUDA events around
replays) and wall time of the blocking call.   python tools/e2e_probe.py [images]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from groomed_nms_b200 import ops, synthetic
from groomed_nms_b200.hostapi import HostRunner
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = 4096
dev = torch.device("cuda", 0)
boxes = np.stack([synthetic.config_c3(seed=3 + 10 * i)[0] for i in range(B)])
scores = np.stack([synthetic.config_c3(seed=3 + 10 * i)[1] for i in range(B)])
grads = np.random.default_rng(0).standard_normal((B, N)).astype(np.float32)
hb, hs, hg = torch.from_numpy(boxes).pin_memory(), torch.from_numpy(scores).pin_memory(), torch.from_numpy(grads).pin_memory()
params = ops.make_params()
for keep_cap, god in ((512, True), (0, False), (512, False), (0, True)):
    r = HostRunner(B, N, dev, params, materialise=False, keep_cap=keep_cap, grad_on_device=god)
    if god:
        r.set_grad(torch.from_numpy(grads).to(dev))
    for _ in range(3):
        r.run_host(hb, hs, None if god else hg)
    st = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(20):
        r.graph.replay()
    e1.record(st)
    e1.synchronize()
    g_us = e0.elapsed_time(e1) / 20 * 1e3
    t0 = time.perf_counter()
    for _ in range(50):
        r.run_host(hb, hs, None if god else hg)
    w_us = (time.perf_counter() - t0) / 50 * 1e6
    print("keep_cap=%d grad_on_device=%d: graph %.1f us per replay, blocking call %.1f us (%.0f M boxes/s)" % (keep_cap, god, g_us, w_us, B * N / w_us))
