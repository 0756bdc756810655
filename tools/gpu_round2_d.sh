set -x
python -m pytest tests/test_gpu_groomed.py tests/test_gpu_classical.py -x -q -k "soft_sort or aploss or inverse_modes or direct" 2>&1 | tail -15
python -m pytest tests/test_gpu_c5_train_step.py -x -q -s 2>&1 | grep -v "^$" | tail -12
python - <<'PY'
import torch, time
from groomed_nms_b200.lib import groomed_nms as G
for n in (1024, 4096):
    s = torch.rand(n, device="cuda"); m = torch.rand(n, n, device="cuda")
    for _ in range(2): G.soft_sort(s, m, 0.01)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): G.soft_sort(s, m, 0.01)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print("soft_sort N=%d: %.2f ms (GEMM %.1f GFLOP -> %.1f TFLOP/s incl. the rest)" % (n, dt * 1e3, 2 * n ** 3 / 1e9, 2 * n ** 3 / dt / 1e12))
PY
