set -x
timeout 600 python -m pytest tests/test_gpu_groomed.py tests/test_gpu_hostapi.py tests/test_gpu_loss_branch.py tests/test_gpu_lossbranch_kernels.py tests/test_gpu_reentrancy.py -x -q -m gpu 2>&1 | tail -15
python tools/elect2_phases.py 1
python tools/chain_phases.py 0
for b in 1 64; do echo "=== images=$b"; timeout 120 python tools/stage_times.py --images $b 2>&1 | grep -v "^matrix=1"; done
