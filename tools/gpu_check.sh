set -x
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -8
