set -x
timeout 900 python -m pytest tests/test_gpu_groomed.py -x -q -m gpu -k "degenerate" 2>&1 | tail -30
