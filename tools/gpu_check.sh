set -x
timeout 600 python -m pytest tests/test_gpu_groomed.py -x -q -m gpu -k "fuzz" 2>&1 | tail -25
