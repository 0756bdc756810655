set -x
./tools/exp/ab_tall.bin 32 2>&1 | tee gpurun_out/ab_tma_r2.txt
python -m pytest tests/test_gpu_overlaps.py -x -q 2>&1 | tail -15
