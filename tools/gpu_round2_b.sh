set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -c 4000 gpurun_out/r2_bench_b.json; tail -5 gpurun_out/r2_bench_b.err
python -m pytest tests/test_gpu_overlaps.py -q -k "kitti" 2>&1 | tail -3
