set -x
python -m pytest tests/test_gpu_head.py tests/test_gpu_reentrancy.py tests/test_gpu_hostapi.py tests/test_gpu_reference_dropin.py -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -c 2500 gpurun_out/r2_bench_b.json; tail -5 gpurun_out/r2_bench_b.err
python bench.py --config c4 --steps 50 --warmup 5 > gpurun_out/r2_c4_n1.json 2> gpurun_out/r2_c4_n1.err; cat gpurun_out/r2_c4_n1.json; tail -5 gpurun_out/r2_c4_n1.err
