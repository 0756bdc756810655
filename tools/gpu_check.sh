set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_groomed.py tests/test_gpu_overlaps.py tests/test_gpu_inference_site.py tests/test_gpu_hostapi.py tests/test_gpu_head.py -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r2_memcheck.txt
cat gpurun_out/r2_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/run_c3_once.py 2 1 2>&1 | tail -12 > gpurun_out/r2_racecheck.txt
cat gpurun_out/r2_racecheck.txt
