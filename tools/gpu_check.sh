set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_final_pytest.log; cat gpurun_out/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench_b64.json 2> gpurun_out/r2_bench_b64.err; tail -c 300 gpurun_out/r2_bench_b64.json; tail -3 gpurun_out/r2_bench_b64.err
