set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r2e_pytest.log
cat gpurun_out/r2e_pytest.log
python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
tail -c 3000 gpurun_out/r2e_bench.json
tail -5 gpurun_out/r2e_bench.err
