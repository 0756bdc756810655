set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r2e_pytest.log
cat gpurun_out/r2e_pytest.log
python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
tail -c 2500 gpurun_out/r2e_bench.json
tail -5 gpurun_out/r2e_bench.err
python bench.py --config c4 --no-cpu > gpurun_out/r2e_c4_n1.json 2> gpurun_out/r2e_c4_n1.err; tail -c 1500 gpurun_out/r2e_c4_n1.json; tail -3 gpurun_out/r2e_c4_n1.err
