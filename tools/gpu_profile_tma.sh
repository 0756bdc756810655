set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tile_tma_kernel -s 2 -c 1 -o gpurun_out/r2_tma_full -f ./tools/exp/ab_tall.bin 32 > gpurun_out/r2_tma_ncu.log 2>&1
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_tma.json 2> gpurun_out/r2_bench_tma.err; tail -c 3000 gpurun_out/r2_bench_tma.json; tail -5 gpurun_out/r2_bench_tma.err
