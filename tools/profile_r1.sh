#!/bin/bash
# Run ON THE GPU BOX (under gpurun): launch list of the bench command + one `--set full` capture of the hot kernels.
#   gpurun --timeout 900 -- 'bash tools/profile_r1.sh'
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"tile_tall_kernel|tile_kernel|sort_kernel|rank_kernel|spatial_kernel|elect_kernel|chain_kernel|has_earlier_kernel|backward_mask_kernel" \
    -s 10 -c 12 -f -o gpurun_out/prof python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out
