#!/bin/bash
# Run ON THE GPU BOX (under gpurun): launch list of the bench command + one `--set full` capture of the dominant kernel
# launched exactly as bench.py launches it (64 images, one 256 x 64 tile per CTA).
#   gpurun --timeout 1200 -- 'bash tools/profile_r2.sh'
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/r2_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tile_tma_kernel" -s 6 -c 1 -f -o gpurun_out/r2_tile_tma_b64 \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/r2_bench_under_ncu2.log 2>&1
ncu --set full --clock-control none -k regex:"sort_kernel|rank_kernel|spatial_kernel|elect_kernel|chain_kernel|backward_mask_kernel|records" \
    -s 12 -c 8 -f -o gpurun_out/r2_small_kernels python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/r2_bench_under_ncu3.log 2>&1
ls -la gpurun_out
