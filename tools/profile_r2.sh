#!/bin/bash
# Run ON THE GPU BOX (under gpurun): launch list of the bench command + `--set full` captures of the dominant kernel launched
# exactly as bench.py launches it (64 images, one 256 x 64 tile per CTA) and of the per-image kernels of the NMS branch.
#   gpurun --timeout 1200 -- 'bash tools/profile_r2.sh'     then, here:   bash tools/profile_r2_summaries.sh
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/r2_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tile_tma_kernel" -s 6 -c 1 -f -o gpurun_out/r2_tile_tma_b64 \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/r2_bench_under_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sort_kernel|rank_kernel|elect2_kernel|backward_mask_kernel|records7" \
    -s 10 -c 5 -f -o gpurun_out/r2_small_kernels python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/r2_bench_under_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"elect2_kernel" -s 2 -c 1 -f -o gpurun_out/r2_b1_kernels \
    python tools/run_c3_once.py 1 4 > gpurun_out/r2_b1_under_ncu.log 2>&1
python bench.py > gpurun_out/r2_bench_b64.json 2> gpurun_out/r2_bench_b64.err
tail -c 600 gpurun_out/r2_bench_b64.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench_b64.err
tail -c 600 gpurun_out/r2_bench_reference_arm.json
ls -la gpurun_out | tail -12
