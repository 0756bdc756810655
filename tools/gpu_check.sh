set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"elect2_kernel" -s 2 -c 1 -f -o gpurun_out/r2_b1_kernels python tools/run_c3_once.py 1 4 > gpurun_out/r2_b1_under_ncu.log 2>&1
tail -3 gpurun_out/r2_b1_under_ncu.log
