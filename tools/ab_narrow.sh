#!/bin/bash
# Run ON THE GPU BOX (under gpurun): first measurement of the experimental 2-rows-per-step matrix-only kernel
# (tile_tall_narrow_kernel, --tall 8, DESIGN.md section 8 lead (b)) and of the kPipe ordering of tile_tall_kernel (--tall 9,
# lead (a)) against the default one (--tall 4).
#   gpurun --timeout 400 -- 'bash tools/ab_narrow.sh'
# 1. its bitwise parity test (skipped unless GNMS_EXPERIMENTAL=1); 2. bench lines of both kernels (roofline.avg_launch_ms is
# the kernel alone, ms_per_step the whole step); 3. if parity holds: one ncu --set full capture of the narrow kernel.
mkdir -p gpurun_out
GNMS_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_overlaps.py -q -k "narrow or tall" 2>&1 | tail -5 | tee gpurun_out/ab_narrow_parity.txt
for tall in 4 8 9 10; do
    timeout 150 python bench.py --no-cpu --tall $tall > gpurun_out/ab_tall_$tall.json 2> gpurun_out/ab_tall_$tall.err
    python - <<PY
import json
d = json.load(open("gpurun_out/ab_tall_$tall.json"))
print("tall=$tall  tile kernel %.1f us  frac %.3f  step %.1f us  value %.1f M boxes/s" % (
    d["roofline"]["avg_launch_ms"] * 1e3, d["roofline"]["frac"], d["ms_per_step"] * 1e3, d["value"] / 1e6))
PY
done | tee gpurun_out/ab_narrow.txt
if grep -q passed gpurun_out/ab_narrow_parity.txt && ! grep -q failed gpurun_out/ab_narrow_parity.txt; then
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tile_tall_narrow_kernel" -s 3 -c 1 -f \
        -o gpurun_out/prof_narrow python bench.py --steps 2 --warmup 3 --no-cpu --tall 8 > gpurun_out/ab_narrow_ncu.log 2>&1
fi
ls -la gpurun_out | tail -8
