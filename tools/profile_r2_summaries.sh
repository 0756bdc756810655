#!/bin/bash
# Run HERE after tools/profile_r2.sh came back: turns gpurun_out/r2_* into the committed summaries under profiles/.
set -e
cd "$(dirname "$0")/.."
HEAD=$(git log -1 --format=%h)
python tools/summarize_launches.py gpurun_out/r2_launches.csv > profiles/r2_launches_bench_b64.txt
cp gpurun_out/r2_launches.csv profiles/r2_launches_bench_b64.csv
python tools/ncu_report.py gpurun_out/r2_tile_tma_b64.ncu-rep tile_tma_kernel \
  "ncu --set full --clock-control none --import-source on -k regex:tile_tma_kernel -s 6 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu --no-extras" \
  "the launch bench.py times: 64 images of N=4096, one 256 x 64 tile per CTA x 4 tiles per CTA; build = $HEAD + working tree of that run (tools/profile_r2.sh)" \
  > profiles/r2_ncu_tile_tma_kernel_b64.txt
for k in sort_kernel rank_kernel elect2_kernel backward_mask_kernel records7; do
  python tools/ncu_report.py gpurun_out/r2_small_kernels.ncu-rep $k "ncu --set full, bench.py step (64 images of N=4096 per launch), build $HEAD" ; echo; echo "=================================================================="; echo
done > profiles/r2_ncu_small_kernels.txt
for k in elect2_kernel; do
  python tools/ncu_report.py gpurun_out/r2_b1_kernels.ncu-rep $k "ncu --set full, ONE image of N=4096 (tools/run_c3_once.py 1 4), build $HEAD" ; echo; echo "=================================================================="; echo
done > profiles/r2_ncu_b1_kernels.txt
python - <<'PY'
import csv, json, subprocess
raw = subprocess.run(["ncu", "-i", "gpurun_out/r2_tile_tma_b64.ncu-rep", "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(raw.splitlines())); h, u = rows[0], rows[1]; r = rows[2]
def val(m):
    v = float(r[h.index(m)]); unit = u[h.index(m)]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
out = {"kernel": r[h.index("Kernel Name")], "images_per_launch": 64, "n_boxes": 4096, "dram_bytes_read": val("dram__bytes_read.sum"),
       "dram_bytes_write": val("dram__bytes_write.sum"), "algorithmic_bytes": 64 * (4 * 4096 ** 2 + 32 * 4096),
       "source": "ncu --set full --clock-control none (tools/profile_r2.sh), one launch inside bench.py, profiles/r2_ncu_tile_tma_kernel_b64.txt"}
json.dump(out, open("profiles/r2_tile_kernel_traffic.json", "w"), indent=1)
print(out)
PY
grep '^{' gpurun_out/r2_bench_b64.json | tail -1 > profiles/r2_bench_b64.json
grep '^{' gpurun_out/r2_bench_reference_arm.json | tail -1 > profiles/r2_bench_reference_arm.json
ls -la profiles | tail -20
