#!/usr/bin/env python
"""Summarise a `--set full` .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small table for profiles/."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
cols = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active"]
idx = [hdr.index(c) for c in cols if c in hdr]
units = rows[1]
print(" | ".join("%s [%s]" % (hdr[i], units[i]) for i in idx))
seen = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    seen[name] = seen.get(name, 0) + 1
    if seen[name] > 1:
        continue
    print(" | ".join(r[i][:70] for i in idx))
