set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_c4_launches.csv python bench.py --config c4 --steps 3 --warmup 1 --no-cpu > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2_c4_launches.csv")) if len(r)>10]
h=rows[0]; ki,vi,gi=h.index("Kernel Name"),h.index("Metric Value"),h.index("Grid Size")
agg=collections.OrderedDict()
for r in rows[1:]:
    try: agg.setdefault((r[ki][:60], r[gi]),[]).append(float(r[vi].replace(",","")))
    except ValueError: pass
for (k,g),v in agg.items(): print("%-62s grid %-16s n=%3d avg %.1f us"%(k,g,len(v),sum(v)/len(v)/1e3))
PY
