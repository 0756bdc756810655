set -x
for i in 1 2; do python bench.py --no-cpu --no-extras > gpurun_out/r2k_bench_$i.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/r2k_bench_$i.json")); e=d["e2e"]
print("run $i: value %.1f M  e2e %.1f M  r1-variant %.1f M  copies-only %.1f M  blocking %.1f M" % (d["value"]/1e6, e["value"]/1e6, e["with_host_gradient_and_int64_keep_lists"]["value"]/1e6, e["copies_only_value"]/1e6, e["blocking_call_value"]/1e6))
PY
done
