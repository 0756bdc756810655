set -x
python tools/e2e_probe.py 64
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_probe_launches.csv python tools/e2e_probe.py 64 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r2_probe_launches.csv 2>/dev/null | head -40
