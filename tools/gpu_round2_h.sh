set -x
timeout 900 python -m pytest tests/test_gpu_groomed.py tests/test_gpu_hostapi.py tests/test_gpu_loss_branch.py tests/test_gpu_lossbranch_kernels.py tests/test_gpu_reentrancy.py -x -q -m gpu 2>&1 | tail -8
for b in 16 64; do echo "=== images=$b"; timeout 120 python tools/stage_times.py --images $b 2>&1 | grep -v "^matrix=1"; done
python bench.py --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']
print('value %.1f M  ms/step %.4f  other_path %.4f ms (%.2f G)  e2e %.1f M  blocking %.1f M' % (d['value']/1e6, d['ms_per_step'], d['other_path']['ms_per_step'], d['other_path']['value']/1e9, e['value']/1e6, e['blocking_call_value']/1e6))"
