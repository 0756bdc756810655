set -x
timeout 600 python -m pytest tests/test_gpu_groomed.py tests/test_gpu_hostapi.py tests/test_gpu_loss_branch.py tests/test_gpu_loss_ref.py tests/test_gpu_classical.py tests/test_gpu_inference_site.py -x -q -m gpu 2>&1 | tail -6
timeout 60 python tools/chain_phases.py 0
timeout 120 python tools/nms_time.py
timeout 300 python bench.py --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']
print('value %.1f M  ms/step %.4f  other_path %.4f ms (%.2f G)  e2e %.1f M' % (d['value']/1e6, d['ms_per_step'], d['other_path']['ms_per_step'], d['other_path']['value']/1e9, e['value']/1e6))
print(json.dumps(d['extra']['latency_us']['c_abi_graph_N4096_3d_matrix_free']))"
