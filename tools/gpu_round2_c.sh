set -x
python -m pytest tests/test_gpu_lossbranch_kernels.py tests/test_gpu_loss_branch.py tests/test_gpu_loss_ref.py tests/test_gpu_inference_site.py tests/test_gpu_hostapi.py -x -q 2>&1 | tail -15
python -m pytest tests/test_gpu_reference_dropin.py tests/test_gpu_c5_train_step.py -x -q -s 2>&1 | grep -v "^$" | tail -25
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench_c.json') if l.startswith('{')][-1]); print(json.dumps(d['e2e'], indent=0)); print(d['value'], d['roofline']['frac'])"; tail -3 gpurun_out/r2_bench_c.err
