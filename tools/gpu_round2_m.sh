# Run under gpurun --gpus 8: what the host side of the box allows (bare cudaMemcpyAsync, no torch), topology, the e2e pipeline
# and C4 at 8 ranks.
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1; lscpu | grep -i "numa\|socket\|^CPU(s)\|model name" >> gpurun_out/r2_topo.txt
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302\|^0x0300" $d/class 2>/dev/null; then echo "$d numa_node=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done >> gpurun_out/r2_topo.txt
cat gpurun_out/r2_topo.txt | tail -30
[ -x tools/exp/h2d_wall.bin ] || (cd tools/exp && nvcc -O3 -o h2d_wall.bin h2d_wall.cu -lpthread); ./tools/exp/h2d_wall.bin 8388608 2228736 400 > gpurun_out/r2_h2d_wall.txt 2>&1; cat gpurun_out/r2_h2d_wall.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu --no-extras > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -c 400 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_n8.json") if l.startswith("{")][-1]); e=d["e2e"]
print("N=8: value %.1f M  e2e %.1f M (samples %s)  copies-only %.1f M  h2d/rank %.1f GB/s  binding %s" % (d["value"]/1e6, e["value"]/1e6, e.get("samples"), e["copies_only_value"]/1e6, e["h2d_GBps_per_rank"], d.get("host_binding")))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --config c4 --gpus 8 --steps 200 --warmup 10 > gpurun_out/r2_c4_n8.json 2> gpurun_out/r2_c4_n8.err
grep '^{' gpurun_out/r2_c4_n8.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C4 n=%d value %.1f M boxes/s  ms/step %.4f'%(d['n_gpus'], d['value']/1e6, d['ms_per_step']), json.dumps(d['extra']))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613 bench.py --config c4 --gpus 4 --steps 200 --warmup 10 > gpurun_out/r2_c4_n4.json 2> gpurun_out/r2_c4_n4.err
grep '^{' gpurun_out/r2_c4_n4.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C4 n=%d value %.1f M boxes/s  ms/step %.4f'%(d['n_gpus'], d['value']/1e6, d['ms_per_step']), json.dumps(d['extra']))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --config c4 --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2_c4_n2.json 2> gpurun_out/r2_c4_n2.err
grep '^{' gpurun_out/r2_c4_n2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C4 n=%d value %.1f M boxes/s  ms/step %.4f'%(d['n_gpus'], d['value']/1e6, d['ms_per_step']), json.dumps(d['extra']))"
python bench.py --config c4 --gpus 1 --steps 200 --warmup 10 --no-cpu > gpurun_out/r2_c4_n1.json 2> gpurun_out/r2_c4_n1.err
grep '^{' gpurun_out/r2_c4_n1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C4 n=%d value %.1f M boxes/s  ms/step %.4f'%(d['n_gpus'], d['value']/1e6, d['ms_per_step']), json.dumps(d['extra']))"
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3
