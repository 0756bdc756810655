# Run under gpurun --gpus G (G = 4 or 8): the batched-image configuration (C4) at 1, 2, 4[, 8] ranks with the in-kernel peer all-reduce,
# NCCL beside it at the largest rank count, + the 2-GPU parity tests.
set -x
G=${1:-4}
mkdir -p gpurun_out
run() {   # ranks collective tag
  if [ $1 -eq 1 ]; then
    python bench.py --config c4 --gpus 1 --steps 200 --warmup 10 --no-cpu > gpurun_out/r2_c4_$3_n$1.json 2> gpurun_out/r2_c4_$3_n$1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2950$1 \
      bench.py --config c4 --gpus $1 --steps 200 --warmup 10 --c4-collective $2 > gpurun_out/r2_c4_$3_n$1.json 2> gpurun_out/r2_c4_$3_n$1.err
  fi
  grep '^{' gpurun_out/r2_c4_$3_n$1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C4 n=%d %s value %.1f M boxes/s  ms/step %.4f'%(d['n_gpus'], d['extra']['strong']['collective'], d['value']/1e6, d['ms_per_step']), json.dumps(d['extra']))"
}
for n in $G $((G/2)) 2 1; do if [ $n -ge 1 ] && [ ! -f gpurun_out/r2_c4_peer_n$n.json.done ]; then run $n peer peer; touch gpurun_out/r2_c4_peer_n$n.json.done; fi; done
run $G nccl nccl
rm -f gpurun_out/*.done
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3
