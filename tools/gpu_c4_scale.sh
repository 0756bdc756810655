# Run under gpurun --gpus G: the batched-image configuration (C4) at 1, 2, 4[, 8] ranks + the 2-GPU parity test.
set -x
G=${1:-4}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3
for n in 1 2 4 8; do
  if [ $n -le $G ]; then
    if [ $n -eq 1 ]; then
      python bench.py --config c4 --gpus 1 --steps 200 --warmup 10 > gpurun_out/r2_c4_n$n.json 2> gpurun_out/r2_c4_n$n.err
      python bench.py --config c4 --gpus 1 --steps 200 --warmup 10 --bucket-mb 48 > gpurun_out/r2_c4_48mb_n$n.json 2>> gpurun_out/r2_c4_n$n.err
    else
      NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n \
        bench.py --config c4 --gpus $n --steps 200 --warmup 10 > gpurun_out/r2_c4_n$n.json 2> gpurun_out/r2_c4_n$n.err
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
        bench.py --config c4 --gpus $n --steps 200 --warmup 10 --bucket-mb 48 > gpurun_out/r2_c4_48mb_n$n.json 2>> gpurun_out/r2_c4_n$n.err
    fi
    grep '^{' gpurun_out/r2_c4_n$n.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C4 n=%d value %.1f M boxes/s  ms/step %.4f'%(d['n_gpus'], d['value']/1e6, d['ms_per_step']), json.dumps(d['extra']))"
    grep '^{' gpurun_out/r2_c4_48mb_n$n.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C4+48MB n=%d value %.1f M boxes/s  ms/step %.4f'%(d['n_gpus'], d['value']/1e6, d['ms_per_step']), json.dumps(d['extra']))"
    grep -i "NVLS\|via P2P\|Connected all" gpurun_out/r2_c4_n$n.err | head -5
  fi
done
tail -3 gpurun_out/r2_c4_n$G.err
