set -x
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_head.py -q -x 2>&1 | tail -15
for c in peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29650 bench.py --config c4 --gpus 2 --steps 200 --warmup 10 --c4-collective $c 2>/tmp/err_$c.txt | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C4 n=%d %s ms/step %.4f'%(d['n_gpus'], d['extra']['strong']['collective'], d['ms_per_step']), json.dumps(d['extra']))"
tail -3 /tmp/err_$c.txt
done
