set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_groomed.py tests/test_gpu_overlaps.py -q -m gpu -k "soft_sort or autograd" 2>&1 | tail -15
for b in 1 4 16; do
  for e in 1 2; do
    echo "=== images=$b election=$e"
    python tools/stage_times.py --images $b --election $e 2>&1 | grep -v "^matrix=1" 
  done
done
echo "=== images=1 rank-by-sort=0 / 1"
python tools/stage_times.py --images 1 --rank-by-sort 0 2>&1 | grep "rank"
python tools/stage_times.py --images 1 --rank-by-sort 1 2>&1 | grep "rank"
