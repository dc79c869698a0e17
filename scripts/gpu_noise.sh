timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "wide-first" 2>&1 | tail -2
for fn in 0 1; do
echo "FUSE=$fn"
MMAE_FUSE_NOISE=$fn timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.3f gemm TF/s %.1f'%(d['ms_per_step'], d['roofline']['achieved']))"
done
