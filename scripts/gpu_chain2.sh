timeout 300 python -m pytest tests/test_gpu_chain.py -q -x 2>&1 | tail -3
timeout 600 python bench.py --workload infer --steps 5 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.2f  chain GB/s %.0f frac %.3f share %.2f'%(d['ms_per_step'], r['achieved'], r['frac'], r['kernel_share_of_step']))"
bash scripts/gpu_trace.sh 2>&1 | grep "tile 5"
