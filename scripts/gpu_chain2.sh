timeout 300 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py -q -x 2>&1 | tail -3
timeout 600 python bench.py --workload infer --steps 5 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.2f  chain GB/s %.0f frac %.3f share %.2f step_frac %.3f launches %d'%(d['ms_per_step'], r['achieved'], r['frac'], r['kernel_share_of_step'], r['step_frac_of_peak'], d['gpu_launches']))"
