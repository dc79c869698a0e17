timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py tests/test_gpu_graphs.py -q -x 2>&1 | tail -4
timeout 600 python bench.py --workload infer --steps 5 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.2f  chain GB/s %.0f frac %.3f step_frac %.3f  e2e %s'%(d['ms_per_step'], r['achieved'], r['frac'], r['step_frac_of_peak'], d['e2e']))"
