for ws in 3 4 5; do echo "WS=$ws"; MMAE_CHAIN_WS=$ws timeout 600 python bench.py --workload infer --steps 5 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.2f  chain GB/s %.0f frac %.3f'%(d['ms_per_step'], r['achieved'], r['frac']))"; done
