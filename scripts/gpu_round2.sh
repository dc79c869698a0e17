mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -2 gpurun_out/all_gpu.log
for wl in small cls; do timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$wl ms/step %.4f launches/step %.1f'%(d['ms_per_step'], d['gpu_launches']/d['steps']))"; done
