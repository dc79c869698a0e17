mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graphs.py -q -x > gpurun_out/graph_test.log 2>&1; echo "graph tests rc=$?"; tail -15 gpurun_out/graph_test.log
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/all_gpu.log 2>&1; echo "all gpu rc=$?"; tail -4 gpurun_out/all_gpu.log
for wl in cls small; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$wl.log 2> gpurun_out/bench_$wl.err; echo "$wl rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$wl.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e'] and d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline'].get('kernel_share_of_step'))
PY
done
