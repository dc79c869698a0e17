mkdir -p gpurun_out
for wl in infer small cls; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_$wl.log 2> gpurun_out/bench_$wl.err; echo "$wl rc=$?"
done
timeout 600 python scripts/time_model.py 2>&1 | tail -6 > gpurun_out/time_model.log; cat gpurun_out/time_model.log
