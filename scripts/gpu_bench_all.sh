mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chain.py -q -x 2>&1 | tail -3
for wl in infer small cls; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_$wl.log 2> gpurun_out/bench_$wl.err; echo "$wl rc=$?"; cat gpurun_out/bench_$wl.log | cut -c1-2500; tail -3 gpurun_out/bench_$wl.err
done
