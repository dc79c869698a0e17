mkdir -p gpurun_out
MMAE_PROFILE_DUMP=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_dump.log 2> gpurun_out/bench_dump.err; echo "dump rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.log
