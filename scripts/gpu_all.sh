mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
tail -3 gpurun_out/all_gpu.log
MMAE_PROFILE_DUMP=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_dump.log 2> gpurun_out/bench_dump.err; echo "dump rc=$?"
tail -c 1800 gpurun_out/bench_dump.log | head -c 300
timeout 600 python bench.py --workload small --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/small.err | cut -c1-700
