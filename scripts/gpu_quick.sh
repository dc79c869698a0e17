mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -k tc > gpurun_out/q_tc.log 2>&1; echo "tc rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k tf32 > gpurun_out/q_parity.log 2>&1; echo "parity tf32 rc=$?"
MMAE_PROFILE_DUMP=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_dump.log 2> gpurun_out/bench_dump.err; echo "dump rc=$?"
tail -c 1500 gpurun_out/bench_dump.log
