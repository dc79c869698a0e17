# full round-style validation: gpu tests, smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/all_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -2 gpurun_out/all_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cat gpurun_out/bench_default.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.log
