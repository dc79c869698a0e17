# compute-sanitizer over a subset of the -m gpu suite sized to fit a few minutes (memcheck slows kernels 10-50x):
# golden vectors (SIMT + Adam + noise + head kernels), tcgen05 GEMM tests, chain / backward chain / grouped wgrad tests.
mkdir -p gpurun_out
SEL='not many and not pipelined and not million and not wide_config'
timeout 480 compute-sanitizer --tool memcheck --error-exitcode 86 --launch-timeout 0 \
  python -m pytest tests/test_golden.py tests/test_gpu_gemm.py tests/test_gpu_chain.py -m gpu -q -x -k "$SEL" > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_memcheck.log | tail -3
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 86 --launch-timeout 0 \
  python -m pytest tests/test_golden.py tests/test_gpu_chain.py -m gpu -q -x -k "tiny or S-untied or S-tied" > gpurun_out/r2_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2_racecheck.log | tail -3
