mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_gemm.py -q -x -k "tc" > gpurun_out/tc2_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -12 gpurun_out/tc2_gemm.log
