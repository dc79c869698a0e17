mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chain.py -q -x > gpurun_out/chain_test.log 2>&1; echo "chain tests rc=$?"; tail -25 gpurun_out/chain_test.log
timeout 300 python scripts/time_small.py 4096 65536 2000000 2>&1 | tail -8
