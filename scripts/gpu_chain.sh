mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chain.py -q -x > gpurun_out/chain_test.log 2>&1; echo "chain tests rc=$?"; tail -5 gpurun_out/chain_test.log
for st in 0 60 120; do echo "stagger=$st"; MMAE_CHAIN_STAGGER=$st timeout 300 python scripts/time_small.py 65536 2000000 2>&1 | grep "chain=1"; done
MMAE_CHAIN_STAGGER=0 bash scripts/gpu_trace.sh 2>&1 | grep "tile [5] "
