for d in 0 1; do echo "dbg=$d"; MMAE_CHAIN_DBG=$d bash scripts/gpu_trace.sh 2>&1 | grep "tile 5 op 0\|fine"; done
