mkdir -p gpurun_out
CMDI="python bench.py --workload infer --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chain_tc|missing_bits|fill_select|transpose_kernel|finalize_scalars|reduce_partials|gemm_" -c 100 --csv --log-file gpurun_out/infer_launches.csv $CMDI > gpurun_out/ncu_infer_launches.log 2>&1; echo "infer rc=$?"
grep -c mmae gpurun_out/infer_launches.csv
