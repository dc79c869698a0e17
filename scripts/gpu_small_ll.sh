mkdir -p gpurun_out
CMD="python bench.py --workload small --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
MMAE_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"_kernel" -c 300 --csv --log-file gpurun_out/small_launches.csv $CMD > gpurun_out/small_ncu.log 2>&1
echo "rc=$?"
