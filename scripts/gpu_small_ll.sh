mkdir -p gpurun_out
for wl in cls; do
CMD="python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
MMAE_GRAPHS=0 $CMD > /dev/null 2>&1 && MMAE_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${wl}_launches.csv $CMD > gpurun_out/${wl}_ncu.log 2>&1
echo "$wl rc=$?"
done
