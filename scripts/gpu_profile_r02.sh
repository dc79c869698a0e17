# Round-2 evidence for profiles/: bench lines, ncu launch lists, --set full captures of the dominant kernels.
mkdir -p gpurun_out
# 1. the default invocation (what the driver runs): wide + others + api + cpu_baseline
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "default rc=$?"
python scripts/show_bench.py gpurun_out/r02_bench_default.json
# 2. reference arm (CPU port on all host cores, full-size steps when they fit the budget)
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r02_bench_reference.json
# 3. grid workload on one GPU
timeout 600 python bench.py --workload grid > gpurun_out/r02_bench_grid_1gpu.json 2> gpurun_out/r02_bench_grid_1gpu.err; echo "grid rc=$?"; cut -c1-500 gpurun_out/r02_bench_grid_1gpu.json
# 4. wide: launch list + full capture of the 17 tcgen05 GEMM launches of one step
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-others"
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mmae -c 400 --csv --log-file gpurun_out/r02_launches_wide.csv $CMD > /dev/null 2>&1
echo "wide launch list rc=$?"
ncu --set full --clock-control none -k regex:gemm_tc -s 17 -c 17 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "wide full capture rc=$?"
ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv > gpurun_out/r02_wide_raw.csv 2>/dev/null; rm -f gpurun_out/prof_gemm.ncu-rep
# 5. small: launch list (eager, so that every kernel is listed) + full capture of one step's chain / wgrad kernels
CMDS="python bench.py --workload small --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
MMAE_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mmae -c 200 --csv --log-file gpurun_out/r02_launches_small.csv $CMDS > /dev/null 2>&1
echo "small launch list rc=$?"
MMAE_GRAPHS=0 ncu --set full --clock-control none -k regex:"chain_tc|wgrad_group|sample_noise|grad_assemble" -s 5 -c 5 -o gpurun_out/prof_small -f $CMDS > gpurun_out/ncu_small_full.log 2>&1
echo "small full capture rc=$?"
ncu -i gpurun_out/prof_small.ncu-rep --page raw --csv > gpurun_out/r02_small_raw.csv 2>/dev/null; rm -f gpurun_out/prof_small.ncu-rep
ls -la gpurun_out | grep r02_
