# ncu evidence for profiles/: launch lists of short bench runs + full captures of the dominant kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mmae -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "wide launch list rc=$?"
# all 17 tcgen05 GEMM launches of one wide step (the second one), full set
$CMD > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none -k regex:gemm_tc -s 17 -c 17 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "wide full capture rc=$?"; tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm_raw.csv 2>/dev/null; rm -f gpurun_out/prof_gemm.ncu-rep   # the report itself exceeds what travels back
# fill-in inference: launch list + full capture of the whole-network kernel
CMDI="python bench.py --workload infer --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMDI > gpurun_out/prof_infer_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mmae -c 100 --csv --log-file gpurun_out/infer_launches.csv $CMDI > gpurun_out/ncu_infer_launches.log 2>&1
echo "infer launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:chain_tc -s 1 -c 1 -o gpurun_out/prof_chain_infer -f $CMDI > gpurun_out/ncu_chain_full.log 2>&1
echo "infer full capture rc=$?"; tail -2 gpurun_out/ncu_chain_full.log
ncu -i gpurun_out/prof_chain_infer.ncu-rep --page raw --csv > gpurun_out/prof_chain_infer_raw.csv 2>/dev/null
ls -la gpurun_out | head -30; du -sh gpurun_out
