mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py > gpurun_out/dp_check_$N.log 2>&1; echo "dp_check rc=$?"
grep dp_check gpurun_out/dp_check_$N.log; tail -5 gpurun_out/dp_check_$N.log
for n in 1 $N; do
if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.log 2> gpurun_out/scale_$n.err;
else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 8 --warmup 3 > gpurun_out/scale_$n.log 2> gpurun_out/scale_$n.err; fi
echo "bench $n rc=$?"; tail -1 gpurun_out/scale_$n.log | cut -c1-400
done
