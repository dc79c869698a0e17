# data-parallel check + scaling of the wide workload on the GPUs of this box
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py > gpurun_out/dp_check_$N.log 2>&1; echo "dp_check rc=$?"; grep "dp_check\[" gpurun_out/dp_check_$N.log | sort -u
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_1.log 2> gpurun_out/scale_1.err; echo "bench 1 rc=$?"; cut -c1-330 gpurun_out/scale_1.log
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.log 2> gpurun_out/scale_$n.err; echo "bench $n rc=$?"; grep '^{' gpurun_out/scale_$n.log | cut -c1-330
  fi
done
