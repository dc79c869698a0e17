mkdir -p gpurun_out
python scripts/gemm_one.py 65536 1024 256 0 0 softsign 3 > gpurun_out/one_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/prof_k256 -f python scripts/gemm_one.py 65536 1024 256 0 0 softsign 3 > gpurun_out/ncu_k256.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/one_plain.log
python scripts/gemm_one.py 65536 2048 4096 0 0 softsign 3
python scripts/gemm_one.py 65536 2048 4096 0 0 linear 3
python scripts/gemm_one.py 65536 1024 256 0 0 linear 3
python scripts/gemm_one.py 4096 2048 65536 1 0 linear 3
