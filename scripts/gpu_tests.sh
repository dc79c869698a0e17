mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r1_gpu.log 2>&1
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -k simt > gpurun_out/r1_simt.log 2>&1; echo "simt rc=$?" >> gpurun_out/r1_rc.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "not tf32" > gpurun_out/r1_parity_fp32.log 2>&1; echo "parity fp32 rc=$?" >> gpurun_out/r1_rc.log
timeout 240 python -m pytest tests/test_gpu_gemm.py -q -s -k "tc" > gpurun_out/r1_tc.log 2>&1; echo "tc rc=$?" >> gpurun_out/r1_rc.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "tf32" > gpurun_out/r1_parity_tf32.log 2>&1; echo "parity tf32 rc=$?" >> gpurun_out/r1_rc.log
cat gpurun_out/r1_rc.log
timeout 900 python -m pytest tests/test_gpu_model.py -q > gpurun_out/r1_model.log 2>&1; echo "model rc=$?" >> gpurun_out/r1_rc.log
tail -1 gpurun_out/r1_rc.log
