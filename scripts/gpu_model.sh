mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -x > gpurun_out/r1_model.log 2>&1; echo "model rc=$?"
tail -30 gpurun_out/r1_model.log
