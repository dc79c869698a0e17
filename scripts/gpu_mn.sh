timeout 200 python -m pytest tests/test_gpu_gemm.py -q -k tc 2>&1 | tail -2
for v in 1 0; do echo "MMAE_TMA_3D=$v"; MMAE_TMA_3D=$v python scripts/gemm_one.py 4096 2048 65536 1 0 linear 5; MMAE_TMA_3D=$v python scripts/gemm_one.py 8192 8192 8192 1 0 linear 5; MMAE_TMA_3D=$v python scripts/gemm_one.py 8192 8192 8192 0 0 linear 5; done
