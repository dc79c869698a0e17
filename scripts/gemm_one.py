"""Runs one GEMM shape through the tcgen05 family a few times (ncu target / quick timing)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalautoencoder_b200 import debug_gemm

M, N, K = (int(x) for x in sys.argv[1:4])
ta, tb = int(sys.argv[4]), int(sys.argv[5])
act = sys.argv[6] if len(sys.argv) > 6 else 'linear'
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
A = torch.randn((K, M) if ta else (M, K), device='cuda')
B = torch.randn((N, K) if tb else (K, N), device='cuda')
bias = torch.randn(N, device='cuda') if act != 'linear' else None
for _ in range(2):
    C = debug_gemm(A, B, bool(ta), bool(tb), bias=bias, activation=act, precision='tf32')
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    C = debug_gemm(A, B, bool(ta), bool(tb), bias=bias, activation=act, precision='tf32')
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print('M=%d N=%d K=%d ta=%d tb=%d act=%s: %.3f ms  %.1f TFLOP/s  out %.1f GB/s' % (M, N, K, ta, tb, act, ms, 2.0 * M * N * K / ms / 1e9, M * N * 4 / ms / 1e6))
