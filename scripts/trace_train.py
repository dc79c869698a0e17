"""MMAE_CHAIN_TRACE=1 timeline of the forward and backward chain kernels inside one small-config train step."""
import os, sys
os.environ['MMAE_CHAIN_TRACE'] = '1'
os.environ['MMAE_GRAPHS'] = '0'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'time_small.py')).read().split("for B in [int(x)")[0]
exec(src)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
X = torch.rand((B, 320), device='cuda')
e = mk(1, B)
for _ in range(2):
    e.gen_noise(B); e.train_step(X, noise=True)
torch.cuda.synchronize()
