MMAE_CHAIN_TRACE=1 timeout 300 python - <<'PY' 2>&1 | tail -45
import os, sys
sys.path.insert(0, os.getcwd())
import torch
src = open('scripts/time_small.py').read().split("for B in [int(x)")[0]
exec(src)
B = 2000000
X = torch.rand((B, 320), device='cuda')
X[::3, 200:220] = -1.0
e = mk(1, B)
out = torch.empty_like(X)
e.forward_into(X, filled=out)
torch.cuda.synchronize()
PY
