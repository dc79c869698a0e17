mkdir -p gpurun_out
cat > /tmp/fwd_once.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'scripts'))
import numpy as np, torch
import importlib.util
spec = importlib.util.spec_from_file_location('ts', 'scripts/time_small.py')
src = open('scripts/time_small.py').read().split("for B in [int(x)")[0]
exec(src)
B = 2000000
X = torch.rand((B, 320), device='cuda')
e = mk(1, B)
for _ in range(3):
    e.forward(X, recon=True, loss=True)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_tc -s 2 -c 1 -o gpurun_out/prof_chain -f python /tmp/fwd_once.py > gpurun_out/ncu_chain.log 2>&1
echo rc=$?; tail -3 gpurun_out/ncu_chain.log
