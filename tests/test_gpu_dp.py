"""Data-parallel parity inside pytest: when the box shows >= 2 GPUs, a 2-rank run of scripts/dp_check.py (NCCL gradient
all-reduce per bucket, pipelined Adam, CUDA-graph replay of the DP step) must reproduce the 1-rank run on the whole
batch: losses to 1e-5 relative, parameter movement to 2e-3 (fp32) / 5e-2 (tf32)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_step_equals_one_rank_step():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    with socket.socket() as s:                       # a free rendezvous port (a fixed one can be taken on a shared box)
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(ROOT, 'scripts', 'dp_check.py')]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith('dp_check[')]
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert len(lines) >= 4 and all(l.endswith('OK') for l in lines), lines
