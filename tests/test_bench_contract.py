"""bench.py's reference arm (the one leg that runs on CPU) prints ONE JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], capture_output=True, text=True, timeout=600,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_line_small():
    d = _run('--impl', 'reference', '--workload', 'small', '--steps', '1', '--warmup', '1')
    assert d['impl'] == 'reference' and d['unit'] == 'samples/s' and d['higher_is_better'] is True
    assert d['metric'].startswith('MMAE train samples/sec') and d['value'] > 0
    assert d['config']['workload'] == 'small'
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'rows' in cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['gpu_launches'] == 0 and d['vs_baseline'] is None


def test_reference_arm_other_ranks_stay_silent(monkeypatch):
    monkeypatch.setenv('RANK', '1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'small',
                          '--steps', '1', '--warmup', '1'], capture_output=True, text=True, timeout=120, cwd=ROOT,
                         env=dict(os.environ, RANK='1'))
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_workload_tables_are_consistent():
    sys.path.insert(0, ROOT)
    import bench
    for name, w in bench.WORKLOADS.items():
        _, starts, names = bench.workload_cfg(name)
        assert starts[0] == 0 and starts[-1] == w['F'] and len(names) == len(w['blocks'])
        assert {'call', 'sms', 'screen', 'location'} <= set(names)          # intelligent noise looks these up by name (:694)
        assert name in bench.FLOPS_PER_SAMPLE and name in bench.METRIC and name in bench.CPU_ROWS
    # GEMM FLOPs per sample = 2*in*out per layer, bwd = 2x fwd minus the first-layer dgrad
    w = bench.WORKLOADS['wide']
    dims = [w['F']] + w['layers']
    fwd = 2 * sum(2 * a * b for a, b in zip(dims[:-1], dims[1:]))
    assert fwd * 3 - 2 * dims[0] * dims[1] == bench.FLOPS_PER_SAMPLE['wide']
