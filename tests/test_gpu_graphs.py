"""CUDA-graph replay of the train steps: after one eager step and one captured step, every further step with the
same (input buffer, batch, flags) is a single graph launch.  Per-step scalars (Philox step, Adam t / alpha) live in
device memory, so replays must be BIT-IDENTICAL to the eager path -- same kernels, same order, same arguments."""
import os

import numpy as np
import pytest
import torch

from oracle import mmae_oracle as O
from tests.helpers import make_cfgs

pytestmark = pytest.mark.gpu


def _engine(ecfg, P, graphs):
    from multimodalautoencoder_b200 import Engine
    old = os.environ.get('MMAE_GRAPHS')
    os.environ['MMAE_GRAPHS'] = '1' if graphs else '0'
    try:
        e = Engine(ecfg)
        e.set_params({k: v.astype(np.float32) for k, v in P.items()})
        e._graph_env_probe = e.graph_replays
        # the switch is read at the first train step
        X0 = torch.zeros((32, ecfg.num_feats), device='cuda')
        e.gen_noise(32)
        e.train_step(X0, noise=True)
        e.set_params({k: v.astype(np.float32) for k, v in P.items()})
    finally:
        if old is None:
            os.environ.pop('MMAE_GRAPHS', None)
        else:
            os.environ['MMAE_GRAPHS'] = old
    return e


@pytest.mark.parametrize('case', [dict(layers=(128, 64), tie=False, B=256, keep=1.0, prec='tf32'),
                                  dict(layers=(200, 100), tie=True, B=100, keep=0.5, prec='tf32'),
                                  dict(layers=(128, 64), tie=False, B=20, keep=1.0, prec='fp32')],
                         ids=['S-chain-256', 'C-dropout-100', 'fp32-batch-20'])
def test_graph_replay_is_bit_identical(case):
    ocfg, ecfg = make_cfgs(precision=case['prec'], layers=case['layers'], tie=case['tie'], lam=0.001, head=[50, 20], seed=3)
    B = case['B']
    rng = np.random.default_rng(21)
    P = O.init_params(ocfg, rng)
    X = torch.as_tensor(rng.uniform(0, 1, (B, 320)).astype(np.float32), device='cuda')
    Y = torch.as_tensor((rng.uniform(size=(B, 3)) < 0.5).astype(np.float32), device='cuda')
    eg, ee = _engine(ecfg, P, True), _engine(ecfg, P, False)
    losses = {0: [], 1: []}
    for k, e in enumerate((eg, ee)):
        for step in range(6):
            e.set_rng_step(100 + step)
            e.gen_noise(B)
            e.train_step(X, noise=True, keep=case['keep'])
            losses[k].append(e.scalars()['recon_loss'])
            e.gen_noise(B)
            e.cls_train_step(X, Y, noise=True, keep=case['keep'])
            losses[k].append(e.scalars()['head_loss'])
    assert eg.graph_replays >= 6, 'later steps of both optimizers should be graph replays'
    assert ee.graph_replays == 0
    assert losses[0] == losses[1]
    for name, _ in eg.variables():
        assert np.array_equal(eg.get_variable(name), ee.get_variable(name)), name
    st_g, st_e = eg.get_opt_state(0, 'weights0'), ee.get_opt_state(0, 'weights0')
    assert st_g[2] == st_e[2] and np.array_equal(st_g[0], st_e[0]) and np.array_equal(st_g[1], st_e[1])
    eg.close(); ee.close()


def test_graph_survives_new_buffers_and_set_variable():
    """A different input tensor (new pointer) or batch size falls back to eager / a new graph; set_variable between
    replays is honoured (weight shadows are refreshed inside the graph)."""
    ocfg, ecfg = make_cfgs(precision='tf32', tie=False, seed=4)
    rng = np.random.default_rng(22)
    P = O.init_params(ocfg, rng)
    eg, ee = _engine(ecfg, P, True), _engine(ecfg, P, False)
    Xs = [torch.as_tensor(rng.uniform(0, 1, (b, 320)).astype(np.float32), device='cuda') for b in (128, 128, 64)]
    for e in (eg, ee):
        for step in range(15):
            X = Xs[step % 3]
            if step == 9:
                e.set_variable('weights0', P['weights0'].astype(np.float32))
            e.set_rng_step(step)
            e.gen_noise(X.shape[0])
            e.train_step(X, noise=True)
    assert eg.graph_replays >= 3 and ee.graph_replays == 0
    for name, _ in eg.variables():
        assert np.array_equal(eg.get_variable(name), ee.get_variable(name)), name
    eg.close(); ee.close()
