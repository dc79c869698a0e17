"""The whole-network tcgen05 kernel (csrc/chain_tc.cu: encoder + decoder + loss in one launch, activations
resident in TMEM) against (a) the fp64 oracle and (b) the per-layer tcgen05 GEMM path on the same inputs.

Tolerances: the tf32 ones of test_gpu_parity.py (loss <= 1e-3 relative, reconstructions <= 5e-3
relative-to-max); chain vs per-layer (both round operands to tf32 the same way, only the accumulation order
inside a tile is shared too) <= 2e-4 relative-to-max on every output.
"""
import os

import numpy as np
import pytest
import torch

from oracle import mmae_oracle as O
from tests.helpers import make_cfgs, rel_err

pytestmark = pytest.mark.gpu


def _engine(ecfg, P, chain):
    from multimodalautoencoder_b200 import Engine
    old = os.environ.get('MMAE_CHAIN')
    os.environ['MMAE_CHAIN'] = '1' if chain else '0'
    try:
        e = Engine(ecfg)
        e.set_params({k: v.astype(np.float32) for k, v in P.items()})
        # the switch is read at the first forward: run one now
        e.forward(np.zeros((32, ecfg.num_feats), np.float32), recon=True)
    finally:
        if old is None:
            os.environ.pop('MMAE_CHAIN', None)
        else:
            os.environ['MMAE_CHAIN'] = old
    return e


CASES = [
    dict(id='S-untied', kw=dict(tie=False), B=384),
    dict(id='S-tied-relu-rmse', kw=dict(tie=True, act='relu', loss='mean_squared'), B=1000),
    dict(id='S-L3-tanh', kw=dict(layers=(128, 64, 32), tie=False, act='tanh'), B=129),
    dict(id='S-one-layer', kw=dict(layers=(64,), tie=True), B=32),
    dict(id='S-many-tiles', kw=dict(tie=False), B=128 * 148 * 2 + 77),
    dict(id='odd-widths', kw=dict(layers=(100, 44), tie=False, act='softplus'), B=257),
]


@pytest.mark.parametrize('case', CASES, ids=[c['id'] for c in CASES])
def test_chain_forward_matches_oracle_and_layerwise(case):
    ocfg, ecfg = make_cfgs(precision='tf32', **case['kw'])
    B = case['B']
    rng = np.random.default_rng(7)
    X = rng.uniform(0, 1, (B, ocfg.num_feats)).astype(np.float32)
    P = O.init_params(ocfg, rng)
    ec = _engine(ecfg, P, True)
    el = _engine(ecfg, P, False)
    n0 = ec.chain_launches
    rc = ec.forward(X, recon=True, embedding=True, loss=True)
    assert ec.chain_launches == n0 + 1, 'the whole-network kernel did not run'
    rl = el.forward(X, recon=True, embedding=True, loss=True)
    assert el.chain_launches == 0
    sc, sl = ec.scalars(), el.scalars()
    c = O.forward(ocfg, P, X.astype(np.float64), X.astype(np.float64))
    assert abs(sc['recon_loss'] - c['recon_loss']) <= 1e-3 * abs(c['recon_loss'])
    assert rel_err(rc['recon'].cpu().numpy(), c['decoded']) <= 5e-3
    assert rel_err(rc['embedding'].cpu().numpy(), c['emb']) <= 5e-3
    assert abs(sc['recon_loss'] - sl['recon_loss']) <= 1e-5 * abs(sl['recon_loss'])
    assert rel_err(rc['recon'].cpu().numpy(), rl['recon'].cpu().numpy()) <= 2e-4
    assert rel_err(rc['embedding'].cpu().numpy(), rl['embedding'].cpu().numpy()) <= 2e-4
    ec.close(); el.close()


@pytest.mark.parametrize('keep', [1.0, 0.5])
def test_chain_train_step_matches_layerwise(keep):
    """Training through the chain (saved activations + delta_L + bias-gradient partials written by the fused
    kernel, per-layer backward) gives the per-layer path's gradients; dropout masks are the same Philox stream."""
    ocfg, ecfg = make_cfgs(precision='tf32', tie=False, lam=0.001, seed=5)
    B = 640
    rng = np.random.default_rng(8)
    X = rng.uniform(0, 1, (B, 320)).astype(np.float32)
    P = O.init_params(ocfg, rng)
    ec = _engine(ecfg, P, True)
    el = _engine(ecfg, P, False)
    for e in (ec, el):
        e.set_rng_step(11)
        e.gen_noise(B)
        e.train_step(X, noise=True, keep=keep)
    assert ec.chain_launches >= 2 and el.chain_launches == 0
    assert abs(ec.scalars()['recon_loss'] - el.scalars()['recon_loss']) <= 1e-5 * el.scalars()['recon_loss']
    for name, _ in ec.variables():
        gc, gl = ec.get_gradient(name).astype(np.float64), el.get_gradient(name).astype(np.float64)
        assert np.linalg.norm(gc - gl) <= 1e-3 * max(np.linalg.norm(gl), 1e-30), name
    ec.close(); el.close()


def test_chain_falls_back_when_it_does_not_fit():
    """[1000, 100] needs more TMEM columns than an SM has: the per-layer GEMMs run, results unchanged."""
    ocfg, ecfg = make_cfgs(precision='tf32', layers=(1000, 100), tie=False, act='relu')
    rng = np.random.default_rng(9)
    X = rng.uniform(0, 1, (256, 320)).astype(np.float32)
    P = O.init_params(ocfg, rng)
    e = _engine(ecfg, P, True)
    r = e.forward(X, recon=True, loss=True)
    assert e.chain_launches == 0
    c = O.forward(ocfg, P, X.astype(np.float64), X.astype(np.float64))
    assert rel_err(r['recon'].cpu().numpy(), c['decoded']) <= 5e-3
    e.close()


@pytest.mark.parametrize('with_loss', [False, True])
def test_chain_fused_fill_in(with_loss):
    """fill_missing_data_in_file (:1167-1187 + data_funcs.py:310-381) as ONE pass: missing-block detection, then the
    whole-network kernel whose last epilogue writes decoded_X into missing blocks and the input everywhere else.
    Untouched cells are bit-exact; filled cells follow the oracle within the tf32 tolerance."""
    ocfg, ecfg = make_cfgs(precision='tf32', tie=False)
    B = 1000
    rng = np.random.default_rng(10)
    X = rng.uniform(0, 1, (B, 320))
    P = O.init_params(ocfg, rng)
    drop = rng.uniform(size=(B, 5)) < 0.2
    for m in range(5):
        X[drop[:, m], ocfg.modality_starts[m]:ocfg.modality_starts[m + 1]] = -1.0
    e = _engine(ecfg, P, True)
    n0, k0 = e.chain_launches, e.kernel_launches
    r = e.forward(X.astype(np.float32), filled=True, loss=with_loss)
    assert e.chain_launches == n0 + 1
    assert e.kernel_launches - k0 <= 3, 'fill-in should be ONE whole-network kernel (+ loss reduction, scalars): detection and select are fused'
    c = O.forward(ocfg, P, X, X)
    want = O.fill_missing(ocfg, X, c['decoded'])
    got = r['filled'].cpu().numpy()
    keep_mask = np.repeat(~drop, np.diff(ocfg.modality_starts), axis=1)
    assert np.array_equal(got[keep_mask], X.astype(np.float32)[keep_mask])
    assert rel_err(got, want) <= 5e-3
    if with_loss:
        assert abs(e.scalars()['recon_loss'] - c['recon_loss']) <= 1e-3 * c['recon_loss']
    e.close()


def test_pipelined_host_inference_matches_single_call():
    """forward_host above 131 072 rows streams chunks through H2D / compute / D2H; the result equals the device call."""
    ocfg, ecfg = make_cfgs(precision='tf32', tie=False)
    rng = np.random.default_rng(12)
    P = O.init_params(ocfg, rng)
    B = 131072 * 2 + 4321
    X = rng.uniform(0, 1, (B, 320)).astype(np.float32)
    X[rng.uniform(size=B) < 0.3, 200:220] = -1.0
    e = _engine(ecfg, P, True)
    want = e.forward(X, filled=True, recon=False)['filled'].cpu().numpy()
    wr = e.forward(X, recon=True, embedding=True)
    out = {'filled': np.empty((B, 320), np.float32)}
    got = e.forward_host(X, filled=True, out=out)['filled']
    assert got is out['filled'] and np.array_equal(got, want)
    r2 = e.forward_host(X, recon=True, embedding=True)
    assert np.array_equal(r2['recon'], wr['recon'].cpu().numpy()) and np.array_equal(r2['embedding'], wr['embedding'].cpu().numpy())
    e.close()


BWD_CASES = [
    dict(id='S-untied', kw=dict(tie=False, lam=0.001), B=640, keep=1.0),
    dict(id='S-untied-dropout', kw=dict(tie=False), B=640, keep=0.5),
    dict(id='S-tied-relu-rmse', kw=dict(tie=True, act='relu', loss='mean_squared', lam=0.001), B=1000, keep=1.0),
    dict(id='S-tied-dropout-tanh', kw=dict(tie=True, act='tanh'), B=300, keep=0.7),
    dict(id='S-L3-softplus', kw=dict(layers=(128, 64, 32), tie=False, act='softplus'), B=129, keep=1.0),
    dict(id='S-L3-tied-dropout', kw=dict(layers=(128, 64, 32), tie=True), B=515, keep=0.5),
    dict(id='odd-widths', kw=dict(layers=(100, 44), tie=False, act='softsign'), B=257, keep=1.0),
    dict(id='S-many-tiles', kw=dict(tie=False), B=128 * 148 + 77, keep=1.0),
]


@pytest.mark.parametrize('case', BWD_CASES, ids=[c['id'] for c in BWD_CASES])
def test_backward_chain_matches_oracle_and_layerwise(case):
    """Every dgrad of the step in one launch (deltas resident in TMEM from layer to layer): gradients against the fp64
    oracle within the tf32 tolerance (5e-3 relative Frobenius) and against the per-layer tcgen05 path (1e-3)."""
    ocfg, ecfg = make_cfgs(precision='tf32', seed=5, **case['kw'])
    B, keep = case['B'], case['keep']
    rng = np.random.default_rng(21)
    X = rng.uniform(0, 1, (B, ocfg.num_feats)).astype(np.float32)
    P = O.init_params(ocfg, rng)
    ec = _engine(ecfg, P, True)
    el = _engine(ecfg, P, False)
    for e in (ec, el):
        e.set_rng_step(11)
        e.gen_noise(B)
        e.train_step(X, noise=True, keep=keep)
    assert ec.backward_chain_launches == 1, 'the backward chain did not run'
    assert ec.wgrad_group_launches == 1, 'the grouped weight-gradient kernel did not run'
    assert el.backward_chain_launches == 0          # (per-layer dgrad GEMMs; its weight gradients may use the grouped launch too)
    zb, mb = ec.get_noise(B)
    noisy = O.noise_from_descriptor(ocfg, X.astype(np.float64), zb, mb)
    from tests.helpers import dropout_masks
    masks = dropout_masks(ocfg, 5, 11, B, keep) if keep < 1.0 else None
    c = O.forward(ocfg, P, noisy, X.astype(np.float64), keep, masks)
    G = O.backward_recon(ocfg, P, c)
    lam, tied = ocfg.weight_penalty, ocfg.tie_weights
    sc = ec.scalars()
    scale = sc['grad_scale'] if ocfg.loss_func == 'mean_squared' else 1.0
    for name, _ in ec.variables():
        gc, gl = ec.get_gradient(name).astype(np.float64), el.get_gradient(name).astype(np.float64)
        assert np.linalg.norm(gc - gl) <= 1e-3 * max(np.linalg.norm(gl), 1e-30), name
        l2 = (2 * lam if tied else lam) if name.startswith('weights') else (lam if 'decode_weights' in name else 0.0)
        want = G[name] - l2 * P[name]                           # the engine folds L2 into Adam
        assert np.linalg.norm(gc * scale - want) <= 5e-3 * max(np.linalg.norm(want), 1e-30), name
    ec.close(); el.close()


@pytest.mark.parametrize('prec', ['tf32', 'fp32'])
def test_noise_drawn_inside_the_step_equals_draw_then_apply(prec):
    """noise='gen' (sample_noise_kernel: descriptor + noisy batch in one pass) is bit-identical to gen_noise followed by
    a step that applies the stored descriptor: same descriptor bytes, same loss, same gradients."""
    from multimodalautoencoder_b200 import Engine
    ocfg, ecfg = make_cfgs(precision=prec, tie=False, seed=9)
    B = 777
    rng = np.random.default_rng(3)
    X = rng.uniform(0, 1, (B, 320)).astype(np.float32)
    P = O.init_params(ocfg, rng)
    runs = []
    for mode in ('two', 'one'):
        e = Engine(ecfg)
        e.set_params({k: v.astype(np.float32) for k, v in P.items()})
        e.set_rng_step(5)
        if mode == 'two':
            e.gen_noise(B)
            e.train_step(X, noise=True)
        else:
            e.train_step(X, noise='gen')
        zb, mb = e.get_noise(B)
        runs.append((zb, mb, e.scalars()['recon_loss'], {n: e.get_gradient(n) for n, _ in e.variables()}))
        e.close()
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])
    assert runs[0][2] == runs[1][2]
    for n in runs[0][3]:
        assert np.array_equal(runs[0][3][n], runs[1][3][n]), n
