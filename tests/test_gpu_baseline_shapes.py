"""Parity on the BASELINE.json shapes themselves (not only on their small cousins), through the C ABI:

  configs[3]  wide MMAE  F = 4096 in 16 blocks of 256, encoder [2048, 1024, 256], untied, softsign, sigmoid-CE: one train
              step at B = 2048 -- loss, every gradient, TF-Adam update -- against the fp64 oracle, fp32 and tf32 engines
  configs[1]  MMAE + mood head  F = 320, [200, 100] + [50, 20] -> 3, relu, B = 4096: reconstruction step and classification step
  configs[4]  fill-in inference  1 048 576 + 77 rows through the whole-network kernel: untouched cells bit-exact, filled cells
              against the oracle on a row sample

Tolerances are test_gpu_parity.py's (fp32: loss 1e-5, gradients 1e-4; tf32: loss 1e-3, gradients 5e-3 relative Frobenius)."""
import numpy as np
import pytest

from oracle import mmae_oracle as O
from oracle import philox_host as PH
from tests.helpers import S_NAMES, S_STARTS, make_cfgs, rel_err
from tests.test_gpu_parity import TOL, _check_adam, _engine, _grad_ok

pytestmark = pytest.mark.gpu

W_NAMES = ['call', 'sms', 'screen', 'location'] + ['phys%02d' % i for i in range(12)]
W_STARTS = [256 * i for i in range(17)]


def _train_step_vs_oracle(ocfg, ecfg, B, prec, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(0.0, 1.0, (B, ocfg.num_feats)).astype(np.float32).astype(np.float64)
    P = O.init_params(ocfg, rng)
    e = _engine(ecfg, P)
    zb, mb = PH.noise_descriptor(0, 0, B, ocfg.num_feats, len(ocfg.modality_names), int(ocfg.num_feats * .05), True,
                                 PH.categorical_thresholds(ocfg.noise_p), e.type_masks, 1)
    noisy = O.noise_from_descriptor(ocfg, X, zb, mb)
    e.set_noise(zb, mb)
    assert np.array_equal(e.apply_noise(X.astype(np.float32)).cpu().numpy(), noisy.astype(np.float32))      # masks: bit-exact
    tol = TOL[prec]
    P2 = {k: v.copy() for k, v in P.items()}
    c, G = O.train_step(ocfg, P2, O.AdamState(), noisy, X)
    e.train_step(X.astype(np.float32), noise=True)
    sc = e.scalars()
    assert abs(sc['recon_loss'] - c['recon_loss']) <= tol['loss'] * abs(c['recon_loss'])
    Geng = {}
    for k, g in G.items():
        got = e.get_gradient(k).astype(np.float64)
        l2 = ocfg.weight_penalty if ('weights' in k) else 0.0
        got = got + l2 * P[k]
        Geng[k] = got
        ok, info = _grad_ok(got, g, tol)
        assert ok, (k, info)
    _check_adam(e, P, Geng, ocfg.learning_rate, G.keys())
    return e, P, X, rng


@pytest.mark.parametrize('prec', ['fp32', 'tf32'])
def test_wide_config_train_step(prec):
    ocfg, ecfg = make_cfgs(num_feats=4096, starts=W_STARTS, names=W_NAMES, layers=(2048, 1024, 256), tie=False,
                           act='softsign', loss='sigmoid_cross_entropy', lam=0.0, lr=1e-3, precision=prec)
    e, P, X, rng = _train_step_vs_oracle(ocfg, ecfg, 2048, prec, 31)
    e.close()


@pytest.mark.parametrize('prec', ['fp32', 'tf32'])
def test_classification_config_both_steps(prec):
    ocfg, ecfg = make_cfgs(layers=(200, 100), tie=False, act='relu', lam=0.001, lr=1e-3, head=[50, 20], num_labels=3,
                           cls_lam=0.001, cls_lr=1e-3, precision=prec)
    B = 4096
    e, P, X, rng = _train_step_vs_oracle(ocfg, ecfg, B, prec, 32)
    # classification step (:647) on the parameters the engine holds now
    Pn = {k: e.get_variable(k).astype(np.float64) for k in P}
    Y = (rng.uniform(size=(B, 3)) < 0.5).astype(np.float64)
    zb, mb = e.get_noise(B)
    noisy = O.noise_from_descriptor(ocfg, X, zb, mb)
    P2 = {k: v.copy() for k, v in Pn.items()}
    c, G = O.cls_train_step(ocfg, P2, O.AdamState(), noisy, Y)
    e.cls_train_step(X.astype(np.float32), Y.astype(np.float32), noise=True)
    sc = e.scalars()
    tol = TOL[prec]
    assert abs(sc['head_loss'] - c['cls_data_loss']) <= max(tol['loss'], 2e-5) * abs(c['cls_data_loss'])
    assert abs(sc['head_acc'] - c['accuracy']) <= (0.0 if prec == 'fp32' else 2e-3) + 1e-9
    for k, g in G.items():
        got = e.get_gradient(k).astype(np.float64)
        if k.startswith('classification_weights'):
            got = got + ocfg.cls_weight_penalty * Pn[k]
        ok, info = _grad_ok(got, g, tol)
        assert ok, (k, info)
    e.close()


def test_fill_in_config_a_million_rows():
    ocfg, ecfg = make_cfgs(tie=False, precision='tf32')
    B = (1 << 20) + 77
    rng = np.random.default_rng(33)
    P = O.init_params(ocfg, rng)
    X = rng.uniform(0, 1, (B, 320)).astype(np.float32)
    drop = rng.uniform(size=(B, 5)) < 0.2
    for m in range(5):
        X[drop[:, m], S_STARTS[m]:S_STARTS[m + 1]] = -1.0
    e = _engine(ecfg, P)
    n0 = e.chain_launches
    got = e.forward(X, filled=True)['filled'].cpu().numpy()
    assert e.chain_launches == n0 + 1, 'fill-in should be one whole-network launch'
    keep = np.repeat(~drop, np.diff(S_STARTS), axis=1)
    assert np.array_equal(got[keep], X[keep])                         # untouched cells: bit-exact
    rows = rng.choice(B, size=20000, replace=False)
    c = O.forward(ocfg, P, X[rows].astype(np.float64), X[rows].astype(np.float64))
    want = O.fill_missing(ocfg, X[rows].astype(np.float64), c['decoded'])
    assert rel_err(got[rows], want) <= 5e-3
    assert np.array_equal(got[-1], np.where(keep[-1], X[-1], got[-1]))  # the ragged last tile took the same path
    e.close()
