"""CUDA engine (through the C ABI) against the fp64 oracle on identical inputs, weights, masks.

Tolerances (SURVEY 8c / north star): masks, indices, predictions bit-exact; fp32 engine <= 1e-5
relative on losses, <= 2e-5 relative-to-max on reconstructions, <= 1e-4 relative-to-max on gradients;
tf32 engine (10-bit-mantissa operands, fp32 accumulate) <= 1e-3 relative on the reconstruction loss and
on the loss curve, <= 5e-3 relative-to-max on reconstructions, and on gradients <= 5e-3 in relative
Frobenius norm with <= 5e-2 relative-to-max per element: with relu, an activation within tf32 rounding
of 0 flips its derivative, which moves a handful of gradient entries by O(1) of their value while the
tensor as a whole stays within 5e-3.
"""
import numpy as np
import pytest
import torch

from oracle import mmae_oracle as O
from oracle import philox_host as PH
from tests.helpers import T_STARTS, S_NAMES, make_cfgs, rel_err, dropout_masks

pytestmark = pytest.mark.gpu

TOL = {'fp32': dict(loss=1e-5, out=2e-5, grad=1e-4, grad_fro=1e-4),
       'tf32': dict(loss=1e-3, out=5e-3, grad=5e-2, grad_fro=5e-3)}


def _grad_ok(got, ref, tol):
    got = np.asarray(got, np.float64)
    fro = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
    return rel_err(got, ref) <= tol['grad'] and fro <= tol['grad_fro'], (rel_err(got, ref), fro)


def _engine(ecfg, P):
    from multimodalautoencoder_b200 import Engine
    e = Engine(ecfg)
    e.set_params({k: v.astype(np.float32) for k, v in P.items()})
    return e


def _data(ocfg, B, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(0.0, 1.0, (B, ocfg.num_feats)).astype(np.float32).astype(np.float64)
    return rng, X



def _check_adam(e, P, G_engine_plus_l2, lr, keys):
    """The fused Adam kernel against the TF formula applied to the engine's own gradient (first step:
    m = 0.1 g, v = 0.001 g^2).  Comparing against the oracle's parameters instead would amplify fp32
    gradient rounding wherever |g| ~ eps, because step 1 of Adam is lr * g / (|g| + eps * sqrt(1000))... ."""
    a = lr * np.sqrt(1 - 0.999) / (1 - 0.9)
    for k in keys:
        g = G_engine_plus_l2[k]
        m, v = 0.1 * g, 0.001 * g * g
        exp = P[k] - a * m / (np.sqrt(v) + 1e-8)
        got = e.get_variable(k).astype(np.float64)
        sens = np.abs(g) < 1e-4 * np.abs(g).max()            # update is ill-conditioned where g ~ 0
        err = np.abs(got - exp)
        assert err[~sens].max() <= 2e-3 * lr + 1e-7 * np.abs(P[k]).max(), (k, err[~sens].max())
        assert err.max() <= 1.01 * lr + 1e-7 * np.abs(P[k]).max(), k

# --------------------------------------------------------------------------- noise
def test_numpy_mode_noise_bit_exact():
    """rng_mode='numpy': same np.random.seed -> the engine masks exactly the reference's cells."""
    from multimodalautoencoder_b200.noise import numpy_descriptor, type_masks_from_names
    ocfg, ecfg = make_cfgs()
    rng, X = _data(ocfg, 257, 0)
    P = O.init_params(ocfg, rng)
    e = _engine(ecfg, P)
    np.random.seed(11)
    want = O.add_noise(ocfg, X)                                   # oracle transliteration, global RandomState
    np.random.seed(11)
    zb, mb = numpy_descriptor(257, 320, 5, True, ocfg.noise_p, type_masks_from_names(ocfg.noise_types, S_NAMES))
    e.set_noise(zb, mb)
    got = e.apply_noise(X.astype(np.float32)).cpu().numpy()
    assert np.array_equal(got, want.astype(np.float32))
    # uniform (non-intelligent) mode, two modalities per row
    ocfg2, ecfg2 = make_cfgs(intelligent=False, num_drop=2)
    np.random.seed(5)
    want2 = O.add_noise(ocfg2, X)
    np.random.seed(5)
    zb, mb = numpy_descriptor(257, 320, 5, False, num_drop=2)
    e.set_noise(zb, mb)
    assert np.array_equal(e.apply_noise(X.astype(np.float32)).cpu().numpy(), want2.astype(np.float32))
    e.close()


@pytest.mark.parametrize('intelligent', [True, False])
def test_philox_noise_matches_host_twin(intelligent):
    ocfg, ecfg = make_cfgs(intelligent=intelligent, num_drop=2, seed=0x1234ABCD5678)
    rng, X = _data(ocfg, 1000, 1)
    e = _engine(ecfg, O.init_params(ocfg, rng))
    e.set_rng_step(7)
    e.gen_noise(1000, first_row=40)
    zb, mb = e.get_noise(1000)
    thr = PH.categorical_thresholds(ocfg.noise_p)
    zb2, mb2 = PH.noise_descriptor(ecfg.seed, 7, 1000, 320, 5, 16, intelligent, thr, e.type_masks, 2, row0=40)
    assert np.array_equal(zb, zb2) and np.array_equal(mb, mb2)
    got = e.apply_noise(X.astype(np.float32)).cpu().numpy()
    assert np.array_equal(got, O.noise_from_descriptor(ocfg, X, zb2, mb2).astype(np.float32))
    if intelligent:   # distribution sanity: noise-type frequencies follow P (:202)
        e.gen_noise(1000)
        freq = np.mean(e.get_noise(1000)[1] == 0)
        assert abs(freq - 0.64) < 0.06
    e.close()


# --------------------------------------------------------------------------- recon step
CASES = [
    dict(id='S-tied-softsign-sce', kw=dict()),
    dict(id='S-untied-relu-rmse', kw=dict(tie=False, act='relu', loss='mean_squared', lam=0.001)),
    dict(id='S-tied-tanh-rmse-L3', kw=dict(layers=(128, 64, 32), act='tanh', loss='mean_squared', lam=0.01)),
    dict(id='S-vae', kw=dict(layers=(128, 64, 32), vae=True, lam=0.001)),
    dict(id='S-untied-softplus', kw=dict(tie=False, act='softplus', lam=0.01)),
    dict(id='grid-1000-100', kw=dict(layers=(1000, 100), tie=False, act='relu', lam=0.001)),
    dict(id='tiny-ragged', kw=dict(num_feats=31, starts=T_STARTS, layers=(12, 6), lam=0.01)),
    dict(id='one-layer', kw=dict(layers=(64,), loss='mean_squared')),
    dict(id='linear-ce', kw=dict(act='linear', layers=(64, 32))),
    # first layer wide enough for the two-SM GEMM: mask + noise applied to the A tile in shared memory (forward and wgrad),
    # modality boundaries on / off 32-column chunks
    dict(id='wide-first-layer-aligned', kw=dict(num_feats=512, starts=[0, 128, 192, 256, 384, 512], layers=(384, 64), tie=False, lam=0.001), B=640),
    dict(id='wide-first-layer-ragged', kw=dict(num_feats=512, starts=[0, 100, 230, 300, 410, 512], layers=(320, 48), tie=True, act='relu', loss='mean_squared'), B=517),
]


@pytest.mark.parametrize('prec', ['fp32', 'tf32'])
@pytest.mark.parametrize('case', CASES, ids=[c['id'] for c in CASES])
def test_forward_backward_parity(case, prec):
    ocfg, ecfg = make_cfgs(precision=prec, **case['kw'])
    B = case.get('B', 384 if ocfg.num_feats == 320 else 37)
    rng, X = _data(ocfg, B, 2)
    P = O.init_params(ocfg, rng)
    e = _engine(ecfg, P)
    zb, mb = PH.noise_descriptor(0, 0, B, ocfg.num_feats, len(ocfg.modality_names), int(ocfg.num_feats * .05), True,
                                 PH.categorical_thresholds(ocfg.noise_p), e.type_masks, 1)
    noisy = O.noise_from_descriptor(ocfg, X, zb, mb)
    e.set_noise(zb, mb)
    eps = None
    if ocfg.variational:
        eps = rng.standard_normal((B, ocfg.layer_sizes[-1])).astype(np.float32)
        e.set_eps(eps)
    tol = TOL[prec]
    # forward fetches (:945): decoded_X, reconstruction_loss, embedding
    r = e.forward(X.astype(np.float32), noise=True, recon=True, embedding=True, loss=True)
    sc = e.scalars()
    c = O.forward(ocfg, P, noisy, X, eps=eps)
    assert abs(sc['recon_loss'] - c['recon_loss']) <= tol['loss'] * abs(c['recon_loss'])
    assert rel_err(r['recon'].cpu().numpy(), c['decoded']) <= tol['out']
    assert rel_err(r['embedding'].cpu().numpy(), c['emb']) <= tol['out']
    if ocfg.variational:
        assert abs(sc['kl_mean'] - np.mean(c['kl'])) <= tol['loss'] * abs(np.mean(c['kl']))
    # one optimizer step (:590): gradients, then parameters after TF-Adam
    P2 = {k: v.copy() for k, v in P.items()}
    st = O.AdamState()
    c2, G = O.train_step(ocfg, P2, st, noisy, X, eps=eps)
    e.train_step(X.astype(np.float32), noise=True)
    sc = e.scalars()
    assert abs(sc['recon_loss'] - c2['recon_loss']) <= tol['loss'] * abs(c2['recon_loss'])
    scale = sc['grad_scale'] if ocfg.loss_func == 'mean_squared' else 1.0
    lam = ocfg.weight_penalty
    Geng = {}
    for k, g in G.items():
        got = e.get_gradient(k).astype(np.float64) * scale
        l2 = 0.0
        if k.startswith('weights'):
            l2 = (2 * lam if ocfg.tie_weights else lam)
        elif k.startswith('decode_weights') or k == 'variance_weights':
            l2 = lam
        got = got + l2 * P[k]                       # the engine folds L2 into Adam, the oracle into G
        Geng[k] = got
        ok, info = _grad_ok(got, g, tol)
        assert ok, (k, info)
    _check_adam(e, P, Geng, ocfg.learning_rate, G.keys())
    untouched = set(P) - set(G)
    for k in untouched:
        assert np.array_equal(e.get_variable(k), P[k].astype(np.float32)), k
    e.close()


@pytest.mark.parametrize('case', [c for c in CASES if c['id'].startswith('wide-first-layer')], ids=lambda c: c['id'])
def test_noise_fused_into_operand_load(case, monkeypatch):
    """MMAE_FUSE_NOISE=1: the two-SM GEMM applies block mask + zero noise to its A tile in shared memory (forward of the
    first encoder layer and its wgrad) instead of reading a materialised noisy X; same loss and gradients."""
    monkeypatch.setenv('MMAE_FUSE_NOISE', '1')
    ocfg, ecfg = make_cfgs(precision='tf32', **case['kw'])
    B = case['B']
    rng, X = _data(ocfg, B, 2)
    P = O.init_params(ocfg, rng)
    e = _engine(ecfg, P)
    zb, mb = PH.noise_descriptor(0, 0, B, ocfg.num_feats, len(ocfg.modality_names), int(ocfg.num_feats * .05), True,
                                 PH.categorical_thresholds(ocfg.noise_p), e.type_masks, 1)
    noisy = O.noise_from_descriptor(ocfg, X, zb, mb)
    e.set_noise(zb, mb)
    P2 = {k: v.copy() for k, v in P.items()}
    c2, G = O.train_step(ocfg, P2, O.AdamState(), noisy, X)
    e.train_step(X.astype(np.float32), noise=True)
    assert e.fused_noise_launches >= 2, 'forward + wgrad of the first layer should have applied the noise in the operand load'
    sc = e.scalars()
    tol = TOL['tf32']
    assert abs(sc['recon_loss'] - c2['recon_loss']) <= tol['loss'] * abs(c2['recon_loss'])
    scale = sc['grad_scale'] if ocfg.loss_func == 'mean_squared' else 1.0
    for k in ('weights0', 'encode_biases0'):
        ok, info = _grad_ok(e.get_gradient(k).astype(np.float64) * scale, G[k] - (0.0 if 'bias' in k else
                            (2 * ocfg.weight_penalty if ocfg.tie_weights else ocfg.weight_penalty) * P[k]), tol)
        assert ok, (k, info)
    e.close()


@pytest.mark.parametrize('prec', ['fp32', 'tf32'])
def test_training_curve_parity(prec):
    """20 steps of noisy training: loss curve and final parameters follow the oracle."""
    ocfg, ecfg = make_cfgs(precision=prec, tie=False, lam=0.001, lr=1e-3)
    B = 256
    rng, Xall = _data(ocfg, 2048, 3)
    P = O.init_params(ocfg, rng)
    P0 = {k: v.copy() for k, v in P.items()}
    e = _engine(ecfg, P)
    st = O.AdamState()
    thr = PH.categorical_thresholds(ocfg.noise_p)
    for step in range(20):
        idx = PH.batch_indices(0, step, B, 2048)
        X = Xall[idx]
        zb, mb = PH.noise_descriptor(0, step, B, 320, 5, 16, True, thr, e.type_masks, 1)
        noisy = O.noise_from_descriptor(ocfg, X, zb, mb)
        c, _ = O.train_step(ocfg, P, st, noisy, X)
        e.set_rng_step(step)
        e.gen_noise(B)
        e.train_step(X.astype(np.float32), noise=True)
        got = e.scalars()['recon_loss']
        assert abs(got - c['recon_loss']) <= (1e-3 if prec == 'tf32' else 2e-5) * c['recon_loss'], step
    for k in P:      # parameter *movement* agrees (Adam's sign-like steps amplify tf32 noise where g ~ 0)
        moved = P[k] - P0[k]
        diff = e.get_variable(k).astype(np.float64) - P[k]
        assert np.linalg.norm(diff) <= (0.1 if prec == 'tf32' else 2e-3) * np.linalg.norm(moved), k
    e.close()


# --------------------------------------------------------------------------- dropout / Philox streams
@pytest.mark.parametrize('prec', ['fp32', 'tf32'])
def test_dropout_masks_match_host_twin(prec):
    ocfg, ecfg = make_cfgs(precision=prec, layers=(128, 64, 32), tie=False, act='relu', lam=0.0, seed=99)
    B, keep = 256, 0.5
    rng, X = _data(ocfg, B, 4)
    P = O.init_params(ocfg, rng)
    e = _engine(ecfg, P)
    e.set_rng_step(3)
    masks = dropout_masks(ocfg, 99, 3, B, keep)
    P2 = {k: v.copy() for k, v in P.items()}
    c, G = O.train_step(ocfg, P2, O.AdamState(), X, X, keep=keep, drop_masks=masks)
    e.train_step(X.astype(np.float32), noise=False, keep=keep)
    tol = TOL[prec]
    assert abs(e.scalars()['recon_loss'] - c['recon_loss']) <= tol['loss'] * c['recon_loss']
    for k, g in G.items():
        ok, info = _grad_ok(e.get_gradient(k), g, tol)
        assert ok, (k, info)
    e.close()


# --------------------------------------------------------------------------- classification head
HEAD_CASES = [
    dict(id='C-sigmoid', kw=dict(layers=(200, 100), head=[50, 20], tie=True, act='relu')),
    dict(id='C-softmax', kw=dict(layers=(200, 100), head=[50, 20], num_labels=None, cls_loss='softmax', tie=False)),
    dict(id='C-vae', kw=dict(layers=(200, 100), head=[25, 10], vae=True, cls_lam=0.001)),
    dict(id='C-quirk-deep-ae', kw=dict(layers=(128, 64, 32, 16), head=[10], cls_act='tanh', cls_lam=0.001)),
]


@pytest.mark.parametrize('prec', ['fp32', 'tf32'])
@pytest.mark.parametrize('case', HEAD_CASES, ids=[c['id'] for c in HEAD_CASES])
def test_classification_step_parity(case, prec):
    ocfg, ecfg = make_cfgs(precision=prec, **case['kw'])
    B = 300
    rng, X = _data(ocfg, B, 6)
    P = O.init_params(ocfg, rng)
    e = _engine(ecfg, P)
    Y = (rng.uniform(size=(B, 3)) < 0.5).astype(np.float64) if ocfg.num_labels else rng.integers(0, 2, B).astype(np.float64)
    eps = None
    if ocfg.variational:
        eps = rng.standard_normal((B, ocfg.layer_sizes[-1])).astype(np.float32)
        e.set_eps(eps)
    tol = dict(TOL[prec])
    if prec == 'tf32' and ocfg.activation == 'relu':
        # small head tensors (B*50 activations): a few relu derivative flips at tf32 noise level weigh more
        tol['grad_fro'] = 2e-2
    r = e.forward(X.astype(np.float32), labels=Y.astype(np.float32), head=True, head_loss=True)
    sc = e.scalars()
    c = O.forward(ocfg, P, X, None, eps=eps, true_Y=Y, want_head=True)
    assert abs(sc['head_loss'] - c['cls_data_loss']) <= tol['loss'] * 5 * abs(c['cls_data_loss'])
    assert rel_err(r['logits'].cpu().numpy(), c['cls_logits']) <= tol['out']
    margin = np.abs(c['cls_logits']).min() if prec == 'tf32' else 0
    if prec == 'fp32' or margin > 1e-2:
        assert np.array_equal(r['preds'].cpu().numpy(), c['predictions'])
        assert abs(sc['head_acc'] - c['accuracy']) < 1e-6
    P2 = {k: v.copy() for k, v in P.items()}
    c2, G = O.cls_train_step(ocfg, P2, O.AdamState(), X, Y, eps=eps)
    e.cls_train_step(X.astype(np.float32), Y.astype(np.float32))
    Geng = {}
    for k, g in G.items():
        got = e.get_gradient(k).astype(np.float64)
        if k.startswith('classification_weights'):
            got = got + ocfg.cls_weight_penalty * P[k]
        Geng[k] = got
        ok, info = _grad_ok(got, g, tol)
        assert ok, (k, info)
    _check_adam(e, P, Geng, ocfg.cls_learning_rate, G.keys())
    for k in set(P) - set(G):     # decoder variables are not touched by classification_opt_step (:443)
        assert np.array_equal(e.get_variable(k), P[k].astype(np.float32)), k
    e.close()


# --------------------------------------------------------------------------- inference (:932-950, :1189-1216)
@pytest.mark.parametrize('prec', ['fp32', 'tf32'])
def test_fill_in_and_per_modality(prec):
    ocfg, ecfg = make_cfgs(precision=prec, loss='mean_squared', tie=False)
    B = 512
    rng, X = _data(ocfg, B, 8)
    P = O.init_params(ocfg, rng)
    e = _engine(ecfg, P)
    Xm = X.copy()
    drop = rng.uniform(size=(B, 5)) < 0.2
    for m in range(5):
        Xm[drop[:, m], ocfg.modality_starts[m]:ocfg.modality_starts[m + 1]] = -1.0
    r = e.forward(Xm.astype(np.float32), recon=True, filled=True, loss=True)
    c = O.forward(ocfg, P, Xm, Xm)
    want = O.fill_missing(ocfg, Xm, c['decoded'])
    got = r['filled'].cpu().numpy()
    keep_mask = np.repeat(~drop, np.diff(ocfg.modality_starts), axis=1)
    assert np.array_equal(got[keep_mask], Xm.astype(np.float32)[keep_mask])        # untouched cells bit-exact
    assert rel_err(got, want) <= TOL[prec]['out']
    assert abs(e.scalars()['recon_loss'] - c['recon_loss']) <= TOL[prec]['loss'] * c['recon_loss']
    # host-buffer variant of the same call (predict(): NumPy in, NumPy out)
    rh = e.forward_host(Xm.astype(np.float32), recon=True, filled=True, loss=True)
    assert np.array_equal(rh['filled'], got) and np.array_equal(rh['recon'], r['recon'].cpu().numpy())
    e.close()


def test_host_step_equals_device_step():
    ocfg, ecfg = make_cfgs(precision='fp32', head=[50, 20])
    rng, X = _data(ocfg, 200, 9)
    P = O.init_params(ocfg, rng)
    Y = (rng.uniform(size=(200, 3)) < 0.5).astype(np.float32)
    a, b = _engine(ecfg, P), _engine(ecfg, P)
    Xf = X.astype(np.float32)
    for s in range(3):
        a.set_rng_step(s); b.set_rng_step(s)
        a.gen_noise(200)
        a.train_step(Xf, noise=True, keep=0.5)
        b.train_step_host(Xf, gen_noise=True, keep=0.5)
        a.cls_train_step(Xf, Y)
        b.cls_train_step_host(Xf, Y)
    b.synchronize()
    for k in P:
        assert np.array_equal(a.get_variable(k), b.get_variable(k)), k
    a.close(); b.close()


def test_resident_dataset_step():
    ocfg, ecfg = make_cfgs(precision='fp32')
    rng, X = _data(ocfg, 1000, 10)
    P = O.init_params(ocfg, rng)
    a, b = _engine(ecfg, P), _engine(ecfg, P)
    a.set_dataset(0, X)
    a.set_rng_step(5); b.set_rng_step(5)
    a.train_step_resident(0, 128, idx=None, gen_noise=True)
    idx = PH.batch_indices(0, 5, 128, 1000)
    b.gen_noise(128)
    b.train_step(X[idx].astype(np.float32), noise=True)
    for k in P:
        assert np.array_equal(a.get_variable(k), b.get_variable(k)), k
    a.close(); b.close()


def test_errors_are_loud():
    from multimodalautoencoder_b200 import Engine
    ocfg, ecfg = make_cfgs()
    with pytest.raises(ValueError):
        bad = make_cfgs(layers=(64,), vae=False)[1]
        bad.variational = True
        Engine(bad)
    e = Engine(ecfg)
    with pytest.raises(ValueError):
        e.set_variable('no_such_var', np.zeros(3))
    with pytest.raises(RuntimeError):
        e.cls_train_step(np.zeros((4, 320), np.float32), np.zeros((4, 3), np.float32))   # no head
    with pytest.raises(RuntimeError):
        e.train_step(np.zeros((4, 320), np.float32), noise=True)                         # no descriptor yet
    e.close()


@pytest.mark.parametrize('prec', ['fp32', 'tf32'])
def test_large_batch_wgrad_splitk(prec):
    """B = 8192: the weight-gradient GEMMs contract over the batch and split K across CTAs in both families
    (fixed-order slice reduction); also exercises decoder wgrads with fewer than 128 output rows."""
    ocfg, ecfg = make_cfgs(precision=prec, tie=False, lam=0.0)
    B = 8192
    rng, X = _data(ocfg, B, 12)
    P = O.init_params(ocfg, rng)
    e = _engine(ecfg, P)
    c, G = O.train_step(ocfg, {k: v.copy() for k, v in P.items()}, O.AdamState(), X, X)
    e.train_step(X.astype(np.float32), noise=False)
    tol = TOL[prec]
    assert abs(e.scalars()['recon_loss'] - c['recon_loss']) <= tol['loss'] * c['recon_loss']
    for k, g in G.items():
        ok, info = _grad_ok(e.get_gradient(k), g, tol)
        assert ok, (k, info)
    e.close()


@pytest.mark.parametrize('prec', ['fp32', 'tf32'])
def test_loss_curve_over_1k_steps(prec):
    """North-star tolerance: 1e-3 relative on the loss curve over 1k steps (and on the final reconstruction RMSE),
    engine vs fp64 oracle from the same weights, batches and masks.  RMSE loss, lr 1e-3, batch 128, block-mask noise."""
    ocfg, ecfg = make_cfgs(precision=prec, tie=False, loss='mean_squared', lam=0.0, lr=1e-3)
    B, N = 128, 4096
    rng, Xall = _data(ocfg, N, 31)
    P = O.init_params(ocfg, rng)
    e = _engine(ecfg, P)
    st = O.AdamState()
    thr = PH.categorical_thresholds(ocfg.noise_p)
    worst = 0.0
    Xd = torch.empty((B, 320), device='cuda')
    for step in range(1000):
        idx = PH.batch_indices(0, step, B, N)
        X = Xall[idx]
        zb, mb = PH.noise_descriptor(0, step, B, 320, 5, 16, True, thr, e.type_masks, 1)
        noisy = O.noise_from_descriptor(ocfg, X, zb, mb)
        c, _ = O.train_step(ocfg, P, st, noisy, X)
        e.set_rng_step(step)
        e.gen_noise(B)
        Xd.copy_(torch.as_tensor(X.astype(np.float32)))          # same device buffer every step: graph replay
        e.train_step(Xd, noise=True)
        if step % 10 == 0 or step == 999:
            got = e.scalars()['recon_loss']
            worst = max(worst, abs(got - c['recon_loss']) / c['recon_loss'])
    assert e.graph_replays > 900
    assert worst <= 1e-3, worst
    # final reconstruction RMSE on held-out rows
    Xv = Xall[:512]
    r = e.forward(Xv.astype(np.float32), recon=True, loss=True)
    cv = O.forward(ocfg, P, Xv, Xv)
    assert abs(e.scalars()['recon_loss'] - cv['recon_loss']) <= 1e-3 * cv['recon_loss']
    e.close()


@pytest.mark.parametrize('prec', ['fp32', 'tf32'])
def test_resident_dataset_view_is_an_index_list(prec):
    """Cross-validation folds as views: base matrix on the device once, a fold = row list.  Sampling through the view is
    bit-identical to uploading the fold's rows as a dataset of their own (Philox indices and explicit ones)."""
    ocfg, ecfg = make_cfgs(precision=prec)
    rng, X = _data(ocfg, 1000, 10)
    P = O.init_params(ocfg, rng)
    view = np.sort(rng.choice(1000, size=700, replace=False)).astype(np.int64)
    a, b = _engine(ecfg, P), _engine(ecfg, P)
    a.set_dataset(0, X); a.set_dataset_view(0, view)
    b.set_dataset(0, X[view])
    for step, idx in ((5, None), (6, rng.integers(0, 700, 128))):
        a.set_rng_step(step); b.set_rng_step(step)
        a.train_step_resident(0, 128, idx=idx, gen_noise=True)
        b.train_step_resident(0, 128, idx=idx, gen_noise=True)
    for k in P:
        assert np.array_equal(a.get_variable(k), b.get_variable(k)), k
    with pytest.raises(ValueError):
        a.train_step_resident(0, 4, idx=np.array([0, 1, 2, 700]), gen_noise=True)       # outside the view
    a.set_dataset_view(0, None)
    a.set_rng_step(9); b.set_dataset(0, X); b.set_rng_step(9)
    a.train_step_resident(0, 64, gen_noise=True); b.train_step_resident(0, 64, gen_noise=True)
    for k in P:
        assert np.array_equal(a.get_variable(k), b.get_variable(k)), k
    a.close(); b.close()


def test_modality_rmse_batched_pass():
    """get_reconstruction_loss_per_modality as one batched device pass (M masked copies stacked, in-kernel reduction)."""
    ocfg, ecfg = make_cfgs(precision='fp32', tie=False, loss='mean_squared')
    rng, X = _data(ocfg, 777, 3)
    P = O.init_params(ocfg, rng)
    e = _engine(ecfg, P)
    k0 = e.kernel_launches
    got = e.modality_rmse(X.astype(np.float32))
    assert e.kernel_launches - k0 <= 12, 'one masked-batch pass: mask, forward, reduction'
    want = O.reconstruction_loss_per_modality(ocfg, P, X.astype(np.float32).astype(np.float64))
    assert np.allclose(got, want, rtol=1e-5)
    e.close()


def test_resident_step_gathers_the_loss_target_of_wide_models():
    """Wide models: the resident-dataset step writes no clean copy of the batch; the loss GEMM's row-layout epilogue reads the
    target rows of the dataset through the sampled index list.  Same result, bit for bit, as the step fed the gathered batch."""
    ocfg, ecfg = make_cfgs(num_feats=1024, starts=[0, 256, 512, 640, 768, 1024], layers=(384, 64), tie=False, precision='tf32')
    rng, X = _data(ocfg, 2000, 13)
    P = O.init_params(ocfg, rng)
    a, b = _engine(ecfg, P), _engine(ecfg, P)
    a.set_dataset(0, X)
    for step in (5, 6):
        a.set_rng_step(step); b.set_rng_step(step)
        a.train_step_resident(0, 512, idx=None, gen_noise=True)
        idx = PH.batch_indices(0, step, 512, 2000)
        b.train_step(X[idx].astype(np.float32), noise='gen')
        assert a.scalars()['recon_loss'] == b.scalars()['recon_loss']
    for k in P:
        assert np.array_equal(a.get_variable(k), b.get_variable(k)), k
    a.close(); b.close()
