"""Pins the fp64 oracle to the reference's OWN code.

`oracle/_ref/` holds the reference's multimodal_autoencoder.py / data_funcs.py converted mechanically to Python 3
(oracle/build_ref.py); it runs on a minimal TF-1 API implemented on torch autograd (oracle/tf1_shim).  What runs
here is therefore the reference's `build_graph` / `encode` / `decode` / `classify` / `add_noise_to_batch` /
`train` / `predict` / `get_reconstruction_loss_per_modality` and its DataLoader's batch sampling and
missing-block rule -- the graph wiring, the RNG call order and the host logic are the reference's; only the
per-op kernels (matmul, softsign, sigmoid-CE, ApplyAdam ...) are restated in the shim.

Tolerances: byte/index tensors bit-exact; fp64 quantities 1e-9 relative (both sides compute in float64).
"""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from oracle import mmae_oracle as O
from oracle.ref_loader import load_reference

REF = load_reference(torch.float64)
pytestmark = pytest.mark.skipif(REF is None, reason='oracle/_ref not built and /root/reference absent')

NAMES = ['phys', 'call', 'sms', 'screen', 'location']
T_STARTS = [0, 11, 15, 19, 24, 31]
S_STARTS = [0, 200, 220, 240, 270, 320]


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def make_pair(F=31, starts=T_STARTS, layers=(12, 6), tie=True, vae=False, act='softsign', loss='sigmoid_cross_entropy',
              lam=0.0, lr=1e-3, keep=1.0, head=None, num_labels=3, cls_loss='sigmoid_cross_entropy', cls_lam=0.0,
              cls_lr=1e-3, intelligent=True, num_drop=1, n_train=64, n_val=40, batch=16, seed=0):
    """(oracle config, oracle params, reference model) with identical injected weights and data."""
    rng = np.random.default_rng(seed)
    ocfg = O.OracleConfig(num_feats=F, layer_sizes=list(layers), modality_starts=list(starts), modality_names=NAMES,
                          tie_weights=tie, variational=vae, activation=act, loss_func=loss, weight_penalty=lam,
                          learning_rate=lr, cls_layer_sizes=head, num_labels=num_labels, cls_activation=act,
                          cls_loss=cls_loss, cls_weight_penalty=cls_lam, cls_learning_rate=cls_lr,
                          intelligent_noise=intelligent, num_modalities_to_drop=num_drop)
    P = O.init_params(ocfg, rng)
    train_X, val_X = rng.uniform(0, 1, (n_train, F)), rng.uniform(0, 1, (n_val, F))
    if head is not None:
        if num_labels is None:
            train_Y, val_Y = rng.integers(0, 2, n_train).astype(np.float64), rng.integers(0, 2, n_val).astype(np.float64)
        else:
            train_Y = (rng.uniform(size=(n_train, num_labels)) < 0.5).astype(np.float64)
            val_Y = (rng.uniform(size=(n_val, num_labels)) < 0.5).astype(np.float64)
    else:
        train_Y = val_Y = None
    dl = REF.make_loader(train_X, val_X, starts, NAMES, train_Y, val_Y, num_labels)
    with quiet():
        m = REF.mmae.MultimodalAutoencoder(
            data_loader=dl, classification_data_loader=dl if head is not None else None, layer_sizes=list(layers),
            variational=vae, tie_weights=tie, batch_size=batch, learning_rate=lr, dropout_prob=keep, weight_penalty=lam,
            activation_func=act, loss_func=loss, classification_layer_sizes=head, weight_initialization='normal',
            intelligent_noise=intelligent, num_modalities_to_drop=num_drop, verbose=False)
        if head is not None:
            m.set_classification_params(learning_rate=cls_lr, weight_penalty=cls_lam, loss_func=cls_loss, suppress_warning=True)
    assert m.tie_weights == ocfg.tie_weights and m.loss_func == ocfg.loss_func          # ctor overrides (:175-179)
    assert {k: tuple(v.value.shape) for k, v in REF.variables(m).items() if v.is_float} == O.param_shapes(ocfg)
    REF.set_variables(m, P)
    return ocfg, P, m, dl, rng


def mask_to_uniform(mask, keep):
    """A U[0,1) draw that makes tf.nn.dropout's floor(keep + U) equal to the given 0/1 mask."""
    return np.where(np.asarray(mask) > 0, 1.0 - 0.5 * keep, 0.5 * (1.0 - keep))


class inject:
    """Context: inject epsilon and dropout masks (by evaluation order) into the shim's random ops."""

    def __init__(self, eps=None, masks=None, keep=1.0):
        self.eps, self.masks, self.keep = eps, masks, keep

    def __enter__(self):
        h = REF.tf.hooks
        h.random_normal = (lambda shp, name: self.eps) if self.eps is not None else None
        h.dropout_uniform = (lambda shp, i: mask_to_uniform(self.masks[i], self.keep)) if self.masks is not None else None

    def __exit__(self, *a):
        REF.tf.hooks.random_normal = None
        REF.tf.hooks.dropout_uniform = None


def drop_masks(ocfg, rng, B, keep, head=False):
    """Random 0/1 masks keyed like the oracle ('enc i' / 'dec j' / 'cls i') + the same masks in graph order."""
    if keep >= 1.0:
        return None, None
    d = [ocfg.num_feats] + list(ocfg.layer_sizes)
    L = ocfg.L
    md, order = {}, []
    for i in range(L - 1):
        md['enc%d' % i] = (rng.uniform(size=(B, d[i + 1])) < keep).astype(np.float64); order.append('enc%d' % i)
    if not head:
        for j in range(L - 1):
            md['dec%d' % j] = (rng.uniform(size=(B, d[L - 1 - j])) < keep).astype(np.float64); order.append('dec%d' % j)
    else:
        for i, (_, dout) in enumerate(ocfg.head_dims()):
            if i < L - 1:                                       # the :533 bound
                md['cls%d' % i] = (rng.uniform(size=(B, dout)) < keep).astype(np.float64); order.append('cls%d' % i)
        for j in range(L - 1):                                  # the oracle's forward always walks the decoder; the
            md['dec%d' % j] = np.ones((B, d[L - 1 - j]))        # classification fetches never reach it
    return md, [md[k] for k in order]


RECON_CASES = {
    'tiny_tied_sce': dict(tie=True, lam=0.01),
    'tiny_untied_sce': dict(tie=False, lam=0.001),
    'tiny_vae': dict(vae=True, lam=0.001, act='relu'),
    'tiny_rmse_tanh': dict(tie=False, loss='mean_squared', act='tanh', lam=0.001),
    'tiny_rmse_tied_softplus': dict(tie=True, loss='mean_squared', act='softplus'),
    'tiny_plain_ce_linear': dict(tie=False, loss='cross_entropy', act='linear'),
    'tiny_dropout_tied': dict(tie=True, keep=0.5, layers=(12, 8, 6)),
    'tiny_dropout_vae_3layer': dict(vae=True, keep=0.7, layers=(12, 8, 6), lam=0.01),
    'tiny_one_layer': dict(tie=True, layers=(7,)),
    'small_untied_rmse': dict(F=320, starts=S_STARTS, layers=(128, 64), tie=False, loss='mean_squared', act='tanh', lam=0.001, batch=48),
    'small_tied_sce_dropout': dict(F=320, starts=S_STARTS, layers=(128, 64, 32), tie=True, keep=0.5, batch=48),
}


@pytest.mark.parametrize('name', list(RECON_CASES))
def test_recon_step_matches_reference_graph(name):
    """Forward fetches, the gradient of total_loss wrt every variable, and three ApplyAdam steps."""
    kw = dict(RECON_CASES[name])
    ocfg, P, m, dl, rng = make_pair(**kw)
    keep = kw.get('keep', 1.0)
    B = m.batch_size
    X = dl.train_X[:B]
    if ocfg.loss_func == 'cross_entropy':
        # plain CE takes log(decoded_X) of the raw output (:386): keep it positive so the case is finite
        for k in P:
            P[k] = np.abs(P[k])
        REF.set_variables(m, P)
    np.random.seed(5)
    noisy = m.add_noise_to_batch(X) if ocfg.loss_func != 'cross_entropy' else X.copy()      # (-1 masks would make the log NaN)
    eps = rng.standard_normal((B, ocfg.layer_sizes[-1])) if ocfg.variational else None

    # ---- keep = 1 fetches (predict :932-950, get_embedding :1062-1080)
    with inject(eps=eps):
        rec, loss, emb = m.session.run([m.decoded_X, m.reconstruction_loss, m.embedding],
                                       {m.noisy_X: noisy, m.true_X: X, m.tf_dropout_prob: 1.0})
    c = O.forward(ocfg, P, noisy, X, eps=eps)
    assert rel(loss, c['recon_loss']) < 1e-12
    assert rel(rec, c['decoded']) < 1e-12 and rel(emb, c['emb']) < 1e-12

    # ---- three optimizer steps (:411, :590), fresh dropout masks each
    st = O.AdamState()
    for s in range(3):
        md, ml = drop_masks(ocfg, rng, B, keep)
        with inject(eps=eps, masks=ml, keep=keep):
            tl, _ = m.session.run([m.total_loss, m.opt_step], {m.noisy_X: noisy, m.true_X: X, m.tf_dropout_prob: keep})
        c, G = O.train_step(ocfg, P, st, noisy, X, keep=keep, drop_masks=md, eps=eps)
        assert rel(tl, c['total_loss']) < 1e-11, s
        got = m.opt_step.opt.last_grads
        assert set(got) == set(G), (sorted(got), sorted(G))           # exactly the variables opt_step touches
        for k in G:
            assert rel(got[k], G[k]) < 1e-9, (s, k)
        now = REF.get_variables(m)
        for k in P:
            assert rel(now[k], P[k]) < 1e-10, (s, k)


CLS_CASES = {
    'sigmoid_head': dict(head=[5, 4], cls_lam=0.001),
    'sigmoid_head_vae': dict(head=[5, 4], vae=True, act='relu', cls_lam=0.001),
    'softmax_head': dict(head=[5], num_labels=None, cls_loss='softmax'),
    'quirk_deep_ae': dict(head=[5], layers=(12, 8, 6), act='tanh'),          # L-1 = 2 >= head depth: logits get act (+dropout)
    'quirk_deep_ae_dropout': dict(head=[5], layers=(12, 8, 6), keep=0.6),
    'head_dropout': dict(head=[5, 4], keep=0.5, tie=False),
}


@pytest.mark.parametrize('name', list(CLS_CASES))
def test_classification_step_matches_reference_graph(name):
    kw = dict(CLS_CASES[name])
    ocfg, P, m, dl, rng = make_pair(**kw)
    keep = kw.get('keep', 1.0)
    B = 16
    X, Y = dl.train_X[:B], dl.train_Y[:B]
    np.random.seed(6)
    noisy = m.add_noise_to_batch(X)
    eps = rng.standard_normal((B, ocfg.layer_sizes[-1])) if ocfg.variational else None
    with inject(eps=eps):
        loss, acc, preds, probs = m.session.run([m.classification_loss, m.accuracy, m.predictions, m.class_probabilities],
                                                {m.noisy_X: noisy, m.true_Y: Y, m.tf_dropout_prob: 1.0})
    c = O.forward(ocfg, P, noisy, None, eps=eps, true_Y=Y, want_head=True)
    assert rel(loss, c['cls_loss']) < 1e-12
    assert np.array_equal(preds, c['predictions']) and preds.dtype == np.int32
    assert abs(float(acc) - c['accuracy']) < 1e-12
    assert rel(probs, c['class_prob']) < 1e-12
    st = O.AdamState()
    dec_before = {k: v.copy() for k, v in REF.get_variables(m).items() if k.startswith('decode_')}
    for s in range(3):
        md, ml = drop_masks(ocfg, rng, B, keep, head=True)
        with inject(eps=eps, masks=ml, keep=keep):
            m.session.run([m.classification_opt_step], {m.noisy_X: noisy, m.true_Y: Y, m.tf_dropout_prob: keep})
        c, G = O.cls_train_step(ocfg, P, st, noisy, Y, keep=keep, drop_masks=md, eps=eps)
        got = m.classification_opt_step.opt.last_grads
        assert set(got) == set(G), (sorted(got), sorted(G))           # encoder (+variance) + head, never the decoder (:443)
        for k in G:
            assert rel(got[k], G[k]) < 1e-9, (s, k)
        now = REF.get_variables(m)
        for k in P:
            assert rel(now[k], P[k]) < 1e-10, (s, k)
    for k, v in dec_before.items():
        assert np.array_equal(REF.get_variables(m)[k], v)


@pytest.mark.parametrize('intelligent,num_drop', [(True, 1), (False, 1), (False, 3)])
@pytest.mark.parametrize('F,starts', [(31, T_STARTS), (320, S_STARTS)])
def test_noise_bytes_match_reference(intelligent, num_drop, F, starts):
    """add_noise_to_batch (:668-702): same np.random stream -> byte-identical noisy batch."""
    ocfg, P, m, dl, rng = make_pair(F=F, starts=starts, layers=(8, 4), intelligent=intelligent, num_drop=num_drop)
    X = dl.train_X[:40]
    np.random.seed(11)
    a = m.add_noise_to_batch(X)
    s1 = np.random.randint(1 << 30)
    np.random.seed(11)
    b = O.add_noise(ocfg, X)
    s2 = np.random.randint(1 << 30)
    assert a.dtype == b.dtype and np.array_equal(a, b)
    assert s1 == s2                                                # the stream was consumed identically
    assert np.array_equal(X, dl.train_X[:40])                      # deep copy (:678): the input is untouched
    if intelligent:                                                # the missing_modes override (:691-692)
        np.random.seed(12)
        a = m.add_noise_to_batch(X, missing_modes=['sms', 'location'])
        np.random.seed(12)
        b = O.add_noise(ocfg, X, missing_modes=['sms', 'location'])
        assert np.array_equal(a, b)


def test_batch_sampling_matches_reference():
    """data_funcs.py:161-195."""
    ocfg, P, m, dl, rng = make_pair(head=[5], n_train=97, n_val=33)
    for fn, n, has_y in ((dl.get_unsupervised_train_batch, 97, False), (dl.get_supervised_train_batch, 97, True),
                         (dl.get_unsupervised_val_batch, 33, False), (dl.get_supervised_val_batch, 33, True)):
        np.random.seed(3)
        got = fn(20)
        np.random.seed(3)
        idx = O.sample_batch_indices(n, 20)
        src = dl.train_X if n == 97 else dl.val_X
        if has_y:
            srcy = dl.train_Y if n == 97 else dl.val_Y
            assert np.array_equal(got[0], src[idx]) and np.array_equal(got[1], srcy[idx])
        else:
            assert np.array_equal(got, src[idx])


def oracle_train_replay(ocfg, P, dl, num_steps, batch, keep, record_every):
    """The reference's train() protocol (:565-590, :704-737) driven by the oracle: RNG call order train-idx,
    train-noise, [val-idx, val-noise at record steps], loss-only runs at record steps, then one optimizer step."""
    st = O.AdamState()
    tr, va = [], []
    for step in range(num_steps):
        X = dl.train_X[O.sample_batch_indices(len(dl.train_X), batch)]
        noisy = O.add_noise(ocfg, X)
        if step % record_every == 0:
            vX = dl.val_X[O.sample_batch_indices(len(dl.val_X), 200)]
            nv = O.add_noise(ocfg, vX)
            tl = O.forward(ocfg, P, noisy, X)['recon_loss']
            vl = O.forward(ocfg, P, nv, vX)['recon_loss']
            if 'entropy' in ocfg.loss_func:
                tl, vl = tl / len(X), vl / len(vX)
            tr.append(tl); va.append(vl)
        O.train_step(ocfg, P, st, noisy, X, keep=keep)
    return tr, va


@pytest.mark.parametrize('kw', [dict(tie=True, lam=0.001), dict(tie=False, loss='mean_squared', act='relu')])
def test_train_loop_matches_reference(kw):
    """The reference's own train() for 40 steps vs the oracle replay of its protocol: loss curves and final weights."""
    ocfg, P, m, dl, rng = make_pair(batch=10, **kw)
    np.random.seed(21)
    with quiet():
        m.train(40, record_every_nth=10, save_every_nth=10 ** 6)
    np.random.seed(21)
    tr, va = oracle_train_replay(ocfg, P, dl, 40, 10, 1.0, 10)
    assert len(m.train_loss) == len(tr) == 4
    assert rel(m.train_loss, tr) < 1e-9 and rel(m.val_loss, va) < 1e-9
    now = REF.get_variables(m)
    for k in P:
        assert rel(now[k], P[k]) < 1e-8, k


def test_train_classification_loop_matches_reference():
    """train_classification (:606-647) + evaluate_classification_performance (:739-764): un-noised val feed, keep = 1."""
    ocfg, P, m, dl, rng = make_pair(head=[5, 4], batch=10, cls_lam=0.001)
    np.random.seed(22)
    with quiet():
        m.train_classification(30, record_every_nth=10, save_every_nth=10 ** 6)
    np.random.seed(22)
    st = O.AdamState()
    tl, ta, vl, va = [], [], [], []
    for step in range(30):
        idx = O.sample_batch_indices(len(dl.train_X), m.classification_batch_size)
        X, Y = dl.train_X[idx], dl.train_Y[idx]
        noisy = O.add_noise(ocfg, X)
        if step % 10 == 0:
            vi = O.sample_batch_indices(len(dl.val_X), 200)
            c1 = O.forward(ocfg, P, noisy, None, true_Y=Y, want_head=True)
            c2 = O.forward(ocfg, P, dl.val_X[vi], None, true_Y=dl.val_Y[vi], want_head=True)
            tl.append(c1['cls_loss']); ta.append(c1['accuracy']); vl.append(c2['cls_loss']); va.append(c2['accuracy'])
        O.cls_train_step(ocfg, P, st, noisy, Y)
    assert rel(m.classification_train_loss, tl) < 1e-9 and rel(m.classification_val_loss, vl) < 1e-9
    assert np.allclose(m.train_acc, ta, atol=1e-12) and np.allclose(m.val_acc, va, atol=1e-12)
    now = REF.get_variables(m)
    for k in P:
        assert rel(now[k], P[k]) < 1e-8, k


def test_inference_surface_matches_reference():
    """predict / get_performance_on_data[_with_noise] / get_embedding / get_classification_predictions /
    get_reconstruction_loss_per_modality (:932-950, :1005-1045, :1062-1080, :1189-1216)."""
    ocfg, P, m, dl, rng = make_pair(head=[5, 4], tie=False, lam=0.001)
    X = dl.val_X
    rec, loss = m.predict(X)
    c = O.forward(ocfg, P, X, X)
    assert rel(rec, c['decoded']) < 1e-12 and rel(loss, c['recon_loss'] / len(X)) < 1e-12
    assert rel(m.get_performance_on_data(X), c['recon_loss'] / len(X)) < 1e-12
    np.random.seed(4)
    ln = m.get_performance_on_data_with_noise(X)
    np.random.seed(4)
    assert rel(ln, O.forward(ocfg, P, O.add_noise(ocfg, X), X)['recon_loss'] / len(X)) < 1e-12
    assert rel(m.get_embedding(X), c['emb']) < 1e-12
    assert np.array_equal(m.get_classification_predictions(X), O.forward(ocfg, P, X, None, want_head=True)['predictions'])
    with quiet():
        rms = m.get_reconstruction_loss_per_modality(X)
    assert rel(rms, O.reconstruction_loss_per_modality(ocfg, P, X)) < 1e-12


def test_missing_block_rule_matches_reference():
    """DataLoader.find_missing_modalities_indices (data_funcs.py:366-381): sum == -width, not all-equal."""
    ocfg, P, m, dl, rng = make_pair()
    X = dl.train_X[:12].copy()
    X[0, 11:15] = -1.0                          # a masked block
    X[1, 0:11] = -1.0; X[1, 24:31] = -1.0
    X[2, 15:19] = [-2.0, 0.0, -1.0, -1.0]       # sums to -4 without being all -1: still "missing" under the rule
    X[3, 19:24] = -0.999999
    miss = O.missing_blocks(ocfg, X)
    for r in range(len(X)):
        want = sorted(set(int(i) for i in dl.find_missing_modalities_indices(X[r])))
        got = sorted(c for mm in range(5) if miss[r, mm] for c in range(T_STARTS[mm], T_STARTS[mm + 1]))
        assert want == got, r
    assert miss[2, 2] and not miss[3, 3]


def test_learning_rate_decay_is_inert():
    """exponential_decay(lr, global_step, ...) with a global_step nobody increments (:356-361, :411)."""
    ocfg, P, m, dl, rng = make_pair()
    X = dl.train_X[:8]
    for _ in range(3):
        m.session.run([m.opt_step], {m.noisy_X: X, m.true_X: X, m.tf_dropout_prob: 1.0})
    assert int(m.session.run(m.global_step)) == 0
    assert float(m.session.run(m.tf_learning_rate)) == m.learning_rate


def test_golden_vectors_were_checked_against_reference():
    """Every committed golden file carries the reference-run values written by make_golden.py."""
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    files = [f for f in os.listdir(gdir) if f.endswith('.npz')]
    assert files
    for f in files:
        z = np.load(os.path.join(gdir, f))
        assert 'ref/losses' in z.files and int(z['ref/verified']) == 1, f
        assert rel(z['ref/losses'], z['losses']) < 1e-9
