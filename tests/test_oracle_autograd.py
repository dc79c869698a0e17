"""The oracle's hand-derived backward vs an independent torch.autograd (fp64) derivation.

The reference cannot run here (Python 2 + TensorFlow 1.x), so two independent
derivations of the same graph (multimodal_autoencoder.py:344-452) must agree.
The torch forward below is written directly from the reference source and does
not call into oracle/ for any arithmetic.
"""
import itertools

import numpy as np
import pytest
import torch

from oracle import mmae_oracle as O

STARTS = [0, 11, 15, 19, 24, 31]
NAMES = ['phys', 'call', 'sms', 'screen', 'location']


def t_act(name, z):
    if name == 'relu':
        return torch.relu(z)
    if name == 'tanh':
        return torch.tanh(z)
    if name == 'softsign':
        return torch.nn.functional.softsign(z)
    if name == 'softplus':
        return torch.nn.functional.softplus(z)
    return z


def torch_graph(cfg, P, noisy, X, keep, masks, eps, Y=None):
    L = cfg.L
    h = noisy
    lv = None
    for i in range(L):
        if cfg.variational and i == L - 1:
            lv = h @ P['variance_weights'] + P['variance_bias']
        h = h @ P['weights%d' % i] + P['encode_biases%d' % i]
        if i < L - 1:
            h = t_act(cfg.activation, h)
            if masks is not None:
                h = h * masks['enc%d' % i] / keep
    emb = h
    if cfg.variational:
        emb = emb + eps * torch.exp(lv)
    dW = [P['weights%d' % i].t() if cfg.tie_weights else P['decode_weights%d' % i] for i in range(L)]
    db = [P['decode_biases%d' % i] for i in range(L)]
    dW.reverse()
    db.reverse()
    x = emb
    for j in range(L):
        x = x @ dW[j] + db[j]
        if j < L - 1:
            x = t_act(cfg.activation, x)
            if masks is not None:
                x = x * masks['dec%d' % j] / keep
    if cfg.loss_func == 'mean_squared':
        rec = torch.sqrt(torch.mean((x - X) ** 2))
    elif cfg.loss_func == 'cross_entropy':
        rec = -torch.sum(X * torch.log(x))
    else:
        rec = torch.nn.functional.binary_cross_entropy_with_logits(x, X, reduction='sum')
    reg = sum(0.5 * (P['weights%d' % i] ** 2).sum() for i in range(L))
    reg = reg + sum(0.5 * (w ** 2).sum() for w in dW)
    if cfg.variational:
        reg = reg + 0.5 * (P['variance_weights'] ** 2).sum()
        kl = -0.5 * torch.sum(1 + 2 * lv - emb ** 2 - torch.exp(2 * lv), dim=1)
        total = torch.mean(rec + kl) + cfg.weight_penalty * reg
    else:
        total = rec + cfg.weight_penalty * reg
    cls = None
    if Y is not None:
        h = emb
        nh = len(cfg.head_dims())
        for i in range(nh):
            h = h @ P['classification_weights%d' % i] + P['classification_biases%d' % i]
            if i < L - 1:
                h = t_act(cfg.cls_activation, h)
                if masks is not None:
                    h = h * masks['cls%d' % i] / keep
        if cfg.cls_loss == 'sigmoid_cross_entropy':
            cls = torch.nn.functional.binary_cross_entropy_with_logits(h, Y, reduction='mean')
        else:
            cls = torch.nn.functional.cross_entropy(h, Y.long(), reduction='mean')
        cls = cls + cfg.cls_weight_penalty * sum(
            0.5 * (P['classification_weights%d' % i] ** 2).sum() for i in range(nh))
    return rec, total, cls


def make_case(seed, layers, tie, vae, act, loss, lam, keep, head=None, num_labels=3, cls_loss='sigmoid_cross_entropy'):
    rng = np.random.default_rng(seed)
    cfg = O.OracleConfig(num_feats=31, layer_sizes=layers, modality_starts=STARTS, modality_names=NAMES,
                         tie_weights=tie, variational=vae, activation=act, loss_func=loss,
                         weight_penalty=lam, cls_layer_sizes=head, num_labels=num_labels,
                         cls_loss=cls_loss, cls_weight_penalty=0.01 if head else 0.0,
                         cls_activation='tanh' if head else act)
    P = O.init_params(cfg, rng, 'normal')
    B = 7
    X = rng.uniform(0.05, 0.95, (B, 31))
    noisy = O.add_noise(cfg, X, np.random.RandomState(seed))
    if loss == 'cross_entropy':
        # raw outputs must stay positive for log(): bias the last decoder bias up
        P['decode_biases0'] = P['decode_biases0'] + 3.0
    masks = None
    if keep < 1.0:
        masks = {}
        d = [31] + layers
        for i in range(len(layers) - 1):
            masks['enc%d' % i] = (rng.uniform(size=(B, d[i + 1])) < keep).astype(float)
        for j in range(len(layers) - 1):
            masks['dec%d' % j] = (rng.uniform(size=(B, d[len(layers) - 1 - j])) < keep).astype(float)
        for i, (_, dout) in enumerate(cfg.head_dims()):
            masks['cls%d' % i] = (rng.uniform(size=(B, dout)) < keep).astype(float)
    eps = rng.standard_normal((B, layers[-1])) if cfg.variational else None
    if head is not None:
        Y = (rng.uniform(size=(B, num_labels)) < 0.5).astype(float) if num_labels else rng.integers(0, 2, B).astype(float)
    else:
        Y = None
    return cfg, P, X, noisy, masks, eps, Y


CASES = []
for tie, act, loss in itertools.product([True, False], ['softsign', 'relu', 'tanh', 'softplus', 'linear'],
                                        ['mean_squared', 'sigmoid_cross_entropy']):
    CASES.append(dict(layers=[12, 6], tie=tie, vae=False, act=act, loss=loss, lam=0.01, keep=1.0))
CASES += [
    dict(layers=[12, 6, 4], tie=True, vae=False, act='softsign', loss='sigmoid_cross_entropy', lam=0.001, keep=0.5),
    dict(layers=[12, 6, 4], tie=False, vae=True, act='softsign', loss='sigmoid_cross_entropy', lam=0.01, keep=1.0),
    dict(layers=[12, 6], tie=False, vae=True, act='relu', loss='sigmoid_cross_entropy', lam=0.0, keep=0.5),
    dict(layers=[9], tie=True, vae=False, act='tanh', loss='mean_squared', lam=0.1, keep=1.0),
    dict(layers=[12, 6], tie=False, vae=False, act='tanh', loss='cross_entropy', lam=0.0, keep=1.0),
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: '-'.join(str(v) for v in c.values()))
def test_recon_backward_matches_autograd(case):
    cfg, P, X, noisy, masks, eps, _ = make_case(3, **case)
    keep = case['keep']
    c = O.forward(cfg, P, noisy, X, keep, masks, eps)
    G = O.backward_recon(cfg, P, c)
    tP = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in P.items()}
    tm = None if masks is None else {k: torch.tensor(v) for k, v in masks.items()}
    rec, total, _ = torch_graph(cfg, tP, torch.tensor(noisy), torch.tensor(X), keep, tm,
                                None if eps is None else torch.tensor(eps))
    total.backward()
    assert abs(rec.item() - c['recon_loss']) <= 1e-10 * max(1, abs(rec.item()))
    assert abs(total.item() - c['total_loss']) <= 1e-10 * max(1, abs(total.item()))
    touched = {k for k, v in tP.items() if v.grad is not None}
    assert touched == set(G.keys())
    for k in G:
        ref = tP[k].grad.numpy()
        assert np.max(np.abs(G[k] - ref)) <= 1e-10 * max(1.0, np.max(np.abs(ref))), k


HEAD_CASES = [
    dict(layers=[12, 6], tie=True, vae=False, act='softsign', loss='sigmoid_cross_entropy', lam=0.01, keep=1.0, head=[5, 4]),
    dict(layers=[12, 6], tie=False, vae=True, act='relu', loss='sigmoid_cross_entropy', lam=0.01, keep=0.5, head=[5, 4]),
    # AE depth 3 > head depth 2+1: activation reaches the logits (reference quirk, :533)
    dict(layers=[12, 8, 6, 5], tie=True, vae=False, act='tanh', loss='mean_squared', lam=0.0, keep=1.0, head=[5]),
    dict(layers=[12, 6], tie=True, vae=False, act='softsign', loss='sigmoid_cross_entropy', lam=0.0, keep=1.0, head=[5, 4],
         num_labels=None, cls_loss='softmax'),
]


@pytest.mark.parametrize('case', HEAD_CASES, ids=lambda c: '-'.join(str(v) for v in c.values()))
def test_cls_backward_matches_autograd(case):
    cfg, P, X, noisy, masks, eps, Y = make_case(5, **case)
    keep = case['keep']
    c = O.forward(cfg, P, noisy, None, keep, masks, eps, true_Y=Y, want_head=True)
    G = O.backward_cls(cfg, P, c)
    tP = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in P.items()}
    tm = None if masks is None else {k: torch.tensor(v) for k, v in masks.items()}
    _, _, cls = torch_graph(cfg, tP, torch.tensor(noisy), torch.tensor(X), keep, tm,
                            None if eps is None else torch.tensor(eps), torch.tensor(Y))
    cls.backward()
    assert abs(cls.item() - c['cls_loss']) <= 1e-10
    touched = {k for k, v in tP.items() if v.grad is not None}
    assert touched == set(G.keys())
    for k in G:
        ref = tP[k].grad.numpy()
        assert np.max(np.abs(G[k] - ref)) <= 1e-10 * max(1.0, np.max(np.abs(ref))), k


def test_tf_adam_formula():
    """TF ApplyAdam (eps outside the bias correction) -- differs from torch.optim.Adam for tiny grads."""
    rng = np.random.default_rng(0)
    P = {'w': rng.standard_normal(50)}
    ref = P['w'].copy()
    st = O.AdamState()
    m = np.zeros(50)
    v = np.zeros(50)
    for t in range(1, 6):
        g = rng.standard_normal(50) * 1e-7
        O.adam_step(P, {'w': g}, st, 1e-3)
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        ref = ref - 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m / (np.sqrt(v) + 1e-8)
    assert np.allclose(P['w'], ref, rtol=0, atol=1e-15)


def test_noise_follows_reference_order():
    """Zeroing first, block mask second (mask wins); RNG order choice -> multinomial per row."""
    cfg = O.OracleConfig(num_feats=31, layer_sizes=[4], modality_starts=STARTS, modality_names=NAMES)
    X = np.full((64, 31), 0.5)
    out = O.add_noise(cfg, X, np.random.RandomState(7))
    rs = np.random.RandomState(7)
    for r in range(64):
        cols = rs.choice(31, size=1)
        k = int(np.argmax(rs.multinomial(1, pvals=cfg.noise_p)))
        exp = X[r].copy()
        exp[cols] = 0
        for n in cfg.noise_types[k]:
            m = NAMES.index(n)
            exp[STARTS[m]:STARTS[m + 1]] = -1.0
        assert np.array_equal(out[r], exp)
    assert np.array_equal(X, np.full((64, 31), 0.5))


def test_fill_missing_rule():
    cfg = O.OracleConfig(num_feats=31, layer_sizes=[4], modality_starts=STARTS, modality_names=NAMES)
    X = np.full((3, 31), 0.25)
    X[0, 11:15] = -1.0
    X[1, 24:31] = -1.0
    X[2, 0:11] = np.array([-2.0, 0.0] + [-1.0] * 9)       # sums to -11 without being all -1: still "missing"
    Xbar = np.full((3, 31), 0.75)
    out = O.fill_missing(cfg, X, Xbar)
    assert np.all(out[0, 11:15] == 0.75) and np.all(out[0, :11] == 0.25)
    assert np.all(out[1, 24:] == 0.75)
    assert np.all(out[2, :11] == 0.75)
