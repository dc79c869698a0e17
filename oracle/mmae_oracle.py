"""fp64 NumPy ORACLE for the MMAE hot path.  TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

PINNING: the reference ships no tests, fixtures or golden vectors and computes in
TensorFlow 1.x, which is neither installed here nor installable (no network).  The
reference's OWN Python is run instead: `oracle/build_ref.py` converts its Python-2
source mechanically to Python 3 (`oracle/_ref/`, git-ignored build product) and
`oracle/tf1_shim` supplies the ~45 TF-1 API names it calls on top of torch autograd.
`tests/test_oracle_vs_ref.py` checks this file against that run -- the reference's
build_graph / encode / decode / classify / add_noise_to_batch / train /
train_classification / predict / per-modality RMSE and its DataLoader's sampling and
missing-block rule -- to 1e-9 (noise, indices, predictions: bit-exact), and
`tests/golden/make_golden.py` stores the reference-run values in every golden file.
What remains restated rather than executed: the TF op KERNELS themselves (items 1-9
below, implemented in the shim from the published TF-1.x semantics); the graph wiring,
variable sets per optimizer, RNG call order and host logic are the reference's code.
It is also cross-checked against an independent torch.autograd fp64 derivation
(tests/test_oracle_autograd.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product path
(multimodalautoencoder_b200/) never does.

Reference lines restated here (all in /root/reference/multimodal_autoencoder.py
unless a file is named):
  parameter shapes / tying / decoder reversal ........ :256-342
  graph: encode, VAE sample, decode .................. :366-378, :454-518
  reconstruction losses .............................. :381-390
  L2 regulariser, KL, total loss ..................... :393-408
  Adam (tf.train.AdamOptimizer defaults) ............. :411, :443
  head: classify, losses, predictions, accuracy ...... :420-452, :520-540
  block-mask noise ................................... :649-702
  batch sampling ..................................... data_funcs.py:161-195
  missing-block rule / fill-in ....................... data_funcs.py:310-381
  per-modality RMSE .................................. :1189-1220

TF-1.x semantics assumed:
  (1) sigmoid_cross_entropy_with_logits(l,z) = max(l,0) - l*z + log1p(exp(-|l|))
  (2) softsign' = 1/(1+|x|)^2 ; softplus' = sigmoid(x)
  (3) l2_loss(w) = sum(w^2)/2
  (4) ApplyAdam: a = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g^2-v)(1-b2);
      theta -= a*m/(sqrt(v)+eps)   (eps NOT bias corrected; t starts at 1)
  (5) exponential_decay is inert (global_step never incremented) -> constant lr
  (6) dropout(x, keep) = x * floor(keep + U[0,1)) / keep
  (7) tf.round = half-to-even, cast float->int32 truncates
  (8) reduce_mean(scalar + vector[B]) broadcasts the scalar
  (9) tied decoder = transpose view: one variable, two gradient paths, double L2
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

F64 = np.float64

DEFAULT_NOISE_P = [0.64018104, 0.03168217, 0.25119437, 0.07694242]  # :202
DEFAULT_NOISE_TYPES = [[], ['call', 'sms', 'screen'], ['location'],
                       ['location', 'call', 'sms', 'screen']]        # :203-206


@dataclass
class OracleConfig:
    num_feats: int
    layer_sizes: List[int]
    modality_starts: List[int]            # length M+1, last == num_feats (data_funcs.py:121-122)
    modality_names: List[str]
    tie_weights: bool = True
    variational: bool = False
    activation: str = 'softsign'
    loss_func: str = 'sigmoid_cross_entropy'
    weight_penalty: float = 0.0
    learning_rate: float = 1e-4
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-8
    mask_with: float = -1.0
    intelligent_noise: bool = True
    num_modalities_to_drop: int = 1
    noise_p: List[float] = field(default_factory=lambda: list(DEFAULT_NOISE_P))
    noise_types: List[List[str]] = field(default_factory=lambda: copy.deepcopy(DEFAULT_NOISE_TYPES))
    # classification head (None -> no head)
    cls_layer_sizes: Optional[List[int]] = None
    num_labels: Optional[int] = 3         # None -> 2 logits, sparse softmax target (:324-327)
    cls_activation: str = 'softsign'
    cls_loss: str = 'sigmoid_cross_entropy'
    cls_weight_penalty: float = 0.0
    cls_learning_rate: float = 1e-4

    def __post_init__(self):
        if self.variational:              # :175-179
            self.tie_weights = False
            self.loss_func = 'sigmoid_cross_entropy'
            if len(self.layer_sizes) < 2:
                raise ValueError('variational MMAE needs >= 2 encoder layers (:299)')

    @property
    def L(self):
        return len(self.layer_sizes)

    def enc_dims(self):
        d = [self.num_feats] + list(self.layer_sizes)
        return [(d[i], d[i + 1]) for i in range(self.L)]

    def head_dims(self):
        if self.cls_layer_sizes is None:
            return []
        out = self.num_labels if self.num_labels is not None else 2
        d = [self.layer_sizes[-1]] + list(self.cls_layer_sizes) + [out]
        return [(d[i], d[i + 1]) for i in range(len(d) - 1)]


# --------------------------------------------------------------------------- params
def param_shapes(cfg: OracleConfig) -> Dict[str, tuple]:
    """Variable names and shapes exactly as the reference creates them (:271-336)."""
    shapes = {}
    for i, (din, dout) in enumerate(cfg.enc_dims()):
        shapes['weights%d' % i] = (din, dout)
        if not cfg.tie_weights:
            shapes['decode_weights%d' % i] = (dout, din)
        shapes['encode_biases%d' % i] = (dout,)
        shapes['decode_biases%d' % i] = (din,)
    if cfg.variational:
        shapes['variance_weights'] = (cfg.layer_sizes[-2], cfg.layer_sizes[-1])
        shapes['variance_bias'] = (cfg.layer_sizes[-1],)
    for i, (din, dout) in enumerate(cfg.head_dims()):
        shapes['classification_weights%d' % i] = (din, dout)
        shapes['classification_biases%d' % i] = (dout,)
    return shapes


def init_params(cfg: OracleConfig, rng: np.random.Generator, init='normal') -> Dict[str, np.ndarray]:
    """Same *distributions* as :22-56 (the TF RNG stream itself is not reproducible)."""
    out = {}
    for name, shp in param_shapes(cfg).items():
        if len(shp) == 1:
            out[name] = np.full(shp, 0.1, F64)                       # :55
        elif init == 'xavier':
            lim = np.sqrt(6.0 / (shp[0] + shp[1]))                   # :40-42
            out[name] = rng.uniform(-lim, lim, shp)
        else:
            sd = 1.0 / np.sqrt(float(shp[0]))                        # :44
            w = rng.standard_normal(shp)
            bad = np.abs(w) > 2.0                                    # truncated normal: redraw
            while bad.any():
                w[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(w) > 2.0
            out[name] = w * sd
    # values live as fp32 in both engine and reference; keep them fp32-representable
    return {k: v.astype(np.float32).astype(F64) for k, v in out.items()}


def decoder_stack(cfg, P):
    """Decoder (weight, bias, enc_index, tied) per decoder layer j, after the :304-305 reversal."""
    out = []
    for j in range(cfg.L):
        i = cfg.L - 1 - j
        W = P['weights%d' % i].T if cfg.tie_weights else P['decode_weights%d' % i]
        out.append((W, P['decode_biases%d' % i], i))
    return out


# --------------------------------------------------------------------------- activations
def act_fwd(name, z):
    if name == 'relu':
        return np.maximum(z, 0.0)
    if name == 'tanh':
        return np.tanh(z)
    if name == 'softsign':
        return z / (1.0 + np.abs(z))
    if name == 'softplus':
        return np.logaddexp(0.0, z)
    return z                                                         # :497 anything else is linear


def act_bwd(name, z):
    if name == 'relu':
        return (z > 0).astype(F64)
    if name == 'tanh':
        return 1.0 - np.tanh(z) ** 2
    if name == 'softsign':
        return 1.0 / (1.0 + np.abs(z)) ** 2
    if name == 'softplus':
        return 1.0 / (1.0 + np.exp(-z))
    return np.ones_like(z)


def sigmoid(x):
    return 0.5 * (1.0 + np.tanh(0.5 * x))


def sigmoid_ce(l, z):
    return np.maximum(l, 0.0) - l * z + np.log1p(np.exp(-np.abs(l)))


# --------------------------------------------------------------------------- forward
def forward(cfg: OracleConfig, P, noisy_X, true_X=None, keep=1.0, drop_masks=None, eps=None,
            true_Y=None, want_head=False):
    """Whole graph (:366-408, :428-452).  drop_masks: dict name -> 0/1 array for
    'enc%d' / 'dec%d' / 'cls%d' activated layers (required when keep < 1)."""
    c = {'keep': keep}
    act = cfg.activation
    a = np.asarray(noisy_X, F64)
    c['a0'] = a

    def drop(h, key):
        if keep >= 1.0 and (drop_masks is None or key not in drop_masks):
            c['m_' + key] = None
            return h
        m = np.asarray(drop_masks[key], F64)
        c['m_' + key] = m
        return h * m / keep

    # encoder (:454-475)
    for i in range(cfg.L):
        last = i == cfg.L - 1
        if cfg.variational and last:
            c['lv'] = a @ P['variance_weights'] + P['variance_bias']
        z = a @ P['weights%d' % i] + P['encode_biases%d' % i]
        c['ez%d' % i] = z
        if not last:
            a = drop(act_fwd(act, z), 'enc%d' % i)
            c['ea%d' % i] = a
    mu = z
    c['mu'] = mu
    if cfg.variational:                                              # :372-375
        e = np.asarray(eps, F64)
        c['eps'] = e
        emb = mu + e * np.exp(c['lv'])
    else:
        emb = mu
    c['emb'] = emb

    # decoder (:499-518)
    u = emb
    dec = decoder_stack(cfg, P)
    for j, (D, b, _) in enumerate(dec):
        c['du%d' % j] = u
        y = u @ D + b
        c['dy%d' % j] = y
        if j < cfg.L - 1:
            u = drop(act_fwd(act, y), 'dec%d' % j)
    logits = y
    c['logits'] = logits

    if true_X is not None:
        X = np.asarray(true_X, F64)
        c['X'] = X
        if cfg.loss_func == 'mean_squared':                          # :383-384 (an RMSE)
            c['recon_loss'] = np.sqrt(np.mean((logits - X) ** 2))
            c['decoded'] = logits
        elif cfg.loss_func == 'cross_entropy':                       # :386
            c['recon_loss'] = -np.sum(X * np.log(logits))
            c['decoded'] = logits
        else:                                                        # :388-390
            c['recon_loss'] = np.sum(sigmoid_ce(logits, X))
            c['decoded'] = sigmoid(logits)
        reg = 0.0                                                    # :394-397
        for i in range(cfg.L):
            reg += 0.5 * np.sum(P['weights%d' % i] ** 2)
        for (D, _, _) in dec:
            reg += 0.5 * np.sum(D ** 2)
        if cfg.variational:
            reg += 0.5 * np.sum(P['variance_weights'] ** 2)
        c['reg_loss'] = cfg.weight_penalty * reg
        if cfg.variational:                                          # :402-406
            lv = c['lv']
            kl = -0.5 * np.sum(1 + 2 * lv - emb ** 2 - np.exp(2 * lv), axis=1)
            c['kl'] = kl
            c['total_loss'] = np.mean(c['recon_loss'] + kl) + c['reg_loss']
        else:
            c['total_loss'] = c['recon_loss'] + c['reg_loss']
    else:
        c['decoded'] = sigmoid(logits) if cfg.loss_func == 'sigmoid_cross_entropy' else logits

    if want_head:
        h = emb
        hd = cfg.head_dims()
        for i in range(len(hd)):
            c['cu%d' % i] = h
            z = h @ P['classification_weights%d' % i] + P['classification_biases%d' % i]
            c['cz%d' % i] = z
            c['c_act%d' % i] = i < cfg.L - 1                         # :533 bound uses AE depth
            if i < cfg.L - 1:
                h = drop(act_fwd(cfg.cls_activation, z), 'cls%d' % i)
            else:
                h = z
        lg = h
        c['cls_logits'] = lg
        c['class_prob'] = sigmoid(lg)                                # :446
        if cfg.cls_loss == 'sigmoid_cross_entropy':
            c['predictions'] = (lg > 0).astype(np.int32)             # round-half-even(sigmoid) :448
        else:
            c['predictions'] = np.argmax(lg, axis=1).astype(np.int32)  # :450
        if true_Y is not None:
            Y = np.asarray(true_Y, F64)
            c['Y'] = Y
            if cfg.cls_loss == 'sigmoid_cross_entropy':              # :432-433
                c['cls_data_loss'] = np.mean(sigmoid_ce(lg, Y))
            else:                                                    # :437-438
                yi = Y.astype(np.int64)
                mx = lg.max(axis=1, keepdims=True)
                lse = mx[:, 0] + np.log(np.sum(np.exp(lg - mx), axis=1))
                c['cls_data_loss'] = np.mean(lse - lg[np.arange(len(lg)), yi])
            regc = sum(0.5 * np.sum(P['classification_weights%d' % i] ** 2) for i in range(len(hd)))
            c['cls_loss'] = c['cls_data_loss'] + cfg.cls_weight_penalty * regc   # :441
            c['accuracy'] = np.mean(c['predictions'] == Y.astype(np.int32))      # :451-452
    return c


# --------------------------------------------------------------------------- backward
def _through(c, key, zkey, actname, g):
    """d/dz of drop(act(z)): same mask, 1/keep, times act'."""
    m = c.get('m_' + key)
    if m is not None:
        g = g * m / c['keep']
    return g * act_bwd(actname, c[zkey])


def backward_recon(cfg: OracleConfig, P, c) -> Dict[str, np.ndarray]:
    """Gradient of total_loss (:406/:408) wrt every variable opt_step touches."""
    G = {}
    lam = cfg.weight_penalty
    X, logits = c['X'], c['logits']
    B = X.shape[0]
    if cfg.loss_func == 'mean_squared':
        N = X.size
        d = (logits - X) / (N * c['recon_loss'])
    elif cfg.loss_func == 'cross_entropy':
        d = -X / logits
    else:
        d = sigmoid(logits) - X
    dec = decoder_stack(cfg, P)
    tied_acc = {}
    for j in range(cfg.L - 1, -1, -1):
        D, _, i = dec[j]
        u = c['du%d' % j]
        dD = u.T @ d
        G['decode_biases%d' % i] = d.sum(axis=0)
        if cfg.tie_weights:
            tied_acc[i] = dD.T + lam * P['weights%d' % i]            # second L2 hit on the tied var
        else:
            G['decode_weights%d' % i] = dD + lam * D
        g_u = d @ D.T
        if j > 0:
            d = _through(c, 'dec%d' % (j - 1), 'dy%d' % (j - 1), cfg.activation, g_u)
    g_e = g_u
    if cfg.variational:
        emb, lv, eps = c['emb'], c['lv'], c['eps']
        g_e = g_e + emb / B
        g_mu = g_e
        g_lv = g_e * eps * np.exp(lv) + (-1.0 + np.exp(2 * lv)) / B
    else:
        g_mu = g_e
        g_lv = None
    _backward_encoder(cfg, P, c, g_mu, g_lv, G, lam, tied_acc)
    return G


def _backward_encoder(cfg, P, c, g_mu, g_lv, G, lam, tied_acc):
    d = g_mu
    for i in range(cfg.L - 1, -1, -1):
        a_in = c['a0'] if i == 0 else c['ea%d' % (i - 1)]
        gW = a_in.T @ d + lam * P['weights%d' % i]
        if tied_acc and i in tied_acc:
            gW = gW + tied_acc[i]
        G['weights%d' % i] = gW
        G['encode_biases%d' % i] = d.sum(axis=0)
        if i == 0:
            break
        g_in = d @ P['weights%d' % i].T
        if i == cfg.L - 1 and g_lv is not None:
            G['variance_weights'] = a_in.T @ g_lv + lam * P['variance_weights']
            G['variance_bias'] = g_lv.sum(axis=0)
            g_in = g_in + g_lv @ P['variance_weights'].T
        d = _through(c, 'enc%d' % (i - 1), 'ez%d' % (i - 1), cfg.activation, g_in)


def backward_cls(cfg: OracleConfig, P, c) -> Dict[str, np.ndarray]:
    """Gradient of classification_loss (:432-441) wrt encoder (+variance) + head variables."""
    G = {}
    lg, Y = c['cls_logits'], c['Y']
    B = lg.shape[0]
    if cfg.cls_loss == 'sigmoid_cross_entropy':
        d = (sigmoid(lg) - Y) / lg.size
    else:
        mx = lg.max(axis=1, keepdims=True)
        p = np.exp(lg - mx)
        p /= p.sum(axis=1, keepdims=True)
        p[np.arange(B), Y.astype(np.int64)] -= 1.0
        d = p / B
    hd = cfg.head_dims()
    for i in range(len(hd) - 1, -1, -1):
        if c['c_act%d' % i]:
            d = _through(c, 'cls%d' % i, 'cz%d' % i, cfg.cls_activation, d)
        W = P['classification_weights%d' % i]
        G['classification_weights%d' % i] = c['cu%d' % i].T @ d + cfg.cls_weight_penalty * W
        G['classification_biases%d' % i] = d.sum(axis=0)
        d = d @ W.T
    g_e = d
    if cfg.variational:
        g_mu = g_e
        g_lv = g_e * c['eps'] * np.exp(c['lv'])
    else:
        g_mu, g_lv = g_e, None
    _backward_encoder(cfg, P, c, g_mu, g_lv, G, 0.0, None)
    return G


# --------------------------------------------------------------------------- Adam
class AdamState:
    def __init__(self):
        self.t = 0
        self.m: Dict[str, np.ndarray] = {}
        self.v: Dict[str, np.ndarray] = {}


def adam_step(P, G, st: AdamState, lr, b1=0.9, b2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer._apply_dense (TF 1.x), see header item (4)."""
    st.t += 1
    a = lr * np.sqrt(1.0 - b2 ** st.t) / (1.0 - b1 ** st.t)
    for k, g in G.items():
        m = st.m.setdefault(k, np.zeros_like(P[k]))
        v = st.v.setdefault(k, np.zeros_like(P[k]))
        m += (g - m) * (1.0 - b1)
        v += (g * g - v) * (1.0 - b2)
        P[k] = P[k] - a * m / (np.sqrt(v) + eps)


def train_step(cfg, P, st, noisy_X, true_X, keep=1.0, drop_masks=None, eps=None):
    c = forward(cfg, P, noisy_X, true_X, keep, drop_masks, eps)
    G = backward_recon(cfg, P, c)
    adam_step(P, G, st, cfg.learning_rate, cfg.beta1, cfg.beta2, cfg.adam_eps)
    return c, G


def cls_train_step(cfg, P, st, noisy_X, true_Y, keep=1.0, drop_masks=None, eps=None):
    c = forward(cfg, P, noisy_X, None, keep, drop_masks, eps, true_Y=true_Y, want_head=True)
    G = backward_cls(cfg, P, c)
    adam_step(P, G, st, cfg.cls_learning_rate, cfg.beta1, cfg.beta2, cfg.adam_eps)
    return c, G


# --------------------------------------------------------------------------- noise / sampling
def sample_batch_indices(n_rows, batch_size, rng=np.random):
    """data_funcs.py:167,176,185,194 -- np.random.choice(n, size=B), with replacement."""
    return rng.choice(n_rows, size=batch_size)


def add_noise(cfg: OracleConfig, X, rng=np.random, missing_modes=()):
    """Block-mask noise, one row at a time, in the reference's RNG call order (:668-702):
    5% of the columns (drawn WITH replacement) -> 0, then whole modality blocks -> mask_with."""
    out = np.array(X, dtype=F64, copy=True)
    nfeat = out.shape[1]
    n_zero = int(nfeat * .05)
    for r in range(out.shape[0]):
        cols = rng.choice(nfeat, size=n_zero)
        out[r, cols] = 0
        if cfg.intelligent_noise:
            k = int(np.argmax(rng.multinomial(1, pvals=cfg.noise_p)))
            names = list(missing_modes) if len(missing_modes) > 0 else cfg.noise_types[k]
            mods = [cfg.modality_names.index(n) for n in names]
        else:
            mods = [int(rng.randint(0, len(cfg.modality_names))) for _ in range(cfg.num_modalities_to_drop)]
        for m in mods:
            out[r, cfg.modality_starts[m]:cfg.modality_starts[m + 1]] = cfg.mask_with
    return out


def noise_from_descriptor(cfg, X, zero_bits, mod_bits):
    """Apply a (zero bitmap [B, ceil(F/32)] uint32, modality bitmask [B] uint32) descriptor:
    the compact form the engine consumes; masked wins over zeroed (:683 then :695)."""
    out = np.array(X, dtype=F64, copy=True)
    B, nfeat = out.shape
    cols = np.arange(nfeat)
    z = (zero_bits[:, cols // 32] >> (cols % 32).astype(np.uint32)) & 1
    out[z.astype(bool)] = 0.0
    for m in range(len(cfg.modality_names)):
        rows = ((mod_bits >> np.uint32(m)) & 1).astype(bool)
        out[rows, cfg.modality_starts[m]:cfg.modality_starts[m + 1]] = cfg.mask_with
    return out


# --------------------------------------------------------------------------- inference helpers
def missing_blocks(cfg, X):
    """data_funcs.py:376-380: block m of a row is missing iff sum(x[s:e]) == -(e-s)."""
    X = np.asarray(X, F64)
    miss = np.zeros((X.shape[0], len(cfg.modality_names)), bool)
    for m in range(len(cfg.modality_names)):
        s, e = cfg.modality_starts[m], cfg.modality_starts[m + 1]
        miss[:, m] = X[:, s:e].sum(axis=1) == -1 * (e - s)
    return miss


def fill_missing(cfg, X, Xbar):
    """data_funcs.py:328-348: reconstruction on missing blocks, original elsewhere."""
    out = np.array(X, dtype=F64, copy=True)
    miss = missing_blocks(cfg, X)
    for m in range(len(cfg.modality_names)):
        s, e = cfg.modality_starts[m], cfg.modality_starts[m + 1]
        out[miss[:, m], s:e] = np.asarray(Xbar, F64)[miss[:, m], s:e]
    return out


def reconstruction_loss_per_modality(cfg, P, X):
    """:1189-1216: per modality, mask it with literal -1.0 for all rows, predict (true_X is
    the *masked* matrix, :941-942), RMSE on that block's columns against the clean X."""
    X = np.asarray(X, F64)
    out = []
    for m in range(len(cfg.modality_names)):
        s, e = cfg.modality_starts[m], cfg.modality_starts[m + 1]
        nz = X.copy()
        nz[:, s:e] = -1.0
        c = forward(cfg, P, nz, nz)
        out.append(float(np.sqrt(np.mean((X[:, s:e] - c['decoded'][:, s:e]) ** 2))))
    return out
