"""Generates the committed golden vectors under tests/golden/ from the fp64 oracle AND checks each of them against
the reference's own graph code.

The reference ships no fixtures.  Its own Python (converted mechanically to Python 3 into oracle/_ref by
oracle/build_ref.py, running on the TF-1 shim in oracle/tf1_shim) is executed here on the same injected weights /
masks / epsilon; the values it produces are stored next to the oracle's under `ref/...` and `ref/verified` is 1 only
when losses, every gradient and the parameters after three Adam steps agree to 1e-9 relative.  The GPU tests then
compare the engine with vectors that the reference's graph wiring has reproduced.
Usage:  python tests/golden/make_golden.py        (needs /root/reference or a prebuilt oracle/_ref)
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mmae_oracle as O            # noqa: E402
from oracle import philox_host as PH           # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
S_STARTS, S_NAMES = [0, 200, 220, 240, 270, 320], ['phys', 'call', 'sms', 'screen', 'location']
T_STARTS = [0, 11, 15, 19, 24, 31]

CASES = {
    'tiny_tied_sce': dict(num_feats=31, starts=T_STARTS, layers=[12, 6], tie=True, B=16, lam=0.01, loss='sigmoid_cross_entropy'),
    'tiny_vae_head': dict(num_feats=31, starts=T_STARTS, layers=[12, 6], vae=True, B=16, lam=0.001, head=[5, 4], act='relu'),
    'small_untied_rmse': dict(num_feats=320, starts=S_STARTS, layers=[128, 64], tie=False, B=48, lam=0.001, loss='mean_squared', act='tanh'),
    'small_tied_sce_dropout': dict(num_feats=320, starts=S_STARTS, layers=[128, 64, 32], tie=True, B=48, lam=0.0, keep=0.5, act='softsign'),
}


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def reference_run(c, cfg, P0, X, noisy, eps, keep, masks_for, Y):
    """The same three optimizer steps (+ one classification step) through the reference's own build_graph."""
    import torch
    from oracle.ref_loader import load_reference
    REF = load_reference(torch.float64)
    if REF is None:
        return None
    dl = REF.make_loader(X, X, c['starts'], S_NAMES, Y, Y, 3 if c.get('head') else None)
    with contextlib.redirect_stdout(io.StringIO()):
        m = REF.mmae.MultimodalAutoencoder(
            data_loader=dl, classification_data_loader=dl if c.get('head') else None, layer_sizes=list(c['layers']),
            variational=cfg.variational, tie_weights=cfg.tie_weights, batch_size=len(X), learning_rate=1e-3,
            dropout_prob=keep, weight_penalty=c['lam'], activation_func=cfg.activation, loss_func=cfg.loss_func,
            classification_layer_sizes=c.get('head'), weight_initialization='normal', verbose=False)
        if c.get('head'):
            m.set_classification_params(learning_rate=1e-3, weight_penalty=0.001, activation_func=cfg.cls_activation,
                                        suppress_warning=True)
    REF.set_variables(m, P0)
    h = REF.tf.hooks
    L = len(c['layers'])
    out = {}
    try:
        h.random_normal = (lambda shp, name: eps) if eps is not None else None
        h.dropout_uniform = None
        out['recon_loss'] = float(m.session.run(m.reconstruction_loss, {m.noisy_X: noisy, m.true_X: X, m.tf_dropout_prob: 1.0}))
        losses = []
        for s in range(3):
            md = masks_for(s)
            if md is not None:
                order = [md['enc%d' % i] for i in range(L - 1)] + [md['dec%d' % j] for j in range(L - 1)]
                h.dropout_uniform = lambda shp, i: np.where(order[i] > 0, 1.0 - 0.5 * keep, 0.5 * (1.0 - keep))
            rl, _ = m.session.run([m.reconstruction_loss, m.opt_step], {m.noisy_X: noisy, m.true_X: X, m.tf_dropout_prob: keep})
            losses.append(float(rl))
            if s == 0:
                out['g'] = dict(m.opt_step.opt.last_grads)
        h.dropout_uniform = None
        out['losses'] = np.asarray(losses)
        out['p3'] = REF.get_variables(m)
        if c.get('head'):
            cl, acc, pred, _ = m.session.run([m.classification_loss, m.accuracy, m.predictions, m.classification_opt_step],
                                             {m.noisy_X: noisy, m.true_Y: Y, m.tf_dropout_prob: 1.0})
            out['cls_total_loss'], out['cls_acc'], out['cls_pred'] = float(cl), float(acc), pred
            out['gh'] = dict(m.classification_opt_step.opt.last_grads)
    finally:
        h.random_normal = None
        h.dropout_uniform = None
    return out


def build(name, c, seed, with_ref=True):
    rng = np.random.default_rng(seed)
    cfg = O.OracleConfig(num_feats=c['num_feats'], layer_sizes=list(c['layers']), modality_starts=list(c['starts']),
                         modality_names=S_NAMES, tie_weights=c.get('tie', False), variational=c.get('vae', False),
                         activation=c.get('act', 'softsign'), loss_func=c.get('loss', 'sigmoid_cross_entropy'),
                         weight_penalty=c['lam'], learning_rate=1e-3, cls_layer_sizes=c.get('head'), cls_weight_penalty=0.001,
                         cls_learning_rate=1e-3)
    P0 = O.init_params(cfg, rng)
    B, F = c['B'], c['num_feats']
    X = rng.uniform(0, 1, (B, F)).astype(np.float32).astype(np.float64)
    type_masks = [sum(1 << S_NAMES.index(n) for n in t) for t in cfg.noise_types]
    zb, mb = PH.noise_descriptor(0, 0, B, F, 5, int(F * .05), True, PH.categorical_thresholds(cfg.noise_p), type_masks, 1)
    noisy = O.noise_from_descriptor(cfg, X, zb, mb)
    keep = c.get('keep', 1.0)

    def masks_for(step):
        """Engine dropout masks of optimizer step `step` (Philox rng_step = 1 + step; slots enc i / 32 + dec j)."""
        if keep >= 1.0:
            return None
        thr = PH.keep_threshold(keep)
        d = [F] + list(c['layers'])
        L = len(c['layers'])
        m = {}
        for i in range(L - 1):
            m['enc%d' % i] = PH.dropout_mask(0, 1 + step, i, B, d[i + 1], thr)
        for j in range(L - 1):
            m['dec%d' % j] = PH.dropout_mask(0, 1 + step, 32 + j, B, d[L - 1 - j], thr)
        return m
    eps = rng.standard_normal((B, c['layers'][-1])).astype(np.float32).astype(np.float64) if cfg.variational else None
    out = {'X': X.astype(np.float32), 'zero_bits': zb, 'mod_bits': mb, 'noisy': noisy.astype(np.float32)}
    if eps is not None:
        out['eps'] = eps.astype(np.float32)
    for k, v in P0.items():
        out['p0/' + k] = v.astype(np.float32)
    P = {k: v.copy() for k, v in P0.items()}
    st = O.AdamState()
    cfwd = O.forward(cfg, P, noisy, X, eps=eps)            # keep = 1 fetches
    out['recon_loss'] = np.float64(cfwd['recon_loss'])
    out['decoded'] = cfwd['decoded'].astype(np.float32)
    out['embedding'] = cfwd['emb'].astype(np.float32)
    losses = []
    for s in range(3):                                     # three optimizer steps on the same batch
        c_, G = O.train_step(cfg, P, st, noisy, X, keep=keep, drop_masks=masks_for(s), eps=eps)
        losses.append(c_['recon_loss'])
        if s == 0:
            for k, g in G.items():
                out['g/' + k] = g.astype(np.float32)
    out['losses'] = np.asarray(losses)
    for k, v in P.items():
        out['p3/' + k] = v.astype(np.float32)
    P3 = {k: v.copy() for k, v in P.items()}
    Y = ch = Gh = None
    if c.get('head'):
        Y = (rng.uniform(size=(B, 3)) < 0.5).astype(np.float64)
        out['Y'] = Y.astype(np.float32)
        st2 = O.AdamState()
        ch, Gh = O.cls_train_step(cfg, P, st2, noisy, Y, eps=eps)
        out['cls_loss'] = np.float64(ch['cls_data_loss'])
        out['cls_acc'] = np.float64(ch['accuracy'])
        out['cls_pred'] = ch['predictions']
        for k, g in Gh.items():
            out['gh/' + k] = g.astype(np.float32)
    if with_ref:
        r = reference_run(c, cfg, P0, X, noisy, eps, keep, masks_for, Y)
        if r is not None:
            errs = [_rel(r['recon_loss'], cfwd['recon_loss']), _rel(r['losses'], losses)]
            G0 = O.backward_recon(cfg, P0, O.forward(cfg, P0, noisy, X, keep, masks_for(0), eps))
            assert set(r['g']) == set(G0)
            errs += [_rel(r['g'][k], G0[k]) for k in G0]
            errs += [_rel(r['p3'][k], P3[k]) for k in P3]
            if ch is not None:
                assert np.array_equal(r['cls_pred'], ch['predictions']) and set(r['gh']) == set(Gh)
                errs += [_rel(r['cls_total_loss'], ch['cls_loss']), _rel(r['cls_acc'], ch['accuracy'])]
                errs += [_rel(r['gh'][k], Gh[k]) for k in Gh]
                out['ref/cls_total_loss'] = np.float64(r['cls_total_loss'])
            out['ref/recon_loss'] = np.float64(r['recon_loss'])
            out['ref/losses'] = r['losses']
            out['ref/max_rel_err'] = np.float64(max(errs))
            out['ref/verified'] = np.int32(1 if max(errs) < 1e-9 else 0)
    return out


if __name__ == '__main__':
    for i, (name, c) in enumerate(CASES.items()):
        blob = build(name, c, 100 + i)
        assert int(blob.get('ref/verified', 0)) == 1, (name, 'reference run missing or disagreeing', blob.get('ref/max_rel_err'))
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **blob)
        print(name, 'loss', float(blob['recon_loss']), 'ref max rel err', float(blob['ref/max_rel_err']),
              'bytes', os.path.getsize(os.path.join(HERE, name + '.npz')))
