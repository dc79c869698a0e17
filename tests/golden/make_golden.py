"""Generates the committed golden vectors under tests/golden/ from the fp64 oracle.

The reference ships no fixtures and cannot run here, so these vectors are produced by OUR
oracle with fixed seeds and injected weights / masks / epsilon; they pin the oracle against regressions and
give the GPU tests size-small, committed expectations.    Usage:  python tests/golden/make_golden.py
"""
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


def build(name, c, seed):
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
    return out


if __name__ == '__main__':
    for i, (name, c) in enumerate(CASES.items()):
        blob = build(name, c, 100 + i)
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **blob)
        print(name, 'loss', float(blob['recon_loss']), 'bytes', os.path.getsize(os.path.join(HERE, name + '.npz')))
