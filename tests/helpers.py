"""Shared builders for the parity tests: one description -> oracle config + engine config."""
import numpy as np

from oracle import mmae_oracle as O
from oracle import philox_host as PH

S_STARTS = [0, 200, 220, 240, 270, 320]                 # SNAPSHOT-shaped small config
S_NAMES = ['phys', 'call', 'sms', 'screen', 'location']
T_STARTS = [0, 11, 15, 19, 24, 31]                      # tiny ragged config (F = 31)


def make_cfgs(num_feats=320, starts=S_STARTS, names=S_NAMES, layers=(128, 64), tie=True, vae=False, act='softsign',
              loss='sigmoid_cross_entropy', lam=0.0, lr=1e-3, head=None, num_labels=3, cls_loss='sigmoid_cross_entropy',
              cls_act=None, cls_lam=0.0, cls_lr=1e-4, intelligent=True, num_drop=1, precision='fp32', seed=0):
    ocfg = O.OracleConfig(num_feats=num_feats, layer_sizes=list(layers), modality_starts=list(starts),
                          modality_names=list(names), tie_weights=tie, variational=vae, activation=act,
                          loss_func=loss, weight_penalty=lam, learning_rate=lr, cls_layer_sizes=head,
                          num_labels=num_labels, cls_activation=cls_act or act, cls_loss=cls_loss,
                          cls_weight_penalty=cls_lam, cls_learning_rate=cls_lr, intelligent_noise=intelligent,
                          num_modalities_to_drop=num_drop)
    from multimodalautoencoder_b200 import EngineConfig
    ecfg = EngineConfig(num_feats=num_feats, layer_sizes=list(layers), modality_starts=list(starts),
                        modality_names=list(names), tie_weights=ocfg.tie_weights, variational=vae, activation=act,
                        loss_func=ocfg.loss_func, weight_penalty=lam, learning_rate=lr, cls_layer_sizes=head,
                        num_labels=num_labels, cls_activation=cls_act or act, cls_loss=cls_loss,
                        cls_weight_penalty=cls_lam, cls_learning_rate=cls_lr, intelligent_noise=intelligent,
                        num_modalities_to_drop=num_drop, seed=seed, precision=precision)
    return ocfg, ecfg


def rel_err(a, ref):
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.max(np.abs(a - ref)) / max(np.max(np.abs(ref)), 1e-30))


def dropout_masks(ocfg, seed, step, B, keep, row0=0):
    """The engine's Philox dropout masks, from the host twin (slots: enc i, 32 + dec j, 64 + head i)."""
    thr = PH.keep_threshold(keep)
    L = ocfg.L
    d = [ocfg.num_feats] + list(ocfg.layer_sizes)
    m = {}
    for i in range(L - 1):
        m['enc%d' % i] = PH.dropout_mask(seed, step, i, B, d[i + 1], thr, row0)
    for j in range(L - 1):
        m['dec%d' % j] = PH.dropout_mask(seed, step, 32 + j, B, d[L - 1 - j], thr, row0)
    for i, (_, dout) in enumerate(ocfg.head_dims()):
        m['cls%d' % i] = PH.dropout_mask(seed, step, 64 + i, B, dout, thr, row0)
    return m
