"""Loads the converted reference (`oracle/_ref`, see build_ref.py) on top of the TF-1 shim.
TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/, smoke() and bench.py's reference arm import it.

    ref = load_reference()                     # None when oracle/_ref cannot be built or found
    dl  = ref.make_loader(train_X, val_X, starts, names, train_Y=..., val_Y=..., num_labels=3)
    m   = ref.mmae.MultimodalAutoencoder(data_loader=dl, layer_sizes=[16, 8], ...)   # the reference's class
    ref.set_variables(m, params)               # inject weights by reference variable name
    m.session.run([m.opt_step], feed)          # the reference's own graph, differentiated by torch.autograd

`make_loader` builds the reference's own DataLoader object without running its CSV/pandas constructor
(`DataLoader.__init__` needs APIs removed from pandas long ago; tests/test_loader_vs_ref.py runs it under
`pandas_compat.legacy_pandas()`): the instance is allocated with
`object.__new__` and given exactly the attributes the model reads, so the batch
sampling (`data_funcs.py:161-195`) and the missing-block rule (`:366-381`) that run are the reference's.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
SHIM_DIR = os.path.join(HERE, 'tf1_shim')

_cached = None


def _load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _stub_matplotlib():
    """The reference imports matplotlib.pyplot at module scope (plots only); absent in this image."""
    mpl = types.ModuleType('matplotlib')
    plt = types.ModuleType('matplotlib.pyplot')

    def _noop(*a, **k):
        return None
    for fn in ('figure', 'plot', 'legend', 'show', 'title', 'xlabel', 'ylabel', 'savefig', 'close', 'scatter', 'hist'):
        setattr(plt, fn, _noop)
    mpl.pyplot = plt
    return mpl, plt


class Reference:
    def __init__(self, tf, mmae, data_funcs, neural_net=None, generic_wrapper=None, helper_funcs=None,
                 autoencoder_wrapper=None, autoencoder_classification_wrapper=None):
        self.tf, self.mmae, self.data_funcs, self.neural_net = tf, mmae, data_funcs, neural_net
        self.generic_wrapper, self.helper_funcs = generic_wrapper, helper_funcs
        self.autoencoder_wrapper = autoencoder_wrapper
        self.autoencoder_classification_wrapper = autoencoder_classification_wrapper

    def make_loader(self, train_X, val_X, modality_starts, modality_names, train_Y=None, val_Y=None, num_labels=None,
                    test_X=None, test_Y=None):
        dl = object.__new__(self.data_funcs.DataLoader)
        dl.train_X = np.asarray(train_X, np.float64)
        dl.val_X = np.asarray(val_X, np.float64)
        dl.test_X = np.asarray(test_X if test_X is not None else val_X, np.float64)
        dl.train_Y = None if train_Y is None else np.asarray(train_Y)
        dl.val_Y = None if val_Y is None else np.asarray(val_Y)
        dl.test_Y = None if test_Y is None else np.asarray(test_Y)
        dl.num_feats = dl.train_X.shape[1]
        dl.num_labels = num_labels
        dl.modality_start_indices = list(modality_starts)
        dl.modality_names = list(modality_names)
        dl.num_modalities = len(modality_names)
        dl.wanted_feats = ['f%d' % i for i in range(dl.num_feats)]
        dl.wanted_labels = []
        dl.fold = None
        return dl

    def nn_variables(self, model):
        return {v.name: v for v in model.graph.variables if v.name and v.name != 'global_step'}

    @staticmethod
    def variables(model):
        return {v.name: v for v in model.graph.variables if v.name}

    def set_variables(self, model, params):
        vs = self.variables(model)
        for k, a in params.items():
            vs[k].load(np.asarray(a, np.float64))

    def get_variables(self, model):
        return {k: v.numpy().astype(np.float64) for k, v in self.variables(model).items() if v.is_float}


def load_reference(dtype=None):
    """Returns a Reference, or None when the converted files are unavailable."""
    global _cached
    if _cached is None:
        sys.path.insert(0, HERE)
        try:
            import build_ref
        finally:
            sys.path.pop(0)
        if not build_ref.build():
            return None
        saved = {k: sys.modules.get(k) for k in ('tensorflow', 'matplotlib', 'matplotlib.pyplot', 'data_funcs', 'helper_funcs',
                                                  'generic_wrapper', 'multimodal_autoencoder')}
        tf = _load_module('mmae_tf1_shim', os.path.join(SHIM_DIR, 'tensorflow', '__init__.py'))
        try:
            sys.modules['tensorflow'] = tf
            try:
                import matplotlib.pyplot  # noqa: F401
            except Exception:
                mpl, plt = _stub_matplotlib()
                sys.modules['matplotlib'] = mpl
                sys.modules['matplotlib.pyplot'] = plt
            df = _load_module('mmae_ref_data_funcs', os.path.join(REF_DIR, 'data_funcs.py'))
            sys.modules['data_funcs'] = df
            mm = _load_module('mmae_ref_multimodal_autoencoder', os.path.join(REF_DIR, 'multimodal_autoencoder.py'))
            nn = gw = hf = None
            try:       # the plain MLP classifier (comparison_algorithms/neural_net.py) and what it imports
                sys.modules['helper_funcs'] = hf = _load_module('mmae_ref_helper_funcs', os.path.join(REF_DIR, 'helper_funcs.py'))
                sys.modules['generic_wrapper'] = gw = _load_module('mmae_ref_generic_wrapper', os.path.join(REF_DIR, 'generic_wrapper.py'))
                nn = _load_module('mmae_ref_neural_net', os.path.join(REF_DIR, 'neural_net.py'))
            except Exception as e:       # noqa: BLE001 -- the MMAE pinning does not depend on it
                print('oracle/_ref: neural_net.py not loadable:', repr(e)[:200])
            aw = acw = None
            try:       # the two grid-search drivers of the MMAE
                sys.modules['multimodal_autoencoder'] = mm
                aw = _load_module('mmae_ref_autoencoder_wrapper', os.path.join(REF_DIR, 'autoencoder_wrapper.py'))
                acw = _load_module('mmae_ref_autoencoder_classification_wrapper',
                                   os.path.join(REF_DIR, 'autoencoder_classification_wrapper.py'))
            except Exception as e:       # noqa: BLE001
                print('oracle/_ref: wrapper modules not loadable:', repr(e)[:200])
        finally:
            for k, v in saved.items():          # leave no fake `tensorflow` / `matplotlib` behind for other importers
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
        _cached = Reference(tf, mm, df, nn, gw, hf, aw, acw)
    if dtype is not None:
        _cached.tf.set_default_dtype(dtype)
    return _cached
