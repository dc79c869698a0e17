"""CPU tests of the host layer: data loader, noise descriptors, grid enumeration, sharding, C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import mmae_oracle as O
from multimodalautoencoder_b200 import _capi
from multimodalautoencoder_b200.data_funcs import DataLoader
from multimodalautoencoder_b200.noise import (apply_descriptor_host, categorical_thresholds, numpy_descriptor,
                                              type_masks_from_names)
from multimodalautoencoder_b200.synthetic import make_frame, wide_blocks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported_and_bound():
    """Every function include/mmae_b200.h declares is exported by the built library and has a ctypes prototype."""
    hdr = open(os.path.join(ROOT, 'include', 'mmae_b200.h')).read()
    declared = set(re.findall(r'\b(mmae_[a-z0-9_]+)\s*\(', hdr))
    declared -= {'mmae_engine', 'mmae_config', 'mmae_outputs'}
    assert len(declared) >= 35
    assert declared == set(_capi.PROTOTYPES), declared ^ set(_capi.PROTOTYPES)
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _capi.load() is not None


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'multimodalautoencoder_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert 'oracle' not in src.replace('# oracle', ''), fn


def test_data_loader_contract():
    df = make_frame(800, seed=3)
    dl = DataLoader(df=df, supervised=True, cross_validation=False, normalize_and_fill=False, suppress_output=True)
    assert dl.modality_names == ['phys', 'call', 'sms', 'screen', 'location']
    assert dl.modality_start_indices == [0, 200, 220, 240, 270, 320]
    assert dl.num_feats == 320 and dl.num_modalities == 5 and dl.num_labels == 3
    assert dl.train_X.dtype == np.float64 and dl.train_Y.shape == (len(dl.train_X), 3)
    assert len(dl.train_X) + len(dl.val_X) + len(dl.test_X) == 800
    assert len(dl.noisy_val_X) + len(dl.clean_val_X) == len(dl.val_X)
    np.random.seed(0)
    idx = np.random.choice(len(dl.train_X), size=7)
    np.random.seed(0)
    assert np.array_equal(dl.get_unsupervised_train_batch(7), dl.train_X[idx])       # data_funcs.py:167
    one = DataLoader(df=df, supervised=True, wanted_label='health_label', normalize_and_fill=False, suppress_output=True)
    assert one.num_labels is None and one.train_Y.ndim == 1
    # cross-validation folds: Test rows are fold -1, every other row is in exactly one validation fold
    cv = DataLoader(df=df, supervised=False, cross_validation=True, normalize_and_fill=False, suppress_output=True)
    sizes = []
    for f in range(5):
        cv.set_to_cross_validation_fold(f)
        sizes.append(len(cv.val_X))
        # a fold is a row list into the base matrix that a device-resident pipeline uploads once
        bX, _ = cv.cross_val_base()
        assert np.array_equal(bX[cv.train_index], cv.train_X) and len(bX) == len(cv.train_X) + len(cv.val_X)
        assert len(cv.train_X) + len(cv.val_X) == int((df['dataset'] != 'Test').sum())
    assert sum(sizes) == int((df['dataset'] != 'Test').sum())


def test_missing_block_rule_and_fill():
    df = make_frame(300, seed=4)
    dl = DataLoader(df=df, supervised=False, normalize_and_fill=False, suppress_output=True)
    cfg = O.OracleConfig(num_feats=320, layer_sizes=[8], modality_starts=dl.modality_start_indices, modality_names=dl.modality_names)
    X = df[dl.wanted_feats].to_numpy()
    assert np.array_equal(dl.missing_modality_mask(X), O.missing_blocks(cfg, X))
    Xbar = np.full_like(X, 0.5)
    out = dl.fill_df_with_reconstruction(df, Xbar)[dl.wanted_feats].to_numpy()
    assert np.array_equal(out, O.fill_missing(cfg, X, Xbar))


@pytest.mark.parametrize('intelligent', [True, False])
def test_numpy_descriptor_matches_reference_noise(intelligent):
    """Host descriptor drawn in the reference's RNG order == the oracle's transliteration of add_noise_to_batch."""
    starts, names = [0, 200, 220, 240, 270, 320], ['phys', 'call', 'sms', 'screen', 'location']
    cfg = O.OracleConfig(num_feats=320, layer_sizes=[8], modality_starts=starts, modality_names=names,
                         intelligent_noise=intelligent, num_modalities_to_drop=2)
    X = np.random.default_rng(0).uniform(size=(64, 320))
    np.random.seed(42)
    want = O.add_noise(cfg, X)
    np.random.seed(42)
    zb, mb = numpy_descriptor(64, 320, 5, intelligent, cfg.noise_p, type_masks_from_names(cfg.noise_types, names), 2)
    got = apply_descriptor_host(X, zb, mb, starts, -1.0)
    assert np.array_equal(got, want)
    assert np.array_equal(got, O.noise_from_descriptor(cfg, X, zb, mb))


def test_thresholds_match_oracle_twin():
    from oracle import philox_host as PH
    p = [0.64018104, 0.03168217, 0.25119437, 0.07694242]
    assert np.array_equal(categorical_thresholds(p), PH.categorical_thresholds(p))


def _fake_loaders():
    df = make_frame(300, seed=5)
    dl = DataLoader(df=df, supervised=False, cross_validation=True, normalize_and_fill=False, suppress_output=True)
    cdl = DataLoader(df=df, supervised=True, cross_validation=True, normalize_and_fill=False, suppress_output=True)
    return dl, cdl


def test_grid_sizes_match_reference(tmp_path):
    """108 settings for MMAEWrapper, 576 for MMAEClassificationWrapper (counted by simulating the
    reference's enumeration)."""
    from multimodalautoencoder_b200.autoencoder_wrapper import MMAEWrapper
    from multimodalautoencoder_b200.autoencoder_classification_wrapper import MMAEClassificationWrapper
    dl, cdl = _fake_loaders()
    w = MMAEWrapper('synthetic.csv', dropbox_path=str(tmp_path) + '/', data_loader=dl, classification_data_loader=cdl)
    assert w.num_settings == 108
    assert sum(1 for s in w.list_of_param_settings if s['variational']) == 36
    keys = set(w.list_of_param_settings[0])
    assert keys == {'architecture', 'tie_weights', 'dropout_prob', 'weight_penalty', 'weight_initialization',
                    'activation_function', 'variational'}
    assert w.optimize_for == 'val_sigmoid_cross_entropy'
    c = MMAEClassificationWrapper('a.csv', 'b.csv', dropbox_path=str(tmp_path) + '/', data_loader=dl,
                                  classification_data_loader=cdl)
    assert c.num_settings == 576
    assert not any(s['variational'] and s['tie_weights'] for s in c.list_of_param_settings)


def test_grid_sharding_is_a_partition(tmp_path):
    from multimodalautoencoder_b200.autoencoder_wrapper import MMAEWrapper
    dl, cdl = _fake_loaders()
    seen = []
    for r in range(8):
        w = MMAEWrapper('synthetic.csv', dropbox_path=str(tmp_path) + '/', data_loader=dl, classification_data_loader=cdl,
                        shard=(r, 8))
        mine = w.my_settings()
        assert abs(len(mine) - 108 / 8) <= 1
        seen += [str(sorted(s.items(), key=str)) for s in mine]
    assert len(seen) == 108 and len(set(seen)) == 108


def test_wide_synthetic_shape():
    blocks = wide_blocks()
    assert len(blocks) == 16 and sum(w for _, w in blocks) == 4096
    assert {'call', 'sms', 'screen', 'location'} <= {n for n, _ in blocks}


def test_svm_scoring_uses_one_embedding_pass(tmp_path):
    """MMAEWrapper.test_embedding_classification_quality: the four row sets (train / val / clean val / noisy val) are
    embedded in ONE session.run and split back in order (the reference ran four, autoencoder_wrapper.py:212-226)."""
    from multimodalautoencoder_b200.autoencoder_wrapper import MMAEWrapper
    dl, cdl = _fake_loaders()
    w = MMAEWrapper('synthetic.csv', dropbox_path=str(tmp_path) + '/', data_loader=dl, classification_data_loader=cdl)
    calls = []

    class FakeSession:
        def run(self, fetch, feed):
            X = feed['noisy_X']
            calls.append(len(X))
            return np.asarray(X)[:, :6] * 2.0                     # a row-wise "embedding"

    class FakeModel:
        val_loss = [1.0]
        embedding, noisy_X, tf_dropout_prob = 'embedding', 'noisy_X', 'keep'
        session = FakeSession()

    w.model = FakeModel()
    seen = {}
    orig = w.svm_pred_best_result

    def spy(svm_model, X, Y, label, best_acc, best_auc):
        seen[len(X)] = X
        return orig(svm_model, X, Y, label, best_acc, best_auc)
    w.svm_pred_best_result = spy
    res = w.test_embedding_classification_quality()
    assert len(calls) == 1 and calls[0] == len(cdl.train_X) + len(cdl.val_X) + len(cdl.clean_val_X) + len(cdl.noisy_val_X)
    assert len(res) == 6 and all(r.shape == (1, 3) for r in res)
    for part in (cdl.val_X, cdl.noisy_val_X, cdl.clean_val_X):     # each SVM sees exactly its own rows' embeddings
        assert np.allclose(seen[len(part)], np.asarray(part)[:, :6] * 2.0)


def test_nn_wrapper_grid_and_fold_protocol(tmp_path, monkeypatch):
    """NNWrapper (comparison_algorithms/neural_net.py:407-631): 4 x 2 x 4 x 1 x 1 = 32 settings by default; one setting =
    one fresh NeuralNetwork per fold (the previous engine closed first), trained num_steps, predicted on that fold's
    validation rows (+ noisy / clean subsets), per-label columns in the result row; test_on_test scores TEST rows."""
    from multimodalautoencoder_b200 import neural_net as nn
    _, cdl = _fake_loaders()
    w = nn.NNWrapper('synthetic.csv', dropbox_path=str(tmp_path) + '/', data_loader=cdl)
    assert w.num_settings == 32
    assert set(w.list_of_param_settings[0]) == {'architecture', 'dropout_prob', 'weight_penalty', 'learning_rate',
                                                'batch_size'}
    assert (w.classifier_name, w.optimize_for, w.min_or_max, w.check_test, w.check_noisy_data) == \
        ('NN', 'val_acc', 'max', True, True)

    log = []

    class FakeNet:
        def __init__(self, **kw):
            self.kw, self.closed = kw, False
            log.append(('new', kw['layer_sizes'], kw['batch_size'], len(kw['data_loader'].val_X)))

        def train(self, num_steps, output_every_nth):
            log.append(('train', num_steps, output_every_nth))

        def predict(self, X):
            log.append(('predict', len(X)))
            return (np.asarray(X)[:, :cdl.num_labels] > 0.5).astype(np.float32)

        def close(self):
            self.closed = True
            log.append(('close',))

    monkeypatch.setattr(nn, 'NeuralNetwork', FakeNet)
    w = nn.NNWrapper('synthetic.csv', layer_sizes=[[8, 4]], dropout_probs=[1.0], weight_penalties=[0.0, .01],
                     num_steps=7, num_cross_folds=3, dropbox_path=str(tmp_path) + '/', data_loader=cdl)
    assert w.num_settings == 2
    row = w.get_cross_validation_results(dict(w.list_of_param_settings[0]))
    assert [e[0] for e in log].count('new') == 3 and [e[0] for e in log].count('close') == 2
    assert [e for e in log if e[0] == 'train'] == [('train', 7, 5001)] * 3
    assert [e[0] for e in log].count('predict') == 9            # val + noisy + clean per fold
    n = cdl.num_labels
    want = {'val_acc', 'val_auc', 'val_f1', 'val_precision', 'val_recall', 'noisy_val_acc', 'noisy_val_auc',
            'clean_val_acc', 'clean_val_auc'}
    for label in nn.LABELS_TO_PREDICT[:n]:
        want |= {p + label for p in ('val_acc_', 'val_auc_', 'noisy_val_acc_', 'noisy_val_auc_', 'clean_val_acc_',
                                     'clean_val_auc_')}
    assert want <= set(row) and 0.0 <= row['val_acc'] <= 1.0
    del log[:]
    preds = w.test_on_test(w.convert_param_dict_for_use({'architecture': '[8, 4]', 'batch_size': '100',
                                                          'learning_rate': .001, 'dropout_prob': 1.0,
                                                          'weight_penalty': 0.0}))
    assert len(preds) == len(cdl.test_X) and log[1][1:3] == ([8, 4], 100)


def test_sweep_resumes_from_results_file(tmp_path, monkeypatch):
    """cont=True (generic_wrapper.py:104-120 in the reference): settings already in the results CSV are skipped, the
    new ones appended to the same file."""
    from multimodalautoencoder_b200 import neural_net as nn
    _, cdl = _fake_loaders()
    trained = []

    class FakeNet:
        def __init__(self, **kw):
            trained.append((tuple(kw['layer_sizes']), kw['weight_penalty']))

        def train(self, num_steps, output_every_nth):
            pass

        def predict(self, X):
            return (np.asarray(X)[:, :cdl.num_labels] > 0.5).astype(np.float32)

        def close(self):
            pass

    monkeypatch.setattr(nn, 'NeuralNetwork', FakeNet)
    kw = dict(layer_sizes=[[8, 4]], dropout_probs=[1.0], num_steps=1, num_cross_folds=2, check_test=False,
              check_noisy_data=False, dropbox_path=str(tmp_path) + '/', data_loader=cdl)
    w = nn.NNWrapper('synthetic.csv', weight_penalties=[0.0, .01], **kw)
    w.run()
    assert len(w.val_results_df) == 2 and len(trained) == 4
    del trained[:]
    w2 = nn.NNWrapper('synthetic.csv', weight_penalties=[0.0, .01, .001], cont=True, **kw)
    assert w2.save_prefix == w.save_prefix and len(w2.val_results_df) == 2
    w2.run()
    assert len(w2.val_results_df) == 3
    assert sorted(set(trained)) == [((8, 4), .001)]
