"""MMAEWrapper / MMAEClassificationWrapper against the reference's OWN classes (oracle/_ref/autoencoder_wrapper.py,
autoencoder_classification_wrapper.py; converted mechanically to Python 3, removed pandas calls restored by
oracle/pandas_compat.py), both reading the same CSV files.  The model class is replaced on both sides by the same
recording stand-in, so what is compared is everything the drivers do around the model: the grid (108 / 576 settings),
the constructor keywords and train() arguments each fit receives, the fold protocol, the SVM scoring of embeddings and
every column of a results row.  CPU only; oracle/ is test infrastructure."""
import contextlib
import io
import os
import types
import warnings

import numpy as np
import pandas as pd
import pytest

from multimodalautoencoder_b200 import autoencoder_classification_wrapper as ours_acw
from multimodalautoencoder_b200 import autoencoder_wrapper as ours_aw
from oracle.pandas_compat import legacy_pandas
from oracle.ref_loader import load_reference

REF = load_reference()
pytestmark = pytest.mark.skipif(REF is None or REF.autoencoder_wrapper is None, reason='oracle/_ref wrappers unavailable')
LABELS = ['tomorrow_Group_Happiness_Evening_Label', 'tomorrow_Group_Health_Evening_Label',
          'tomorrow_Group_Calmness_Evening_Label']


def cleaned_frame(seed, n=160):
    """A 'Data/Cleaned/'-style file: normalised, filled, three label columns, split and noise bookkeeping."""
    rng = np.random.default_rng(seed)
    cols = {'timestamp': np.arange(n)}
    for p, w in (('phys', 6), ('call', 4), ('sms', 3), ('screen', 2), ('location', 5)):
        for j in range(w):
            cols['%s_f%d' % (p, j)] = rng.random(n)
    df = pd.DataFrame(cols, index=pd.Index(['u%d' % i for i in range(n)], name='user_id'))
    for k, lab in enumerate(LABELS):
        df[lab] = (df['phys_f%d' % k] + 0.3 * rng.random(n) > 0.65).astype(float)
    df['dataset'] = rng.choice(['Train', 'Val', 'Test'], n, p=[.6, .2, .2])
    df['logistics_noisy'] = rng.random(n) < 0.3
    return df


def make_fake(record):
    class FakeModel:
        """Stands in for MultimodalAutoencoder on both sides; every call is recorded."""
        noisy_X, tf_dropout_prob, embedding = 'noisy_X', 'keep', 'embedding'

        def __init__(self, **kw):
            self.kw = kw
            self.data_loader = kw['data_loader']
            self.val_loss = []
            self.session = self
            self._scale = 1.0 + 0.01 * sum(kw['layer_sizes']) / 1000.0 + (0.5 if kw['variational'] else 0.0) + kw['weight_penalty']
            record.append(('init', {k: (os.path.basename(os.path.normpath(v)) if k == 'checkpoint_dir' else v)
                                    for k, v in kw.items() if k not in ('data_loader', 'classification_data_loader')}))

        def train(self, num_steps, record_every_nth=None, save_every_nth=None):
            self.val_loss.append(1.0)
            record.append(('train', num_steps, float(record_every_nth), save_every_nth, len(self.data_loader.train_X)))

        def train_classification(self, num_steps, record_every_nth=None, save_every_nth=None):
            record.append(('train_classification', num_steps, float(record_every_nth), save_every_nth))

        def get_performance_on_data_with_noise(self, X):
            return float(np.mean(X)) * self._scale

        def set_classification_params(self, **kw):
            record.append(('set_classification_params', kw))

        def get_performance_on_data(self, X):
            record.append(('get_performance_on_data', len(X)))
            return float(np.mean(X)) * self._scale * 0.9

        def run(self, fetch, feed):                       # session.run(embedding, {noisy_X: X, keep: 1.0})
            X = np.asarray(feed[self.noisy_X], np.float64)
            return np.tanh(X[:, :8] * self._scale)

        def get_classification_predictions(self, X):
            X = np.asarray(X)
            n = getattr(self.kw.get('classification_data_loader'), 'num_labels', None)
            return (X[:, :3] * self._scale > 0.6).astype(float) if n else (X[:, 0] * self._scale > 0.6).astype(float)

        def close(self):
            pass
    return FakeModel


def quiet(fn, *a, **k):
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter('ignore')
        return fn(*a, **k)


def write_files(tmp_path):
    dirs = []
    for side in ('ref', 'ours'):
        d = str(tmp_path / side) + '/'
        os.makedirs(d + 'Data/Cleaned/')
        cleaned_frame(1).to_csv(d + 'Data/Cleaned/all_modalities_present.csv')
        cleaned_frame(2).to_csv(d + 'Data/Cleaned/modalities_missing.csv')
        dirs.append(d)
    return dirs


def as_set(settings):
    return {str(sorted(d.items(), key=str)) for d in settings}


def same_row(a, b):
    assert sorted(a) == sorted(b)
    for k in a:
        if isinstance(a[k], (float, np.floating)):
            assert np.array_equal(np.float64(a[k]), np.float64(b[k]), equal_nan=True), (k, a[k], b[k])
        else:
            assert a[k] == b[k], (k, a[k], b[k])


def test_mmae_wrapper_grid_protocol_and_result_row(tmp_path, monkeypatch):
    rdir, odir = write_files(tmp_path)
    rec_r, rec_o = [], []
    monkeypatch.setattr(REF.autoencoder_wrapper, 'mmae', types.SimpleNamespace(MultimodalAutoencoder=make_fake(rec_r)))
    monkeypatch.setattr(ours_aw, 'mmae', types.SimpleNamespace(MultimodalAutoencoder=make_fake(rec_o)))
    with legacy_pandas():
        np.random.seed(4)
        r = quiet(REF.autoencoder_wrapper.MMAEWrapper, 'all_modalities_present.csv', dropbox_path=rdir, num_cross_folds=2)
    np.random.seed(4)
    o = quiet(ours_aw.MMAEWrapper, 'all_modalities_present.csv', dropbox_path=odir, num_cross_folds=2)
    # the grid: 72 plain settings + 36 variational ones
    assert r.num_settings == o.num_settings == 108
    assert as_set(r.list_of_param_settings) == as_set(o.list_of_param_settings)
    for a in ('loss_func', 'learning_rate', 'clip_gradients', 'normalization', 'mask_with', 'fill_missing', 'decay',
              'decay_steps', 'decay_rate', 'batch_size', 'num_steps', 'optimize_for', 'min_or_max', 'save_prefix',
              'classifier_name', 'check_test'):
        assert getattr(r, a) == getattr(o, a), a
    assert o.results_path[len(odir):] == r.results_path[len(rdir):]
    # both loaders read the files the same way (folds drawn by the same np.random stream and persisted)
    for name in ('data_loader', 'classification_data_loader'):
        a, b = getattr(r, name), getattr(o, name)
        assert np.array_equal(a.train_X, b.train_X) and np.array_equal(a.val_X, b.val_X) and np.array_equal(a.test_X, b.test_X)
    assert np.array_equal(r.classification_data_loader.train_Y, o.classification_data_loader.train_Y)
    # one setting through the cross-validation protocol: fits, folds, SVM scoring, result columns
    setting = {'architecture': [300, 100], 'tie_weights': False, 'dropout_prob': 0.5, 'weight_penalty': .001,
               'weight_initialization': 'normal', 'activation_function': 'softsign', 'variational': True}
    r.num_steps = o.num_steps = 50
    with legacy_pandas():
        row_r = quiet(r.get_cross_validation_results, dict(setting))
    row_o = quiet(o.get_cross_validation_results, dict(setting))
    same_row(row_r, row_o)
    assert len(row_o) == len(setting) + 1 + 6 * 3 + 6 and 'val_sigmoid_cross_entropy' in row_o
    # what each fit was given
    assert [e[0] for e in rec_r] == [e[0] for e in rec_o] == ['init', 'train'] * 2
    for (_, kr), (_, ko) in zip(rec_r[0::2], rec_o[0::2]):
        extra = set(ko) - set(kr)
        assert extra <= {'rng_mode', 'precision', 'seed', 'device'}           # keyword-only additions of the package
        assert {k: ko[k] for k in kr} == kr
    assert rec_r[1::2] == rec_o[1::2]
    # the held-out test protocol
    with legacy_pandas():
        t_r = quiet(r.test_on_test, dict(setting))
    t_o = quiet(o.test_on_test, dict(setting))
    assert t_r is None and isinstance(t_o, float)                 # the reference only prints the test loss
    assert rec_r[-1] == rec_o[-1] == ('get_performance_on_data', len(o.data_loader.test_X))


@pytest.mark.parametrize('wanted_label', [None, LABELS[1]])
def test_mmae_classification_wrapper_grid_protocol_and_result_row(tmp_path, monkeypatch, wanted_label):
    rdir, odir = write_files(tmp_path)
    rec_r, rec_o = [], []
    monkeypatch.setattr(REF.autoencoder_classification_wrapper, 'mmae',
                        types.SimpleNamespace(MultimodalAutoencoder=make_fake(rec_r)))
    monkeypatch.setattr(ours_acw, 'mmae', types.SimpleNamespace(MultimodalAutoencoder=make_fake(rec_o)))
    kw = dict(num_cross_folds=2, wanted_label=wanted_label)
    with legacy_pandas():
        np.random.seed(9)
        r = quiet(REF.autoencoder_classification_wrapper.MMAEClassificationWrapper, 'all_modalities_present.csv',
                  'modalities_missing.csv', dropbox_path=rdir, **kw)
    np.random.seed(9)
    o = quiet(ours_acw.MMAEClassificationWrapper, 'all_modalities_present.csv', 'modalities_missing.csv', dropbox_path=odir,
              **kw)
    assert r.num_settings == o.num_settings == 576
    assert as_set(r.list_of_param_settings) == as_set(o.list_of_param_settings)
    for a in ('mmae_loss_func', 'mmae_learning_rate', 'mmae_num_steps', 'mmae_batch_size', 'classification_learning_rate',
              'classification_num_steps', 'classification_batch_size', 'optimize_for', 'min_or_max', 'save_prefix',
              'classifier_name', 'check_noisy_data'):
        assert getattr(r, a) == getattr(o, a), a
    a, b = r.classification_data_loader, o.classification_data_loader
    assert np.array_equal(a.train_X, b.train_X) and np.array_equal(a.train_Y, b.train_Y) and a.num_labels == b.num_labels
    setting = dict(r.list_of_param_settings[17])
    r.mmae_num_steps = o.mmae_num_steps = 40
    r.classification_num_steps = o.classification_num_steps = 30
    with legacy_pandas():
        row_r = quiet(r.get_cross_validation_results, dict(setting))
    row_o = quiet(o.get_cross_validation_results, dict(setting))
    same_row(row_r, row_o)
    assert [e[0] for e in rec_r] == [e[0] for e in rec_o] == \
        ['init', 'set_classification_params', 'train', 'train_classification'] * 2
    for (_, kr), (_, ko) in zip(rec_r[0::4], rec_o[0::4]):
        assert set(ko) - set(kr) <= {'rng_mode', 'precision', 'seed', 'device'}
        assert {k: ko[k] for k in kr} == kr
    for i in (1, 2, 3):
        assert rec_r[i::4] == rec_o[i::4]


@pytest.mark.skipif(REF is None or REF.neural_net is None, reason='oracle/_ref/neural_net.py unavailable')
def test_nn_wrapper_grid_protocol_and_result_row(tmp_path, monkeypatch):
    """NNWrapper (comparison_algorithms/neural_net.py:407-631) with the network replaced by the same stand-in."""
    from multimodalautoencoder_b200 import neural_net as ours_nn
    rdir, odir = write_files(tmp_path)
    rec_r, rec_o = [], []

    def make_net(record):
        class FakeNet:
            def __init__(self, **kw):
                self.dl = kw['data_loader']
                self._scale = 1.0 + 0.001 * sum(kw['layer_sizes']) + kw['weight_penalty'] + 0.1 * kw['dropout_prob']
                record.append(('init', {k: v for k, v in kw.items()
                                        if k in ('layer_sizes', 'batch_size', 'learning_rate', 'dropout_prob', 'weight_penalty', 'verbose')}))

            def train(self, num_steps, output_every_nth):
                record.append(('train', num_steps, output_every_nth, len(self.dl.train_X)))

            def predict(self, X):
                return (np.asarray(X)[:, :3] * self._scale > 0.7).astype(float)

            def close(self):
                pass
        return FakeNet
    monkeypatch.setattr(REF.neural_net, 'NeuralNetwork', make_net(rec_r))
    monkeypatch.setattr(ours_nn, 'NeuralNetwork', make_net(rec_o))
    with legacy_pandas():
        np.random.seed(6)
        r = quiet(REF.neural_net.NNWrapper, 'modalities_missing.csv', dropbox_path=rdir, num_cross_folds=3)
    np.random.seed(6)
    o = quiet(ours_nn.NNWrapper, 'modalities_missing.csv', dropbox_path=odir, num_cross_folds=3)
    assert r.num_settings == o.num_settings == 32 and as_set(r.list_of_param_settings) == as_set(o.list_of_param_settings)
    for a in ('num_steps', 'output_every_nth', 'optimize_for', 'min_or_max', 'check_test', 'check_noisy_data', 'save_prefix',
              'classifier_name', 'normalization', 'normalize_and_fill'):
        assert getattr(r, a) == getattr(o, a), a
    assert np.array_equal(r.data_loader.train_X, o.data_loader.train_X)
    setting = dict(r.list_of_param_settings[5])
    with legacy_pandas():
        row_r = quiet(r.get_cross_validation_results, dict(setting))
    row_o = quiet(o.get_cross_validation_results, dict(setting))
    same_row(row_r, row_o)
    assert rec_r == rec_o and [e[0] for e in rec_o] == ['init', 'train'] * 3
