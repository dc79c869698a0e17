"""The package's grid-search drivers (generic_wrapper.py) against the reference's OWN Wrapper / ClassificationWrapper
classes (oracle/_ref/generic_wrapper.py, converted mechanically to Python 3; removed pandas calls restored by
oracle/pandas_compat.py).  Both run the same sweep with the fit replaced by a deterministic function of the setting
and the fold: the set of settings enumerated, the rows and columns of the results CSV, the best setting, the save
prefix and the classification metrics must agree.  CPU only; oracle/ is test infrastructure."""
import contextlib
import io
import os

import numpy as np
import pandas as pd
import pytest

from multimodalautoencoder_b200 import generic_wrapper as ours_gw
from multimodalautoencoder_b200 import helper_funcs as ours_helper
from oracle.pandas_compat import legacy_pandas
from oracle.ref_loader import load_reference

REF = load_reference()
pytestmark = pytest.mark.skipif(REF is None or REF.generic_wrapper is None, reason='oracle/_ref/generic_wrapper.py unavailable')

PARAMS = {'architecture': [[300, 100], [128, 64], [1000, 100]], 'dropout_prob': [1.0, 0.5], 'weight_penalty': [0.0, .001, .01],
          'activation_function': ['softsign', 'relu'], 'tie_weights': [True, False]}


def score_of(param_dict, fold):
    """Deterministic stand-in for a fit: depends on every hyper-parameter and on the fold."""
    a = param_dict['architecture']
    return (sum(a) % 97) / 97.0 + param_dict['dropout_prob'] * 0.37 + param_dict['weight_penalty'] * 11 + \
        (0.05 if param_dict['activation_function'] == 'relu' else 0.0) + (0.021 if param_dict['tie_weights'] else 0.0) + fold * 1e-3


class FoldOnlyLoader:
    fold = 0

    def set_to_cross_validation_fold(self, f):
        self.fold = f


def make_sweeper(base):
    class Sweeper(base):
        def define_params(self):
            self.params = {k: list(v) for k, v in PARAMS.items()}

        def load_data(self):
            self.data_loader = FoldOnlyLoader()

        def train_and_predict(self, param_dict):
            return score_of(param_dict, self.data_loader.fold)

        def test_on_test(self, param_dict):
            return score_of(param_dict, 99)
    return Sweeper


def canon(df, keys):
    df = df.copy()
    for k in keys:
        df[k] = df[k].map(str)
    return df.sort_values(list(keys)).reset_index(drop=True)[sorted(df.columns)]


def run_quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_sweep_rows_best_setting_and_resume_match_the_reference(tmp_path):
    rdir, odir = str(tmp_path / 'ref') + '/', str(tmp_path / 'ours') + '/'
    for d in (rdir, odir):
        os.makedirs(d + 'Results/MMAE/')
    kw = dict(num_cross_folds=3, optimize_for='val_score', min_or_max='max', save_results_every_nth=5)
    with legacy_pandas():
        r = run_quiet(make_sweeper(REF.generic_wrapper.Wrapper), 'snapshot.csv', dropbox_path=rdir, **kw)
        run_quiet(r.run)
    o = run_quiet(make_sweeper(ours_gw.Wrapper), 'snapshot.csv', dropbox_path=odir, **kw)
    run_quiet(o.run)
    keys = sorted(PARAMS)
    # 1. the enumeration: every combination exactly once, on both sides (the reference's pop/recurse walk, :151-187)
    assert r.num_settings == o.num_settings == 3 * 2 * 3 * 2 * 2
    as_set = lambda lst: {str(sorted(d.items(), key=str)) for d in lst}        # noqa: E731
    assert as_set([{k: d[k] for k in keys} for d in r.list_of_param_settings]) == \
        as_set([{k: d[k] for k in keys} for d in o.list_of_param_settings])
    # 2. file name and contents of the results CSV (row order is the enumeration's and is not a contract)
    assert r.save_prefix == o.save_prefix == 'MMAE-snapshot'
    rf, of = rdir + 'Results/MMAE/MMAE-snapshot.csv', odir + 'Results/MMAE/MMAE-snapshot.csv'
    a, b = pd.read_csv(rf, index_col=0), pd.read_csv(of, index_col=0)
    assert sorted(a.columns) == sorted(b.columns) == sorted(keys + ['val_score'])
    ca, cb = canon(a, keys), canon(b, keys)
    assert ca[keys].equals(cb[keys])
    assert np.array_equal(ca['val_score'].to_numpy(), cb['val_score'].to_numpy())       # np.mean over the same 3 folds
    # 3. the best setting
    with legacy_pandas():
        br = run_quiet(r.find_best_setting)
    bo = run_quiet(o.find_best_setting)
    assert {k: str(br[k]) for k in keys} == {k: str(bo[k]) for k in keys} and br['val_score'] == bo['val_score']
    with legacy_pandas():
        lo_r = run_quiet(r.find_best_setting, min_or_max='min')
    lo_o = run_quiet(o.find_best_setting, min_or_max='min')
    assert lo_r['val_score'] == lo_o['val_score'] < bo['val_score'] and str(lo_r['architecture']) == str(lo_o['architecture'])
    # 4. a second run without cont must not overwrite: prefix + '2' (:189-205)
    with legacy_pandas():
        r2 = run_quiet(make_sweeper(REF.generic_wrapper.Wrapper), 'snapshot.csv', dropbox_path=rdir, **kw)
    o2 = run_quiet(make_sweeper(ours_gw.Wrapper), 'snapshot.csv', dropbox_path=odir, **kw)
    assert r2.save_prefix == o2.save_prefix == 'MMAE-snapshot2'
    # 5. cont=True on a truncated results file: both finish exactly the missing settings
    a.iloc[:50].to_csv(rf)
    b.iloc[:50].to_csv(of)
    with legacy_pandas():
        r3 = run_quiet(make_sweeper(REF.generic_wrapper.Wrapper), 'snapshot.csv', dropbox_path=rdir, cont=True, **kw)
        assert r3.started_from == 50
        run_quiet(r3.sweep_all_parameters)
    o3 = run_quiet(make_sweeper(ours_gw.Wrapper), 'snapshot.csv', dropbox_path=odir, cont=True, **kw)
    assert o3.started_from == 50
    run_quiet(o3.sweep_all_parameters)
    a3, b3 = pd.read_csv(rf, index_col=0), pd.read_csv(of, index_col=0)
    assert len(a3) == len(b3) == 72
    assert canon(a3, keys)[keys].equals(canon(b3, keys)[keys])
    assert np.allclose(canon(a3, keys)['val_score'], canon(b3, keys)['val_score'], rtol=0, atol=0)
    # 6. values read back from the CSV become objects again
    row = dict(b3.iloc[0])
    with legacy_pandas():
        cr = run_quiet(r3.convert_param_dict_for_use, dict(row, batch_size='20'))
    co = o3.convert_param_dict_for_use(dict(row, batch_size='20'))
    assert cr['architecture'] == co['architecture'] and isinstance(co['architecture'], list)
    assert cr['batch_size'] == co['batch_size'] == 20


class LabelLoader(FoldOnlyLoader):
    """Five folds of fixed binary labels with noisy / clean subsets."""

    def __init__(self):
        rng = np.random.default_rng(5)
        self._Y = [(rng.random(40 + 3 * f) < 0.45).astype(float) for f in range(5)]
        self._X = [rng.random((40 + 3 * f, 6)) for f in range(5)]
        self.set_to_cross_validation_fold(0)

    def set_to_cross_validation_fold(self, f):
        self.fold = f
        self.val_X, self.val_Y = self._X[f], self._Y[f]
        n = len(self.val_Y) // 3
        self.noisy_val_X, self.noisy_val_Y = self.val_X[:n], self.val_Y[:n]
        self.clean_val_X, self.clean_val_Y = self.val_X[n:], self.val_Y[n:]


def make_classifier(base):
    class Clf(base):
        def define_params(self):
            self.params = {'C': [0.1, 1.0, 10.0], 'kernel': ['linear', 'rbf']}

        def load_data(self):
            self.data_loader = LabelLoader()

        def predict_on_data(self, X):
            return (np.asarray(X)[:, 0] * self._c > 0.4).astype(float)

        def train_and_predict(self, param_dict):
            self._c = {0.1: 0.7, 1.0: 1.0, 10.0: 1.4}[param_dict['C']] * (1.1 if param_dict['kernel'] == 'rbf' else 1.0)
            return self.predict_on_data(self.data_loader.val_X)
    return Clf


@pytest.mark.parametrize('noisy', [True, False])
def test_classification_sweep_metrics_match_the_reference(tmp_path, noisy):
    rdir, odir = str(tmp_path / 'ref') + '/', str(tmp_path / 'ours') + '/'
    for d in (rdir, odir):
        os.makedirs(d + 'Results/SVM/')
    kw = dict(wanted_label='tomorrow_Group_Happiness_Evening_Label', check_noisy_data=noisy, num_cross_folds=5)
    with legacy_pandas():
        r = run_quiet(make_classifier(REF.generic_wrapper.ClassificationWrapper), 'snapshot.csv', dropbox_path=rdir, **kw)
        run_quiet(r.sweep_all_parameters)
    o = run_quiet(make_classifier(ours_gw.ClassificationWrapper), 'snapshot.csv', dropbox_path=odir, **kw)
    run_quiet(o.sweep_all_parameters)
    assert r.save_prefix == o.save_prefix and r.save_prefix.startswith('SVM-snapshot-')
    a = pd.read_csv(rdir + 'Results/SVM/' + r.save_prefix + '.csv', index_col=0)
    b = pd.read_csv(odir + 'Results/SVM/' + o.save_prefix + '.csv', index_col=0)
    assert sorted(a.columns) == sorted(b.columns)
    want = {'val_acc', 'val_auc', 'val_f1', 'val_precision', 'val_recall'} | \
        ({'noisy_val_acc', 'noisy_val_auc', 'clean_val_acc', 'clean_val_auc'} if noisy else set())
    assert want <= set(b.columns)
    ca, cb = canon(a, ['C', 'kernel']), canon(b, ['C', 'kernel'])
    for c in sorted(want):
        assert np.array_equal(ca[c].to_numpy(), cb[c].to_numpy()), c


def test_metric_helpers_match_the_reference():
    rng = np.random.default_rng(11)
    for n in (1, 7, 200):
        y = (rng.random(n) < 0.4).astype(float)
        p = (rng.random(n) < 0.5).astype(float)
        with contextlib.redirect_stdout(io.StringIO()):
            a = REF.generic_wrapper.compute_all_classification_metrics(p, y)
            b = ours_gw.compute_all_classification_metrics(p, y)
        assert np.array_equal(np.asarray(a, float), np.asarray(b, float), equal_nan=True)     # n = 1: AUC is nan on both sides
        assert REF.generic_wrapper.get_baseline(y) == ours_gw.get_baseline(y)
        assert REF.generic_wrapper.binary_accuracy(y, p) == ours_gw.binary_accuracy(y, p)
    for secs in (0, 59, 60, 3599, 3600, 86399, 90061):
        # the reference divides ints with Python 2's `/` (floor); the converted file runs the same text under Python 3
        ref = tuple(int(v) for v in REF.helper_funcs.get_secs_mins_hours_from_secs(secs))
        assert ref == ours_helper.get_secs_mins_hours_from_secs(secs)
    for col in ('tomorrow_Group_Happiness_Evening_Label', 'tomorrow_Group_Health_Evening_Label',
                'tomorrow_Group_Calmness_Evening_Label', 'something_else_label', 'Happiness'):
        assert REF.helper_funcs.get_friendly_label_name(col) == ours_helper.get_friendly_label_name(col)
