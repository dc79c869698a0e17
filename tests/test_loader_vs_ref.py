"""The package's DataLoader against the reference's OWN DataLoader (oracle/_ref/data_funcs.py, converted mechanically
to Python 3 by oracle/build_ref.py) reading the same CSV files.  pandas removed DataFrame.from_csv / .as_matrix / .ix
long ago; oracle/pandas_compat.py restores them for the duration of the reference's calls.  Checked bit-for-bit:
normalisation (z-score / min-max on Train statistics), whole-modality gap filling with the reference's block rules,
fillna, the shuffle, label handling, Train / Val / Test and noisy / clean matrices, cross-validation fold assignment
(persisted into the CSV, as the reference does) and the per-fold matrices, modality discovery, batch sampling.

CPU only.  oracle/ is test infrastructure: nothing under multimodalautoencoder_b200/ imports it."""
import contextlib
import io
import os
import warnings

import numpy as np
import pandas as pd
import pytest

from multimodalautoencoder_b200.data_funcs import DataLoader
from oracle.pandas_compat import legacy_pandas
from oracle.ref_loader import load_reference

REF = load_reference()
pytestmark = pytest.mark.skipif(REF is None, reason='oracle/_ref unavailable')

MATS = ('train_X', 'val_X', 'test_X', 'train_Y', 'val_Y', 'test_Y', 'noisy_train_X', 'clean_train_X', 'noisy_val_X',
        'clean_val_X', 'noisy_val_Y', 'clean_val_Y', 'noisy_test_X', 'clean_test_X', 'noisy_test_Y', 'clean_test_Y')


def make_csv_frame(seed=0, n=90, colon=False):
    """A SNAPSHOT-shaped frame: user_id index, timestamp, prefixed feature columns with scattered NaNs and a few rows that
    miss a whole modality, two label columns (one NaN), dataset split, logistics_noisy."""
    rng = np.random.default_rng(seed)
    cols = {'timestamp': np.arange(n)}
    for p, w in (('phys', 6), ('call', 4), ('sms', 3), ('screen', 2), ('location', 5)):
        for j in range(w):
            name = 'phys_%s:f%d' % (('0-8H', '8-16H')[j // 3], j) if (colon and p == 'phys') else '%s_f%d' % (p, j)
            cols[name] = rng.normal(size=n) * 3 + 1
    df = pd.DataFrame(cols, index=pd.Index(['u%d' % i for i in range(n)], name='user_id'))
    feats = [c for c in df.columns if c != 'timestamp']
    df[feats] = df[feats].mask(rng.random((n, len(feats))) < 0.1)
    for r, p in ((3, 'call'), (7, 'location'), (9, 'phys'), (11, 'sms'), (12, 'screen'), (13, 'call')):
        df.loc[df.index[r], [c for c in feats if c.startswith(p)]] = np.nan
    df['happiness_label'] = rng.integers(0, 2, n).astype(float)
    df['health_label'] = rng.integers(0, 2, n).astype(float)
    df.loc[df.index[5], 'health_label'] = np.nan
    df['dataset'] = rng.choice(['Train', 'Val', 'Test'], n, p=[.6, .2, .2])
    df['logistics_noisy'] = rng.random(n) < 0.3
    return df


def load_both(df, tmp_path, seed=1, **kw):
    fr, fo = str(tmp_path / 'ref.csv'), str(tmp_path / 'ours.csv')
    df.to_csv(fr)
    df.to_csv(fo)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter('ignore')
        with legacy_pandas():
            np.random.seed(seed)
            ref = REF.data_funcs.DataLoader(fr, **kw)
        np.random.seed(seed)
        ours = DataLoader(fo, **kw)
    return ref, ours, fr, fo


def assert_same(ref, ours, names):
    for k in names:
        if not hasattr(ref, k):          # e.g. clean_train_Y / clean_val_Y outside cross validation: the reference keeps
            continue                     # them in locals (data_funcs.py:155-158); the package's extra attributes are harmless
        a, b = getattr(ref, k), getattr(ours, k, None)
        if a is None:
            assert b is None, k
            continue
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape and a.dtype == b.dtype, (k, a.shape, b.shape, a.dtype, b.dtype)
        assert np.array_equal(a, b), k


CONFIGS = {
    'zscore_gaps': dict(normalize_and_fill=True, normalization='z_score', fill_gaps_with=-1),
    'minmax': dict(normalize_and_fill=True, normalization='between_0_and_1'),
    'raw_unsupervised': dict(normalize_and_fill=True, normalization=None, fill_missing_with=-5, supervised=False),
    'one_label_signed': dict(normalize_and_fill=True, wanted_label='health_label', labels_to_sign=True),
    'cleaned_file': dict(normalize_and_fill=False),
}


@pytest.mark.parametrize('name', sorted(CONFIGS))
def test_loaded_matrices_are_bit_identical(name, tmp_path):
    kw = dict(CONFIGS[name], suppress_output=True)
    df = make_csv_frame()
    if name == 'cleaned_file':                       # the 'Data/Cleaned/' convention: already normalised and filled
        feats = [c for c in df.columns if '_f' in c]
        df[feats] = df[feats].fillna(0.0)
        df = df.dropna(subset=['health_label'])
    ref, ours, _, _ = load_both(df, tmp_path, **kw)
    assert ref.wanted_feats == ours.wanted_feats and ref.wanted_labels == ours.wanted_labels
    assert ref.modality_names == ours.modality_names and ref.modality_start_indices == ours.modality_start_indices
    assert ref.num_feats == ours.num_feats and ref.num_modalities == ours.num_modalities
    assert getattr(ref, 'num_labels', None) == ours.num_labels
    assert_same(ref, ours, MATS if kw.get('supervised', True) else MATS[:3])
    # the batch calls the train loops make (data_funcs.py:161-195): same np.random stream -> same rows
    for call, args in (('get_unsupervised_train_batch', (7,)), ('get_unsupervised_val_batch', (5,))) + \
            ((('get_supervised_train_batch', (6,)), ('get_supervised_val_batch', (4,))) if kw.get('supervised', True) else ()):
        np.random.seed(3)
        a = getattr(ref, call)(*args)
        np.random.seed(3)
        b = getattr(ours, call)(*args)
        for x, y in zip(a if isinstance(a, tuple) else (a,), b if isinstance(b, tuple) else (b,)):
            assert np.array_equal(np.asarray(x), np.asarray(y)), call


def test_gap_filling_keeps_the_reference_block_rules(tmp_path):
    """data_funcs.py:712-769 as it behaves, not as its docstring reads: the block walk starts at column position 2, so the
    first feature column of the first modality is not part of its fill range, and the last modality is never filled
    (nothing closes the last block)."""
    df = make_csv_frame()
    ref, ours, _, _ = load_both(df, tmp_path, normalize_and_fill=True, normalization=None, fill_gaps_with=-1,
                                fill_missing_with=0, suppress_output=True)
    for dl in (ref, ours):
        row = dl.df.loc['u9', dl.wanted_feats].to_numpy(dtype=float)             # misses all of phys
        s, e = dl.modality_start_indices[0], dl.modality_start_indices[1]
        assert row[s] == 0.0 and np.all(row[s + 1:e] == -1.0)
        row = dl.df.loc['u7', dl.wanted_feats].to_numpy(dtype=float)             # misses all of location (last block)
        assert np.all(row[dl.modality_start_indices[-2]:] == 0.0)
        row = dl.df.loc['u3', dl.wanted_feats].to_numpy(dtype=float)             # misses all of call
        assert np.all(row[dl.modality_start_indices[1]:dl.modality_start_indices[2]] == -1.0)
    assert_same(ref, ours, MATS)


def test_subdivided_physiology_modalities(tmp_path):
    """subdivide_physiology_features=True: 'phys_0-8H:...' / 'phys_8-16H:...' become modalities of their own (the time-of-day
    naming the reference's get_start_index expects, :659-694)."""
    df = make_csv_frame(seed=4, colon=True)
    ref, ours, _, _ = load_both(df, tmp_path, normalize_and_fill=True, subdivide_physiology_features=True,
                                suppress_output=True)
    assert ours.modality_names == ref.modality_names and len(ours.modality_names) == 6
    assert ours.modality_start_indices == ref.modality_start_indices
    assert_same(ref, ours, MATS)


def test_cross_validation_folds_and_their_persistence(tmp_path):
    """Folds are drawn row by row with np.random.randint(0, 5) (Test rows: -1), written back into the CSV (:212-224) and
    reused by every later loader of the file; every fold's matrices are the reference's."""
    df = make_csv_frame(seed=2)
    kw = dict(normalize_and_fill=True, cross_validation=True, fill_gaps_with=-1, suppress_output=True)
    ref, ours, fr, fo = load_both(df, tmp_path, **kw)
    assert open(fr).read() == open(fo).read() and 'logistics_cv_fold' in pd.read_csv(fo).columns
    for f in range(5):
        with legacy_pandas():
            ref.set_to_cross_validation_fold(f)
        ours.set_to_cross_validation_fold(f)
        assert ref.fold == ours.fold == f
        assert_same(ref, ours, ('train_X', 'train_Y', 'val_X', 'val_Y', 'noisy_train_X', 'clean_train_X', 'noisy_val_X',
                                'clean_val_X', 'noisy_val_Y', 'clean_val_Y'))
        base_X, _ = ours.cross_val_base()
        assert np.array_equal(base_X[ours.train_index], ours.train_X)
    # a second loader of the same file (the wrappers build an unsupervised and a supervised one) sees the same folds
    with contextlib.redirect_stdout(io.StringIO()):
        np.random.seed(77)
        again = DataLoader(fo, normalize_and_fill=False, cross_validation=True, supervised=False, suppress_output=True)
    a = again.df['logistics_cv_fold'].sort_index()
    b = ours.df['logistics_cv_fold'].sort_index()
    assert a.index.equals(b.index) or len(a) >= len(b)
    assert (a.loc[b.index] == b).all()
    # opting out leaves the input file alone
    f2 = str(tmp_path / 'untouched.csv')
    df.to_csv(f2)
    before = open(f2).read()
    with contextlib.redirect_stdout(io.StringIO()):
        DataLoader(f2, normalize_and_fill=False, cross_validation=True, supervised=False, suppress_output=True,
                   persist_folds=False)
    assert open(f2).read() == before


def test_null_column_removal_where_the_reference_crashes(tmp_path):
    """A feature that is entirely null in one split: the reference's remove_null_cols calls an undefined `dropCols`
    (data_funcs.py:619) and dies with NameError; the package does what that function's docstring says."""
    df = make_csv_frame(seed=5)
    df.loc[df['dataset'] == 'Val', 'sms_f2'] = np.nan
    fr = str(tmp_path / 'ref.csv')
    df.to_csv(fr)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()), legacy_pandas():
        warnings.simplefilter('ignore')
        with pytest.raises(NameError):
            REF.data_funcs.DataLoader(fr, normalize_and_fill=True, suppress_output=True)
    with contextlib.redirect_stdout(io.StringIO()):
        ours = DataLoader(fr, normalize_and_fill=True, suppress_output=True)
    assert 'sms_f2' not in ours.wanted_feats and ours.num_feats == 19
    assert ours.modality_start_indices == [0, 6, 10, 12, 14, 19]


def test_fill_df_with_reconstruction_touches_only_missing_blocks(tmp_path):
    """data_funcs.py:310-381: a modality block is replaced by the reconstruction iff its values sum to -width; every
    other cell keeps its bits.  The reference's per-row loop and the package's vectorised select give the same frame."""
    df = make_csv_frame(seed=6)
    ref, ours, _, _ = load_both(df, tmp_path, normalize_and_fill=True, normalization='between_0_and_1', fill_gaps_with=-1,
                                fill_missing_with=-1, suppress_output=True)
    Xbar = np.random.default_rng(0).random((len(ours.df), ours.num_feats))
    with contextlib.redirect_stdout(io.StringIO()):
        with legacy_pandas():
            a = ref.fill_df_with_reconstruction(ref.df.copy(), Xbar, plot_to_debug=False)
        b = ours.fill_df_with_reconstruction(ours.df.copy(), Xbar)
    A, B = a[ref.wanted_feats].to_numpy(), b[ours.wanted_feats].to_numpy()
    before = ours.df[ours.wanted_feats].to_numpy()
    assert np.array_equal(A, B)
    changed = A != before
    assert changed.any() and np.array_equal(A[changed], Xbar[changed])
    assert a.drop(columns=ref.wanted_feats).equals(b.drop(columns=ours.wanted_feats))
