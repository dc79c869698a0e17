"""Data layer of the MMAE: CSV / DataFrame -> Train / Val / Test (or cross-validation fold) matrices,
modality discovery from column-name prefixes, random batch sampling, fill-in of missing blocks.

Host-side glue with the reference's public surface (reference: data_funcs.py): the class name,
constructor keywords, attributes (train_X, val_X, test_X, *_Y, num_feats, num_labels, wanted_feats,
wanted_labels, modality_names, modality_start_indices, num_modalities, fold, df, clean_/noisy_*) and
method names are the ones multimodal_autoencoder.py and the wrappers rely on.  Written for Python 3 and
pandas >= 2 (the reference used DataFrame.from_csv / .as_matrix / .ix, all removed since).

Differences, all opt-in or fixes:
  * DataLoader(df=...) accepts an in-memory DataFrame (synthetic data never touches the disk);
  * persist_folds=False opts out of the reference's rewrite of the input CSV after it assigns folds
    (data_funcs.py:220-222; the default keeps it: a second loader of the same file must see the same folds);
  * per-row Python loops of the reference are vectorised where that cannot change results.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

NUM_CROSS_VAL_FOLDS = 5
_NON_FEATURE_TOKENS = ('user_id', 'timestamp', 'label', 'Label', 'dataset', 'logistics', 'ppt_id')


# --------------------------------------------------------------------------- column bookkeeping
def get_wanted_feats_from_df(df):
    """Feature columns = every column whose name contains none of the bookkeeping tokens (data_funcs.py:461-467)."""
    return [c for c in df.columns.values if not any(tok in c for tok in _NON_FEATURE_TOKENS)]


def get_feat_prefix(feat_name, subdivide_phys=False):
    """Modality of a column: text before the first '_', or before ':' for phys* when subdividing (:676-694)."""
    prefix = feat_name[0:feat_name.find('_')]
    if subdivide_phys and prefix == 'phys':
        return feat_name[0:feat_name.find(':')]
    return prefix


def get_start_index(wanted_feats, modality):
    """Index of the first column of a modality (:659-674); columns of one modality are assumed contiguous."""
    sep = ':' if (modality[0:4] == 'phys' and 'H' in modality and modality != 'physTemp') else '_'
    for i, s in enumerate(wanted_feats):
        if modality + sep in s:
            return i
    return None


def get_modality_dict(wanted_feats, subdivide_phys=False):
    mods = {get_feat_prefix(f, subdivide_phys=subdivide_phys) for f in wanted_feats}
    return {m: get_start_index(wanted_feats, m) for m in mods}


def get_modality_names_indices(modality_dict):
    """Names and start indices, sorted by start index (:696-710)."""
    pairs = sorted(modality_dict.items(), key=lambda kv: kv[1])
    return [n for n, _ in pairs], [i for _, i in pairs]


def convert_matrix_tf_format(X):
    return np.asarray(X).astype(np.float64)           # the reference hands float64 to TensorFlow (:534-545)


def get_matrices_for_dataset(data_df, wanted_feats, wanted_labels, dataset=None, labels_to_sign=False):
    """(X, Y) of one split; Y is 1-D for a single label, 2-D otherwise, None when unsupervised (:494-532)."""
    part = data_df if dataset is None else data_df[data_df['dataset'] == dataset]
    X = convert_matrix_tf_format(part[wanted_feats].astype(float).to_numpy())
    if wanted_labels is None:
        return X, None
    if len(wanted_labels) == 1:
        y = np.asarray(part[wanted_labels[0]].tolist())
    else:
        y = np.asarray(part[wanted_labels].to_numpy())
    if labels_to_sign:
        y = 2 * y - 1
    return X, y


def get_matrix_for_dataset(data_df, wanted_feats, dataset):
    return get_matrices_for_dataset(data_df, wanted_feats, None, dataset)[0]


def remove_rows_with_no_label(data_df, wanted_labels, suppress_output=False):
    if wanted_labels is None:
        return data_df
    out = data_df.dropna(subset=wanted_labels, how='any')
    if not suppress_output:
        print("Rows with every wanted label present: %d of %d" % (len(out), len(data_df)))
    return out


# --------------------------------------------------------------------------- normalisation / filling
def normalize_columns(df, wanted_feats, normalization='z_score'):
    """Column-wise z-score or min-max using statistics of the Train split only (:547-572)."""
    df = df.copy()
    train = df[df['dataset'] == 'Train']
    for feat in wanted_feats:
        vals = train[feat].dropna().to_numpy(dtype=float)
        if normalization == 'z_score':
            df[feat] = (df[feat] - np.mean(vals)) / np.std(vals)
        else:
            lo, hi = vals.min(), vals.max()
            df[feat] = (df[feat] - lo) / (hi - lo)
    return df


def find_null_columns(df, features):
    return [f for f in features if len(df) == int(df[f].isnull().sum())]


def remove_null_cols(df, features):
    """Drops features that are entirely null in any of Train / Test / Val (:591-621)."""
    bad = []
    for split in ('Train', 'Test', 'Val'):
        for f in find_null_columns(df[df['dataset'] == split], features):
            if f not in bad:
                bad.append(f)
    if bad:
        print("Removing %d completely-null columns: %s" % (len(bad), bad))
        df = df.drop(columns=bad)
        features = [f for f in features if f not in bad]
    return df, features


def _gap_blocks(columns):
    """The column blocks the reference's row loop walks (:733-766), which depend on the column layout only:
    (first position, end position, feature columns counted) per run of equal feature prefixes.  Kept as the reference
    computes them: the first block starts at POSITION 2 whatever precedes it (the loop assumes two leading bookkeeping
    columns), a block's fill range is positional and so takes in any non-feature column that sits inside it, and the
    LAST block is never closed -- its rows are never filled."""
    skip = ('user_id', 'timestamp', 'logistics', 'label', 'dataset', 'Label')
    columns = list(columns)
    current, start, members, blocks = get_feat_prefix(columns[2], subdivide_phys=True), 2, [], []
    for pos, feat in enumerate(columns):
        if any(tok in feat for tok in skip):
            continue
        prefix = get_feat_prefix(feat, subdivide_phys=True)
        if prefix != current:
            if not members:
                raise ZeroDivisionError("float division by zero")     # the reference divides by the (empty) block's size
            blocks.append((start, pos, members))
            start, members = pos, []
        members.append(feat)
        current = prefix
    return blocks


def fill_gaps_in_modalities(df, fill_value, suppress_output=False, verbose=False):
    """Rows missing > 80 % of one modality get that modality's column range set to fill_value (:712-769) -- the
    reference's per-row, per-column Python loop as one vectorised pass per block, with its block rules (_gap_blocks)."""
    df = df.copy()
    filled = 0
    for start, end, members in _gap_blocks(df.columns.values):
        rows = (df[members].isnull().sum(axis=1) / float(len(members)) > 0.8).to_numpy()
        if rows.any():
            for pos in range(start, end):
                col = df.columns[pos]
                if not pd.api.types.is_float_dtype(df[col].dtype):
                    df[col] = df[col].astype(object)
                df.iloc[np.nonzero(rows)[0], pos] = fill_value
            filled += int(rows.sum())
            if verbose:
                print("Filled", get_feat_prefix(members[0], subdivide_phys=True), "(index", start, "to", end, ") in",
                      int(rows.sum()), "rows")
    if not suppress_output:
        print("Filled gaps in", filled, "rows with", fill_value)
    return df


def normalize_fill_df(data_df, wanted_feats, normalization='z_score', suppress_output=False, remove_cols=True,
                      fill_missing=0.0, fill_gaps=None):
    """normalise -> drop null columns -> fill whole-modality gaps -> fillna -> shuffle (:385-426)."""
    if normalization is not None:
        data_df = normalize_columns(data_df, wanted_feats, normalization)
    if remove_cols:
        data_df, wanted_feats = remove_null_cols(data_df, list(wanted_feats))
    if fill_gaps is not None:
        data_df = fill_gaps_in_modalities(data_df, fill_gaps, suppress_output=suppress_output)
    data_df = data_df.fillna(fill_missing)
    return data_df.sample(frac=1)


def assign_cv_fold(row, num_folds=NUM_CROSS_VAL_FOLDS):
    """-1 for Test rows, otherwise np.random.randint(0, 5) -- the literal 5 is the reference's (:623-635)."""
    return -1 if row['dataset'] == 'Test' else np.random.randint(0, 5)


# --------------------------------------------------------------------------- the loader
class DataLoader:
    def __init__(self, filename=None, supervised=True, suppress_output=False, cross_validation=False,
                 normalize_and_fill=True, normalization='between_0_and_1', fill_missing_with=0,
                 fill_gaps_with=None, extract_modalities=True, subdivide_physiology_features=False,
                 wanted_label=None, labels_to_sign=False, separate_noisy_data=True, df=None, persist_folds=True):
        self.filename = filename
        self.supervised = supervised
        self.normalize_and_fill = normalize_and_fill
        self.normalization = normalization
        self.cross_validation = cross_validation
        self.subdivide_phys = subdivide_physiology_features
        self.suppress_output = suppress_output
        self.extract_modalities = extract_modalities
        self.labels_to_sign = labels_to_sign
        self.fill_missing_with = fill_missing_with
        self.fill_gaps_with = fill_gaps_with
        self.persist_folds = persist_folds

        if df is not None:
            self.df = df.copy()
        elif filename is not None:
            self.df = pd.read_csv(filename, index_col=0)
        else:
            raise ValueError("DataLoader needs a filename or a DataFrame")
        self.separate_noisy_data = separate_noisy_data and 'logistics_noisy' in self.df.columns
        if cross_validation:
            self.df = self.assign_cross_val_folds(self.df)
            self.fold = 0
        self.wanted_feats = get_wanted_feats_from_df(self.df)

        if not supervised:
            self.wanted_labels = None
            self.num_labels = None
        elif wanted_label is not None:
            self.wanted_labels = [wanted_label]
            self.num_labels = None                     # -> 2 logits + sparse softmax in the MMAE (:324-327)
        else:
            self.wanted_labels = [c for c in self.df.columns.values if 'label' in c or 'Label' in c]
            self.num_labels = len(self.wanted_labels)
            if len(self.wanted_labels) == 1:
                self.num_classes = len(self.df[self.wanted_labels[0]].unique())
        self.df = remove_rows_with_no_label(self.df, self.wanted_labels, suppress_output=True)

        if normalize_and_fill:
            self.df = normalize_fill_df(self.df, self.wanted_feats, suppress_output=suppress_output, remove_cols=True,
                                        normalization=normalization, fill_missing=fill_missing_with,
                                        fill_gaps=fill_gaps_with)
            self.wanted_feats = [f for f in self.wanted_feats if f in self.df.columns]

        self.get_matrices_from_df()
        self.num_feats = self.get_feature_size()
        if extract_modalities:
            self.modality_dict = get_modality_dict(self.wanted_feats, subdivide_phys=self.subdivide_phys)
            self.modality_names, self.modality_start_indices = get_modality_names_indices(self.modality_dict)
            self.modality_start_indices.append(self.num_feats)
            self.num_modalities = len(self.modality_dict)
        if not suppress_output:
            print("%d train / %d val / %d test rows, %d features" % (len(self.train_X), len(self.val_X),
                                                                    len(self.test_X), self.num_feats))

    # ---- matrices
    def _split(self, frame, dataset):
        return get_matrices_for_dataset(frame, self.wanted_feats, self.wanted_labels, dataset,
                                        labels_to_sign=self.labels_to_sign)

    def get_matrices_from_df(self):
        self.test_X, self.test_Y = self._split(self.df, 'Test')
        if self.separate_noisy_data:
            (self.clean_test_X, self.clean_test_Y, self.noisy_test_X,
             self.noisy_test_Y) = self.get_noisy_clean_data_for_dataset('Test')
        if not self.cross_validation:
            self.train_X, self.train_Y = self._split(self.df, 'Train')
            self.val_X, self.val_Y = self._split(self.df, 'Val')
            if self.separate_noisy_data:
                (self.clean_train_X, self.clean_train_Y, self.noisy_train_X,
                 self.noisy_train_Y) = self.get_noisy_clean_data_for_dataset('Train')
                (self.clean_val_X, self.clean_val_Y, self.noisy_val_X,
                 self.noisy_val_Y) = self.get_noisy_clean_data_for_dataset('Val')
        else:
            self.set_to_cross_validation_fold(0)

    # ---- batch sampling: np.random.choice(n, size=B), with replacement (data_funcs.py:161-195)
    def get_unsupervised_train_batch(self, batch_size):
        return self.train_X[np.random.choice(len(self.train_X), size=batch_size)]

    def get_supervised_train_batch(self, batch_size):
        idx = np.random.choice(len(self.train_X), size=batch_size)
        return self.train_X[idx], self.train_Y[idx]

    def get_unsupervised_val_batch(self, batch_size):
        return self.val_X[np.random.choice(len(self.val_X), size=batch_size)]

    def get_supervised_val_batch(self, batch_size):
        idx = np.random.choice(len(self.val_X), size=batch_size)
        return self.val_X[idx], self.val_Y[idx]

    def get_val_data(self):
        return self.val_X, self.val_Y

    def get_feature_size(self):
        return np.shape(self.train_X)[1]

    # ---- cross validation
    def assign_cross_val_folds(self, df):
        if 'logistics_cv_fold' not in df.columns.values:
            df = df.copy()
            df['logistics_cv_fold'] = [assign_cv_fold(r) for _, r in df[['dataset']].iterrows()]
            if self.persist_folds and self.filename is not None:
                df.to_csv(self.filename)
        return df

    def _fold_frames(self, fold):
        cv = self.df['logistics_cv_fold']
        return self.df[(cv != fold) & (cv != -1)], self.df[cv == fold]

    def get_cross_val_data_for_fold(self, fold):
        train_df, val_df = self._fold_frames(fold)
        train_X, train_Y = self._split(train_df, None)
        val_X, val_Y = self._split(val_df, None)
        return train_X, train_Y, val_X, val_Y

    def cross_val_base(self):
        """(X, Y) of every row that takes part in cross validation (fold != -1), in DataFrame order: the matrix a
        device-resident pipeline uploads ONCE; each fold is then the row list `train_index` into it."""
        if getattr(self, '_cv_base', None) is None:
            cv = self.df['logistics_cv_fold']
            self._cv_base = self._split(self.df[cv != -1], None)
            self._cv_fold_of_row = cv[cv != -1].to_numpy()
        return self._cv_base

    def set_to_cross_validation_fold(self, fold):
        self.fold = fold
        self.train_X, self.train_Y, self.val_X, self.val_Y = self.get_cross_val_data_for_fold(fold)
        self.cross_val_base()
        # rows of the base matrix this fold trains on, in the order of train_X (train_X == base_X[train_index])
        self.train_index = np.nonzero(self._cv_fold_of_row != fold)[0].astype(np.int64)
        if self.separate_noisy_data:
            self.set_noisy_clean_data_for_fold(fold)

    # ---- noisy / clean partitions (logistics_noisy column)
    def get_noisy_or_clean_data_matrices(self, df, noisy=True):
        return self._split(df[df['logistics_noisy'] == noisy], None)

    def get_noisy_clean_data_for_dataset(self, dset):
        clean_X, clean_Y = self._split(self.df[self.df['logistics_noisy'] == False], dset)   # noqa: E712
        noisy_X, noisy_Y = self._split(self.df[self.df['logistics_noisy'] == True], dset)    # noqa: E712
        return clean_X, clean_Y, noisy_X, noisy_Y

    def set_noisy_clean_data_for_fold(self, fold):
        train_df, val_df = self._fold_frames(fold)
        self.noisy_train_X, self.noisy_train_Y = self.get_noisy_or_clean_data_matrices(train_df, True)
        self.clean_train_X, self.clean_train_Y = self.get_noisy_or_clean_data_matrices(train_df, False)
        self.noisy_val_X, self.noisy_val_Y = self.get_noisy_or_clean_data_matrices(val_df, True)
        self.clean_val_X, self.clean_val_Y = self.get_noisy_or_clean_data_matrices(val_df, False)

    # ---- fill-in (data_funcs.py:310-381)
    def missing_modality_mask(self, X):
        """[rows, M] bool: modality m of a row is missing iff sum(x[s:e]) == -(e-s) (:376-380)."""
        X = np.asarray(X, np.float64)
        out = np.zeros((X.shape[0], self.num_modalities), bool)
        for m in range(self.num_modalities):
            s, e = self.modality_start_indices[m], self.modality_start_indices[m + 1]
            out[:, m] = X[:, s:e].sum(axis=1) == -1 * (e - s)
        return out

    def find_missing_modalities_indices(self, x):
        miss = self.missing_modality_mask(np.asarray(x, np.float64)[None, :])[0]
        idx = []
        for m in np.nonzero(miss)[0]:
            idx.extend(range(self.modality_start_indices[m], self.modality_start_indices[m + 1]))
        return idx

    def fill_df_with_reconstruction(self, df, Xbar, plot_to_debug=False):
        """Reconstruction on missing modality blocks, original values elsewhere; returns the DataFrame."""
        X = df[self.wanted_feats].to_numpy(dtype=np.float64)
        miss = self.missing_modality_mask(X)
        cols = np.repeat(miss, np.diff(self.modality_start_indices), axis=1)
        filled = np.where(cols, np.asarray(Xbar, np.float64), X)
        df = df.copy()
        df.loc[:, self.wanted_feats] = filled
        n = int(miss.any(axis=1).sum())
        print("Filled %d rows with reconstruction (%.1f%%)" % (n, 100.0 * n / max(len(df), 1)))
        return df
