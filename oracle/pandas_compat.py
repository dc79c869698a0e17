"""pandas APIs the reference's data_funcs.py calls and pandas removed since (DataFrame.from_csv, .as_matrix, .ix),
restored for the duration of a `with legacy_pandas():` block so that the reference's OWN loader code can run.
TEST INFRASTRUCTURE -- NOT PRODUCT CODE (see ref_loader.py); the product's loader is written for current pandas.

Semantics follow the pandas 0.2x documentation of the removed calls:
  * DataFrame.from_csv(path)  == read_csv(path, index_col=0, parse_dates=True)
  * df.as_matrix()            == df.values
  * df.ix[row, col]           -- label based, with positional fallback for an integer key on a non-integer axis
                                  (the reference writes `df.ix[row_label, column_position] = v`, data_funcs.py:754)
  * df.append(row, ignore_index=True) == concat([df, DataFrame([row])], ignore_index=True)   (generic_wrapper.py:276)
  * series[[i, j, ...]] = v   -- integer keys on a non-integer index are positions
                                  (the reference writes `xfill[missing_idxs] = Xbar[i, missing_idxs]`, data_funcs.py:338)
"""
from __future__ import annotations

import contextlib

import numpy as np
import pandas as pd


class _Ix:
    def __init__(self, df):
        self.df = df

    def _split(self, key):
        r, c = key if isinstance(key, tuple) else (key, slice(None))
        return r, c

    def _col(self, c):
        cols = self.df.columns
        if isinstance(c, (int, np.integer)) and not pd.api.types.is_integer_dtype(cols.dtype):
            return cols[c]                      # positional fallback
        return c

    def _row(self, r):
        idx = self.df.index
        if isinstance(r, (int, np.integer)) and not pd.api.types.is_integer_dtype(idx.dtype):
            return idx[r]
        return r

    def __getitem__(self, key):
        r, c = self._split(key)
        return self.df.loc[self._row(r), self._col(c)]

    def __setitem__(self, key, value):
        r, c = self._split(key)
        self.df.loc[self._row(r), self._col(c)] = value


def _from_csv(path, **kw):
    return pd.read_csv(path, index_col=0, parse_dates=True, **kw)


def _positional_setitem(orig):
    def setitem(self, key, value):
        ints = isinstance(key, (list, np.ndarray)) and len(key) > 0 and all(isinstance(k, (int, np.integer)) for k in key)
        if ints and not pd.api.types.is_integer_dtype(self.index.dtype):
            self.iloc[list(key)] = value
        else:
            orig(self, key, value)
    return setitem


@contextlib.contextmanager
def legacy_pandas():
    added = []
    series_setitem = pd.Series.__setitem__
    try:
        pd.Series.__setitem__ = _positional_setitem(series_setitem)
        if not hasattr(pd.DataFrame, 'from_csv'):
            pd.DataFrame.from_csv = staticmethod(_from_csv); added.append('from_csv')
        if not hasattr(pd.DataFrame, 'as_matrix'):
            pd.DataFrame.as_matrix = lambda self, columns=None: (self if columns is None else self[columns]).values
            added.append('as_matrix')
        if not hasattr(pd.DataFrame, 'append'):
            def _append(self, other, ignore_index=False):
                other = pd.DataFrame([other]) if isinstance(other, (dict, pd.Series)) else other
                return pd.concat([self, other], ignore_index=ignore_index)
            pd.DataFrame.append = _append; added.append('append')
        if not hasattr(pd.DataFrame, 'ix'):
            pd.DataFrame.ix = property(_Ix); added.append('ix')
        yield
    finally:
        pd.Series.__setitem__ = series_setitem
        for name in added:
            delattr(pd.DataFrame, name)
