"""Synthetic SNAPSHOT-shaped data: the reference ships no data, so benches and tests
build frames with the column conventions its DataLoader expects -- modality-prefixed feature columns
(phys_/call_/sms_/screen_/location_), '*_label' columns, 'dataset' in {Train, Val, Test} and
'logistics_noisy' marking rows that genuinely miss whole modalities (filled with -1)."""
from __future__ import annotations

import numpy as np
import pandas as pd

SMALL_BLOCKS = [('phys', 200), ('call', 20), ('sms', 20), ('screen', 30), ('location', 50)]
LABELS = ['happiness_label', 'health_label', 'calmness_label']


def wide_blocks(n_blocks=16, width=256):
    names = ['call', 'sms', 'screen', 'location'] + ['phys%02d' % i for i in range(n_blocks - 4)]
    return [(n, width) for n in names[:n_blocks]]


def make_frame(n_rows=2000, blocks=SMALL_BLOCKS, n_labels=3, seed=1234, noisy_fraction=0.3, splits=(0.7, 0.15, 0.15)):
    rng = np.random.default_rng(seed)
    cols, starts = [], [0]
    for name, width in blocks:
        cols += ['%s_f%03d' % (name, i) for i in range(width)]
        starts.append(starts[-1] + width)
    X = rng.uniform(0.0, 1.0, (n_rows, starts[-1])).astype(np.float32).astype(np.float64)
    noisy = rng.uniform(size=n_rows) < noisy_fraction
    droppable = [i for i, (n, _) in enumerate(blocks) if n in ('call', 'sms', 'screen', 'location')] or list(range(len(blocks)))
    for r in np.nonzero(noisy)[0]:
        for m in rng.choice(droppable, size=rng.integers(1, len(droppable) + 1), replace=False):
            X[r, starts[m]:starts[m + 1]] = -1.0
    df = pd.DataFrame(X, columns=cols)
    df.insert(0, 'user_id', rng.integers(0, 200, n_rows))
    df.insert(1, 'timestamp', np.arange(n_rows))
    u = rng.uniform(size=n_rows)
    df['dataset'] = np.where(u < splits[0], 'Train', np.where(u < splits[0] + splits[1], 'Val', 'Test'))
    df['logistics_noisy'] = noisy
    for j in range(n_labels):
        df[LABELS[j] if j < len(LABELS) else 'extra%d_label' % j] = (rng.uniform(size=n_rows) < 0.5).astype(float)
    return df
