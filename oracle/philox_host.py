"""Host twin (NumPy, vectorised) of the engine's counter-based generator.
TEST INFRASTRUCTURE -- only tests/, smoke() and bench.py's checker legs import this.

The engine's rng_mode='philox' draws every random decision of the hot path
(batch rows, 5 % zero-noise columns, dropped modality blocks, dropout masks,
VAE epsilon) from Philox4x32-10 keyed by (seed) and countered by
(element index, stream tag, step) -- see csrc/philox.cuh, which this file must
match bit for bit.  The *distributions* are those of the reference's NumPy
legacy-RandomState calls (multimodal_autoencoder.py:682,689,699;
data_funcs.py:167); the *stream* is ours, because MT19937 + rejection sampling
is sequential and cannot be reproduced in parallel on a GPU (SURVEY.md section 7,
hard part 2).
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
U64 = np.uint64
M0, M1 = U64(0xD2511F53), U64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85

# stream tags (csrc/philox.cuh)
STREAM_BATCH = 1
STREAM_ZERO = 2
STREAM_MOD = 3
STREAM_EPS = 4
STREAM_DROP = 16          # + layer slot


def philox4x32(c0, c1, c2, c3, k0, k1):
    """10-round Philox-4x32.  All inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x, U32) for x in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0.astype(U64)
        p1 = M1 * c2.astype(U64)
        hi0, lo0 = (p0 >> U64(32)).astype(U32), p0.astype(U32)
        hi1, lo1 = (p1 >> U64(32)).astype(U32), p1.astype(U32)
        c0, c1, c2, c3 = hi1 ^ c1 ^ U32(k0), lo1, hi0 ^ c3 ^ U32(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def words(seed, stream, step, n_words, first_counter=0):
    """Flat uint32 stream: word j comes from counter (first_counter + j//4), lane j%4."""
    nc = (n_words + 3) // 4
    idx = np.arange(nc, dtype=U64) + U64(first_counter)
    w = philox4x32(idx.astype(U32), (idx >> U64(32)).astype(U32), U32(stream), U32(step),
                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(w, axis=1).reshape(-1)[:n_words]


def mulhi(w, n):
    return ((w.astype(U64) * U64(n)) >> U64(32)).astype(np.int64)


def batch_indices(seed, step, batch, n_rows):
    return mulhi(words(seed, STREAM_BATCH, step, batch), n_rows)


def noise_descriptor(seed, step, batch, num_feats, n_mod, n_zero, intelligent, thresholds=None,
                     type_masks=None, num_drop=1, row0=0):
    """(zero_bits [B, ceil(F/32)] uint32, mod_bits [B] uint32) for rows row0..row0+batch-1."""
    zw = (num_feats + 31) // 32
    q = (n_zero + 3) // 4                       # counters per row for the zero-noise stream
    zero_bits = np.zeros((batch, zw), U32)
    if n_zero > 0:
        w = words(seed, STREAM_ZERO, step, batch * q * 4, first_counter=row0 * q).reshape(batch, q * 4)
        cols = mulhi(w[:, :n_zero], num_feats)
        for j in range(n_zero):
            c = cols[:, j]
            np.bitwise_or.at(zero_bits, (np.arange(batch), c // 32), (U32(1) << (c % 32).astype(U32)))
    w = words(seed, STREAM_MOD, step, batch * 4, first_counter=row0).reshape(batch, 4)
    mod_bits = np.zeros(batch, U32)
    if intelligent:
        thr = np.asarray(thresholds, dtype=np.uint64)
        k = (w[:, 0].astype(np.uint64)[:, None] >= thr[None, :]).sum(axis=1)
        k = np.minimum(k, len(type_masks) - 1)
        mod_bits = np.asarray(type_masks, U32)[k]
    else:
        assert num_drop <= 4
        for d in range(num_drop):
            m = mulhi(w[:, d], n_mod)
            mod_bits |= (U32(1) << m.astype(U32))
    return zero_bits, mod_bits


def dropout_mask(seed, step, slot, rows, width, keep_thr, row0=0):
    """1.0 where kept.  Element e = row*width + col uses counter e//4, lane e%4."""
    e0 = row0 * width
    assert e0 % 4 == 0 or rows * width == 0 or width % 4 == 0
    first = e0 // 4
    off = e0 - first * 4
    w = words(seed, STREAM_DROP + slot, step, rows * width + off, first_counter=first)[off:]
    return ((w >> U32(8)) < U32(keep_thr)).astype(np.float64).reshape(rows, width)


def keep_threshold(keep):
    import math
    return min(int(math.ceil(float(keep) * (1 << 24))), 1 << 24)


def categorical_thresholds(p):
    """Cumulative probabilities as uint32 thresholds (first K-1 used; word >= thr[k] -> class > k)."""
    cum = np.cumsum(np.asarray(p, np.float64))
    cum = cum / cum[-1]
    t = np.floor(cum[:-1] * 4294967296.0)
    return np.minimum(t, 4294967295).astype(np.uint64).astype(np.uint32)
