"""Host side of the block-mask noise (multimodal_autoencoder.py:649-702).

rng_mode='numpy': the descriptor (zero bitmap, modality bitmask) is drawn here from NumPy's
legacy global RandomState in exactly the reference's call order -- per row
``choice(F, size=int(F*.05))`` then ``multinomial(1, P)`` (or ``randint(0, M)`` x
num_modalities_to_drop) -- so that, under the same np.random.seed, the engine masks the same
cells the reference would.  The device applies the descriptor while loading X.
rng_mode='philox': the descriptor is drawn on the device (csrc/kernels.cuh noise_gen_kernel).
"""
from __future__ import annotations

import numpy as np

DEFAULT_NOISE_P = [0.64018104, 0.03168217, 0.25119437, 0.07694242]        # :202
DEFAULT_NOISE_TYPES = [[], ['call', 'sms', 'screen'], ['location'],
                       ['location', 'call', 'sms', 'screen']]              # :203-206


def type_masks_from_names(noise_types, modality_names):
    """Modality bitmask per noise type; names are looked up like :694 (ValueError if absent)."""
    masks = []
    for names in noise_types:
        m = 0
        for n in names:
            m |= 1 << modality_names.index(n)
        masks.append(m)
    return masks


def categorical_thresholds(p):
    """Cumulative P as uint32 thresholds: Philox word >= thr[k] -> class index > k."""
    cum = np.cumsum(np.asarray(p, np.float64))
    cum = cum / cum[-1]
    t = np.floor(cum[:-1] * 4294967296.0)
    return np.minimum(t, 4294967295).astype(np.uint64).astype(np.uint32)


def numpy_descriptor(batch, num_feats, n_mod, intelligent, noise_p=None, type_masks=None, num_drop=1,
                     override_mask=None, rng=np.random):
    """Descriptor for `batch` rows drawn in the reference's RNG order (one row at a time)."""
    zw = (num_feats + 31) // 32
    n_zero = int(num_feats * .05)
    zero_bits = np.zeros((batch, zw), np.uint32)
    mod_bits = np.zeros(batch, np.uint32)
    # The RNG calls stay one row at a time, in the reference's order; everything else is batched.  In the legacy
    # RandomState stream choice(F, size=k) IS randint(0, F, size=k) (choice's replace=True, p=None branch), which
    # skips choice's argument handling (13 -> 9 us per row here).
    cols_all = np.empty((batch, n_zero), np.int64)
    randint, multinomial = rng.randint, rng.multinomial
    if intelligent:
        ks = np.empty(batch, np.int64)
        for r in range(batch):
            cols_all[r] = randint(0, num_feats, n_zero)
            ks[r] = multinomial(1, noise_p).argmax()
        if override_mask is None:
            mod_bits[:] = np.asarray(type_masks, np.uint32)[ks]
        else:
            mod_bits[:] = override_mask                                                     # :691-692
    else:
        for r in range(batch):
            cols_all[r] = randint(0, num_feats, n_zero)
            m = 0
            for _ in range(num_drop):
                m |= 1 << int(randint(0, n_mod))
            mod_bits[r] = m
    if n_zero:
        flat = cols_all.ravel()
        rows = np.repeat(np.arange(batch), n_zero)
        np.bitwise_or.at(zero_bits, (rows, flat >> 5), (np.uint32(1) << (flat & 31).astype(np.uint32)))
    return zero_bits, mod_bits


def apply_descriptor_host(X, zero_bits, mod_bits, modality_starts, mask_with):
    """noisy_X on the host from a descriptor (used where the caller wants the NumPy array itself)."""
    out = np.array(X, dtype=np.float64, copy=True)
    cols = np.arange(out.shape[1])
    z = (zero_bits[:, cols >> 5] >> (cols & 31).astype(np.uint32)) & 1
    out[z.astype(bool)] = 0.0
    for m in range(len(modality_starts) - 1):
        rows = ((mod_bits >> np.uint32(m)) & 1).astype(bool)
        out[rows, modality_starts[m]:modality_starts[m + 1]] = mask_with
    return out
