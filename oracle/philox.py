"""Philox4x32-10 counter-based generator (Salmon et al., SC'11 "Parallel random numbers: as easy as 1, 2, 3"),
vectorised in numpy.  TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference seeds every Dropout layer with ``seed=self.seed`` (reference multinet.py:140) and lets TensorFlow
draw the masks; TensorFlow's stream cannot be reproduced without TensorFlow, so the dropout mask is *defined*
here, identically to ``deepimpute_b200/csrc/common.cuh``:

    element (b, j) of sub-network s at optimiser step t  (b = row in the batch, j = hidden unit)
    words  = philox4x32_10(counter = (j, b >> 2, s, t), key = (seed & 0xffffffff, seed >> 32))
    keep   = words[b & 3] >= floor(rate * 2**32)
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Ten rounds of Philox-4x32.  Counters: broadcastable uint32 arrays; keys: Python ints.  Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3)])
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def dropout_threshold(rate):
    """uint32 threshold; the rate is taken at float32 precision because that is what ``di_config`` carries."""
    return min(int(float(np.float32(rate)) * 4294967296.0), 0xFFFFFFFF)


def dropout_keep_mask(seed, subnet, step, n_rows, n_hidden, rate):
    """Boolean [n_rows, n_hidden] keep-mask of sub-network ``subnet`` at optimiser step ``step``."""
    if rate <= 0.0:
        return np.ones((n_rows, n_hidden), dtype=bool)
    b = np.arange(n_rows, dtype=np.uint64)[:, None]
    j = np.arange(n_hidden, dtype=np.uint64)[None, :]
    words = philox4x32_10(j, b >> np.uint64(2), np.uint64(subnet), np.uint64(step & 0xFFFFFFFF),
                          seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    sel = np.broadcast_to(b & np.uint64(3), words[0].shape)
    w = np.choose(sel.astype(np.int64), words)
    return w >= np.uint32(dropout_threshold(rate))
