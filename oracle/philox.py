"""Philox4x32-10 counter-based RNG restated in NumPy (TEST INFRASTRUCTURE).

The reference draws its noise on the host with NumPy's global Mersenne Twister and never seeds
it (priors.py:67-68, includes/utils.py:17-19), so there is nothing to reproduce bit-for-bit; the
B200 path replaces it with a counter-based generator.  This file restates
the published Philox4x32-10 algorithm (Salmon et al., SC'11; Random123 v1.14 ``philox.h``) and the
counter layout of ``csrc/reparam.cu`` so that the CUDA generator can be checked bit-exactly on the
integer stream and to fp32 tolerance on the transformed normals / Gumbels.

Counter layout (must match csrc/reparam.cu):
    counter = (global_row, column_block, step, stream)   key = (seed_lo, seed_hi)
    stream 0: Gaussian eps [B, L]  - block j covers columns 4j..4j+3 (two Box-Muller pairs)
    stream 1: uniform U for Gumbel [B, K] - block j covers columns 4j..4j+3
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over arrays of uint32 counters; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & MASK, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def u01(x):
    """uint32 -> (0,1): top 24 bits, centred: ((x >> 8) + 0.5) * 2^-24."""
    return ((np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float64) + 0.5) * (1.0 / 16777216.0)


def _blocks(rows, ncol, row_offset, step, stream, seed):
    nblk = (ncol + 3) // 4
    r = (np.arange(rows, dtype=np.uint64) + np.uint64(row_offset))[:, None]
    j = np.arange(nblk, dtype=np.uint64)[None, :]
    r, j = np.broadcast_arrays(r, j)
    return philox4x32_10(r, j, np.uint64(step), np.uint64(stream), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def normal(rows, ncol, seed, step, row_offset=0):
    """eps [rows, ncol] float64 ~ N(0,1): Box-Muller on (x0,x1) -> cols 4j,4j+1 and (x2,x3) -> 4j+2,4j+3."""
    x0, x1, x2, x3 = _blocks(rows, ncol, row_offset, step, 0, seed)
    out = np.empty((rows, x0.shape[1] * 4), np.float64)
    for a, b, col in ((x0, x1, 0), (x2, x3, 2)):
        rad = np.sqrt(-2.0 * np.log(u01(a)))
        ang = 2.0 * np.pi * u01(b)
        out[:, col::4] = rad * np.cos(ang)
        out[:, col + 1::4] = rad * np.sin(ang)
    return out[:, :ncol]


def uniform(rows, ncol, seed, step, row_offset=0):
    xs = _blocks(rows, ncol, row_offset, step, 1, seed)
    out = np.empty((rows, xs[0].shape[1] * 4), np.float64)
    for i, x in enumerate(xs):
        out[:, i::4] = u01(x)
    return out[:, :ncol]


def gumbel(rows, ncol, seed, step, row_offset=0, eps=1e-20):
    """includes/utils.py:17-19 applied to the Philox uniforms."""
    U = uniform(rows, ncol, seed, step, row_offset)
    return -np.log(eps - np.log(U + eps))
