"""TEST INFRASTRUCTURE ONLY -- CPU statement of what `fsb_gram` returns on the FSB_GRAM_INT8 path
(fitsnap_b200/csrc/gram_i8.cu), in exact Python integers.

The reference forms  aw = w[:, None] * a,  bw = w * b  (solvers/svd.py:42-46) and hands them to LAPACK; the
transpose trick forms aw.T @ aw, aw.T @ bw (examples/library/transpose_trick/example.py:226-246).  The int8
tensor-core path computes that Gram as integer arithmetic; integer work has a bit-exact bar, so this module
restates it with unbounded integers and the GPU test compares bit for bit:

  per slab of rows (<= 2^18, the device's slab rule):
    m_c  = max_r |fl(w_r * a_rc)|              (augmented column k: fl(w_r * b_r))
    e_c  = 52 - ilogb(m_c)   (0 for an all-zero column; clamped to +-1000)
    q_rc = rint(fl(w_r * a_rc) * 2^e_c)        (|q| < 2^53: an exact integer)
    G'   = q^T q                               (exact; the device gets it through 16 residues + CRT)
    G_slab[i][j] = round_to_nearest_even(G'[i][j]) * 2^-(e_i + e_j)
  G = sum of the slabs in row order, in fp64.
"""
import math

import numpy as np

from oracle.linear_fit import weighted_system

SLAB_ROWS = 262144
BETA = 53


def slab_rows_for(n_rows):
    """Device rule (plan_i8): equal slabs of at most 2^18 rows, rounded up to 128."""
    n = max(int(n_rows), 1)
    nslab = -(-n // SLAB_ROWS)
    per = -(-n // nslab)
    return -(-per // 128) * 128


def _slab_gram(aug):
    k1 = aug.shape[1]
    m = np.abs(aug).max(axis=0) if aug.shape[0] else np.zeros(k1)
    e = np.zeros(k1, dtype=np.int64)
    for c in range(k1):
        if m[c] > 0.0:
            e[c] = min(1000, max(-1000, BETA - 1 - math.frexp(m[c])[1] + 1))   # ilogb(x) = frexp exponent - 1
    q = np.rint(aug * np.ldexp(1.0, e)[None, :])
    qi = np.array([[int(v) for v in row] for row in q], dtype=object).reshape(q.shape[0], k1)
    gint = qi.T.dot(qi) if q.shape[0] else np.zeros((k1, k1), dtype=object)
    out = np.empty((k1, k1))
    for i in range(k1):
        for j in range(k1):
            out[i, j] = math.ldexp(float(int(gint[i, j])), -int(e[i] + e[j]))    # int -> float rounds to nearest even
    return out


def quantised_gram(a, b, w, testing=None):
    """(k+1) x (k+1) augmented Gram exactly as the int8 path defines it (training rows only)."""
    a = np.asarray(a, dtype=np.float64)
    n, k = a.shape
    wv = np.asarray(w, dtype=np.float64).copy()
    if testing is not None:
        wv[np.asarray(testing, dtype=bool)] = 0.0        # the device masks by weight 0, rows stay in place
    aug = np.concatenate([wv[:, None] * a, (wv * np.asarray(b, dtype=np.float64))[:, None]], axis=1)
    step = slab_rows_for(n)
    total = None
    for r0 in range(0, max(n, 1), step):
        g = _slab_gram(aug[r0:r0 + step])
        total = g if total is None else total + g
    return total


# --------------------------------------------------------------------------------------------------
# The same definition, fast enough for BASELINE-sized matrices: the 53-bit integers q are cut into four
# 14-bit limbs; a limb-by-limb product summed over at most 2^18 rows stays below 2^46, so every one of the
# 16 limb Grams is EXACT in an fp64 dgemm.  Only the final recombination  G' = sum 2^(14(p+q)) P_pq  runs in
# Python integers (k^2 of them).  tests/test_oracle.py pins this function to `quantised_gram` bit for bit.
_LIMB_BITS = 14
_NLIMB = 4


def _slab_gram_fast(aug):
    k1 = aug.shape[1]
    n = aug.shape[0]
    assert n <= (1 << 18)
    m = np.abs(aug).max(axis=0) if n else np.zeros(k1)
    e = np.zeros(k1, dtype=np.int64)
    nz = m > 0.0
    e[nz] = np.clip(BETA - 1 - (np.frexp(m[nz])[1] - 1), -1000, 1000)
    q = np.rint(aug * np.ldexp(1.0, e)[None, :])                  # integer-valued, |q| < 2^53 (one value may be 2^53)
    sign = np.sign(q)
    mag = np.abs(q)
    limbs = []
    base = float(1 << _LIMB_BITS)
    for _ in range(_NLIMB):                                       # exact: mag is an integer below 2^54
        hi = np.floor(mag / base)
        limbs.append(sign * (mag - hi * base))
        mag = hi
    assert not mag.any()
    gint = np.zeros((k1, k1), dtype=object)
    for p in range(_NLIMB):
        for r in range(p, _NLIMB):
            blk = limbs[p].T @ limbs[r]                           # exact in fp64 (< 2^46)
            if r != p:
                blk = blk + blk.T
            gint += np.vectorize(int, otypes=[object])(blk) * (1 << (_LIMB_BITS * (p + r)))
    out = np.empty((k1, k1))
    for i in range(k1):
        ei = int(e[i])
        for j in range(k1):
            out[i, j] = math.ldexp(float(int(gint[i, j])), -(ei + int(e[j])))
    return out


def quantised_gram_fast(a, b, w, testing=None):
    """`quantised_gram`, through exact fp64 limb products (BLAS) instead of Python integers."""
    a = np.asarray(a, dtype=np.float64)
    n, k = a.shape
    wv = np.asarray(w, dtype=np.float64).copy()
    if testing is not None:
        wv[np.asarray(testing, dtype=bool)] = 0.0
    aug = np.concatenate([wv[:, None] * a, (wv * np.asarray(b, dtype=np.float64))[:, None]], axis=1)
    step = slab_rows_for(n)
    total = None
    for r0 in range(0, max(n, 1), step):
        g = _slab_gram_fast(aug[r0:r0 + step])
        total = g if total is None else total + g
    return total
