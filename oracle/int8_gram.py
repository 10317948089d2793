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
