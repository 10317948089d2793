"""Seeded synthetic linear systems shared by the golden generator and the tests
(SURVEY 8d generator).  numpy's default_rng streams are stable across versions, so the
systems are reproducible on the GPU box; tests verify that through stored checksums."""
import numpy as np

SOLVE_CASES = {
    # cond(w*A) ~ 1e3: parity on these alone would be vacuous ...
    "well": dict(seed=2024, n=4000, k=37, col_decades=3),
    # ... so: column scales over 5 decades + two nearly collinear columns, cond(w*A) ~ 5e5
    "ill": dict(seed=2025, n=6000, k=64, col_decades=5, collinear=1e-2),
    # cond(w*A) ~ 1e7-1e8: outside the normal-equation comfort zone, needs adaptive refinement
    "hard": dict(seed=2028, n=6000, k=48, col_decades=5, collinear=3e-4),
    # rank-deficient by construction (all-zero columns, as mixed twojmax produces)
    "zerocol": dict(seed=2026, n=3000, k=45, col_decades=3, zero_cols=3),
    # k + 1 > 128: several Gram super-tiles, several Cholesky panels
    "wide": dict(seed=2027, n=5000, k=150, col_decades=3),
}


def synth_system(seed, n, k, col_decades, collinear=0.0, zero_cols=0):
    """A = Z * column scales 10^U(-d,0), Z ~ N(0,1) (optionally with two nearly dependent
    columns, perturbation size `collinear`); b = A x_true + 1e-3 N(0,1);
    w in {1e-2, 1, 100} by row class 5/85/10 %; ~10 % test rows."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((n, k))
    if collinear:
        z[:, 1] = 0.7 * z[:, 0] + collinear * z[:, 1]
        z[:, k - 1] = z[:, k - 2] - 2.0 * z[:, 2] + collinear * z[:, k - 1]
    a = z * 10.0 ** rng.uniform(-col_decades, 0, k)
    if zero_cols:
        a[:, rng.choice(k, zero_cols, replace=False)] = 0.0
    x_true = rng.standard_normal(k)
    b = a @ x_true + 1e-3 * rng.standard_normal(n)
    cls = rng.choice(3, n, p=[0.05, 0.85, 0.10])
    w = np.array([1e-2, 1.0, 100.0])[cls]
    testing = rng.random(n) < 0.1
    return a, b, w, testing
