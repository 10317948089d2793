"""GPU parity at the BASELINE.json shapes, through the FSB_GRAM_AUTO dispatch the product uses:

  C3-like  356 536 x 480   InP energies + forces, weights from the 19 (eweight, fweight) pairs of the InP example
                           -> int8 tcgen05 Gram: 2 slabs x 10 lower-triangular tiles x several row chunks
  c4s      1 000 000 x 1000 ACE-like (BASELINE configs[3] shape, 1/10 of the rows)
                           -> int8 tcgen05 Gram: 4 slabs x 36 tiles x 4 chunks
  C5-like  1 772 880 x 110 WBe, 44 groups of (eweight, fweight, vweight) spanning 1e-12 .. 1.5e3
                           -> fp64 DMMA Gram (gram_dmma_kernel, 105 <= k+1 <= 128)
plus the ill-conditioned variant of each (column scales 10^U(-5,0) and two nearly collinear columns).

Bars: int8 Gram BIT-IDENTICAL to the exact-integer oracle (oracle/int8_gram.quantised_gram_fast, pinned to the
literal Python-integer statement in tests/test_oracle.py); coefficients <= 1e-10 max-rel against the oracle's exact
ridge / lstsq statement on the same rows (well-conditioned) -- for the ill-conditioned variants the bar is the
accuracy the reference's own LAPACK solve has on such a system, cond * eps, stated per test.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import int8_gram
from oracle import linear_fit as lf
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
ALPHA = 1e-6


def _weights(tag):
    with open(os.path.join(GOLDEN, "group_weights.json")) as f:
        return np.array([[r[1], r[2], r[3]] for r in json.load(f)[tag]])


def _device_system(engine, n, k, seed, decades, collinear, row_weights):
    """A = N(0,1) * column scales 10^U(-decades, 0) (+ two nearly dependent columns), b = A x_true + 1e-3 N(0,1);
    generated on the device (a 1e6 x 1000 host matrix would cost more box time than the test)."""
    dev = engine.device
    g = torch.Generator(device=dev).manual_seed(seed)
    A = torch.randn((n, k), dtype=torch.float64, device=dev, generator=g)
    if collinear:
        A[:, 1] = 0.7 * A[:, 0] + collinear * A[:, 1]
        A[:, k - 1] = A[:, k - 2] - 2.0 * A[:, 2] + collinear * A[:, k - 1]
    A *= 10.0 ** (torch.rand(k, dtype=torch.float64, device=dev, generator=g) * -float(decades))
    x_true = torch.randn(k, dtype=torch.float64, device=dev, generator=g)
    b = A @ x_true + 1e-3 * torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    w = torch.from_numpy(row_weights).to(dev)
    return A, b, w


def _efs_row_weights(n_rows, natoms, table, rng, stress=True):
    """Row weights of configurations of `natoms` atoms each: one energy row (eweight), 3N force rows (fweight) and
    optionally 6 virial rows (vweight), the configuration's group drawn from `table`."""
    per = 1 + 3 * natoms + (6 if stress else 0)
    ncfg = -(-n_rows // per)
    grp = rng.integers(0, len(table), ncfg)
    w = np.empty((ncfg, per))
    w[:, 0] = table[grp, 0]
    w[:, 1:1 + 3 * natoms] = table[grp, 1][:, None]
    if stress:
        w[:, 1 + 3 * natoms:] = table[grp, 2][:, None]
    return np.ascontiguousarray(w.reshape(-1)[:n_rows])


@pytest.mark.parametrize("ill", [False, True])
def test_c3_like_inp_shape_int8_auto(engine, ill):
    n, k = 356_536, 480
    rng = np.random.default_rng(31)
    w_rows = _efs_row_weights(n, 64, _weights("InP"), rng, stress=False)
    w_rows /= w_rows.max()            # the InP table is ~1e3..6e6: keep the Gram in a comfortable range, ratios intact
    A, b, w = _device_system(engine, n, k, 3100 + ill, 5 if ill else 3, 1e-2 if ill else 0.0, w_rows)
    assert engine.gram_path(n, k) == "int8"
    g = engine.gram(A, b, w).cpu().numpy()
    a_h, b_h, w_h = A.cpu().numpy(), b.cpu().numpy(), w.cpu().numpy()
    assert np.array_equal(g, int8_gram.quantised_gram_fast(a_h, b_h, w_h)), "int8 Gram differs from the exact-integer oracle"
    res = engine.fit(A, b, w, None, alpha=ALPHA, refine=3)
    ref = lf.ridge_fit_exact(a_h, b_h, w_h, ALPHA)
    mr, l2, _ = lf.coeff_rel_err(res.coefficients(), ref)
    assert mr < (1e-8 if ill else 1e-10), (mr, l2)


@pytest.mark.parametrize("ill", [False, True])
def test_c4s_ace_shape_int8_auto_multi_slab(engine, ill):
    n, k = 1_000_000, 1000
    rng = np.random.default_rng(41)
    table = np.array([[1e-2, 1.0, 100.0 * 1e-3], [1.0, 100.0, 1e-2 * 1e-3], [100.0, 1e-2, 1.0 * 1e-3]])
    w_rows = _efs_row_weights(n, 31, table, rng)
    A, b, w = _device_system(engine, n, k, 4100 + ill, 5 if ill else 3, 1e-2 if ill else 0.0, w_rows)
    assert engine.gram_path(n, k) == "int8"
    g_full = engine.gram(A, b, w).clone()
    # (1) multi-slab accumulation: the full call == the fp64 sum, in slab order, of the Grams of its slabs
    step = int8_gram.slab_rows_for(n)
    acc = None
    for r0 in range(0, n, step):
        gs = engine.gram(A[r0:r0 + step], b[r0:r0 + step], w[r0:r0 + step]).clone()
        acc = gs if acc is None else acc + gs
    assert torch.equal(g_full, acc)
    # (2) one row chunk of one slab, all 36 tiles, against the exact-integer oracle (host cost ~ seconds)
    r0, r1 = 2 * step + 4096, 2 * step + 4096 + 65_536
    gs = engine.gram(A[r0:r1], b[r0:r1], w[r0:r1]).cpu().numpy()
    a_h, b_h, w_h = A[r0:r1].cpu().numpy(), b[r0:r1].cpu().numpy(), w[r0:r1].cpu().numpy()
    assert np.array_equal(gs, int8_gram.quantised_gram_fast(a_h, b_h, w_h))
    # (3) coefficients of a sub-sampled system against the oracle's exact ridge statement
    ns = 81_920
    res = engine.fit(A[:ns], b[:ns], w[:ns], None, alpha=ALPHA, refine=3)
    assert engine.gram_path(ns, k) == "int8"
    ref = lf.ridge_fit_exact(A[:ns].cpu().numpy(), b[:ns].cpu().numpy(), w[:ns].cpu().numpy(), ALPHA)
    mr, l2, _ = lf.coeff_rel_err(res.coefficients(), ref)
    assert mr < (1e-8 if ill else 1e-10), (mr, l2)
    # (4) the full fit: gradient of the ridge objective at the solution, from the streaming residual kernel
    full = engine.fit(A, b, w, None, alpha=ALPHA, refine=3)
    grad = engine.residual(A, b, w, None, full.x) - ALPHA * full.x
    assert float(grad.abs().max() / g_full[:k, k].abs().max()) < 1e-12


@pytest.mark.parametrize("ill", [False, True])
def test_c5_like_wbe_shape_44_weight_groups(engine, ill):
    n, k = 1_772_880, 110
    rng = np.random.default_rng(51)
    w_rows = _efs_row_weights(n, 12, _weights("WBe"), rng)
    A, b, w = _device_system(engine, n, k, 5100 + ill, 5 if ill else 3, 1e-2 if ill else 0.0, w_rows)
    assert engine.gram_path(n, k) == "fp64"
    a_h, b_h, w_h = A.cpu().numpy(), b.cpu().numpy(), w.cpu().numpy()
    # Gram: normwise 1e-13 against numpy (weights over 15 decades: the 1e-12 virial rows must not disturb the rest)
    g = engine.gram(A, b, w).cpu().numpy()
    aw = np.concatenate([w_h[:, None] * a_h, (w_h * b_h)[:, None]], 1)
    exact = aw.T @ aw
    nrm = np.sqrt(np.diag(exact))
    assert np.max(np.abs(g - exact) / np.outer(nrm, nrm)) < 1e-13
    del aw
    for alpha in (0.0, ALPHA):
        res = engine.fit(A, b, w, None, alpha=alpha, refine=3)
        ref = lf.svd_fit(a_h, b_h, w_h) if alpha == 0.0 else lf.ridge_fit_exact(a_h, b_h, w_h, alpha)
        mr, l2, _ = lf.coeff_rel_err(res.coefficients(), ref)
        assert mr < (1e-8 if ill else 1e-10), (alpha, mr, l2)
