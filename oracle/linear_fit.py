"""CPU oracle (TEST INFRASTRUCTURE, not product code): numpy restatement of the
FitSNAP linear-fit hot path.  See oracle/__init__.py for who may import this.

Every function cites the reference file:line (relative to /root/reference) whose
arithmetic it restates.  The heavy third-party calls of the reference
(`scipy.linalg.lstsq`, `sklearn.linear_model.Ridge/Lasso`; versions unpinned in the
reference's pyproject.toml:25-36) are called exactly the way the reference calls
them -- they ARE the reference arithmetic at that boundary.

Parity status: pinned (tests/test_oracle.py) against the reference's golden Ta
triple and against fixtures generated from the unmodified reference classes.
"""
from __future__ import annotations

import numpy as np

# lammps_snap.py:526 / lammps_pace.py:479 -- eV/A^3 -> bar conversion used for virial rows
VIRIAL_UNIT = 1.6021765e6
# lammps_snap.py:541 -- order in which the 3x3 stress tensor is flattened to 6 rows
VOIGT_I = (0, 1, 2, 1, 0, 0)
VOIGT_J = (0, 1, 2, 2, 2, 1)


# --------------------------------------------------------------------------- rows
def rows_per_config(natoms, energy=True, force=True, stress=True):
    """calculator.py:261-272 (a_len): rows contributed by one configuration."""
    return int(bool(energy)) + 3 * int(natoms) * int(bool(force)) + 6 * int(bool(stress))


def descriptor_width(ncoeff, numtypes, bzeroflag):
    """lammps_snap.py:15-23 / lammps_pace.py:15-23 get_width() (linear branch)."""
    return ncoeff * numtypes + (0 if bzeroflag else numtypes)


def config_rows(block, natoms, volume, energy, forces, stress, eweight, fweight, vweight,
                type_fraction, numtypes, ncoeff, bzeroflag, blank2j,
                use_energy=True, use_force=True, use_stress=True):
    """Rows of (A, b, w) for ONE configuration.

    Restates lammps_snap.py:391-549 (`LammpsSnap._collect_lammps`) and
    lammps_pace.py:369-501 (`LammpsPace._collect_lammps`), bikflag=0 (the linear path).

    block         (1+3N+6, ncoeff*numtypes+1) raw LAMMPS compute array; last column is
                  the reference-potential energy / force / virial.
    type_fraction (numtypes,) fraction of atoms of each type (lammps_snap.py:459-462);
                  only used when bzeroflag == 0.
    blank2j       (K,) column mask/prefactor ([BISPECTRUM]/[ACE] blank2J).
    """
    block = np.asarray(block, dtype=np.float64)
    n = int(natoms)
    kraw = ncoeff * numtypes
    assert block.shape == (1 + 3 * n + 6, kraw + 1), block.shape
    k = descriptor_width(ncoeff, numtypes, bzeroflag)
    blank2j = np.asarray(blank2j, dtype=np.float64)
    assert blank2j.shape == (k,)

    def widen(rows, first_col):
        # insert one column per type in front of each type's ncoeff block when
        # bzeroflag == 0 (lammps_snap.py:455-464 energy, :495-499 force, :528-532 virial)
        if bzeroflag:
            return rows
        r = rows.reshape(rows.shape[0], numtypes, ncoeff)
        lead = np.broadcast_to(np.asarray(first_col, dtype=np.float64).reshape(-1, numtypes, 1),
                               (rows.shape[0], numtypes, 1))
        return np.concatenate([lead, r], axis=2).reshape(rows.shape[0], k)

    a_parts, b_parts, w_parts = [], [], []
    if use_energy:
        # lammps_snap.py:435 (R/N), :466-467 (mask), :469-473 (b), :476 (w)
        e_row = block[0:1, :kraw] / n
        e_row = widen(e_row, np.asarray(type_fraction, dtype=np.float64).reshape(1, numtypes))
        a_parts.append(e_row * blank2j[np.newaxis, :])
        b_parts.append(np.array([(energy - block[0, kraw]) / n]))
        w_parts.append(np.array([eweight], dtype=np.float64))
    if use_force:
        # lammps_snap.py:493-511; the reference multiplies by diag(blank2J) with matmul,
        # which equals the element-wise product for finite inputs.
        f_rows = widen(block[1:1 + 3 * n, :kraw], np.zeros((3 * n, numtypes)))
        a_parts.append(f_rows * blank2j[np.newaxis, :])
        b_parts.append(np.asarray(forces, dtype=np.float64).ravel() - block[1:1 + 3 * n, kraw])
        w_parts.append(np.full(3 * n, fweight, dtype=np.float64))
    if use_stress:
        # lammps_snap.py:526 ((1.6021765e6*R)/V in that order), :535-543
        v_rows = VIRIAL_UNIT * block[1 + 3 * n:, :kraw] / volume
        v_rows = widen(v_rows, np.zeros((6, numtypes)))
        a_parts.append(v_rows * blank2j[np.newaxis, :])
        s = np.asarray(stress, dtype=np.float64)
        b_parts.append(s[list(VOIGT_I), list(VOIGT_J)].ravel() - block[1 + 3 * n:, kraw])
        w_parts.append(np.full(6, vweight, dtype=np.float64))
    if not a_parts:
        return np.zeros((0, k)), np.zeros(0), np.zeros(0)
    return np.concatenate(a_parts, 0), np.concatenate(b_parts), np.concatenate(w_parts)


def assemble(configs, numtypes, ncoeff, bzeroflag, blank2j,
             use_energy=True, use_force=True, use_stress=True):
    """Stack `config_rows` over configurations in order (single rank:
    calculator.py:287-291 + lammps_snap.py:488-549 interleave E/F/S per config).

    `configs` is a list of dicts with keys block, natoms, volume, energy, forces, stress,
    eweight, fweight, vweight, type_fraction.
    """
    parts = [config_rows(c["block"], c["natoms"], c["volume"], c["energy"], c["forces"], c["stress"],
                         c["eweight"], c["fweight"], c["vweight"], c.get("type_fraction"),
                         numtypes, ncoeff, bzeroflag, blank2j, use_energy, use_force, use_stress)
             for c in configs]
    k = descriptor_width(ncoeff, numtypes, bzeroflag)
    if not parts:
        return np.zeros((0, k)), np.zeros(0), np.zeros(0)
    return (np.concatenate([p[0] for p in parts], 0), np.concatenate([p[1] for p in parts]),
            np.concatenate([p[2] for p in parts]))


def config_rows_single(block, natoms, volume, energy, forces, stress, eweight, fweight, vweight,
                       type_fraction, numtypes, ncoeff, bzeroflag, blank2j,
                       use_energy=True, use_force=True, use_stress=True):
    """(a, b, w) of `process_single` for ONE configuration: lammps_snap.py:224-389 /
    lammps_pace.py:197-366 (`_collect_lammps_single`, bikflag = 0).  Same row arithmetic as
    `config_rows`, different layout: `a` always holds the energy row and the 3N force rows
    (+ the 6 virial rows iff `use_stress`: `na = rows of the block, minus 6 without stress`,
    :266-269), `irow` advances past a family whether or not it is assembled (:341, :364), so
    the rows of a switched-off family stay zero; weights default to 1.0 when the data
    dictionary has no eweight / fweight / vweight key (:337, :360, :383) -- pass None."""
    n = int(natoms)
    k = descriptor_width(ncoeff, numtypes, bzeroflag)
    na = 1 + 3 * n + (6 if use_stress else 0)
    a, b, w = np.zeros((na, k)), np.zeros(na), np.zeros(na)
    ew, fw, vw = (1.0 if v is None else v for v in (eweight, fweight, vweight))
    for on, dst0, cnt, sel in ((use_energy, 0, 1, (True, False, False)), (use_force, 1, 3 * n, (False, True, False)),
                               (use_stress, 1 + 3 * n, 6, (False, False, True))):
        if on:
            ra, rb, rw = config_rows(block, natoms, volume, energy, forces, stress, ew, fw, vw, type_fraction,
                                     numtypes, ncoeff, bzeroflag, blank2j, *sel)
            a[dst0:dst0 + cnt], b[dst0:dst0 + cnt], w[dst0:dst0 + cnt] = ra, rb, rw
    return a, b, w


# ------------------------------------------------- the reference's CPU path, step for step
# Used by bench.py's reference arm / cpu_baseline: the functions above restate WHAT the reference
# computes; these two also do it HOW the reference does it (the numpy calls that dominate its
# time), so that timing them is timing the reference's algorithm, not a tidied-up version.
def config_rows_as_reference(block, natoms, volume, energy, forces, stress, eweight, fweight, vweight,
                             type_fraction, numtypes, ncoeff, bzeroflag, blank2j, out_a, out_b, out_w, index):
    """lammps_snap.py:430-549 with its own operations: in-place `/= num_atoms` on the block view (:435),
    `np.concatenate` of the one-hot / zero lead columns (:457-464, :495-499, :528-532), the force and
    virial rows multiplied by the DENSE `np.diag(blank2J)` with `np.matmul` (:501-502, :535-536: an
    O(rows K^2) product for a column mask), rows written into the shared arrays at `index`.
    Returns the next index.  Bit-identical to `config_rows` (asserted in tests/test_oracle.py)."""
    n = int(natoms)
    kraw = ncoeff * numtypes
    k = descriptor_width(ncoeff, numtypes, bzeroflag)
    lmp = np.array(block, dtype=np.float64)          # the LAMMPS array of this configuration (:423)
    if np.isinf(lmp).any() or np.isnan(lmp).any():   # :426-428
        raise ValueError("Nan in computed data")
    irow = 0
    b_sum = lmp[irow:irow + 1, :kraw]
    b_sum /= n                                       # :435
    if not bzeroflag:                                # :455-464
        b_sum = b_sum.reshape(numtypes, ncoeff)
        onehot = np.asarray(type_fraction, dtype=np.float64).reshape(numtypes, 1)
        b_sum = np.concatenate((onehot, b_sum), axis=1).reshape(k)
    out_a[index:index + 1] = b_sum * blank2j[np.newaxis, :]          # :466-467
    out_b[index] = (energy - lmp[irow, kraw]) / n                    # :469-473
    out_w[index] = eweight
    index += 1
    irow += 1
    nf = 3 * n
    db = lmp[irow:irow + nf, :kraw]
    if not bzeroflag:                                # :495-499
        db = db.reshape(nf, numtypes, ncoeff)
        db = np.concatenate([np.zeros((nf, numtypes, 1)), db], axis=2).reshape(nf, k)
    out_a[index:index + nf] = np.matmul(db, np.diag(blank2j))        # :501-502
    out_b[index:index + nf] = np.asarray(forces, dtype=np.float64).ravel() - lmp[irow:irow + nf, kraw]
    out_w[index:index + nf] = fweight
    index += nf
    irow += nf
    vb = VIRIAL_UNIT * lmp[irow:irow + 6, :kraw] / volume             # :526
    if not bzeroflag:
        vb = vb.reshape(6, numtypes, ncoeff)
        vb = np.concatenate([np.zeros((6, numtypes, 1)), vb], axis=2).reshape(6, k)
    out_a[index:index + 6] = np.matmul(vb, np.diag(blank2j))         # :535-536
    s = np.asarray(stress, dtype=np.float64)
    out_b[index:index + 6] = s[list(VOIGT_I), list(VOIGT_J)].ravel() - lmp[irow:irow + 6, kraw]
    out_w[index:index + 6] = vweight
    return index + 6


def assemble_as_reference(configs, numtypes, ncoeff, bzeroflag, blank2j):
    """calculator.py:287-291 (allocate the shared a, b, w) + one `_collect_lammps` per configuration, all three
    row families on (the bench workloads)."""
    k = descriptor_width(ncoeff, numtypes, bzeroflag)
    n_rows = sum(rows_per_config(c["natoms"]) for c in configs)
    a, b, w = np.zeros((n_rows, k)), np.zeros(n_rows), np.zeros(n_rows)
    blank2j = np.asarray(blank2j, dtype=np.float64)
    index = 0
    for c in configs:
        index = config_rows_as_reference(c["block"], c["natoms"], c["volume"], c["energy"], c["forces"], c["stress"],
                                         c["eweight"], c["fweight"], c["vweight"], c.get("type_fraction"),
                                         numtypes, ncoeff, bzeroflag, blank2j, a, b, w, index)
    return a, b, w


def ridge_perform_fit_as_reference(a, b, w, alpha, testing=None):
    """ridge.py:24-60 step for step: the training mask as a PYTHON LIST of bools (:28-33 -- numpy turns the
    list into an index array on every one of the three fancy-indexing copies), `w[:, None] * a[training]`
    (a second full copy of A, :39), sklearn Ridge (:49-57) and the residual product `aw @ coef - bw` (:60)."""
    from sklearn.linear_model import Ridge
    if testing is not None:
        training = [not elem for elem in testing]
    else:
        training = [True] * np.shape(a)[0]
    w = w[training]                 # ridge.py:36 (the shared-array branch; the explicit-array branch :39 forgets this
    aw, bw = w[:, np.newaxis] * a[training], w * b[training]      # mask and only works when nothing is masked)
    reg = Ridge(alpha=alpha, fit_intercept=False)
    reg.fit(aw, bw)
    residues = np.matmul(aw, reg.coef_) - bw
    return reg.coef_, residues


def svd_perform_fit_as_reference(a, b, w, testing=None):
    """svd.py:31-54 step for step (list mask, two copies of A, scipy.linalg.lstsq(aw, bw, 1e-13))."""
    from scipy.linalg import lstsq
    if testing is not None:
        training = [not elem for elem in testing]
    else:
        training = [True] * np.shape(a)[0]
    w = w[training]                 # svd.py:43
    aw, bw = w[:, np.newaxis] * a[training], w * b[training]
    fit, _residues, _rank, _s = lstsq(aw, bw, 1.0e-13)
    return fit


# ------------------------------------------------------------------------ solvers
def weighted_system(a, b, w, testing=None):
    """svd.py:35-46 / ridge.py:28-39: boolean-mask the training rows, then
    aw = w[:,None]*A, bw = w*b."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    if testing is not None:
        keep = ~np.asarray(testing, dtype=bool)
        a, b, w = a[keep], b[keep], w[keep]
    return w[:, np.newaxis] * a, w * b


def svd_fit(a, b, w, testing=None, apply_transpose=False):
    """svd.py:18-54 `SVD.perform_fit`: scipy.linalg.lstsq(aw, bw, 1.0e-13)
    (LAPACK gelsd, min-norm).  apply_transpose: svd.py:48-53."""
    from scipy.linalg import lstsq
    aw, bw = weighted_system(a, b, w, testing)
    if apply_transpose:
        if np.linalg.cond(aw) ** 2 < 1.0 / np.finfo(np.float64).eps:
            bw = aw.T @ bw
            aw = aw.T @ aw
    x, _res, _rank, _s = lstsq(aw, bw, 1.0e-13)
    return x


def ridge_fit(a, b, w, alpha, testing=None, local_solver=False, apply_transpose=False):
    """ridge.py:11-60 `RIDGE.perform_fit`: sklearn Ridge(alpha, fit_intercept=False)
    (ridge.py:49-50) or `Local_Ridge` = inv(XtX + alpha I) @ Xty
    (lib/ridge_solver/regressor.py:10-16) when [RIDGE] local_solver = 1."""
    aw, bw = weighted_system(a, b, w, testing)
    if apply_transpose:  # ridge.py:41-43
        bw = aw.T @ bw
        aw = aw.T @ aw
    if local_solver:
        xtx = aw.T @ aw
        return np.linalg.inv(xtx + alpha * np.eye(xtx.shape[0])) @ (aw.T @ bw)
    from sklearn.linear_model import Ridge
    reg = Ridge(alpha=alpha, fit_intercept=False)
    reg.fit(aw, bw)
    return reg.coef_


def ridge_fit_exact(a, b, w, alpha, testing=None):
    """Higher-accuracy statement of the SAME minimiser as ridge.py:49-57
    (argmin |aw x - bw|^2 + alpha |x|^2) through an orthogonal solve of the augmented
    system [aw; sqrt(alpha) I] x = [bw; 0]  (sklearn's own Cholesky result is
    only ~cond*eps accurate, so parity is judged against both)."""
    from scipy.linalg import lstsq
    aw, bw = weighted_system(a, b, w, testing)
    k = aw.shape[1]
    aug = np.concatenate([aw, np.sqrt(alpha) * np.eye(k)], 0)
    rhs = np.concatenate([bw, np.zeros(k)])
    return lstsq(aug, rhs)[0]


def anl_fit(a, b, w, cov_nugget=0.0, testing=None):
    """anl.py:19-58 `ANL.perform_fit`: posterior mean pinv(aw^T aw + nugget I) aw^T bw (symmetrised
    inverse, anl.py:41-44), data-noise estimate sigmahat = (|res|^2 / 2) / ((npt - nbas)/2 - 1)
    (anl.py:48-52) and the posterior covariance sigmahat * inverse (anl.py:56).  Returns (mean, cov)."""
    aw, bw = weighted_system(a, b, w, testing)
    npt, nbas = aw.shape
    invptp = np.linalg.pinv(aw.T @ aw + cov_nugget * np.diag(np.ones((nbas,))))
    invptp = invptp * 0.5 + invptp.T * 0.5
    mean = invptp @ (aw.T @ bw)
    res = bw - aw @ mean
    bp = (res @ res) / 2.0
    ap = (npt - nbas) / 2.0
    return mean, (bp / (ap - 1.0)) * invptp


def lasso_fit(a, b, w, alpha, max_iter, testing=None, apply_transpose=False, tol=1e-4):
    """lasso.py:15-30 `LASSO.perform_fit`: sklearn Lasso(alpha, fit_intercept=False,
    max_iter) = argmin 1/(2 n) |bw - aw x|^2 + alpha |x|_1 (coordinate descent)."""
    from sklearn.linear_model import Lasso
    aw, bw = weighted_system(a, b, w, testing)
    if apply_transpose:  # lasso.py:22-24
        bw = aw.T @ bw
        aw = aw.T @ aw
    reg = Lasso(alpha=alpha, fit_intercept=False, max_iter=max_iter, tol=tol)
    reg.fit(aw, bw)
    return reg.coef_


def lasso_objective(aw, bw, x, alpha):
    """sklearn Lasso objective, used to compare minimisers by value."""
    r = bw - aw @ x
    return 0.5 * float(r @ r) / aw.shape[0] + alpha * float(np.abs(x).sum())


# --------------------------------------------------------------- Gram (transpose trick)
def gram(a, b, w, testing=None):
    """examples/library/transpose_trick/example.py:226-246: C = aw^T aw, d = aw^T bw
    (plus bw^T bw and the training row count, which LASSO's objective needs)."""
    aw, bw = weighted_system(a, b, w, testing)
    return aw.T @ aw, aw.T @ bw, float(bw @ bw), aw.shape[0]


# ------------------------------------------------------------------ error analysis
def group_errors(truths, preds, weights):
    """solver.py:108-133 `_ncount_mae_rmse_rsq_unweighted_and_weighted` for one group."""
    t = np.asarray(truths, dtype=np.float64)
    p = np.asarray(preds, dtype=np.float64)
    w = np.asarray(weights, dtype=np.float64)
    res = t - p
    n = len(t)
    ssr = np.square(res).sum()
    out = {"ncount": n, "mae": np.mean(np.abs(res)), "rmse": np.sqrt(ssr / n),
           "rsq": 1 - ssr / np.sum(np.square(t - (t / n).sum()))}
    wres = w * res
    wn = int(np.count_nonzero(w))
    wssr = np.square(wres).sum()
    with np.errstate(divide="ignore", invalid="ignore"):
        out.update({"w_ncount": wn, "w_mae": np.mean(np.abs(wres)),
                    "w_rmse": np.sqrt(wssr / wn) if wn else np.nan,
                    "w_rsq": 1 - wssr / np.sum(np.square(w * t - (w * t / wn).sum())) if wn else np.nan})
    return out


def predictions(a, x):
    """solver.py:377 `df['preds'] = a @ self.fit`."""
    return np.asarray(a, dtype=np.float64) @ np.asarray(x, dtype=np.float64)


def offset_fit(x, numtypes, ncoeff):
    """solver.py:78-102 `_offset`: re-insert a zero B0 per type when bzeroflag=1."""
    x = np.asarray(x, dtype=np.float64)
    if numtypes > 1:
        x = x.reshape(numtypes, ncoeff)
        return np.concatenate([np.zeros((numtypes, 1)), x], axis=1).reshape((-1, 1))
    return np.insert(x, 0, 0)


# ----------------------------------------------------------------- parity metric
def coeff_rel_err(x, x_ref, floor=1e-12):
    """SURVEY 8c parity definition: max_i |x_i-xref_i|/|xref_i| over coefficients with
    |xref_i| > floor*|xref|_inf; the others must satisfy |x_i| <= floor*|xref|_inf*1e3.
    Returns (max_rel, l2_rel, max_small_abs_over_scale)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    x_ref = np.asarray(x_ref, dtype=np.float64).ravel()
    scale = np.max(np.abs(x_ref)) if x_ref.size else 0.0
    big = np.abs(x_ref) > floor * scale
    max_rel = float(np.max(np.abs(x[big] - x_ref[big]) / np.abs(x_ref[big]))) if big.any() else 0.0
    small = float(np.max(np.abs(x[~big])) / scale) if (~big).any() and scale > 0 else 0.0
    nrm = np.linalg.norm(x_ref)
    l2 = float(np.linalg.norm(x - x_ref) / nrm) if nrm > 0 else float(np.linalg.norm(x))
    return max_rel, l2, small
