"""CPU: pin the oracle (oracle/linear_fit.py) to the reference.

(1) the reference's own golden triple -> Ta_pot.snapcoeff (the assertion the reference's
    tests/example_checker.py:62 makes is max(test - std) < 1e-6; we hold the oracle to 1e-12);
(2) fixtures produced by the unmodified reference classes (oracle/make_golden.py);
(3) when /root/reference is present (build container), the live reference itself.
"""
import numpy as np
import pytest

from oracle import linear_fit as lf
from oracle import ref_driver as rd
from tests.conftest import load_golden
from tests.synth import SOLVE_CASES, synth_system

SCATTER = ["snap_b0_efs", "snap_b1_efs", "snap_b0_ef", "snap_b1_es", "snap_b0_f", "pace_b0_efs", "pace_b1_efs"]


def test_oracle_svd_reproduces_reference_golden_coefficients(ta):
    x = lf.svd_fit(ta["a"], ta["b"], ta["w"])
    assert np.max(np.abs(x - ta["snapcoeff"])) < 1e-12          # golden file of the reference
    assert np.max(x - ta["snapcoeff"]) < 1e-6                   # the reference's own criterion
    assert np.array_equal(x, ta["ref_svd"])                     # reference class, same container


def test_oracle_ridge_lasso_match_reference_classes(ta):
    a, b, w = ta["a"], ta["b"], ta["w"]
    assert np.allclose(lf.ridge_fit(a, b, w, 1e-6), ta["ref_ridge_1e6"], rtol=0, atol=1e-13)
    assert np.allclose(lf.ridge_fit(a, b, w, 1e-6, local_solver=True), ta["ref_ridge_local_1e6"], rtol=0, atol=1e-13)
    assert np.allclose(lf.lasso_fit(a, b, w, 1e-6, 20000), ta["ref_lasso_1e6"], rtol=0, atol=1e-13)


def test_oracle_training_mask(ta):
    x = lf.svd_fit(ta["a"], ta["b"], ta["w"], testing=ta["testing"])
    assert np.allclose(x, ta["ref_svd_split"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("name", list(SOLVE_CASES))
def test_oracle_synthetic_cases(name):
    g = load_golden("solve_%s.npz" % name)
    a, b, w, t = synth_system(**SOLVE_CASES[name])
    assert np.allclose([a.sum(), b.sum(), w.sum(), t.sum()], g["checksum"], rtol=1e-13)   # generator is reproducible
    tol = 1e-9 if name == "hard" else 1e-12
    assert lf.coeff_rel_err(lf.svd_fit(a, b, w, t), g["ref_svd"])[1] < tol
    assert lf.coeff_rel_err(lf.svd_fit(a, b, w), g["ref_svd_all"])[1] < tol
    assert lf.coeff_rel_err(lf.ridge_fit(a, b, w, 1e-6, t), g["ref_ridge_1e6"])[1] < 1e-9


@pytest.mark.parametrize("tag", SCATTER)
def test_oracle_scatter_bit_exact(tag):
    g = load_golden("scatter_%s.npz" % tag)
    nat = g["natoms"]
    roff = np.concatenate([[0], np.cumsum(7 + 3 * nat.astype(np.int64))])
    aoff = np.concatenate([[0], np.cumsum(nat.astype(np.int64))])
    cfgs = []
    for c in range(len(nat)):
        cfgs.append(dict(block=g["raw"][roff[c]:roff[c + 1]], natoms=int(nat[c]), volume=float(g["volume"][c]),
                         energy=float(g["energy"][c]), forces=g["forces"][3 * aoff[c]:3 * aoff[c + 1]],
                         stress=g["stress"][c], eweight=float(g["eweight"][c]), fweight=float(g["fweight"][c]),
                         vweight=float(g["vweight"][c]), type_fraction=g["type_fraction"][c]))
    a, b, w = lf.assemble(cfgs, int(g["numtypes"]), int(g["ncoeff"]), int(g["bzeroflag"]), g["blank2j"],
                          int(g["use_energy"]), int(g["use_force"]), int(g["use_stress"]))
    assert np.array_equal(a, g["ref_a"]) and np.array_equal(b, g["ref_b"]) and np.array_equal(w, g["ref_w"])


def test_group_errors_match_reference_metrics_file(ta):
    """Ta_metrics.md of the reference's golden run: '*ALL Unweighted Training Energy' mae/rmse
    (0.112787 / 0.379769).  The legacy golden arrays hold the 363 energy rows first."""
    x = lf.svd_fit(ta["a"], ta["b"], ta["w"])
    p = lf.predictions(ta["a"], x)
    e = lf.group_errors(ta["b"][:363], p[:363], ta["w"][:363])
    assert abs(e["mae"] - 0.112787) < 5e-7 and abs(e["rmse"] - 0.379769) < 5e-7


@pytest.mark.skipif(not rd.reference_available(), reason="reference tree only exists in the build container")
def test_oracle_against_live_reference():
    rng = np.random.default_rng(5)
    a = rng.standard_normal((500, 12)) * 10.0 ** rng.uniform(-2, 0, 12)
    b = rng.standard_normal(500)
    w = 10.0 ** rng.uniform(-2, 2, 500)
    t = rng.random(500) < 0.2
    assert np.array_equal(rd.ref_fit("SVD", a, b, w, testing=t)[0], lf.svd_fit(a, b, w, t))
    assert np.array_equal(rd.ref_fit("RIDGE", a, b, w, testing=t, ridge_alpha=1e-4)[0], lf.ridge_fit(a, b, w, 1e-4, t))
    x_t = rd.ref_fit("SVD", a, b, w, apply_transpose=1)[0]
    assert np.array_equal(x_t, lf.svd_fit(a, b, w, apply_transpose=True))


@pytest.mark.parametrize("name", ["well", "zerocol"])
def test_oracle_anl_matches_reference_fixture(name):
    """oracle.linear_fit.anl_fit restates anl.py:19-58; fixture written by the unmodified reference."""
    from tests.synth import SOLVE_CASES, synth_system
    g = load_golden("anl_%s.npz" % name)
    a, b, w, t = synth_system(**SOLVE_CASES[name])
    assert np.allclose([a.sum(), b.sum(), w.sum(), t.sum()], g["checksum"], rtol=1e-13)
    mean, cov = lf.anl_fit(a, b, w, float(g["cov_nugget"]), t)
    assert np.max(np.abs(mean - g["ref_mean"])) <= 1e-9 * np.max(np.abs(g["ref_mean"]))
    assert np.max(np.abs(cov - g["ref_cov"])) <= 1e-8 * np.max(np.abs(g["ref_cov"]))


def test_int8_gram_oracle_is_the_correctly_rounded_quantised_gram():
    """oracle/int8_gram.py (exact-integer statement of the tensor-core Gram): equals aw^T aw within half an ulp-ish of
    an extended-precision product, is EXACT on dyadic inputs, and follows the device's slab rule."""
    from oracle.int8_gram import quantised_gram, slab_rows_for
    rng = np.random.default_rng(0)
    n, k = 1500, 20
    a = rng.standard_normal((n, k)) * 10.0 ** rng.uniform(-3, 0, k)
    b, w = rng.standard_normal(n), 10.0 ** rng.uniform(-2, 2, n)
    t = rng.random(n) < 0.2
    g = quantised_gram(a, b, w, t)
    aw, bw = lf.weighted_system(a, b, w, t)
    aug = np.concatenate([aw, bw[:, None]], 1).astype(np.longdouble)
    exact = np.asarray(aug.T @ aug, dtype=np.float64)
    d = np.sqrt(np.diag(exact))
    assert np.max(np.abs(g - exact) / np.outer(d, d)) < 1e-15 and np.array_equal(g, g.T)
    a2 = rng.integers(-1000, 1001, (300, 7)).astype(float) * 2.0 ** rng.integers(-20, 20, 7)
    b2, w2 = rng.integers(-50, 51, 300).astype(float), 2.0 ** rng.integers(-3, 4, 300)
    aug2 = np.concatenate([a2 * w2[:, None], (w2 * b2)[:, None]], 1)
    assert np.array_equal(quantised_gram(a2, b2, w2), aug2.T @ aug2)
    assert slab_rows_for(1) == 128 and slab_rows_for(262144) == 262144 and slab_rows_for(600000) == 200064


SINGLE = ["snap_b0_efs", "snap_b1_ef", "snap_b0_es", "snap_b1_fs", "pace_b0_efs", "pace_b1_ef"]


@pytest.mark.parametrize("tag", SINGLE)
def test_oracle_process_single_bit_exact(tag):
    """`config_rows_single` == the unmodified reference's `process_single` output (fixtures written by
    oracle/make_golden.py): zero rows for switched-off families, default weights of 1.0."""
    g = load_golden("single_%s.npz" % tag)
    nat = g["natoms"]
    roff = np.concatenate([[0], np.cumsum(7 + 3 * nat.astype(np.int64))])
    aoff = np.concatenate([[0], np.cumsum(nat.astype(np.int64))])
    ooff = np.concatenate([[0], np.cumsum(g["rows_per_config"])])
    drop = bool(g["weights_dropped"])
    for c in range(len(nat)):
        a, b, w = lf.config_rows_single(
            g["raw"][roff[c]:roff[c + 1]], nat[c], g["volume"][c], g["energy"][c], g["forces"][3 * aoff[c]:3 * aoff[c + 1]],
            g["stress"][c], None if drop else g["eweight"][c], None if drop else g["fweight"][c],
            None if drop else g["vweight"][c], g["type_fraction"][c], int(g["numtypes"]), int(g["ncoeff"]),
            int(g["bzeroflag"]), g["blank2j"], bool(g["use_energy"]), bool(g["use_force"]), bool(g["use_stress"]))
        sl = slice(ooff[c], ooff[c + 1])
        assert np.array_equal(a, g["ref_a"][sl]) and np.array_equal(b, g["ref_b"][sl]) and np.array_equal(w, g["ref_w"][sl])


def test_reference_style_assembly_and_fit_equal_the_plain_restatement():
    """The timing-faithful variants used by bench.py's reference arm (dense diag(blank2J) matmul, list masks, residual
    product) compute the same numbers as the plain restatement."""
    g = load_golden("scatter_snap_b0_efs.npz")
    nat = g["natoms"]
    roff = np.concatenate([[0], np.cumsum(7 + 3 * nat.astype(np.int64))])
    aoff = np.concatenate([[0], np.cumsum(nat.astype(np.int64))])
    cfgs = [dict(block=g["raw"][roff[c]:roff[c + 1]], natoms=int(nat[c]), volume=g["volume"][c], energy=g["energy"][c],
                 forces=g["forces"][3 * aoff[c]:3 * aoff[c + 1]], stress=g["stress"][c], eweight=g["eweight"][c],
                 fweight=g["fweight"][c], vweight=g["vweight"][c], type_fraction=g["type_fraction"][c])
            for c in range(len(nat))]
    a, b, w = lf.assemble_as_reference(cfgs, int(g["numtypes"]), int(g["ncoeff"]), int(g["bzeroflag"]), g["blank2j"])
    assert np.array_equal(a, g["ref_a"]) and np.array_equal(b, g["ref_b"]) and np.array_equal(w, g["ref_w"])
    t = g["ref_testing"]
    x1, res = lf.ridge_perform_fit_as_reference(a, b, w, 1e-6, [bool(v) for v in t])
    assert np.array_equal(x1, lf.ridge_fit(a, b, w, 1e-6, t)) and res.shape[0] == int((~t).sum())
    assert np.array_equal(lf.svd_perform_fit_as_reference(a, b, w, [bool(v) for v in t]), lf.svd_fit(a, b, w, t))
