"""GPU parity of the int8 tcgen05 Gram (csrc/gram_i8.cu, FSB_GRAM_INT8) against the oracle Gram, the exact
(extended precision) Gram and the fp64 DMMA path.  Run with `pytest -m gpu` on a B200.

Tolerance: the same normwise 1e-13 as the fp64 path (tests/test_gpu_parity.py); the integer path is in fact
correctly rounded per slab, which the order-independence test pins bit for bit.
"""
import numpy as np
import pytest
import torch

from oracle import linear_fit as lf
from tests.synth import SOLVE_CASES, synth_system
from tests.test_gpu_parity import dev, gram_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine8():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from fitsnap_b200.engine import Engine
    eng = Engine(0)
    eng.set_gram_path("int8")
    return eng


@pytest.mark.parametrize("n,k", [(1, 1), (7, 3), (100, 31), (1000, 127), (1000, 128), (2500, 129), (600, 300),
                                 (3000, 520)])
def test_int8_gram_shapes(engine8, n, k):
    assert engine8.gram_path(n, k) == "int8"
    rng = np.random.default_rng(n * 1000 + k)
    a = rng.standard_normal((n, k)) * 10.0 ** rng.uniform(-3, 0, k)
    b = rng.standard_normal(n)
    w = 10.0 ** rng.uniform(-2, 2, n)
    t = rng.random(n) < 0.15
    A, B, W, T = dev(engine8, a, b, w, t)
    gram_close(engine8.gram(A, B, W, T).cpu().numpy(), a, b, w, t)
    gram_close(engine8.gram(A, B, W, None).cpu().numpy(), a, b, w, None)


def test_int8_gram_is_correctly_rounded_on_small_input(engine8):
    """Entries whose exact value fits: small integers times powers of two -> the Gram must be EXACT."""
    rng = np.random.default_rng(5)
    n, k = 900, 37
    a = rng.integers(-1000, 1001, (n, k)).astype(np.float64) * 2.0 ** rng.integers(-20, 20, k)
    b = rng.integers(-50, 51, n).astype(np.float64)
    w = 2.0 ** rng.integers(-3, 4, n)
    A, B, W, _ = dev(engine8, a, b, w)
    g = engine8.gram(A, B, W).cpu().numpy()
    aug = np.concatenate([a * w[:, None], (w * b)[:, None]], axis=1)
    exact = aug.T @ aug          # every product and partial sum is an exactly representable dyadic number
    assert np.array_equal(g, exact)


def test_int8_gram_vs_extended_precision_with_wide_weights(engine8):
    """WBe-like group weights span 1e-12 .. 1.5e3 (examples/WBe_PRB2019/WBe-example.in [GROUPS])."""
    rng = np.random.default_rng(6)
    n, k = 4000, 55
    a = rng.standard_normal((n, k)) * 10.0 ** rng.uniform(-4, 1, k)
    b = rng.standard_normal(n)
    w = 10.0 ** rng.uniform(-12, 3.2, n)
    A, B, W, _ = dev(engine8, a, b, w)
    g = engine8.gram(A, B, W).cpu().numpy()
    aug = np.concatenate([a * w[:, None], (w * b)[:, None]], axis=1).astype(np.longdouble)
    exact = np.asarray(aug.T @ aug, dtype=np.float64)
    d = np.sqrt(np.diag(exact))
    assert np.max(np.abs(g - exact) / np.outer(d, d)) < 1e-15


def test_int8_gram_is_independent_of_row_order(engine8):
    """Integer accumulation: permuting the rows of a slab must not change a single bit."""
    rng = np.random.default_rng(7)
    n, k = 30000, 140
    a = rng.standard_normal((n, k)) * 10.0 ** rng.uniform(-3, 0, k)
    b, w = rng.standard_normal(n), 10.0 ** rng.uniform(-2, 2, n)
    perm = rng.permutation(n)
    A, B, W, _ = dev(engine8, a, b, w)
    g1 = engine8.gram(A, B, W).clone()
    A2, B2, W2, _ = dev(engine8, a[perm], b[perm], w[perm])
    g2 = engine8.gram(A2, B2, W2)
    assert torch.equal(g1, g2)


def test_int8_gram_several_slabs_matches_fp64_path(engine, engine8):
    """600k rows = 3 slabs of the int8 path, compared on the device with the DMMA path."""
    n, k = 600000, 40
    gen = torch.Generator(device=engine8.device).manual_seed(12)
    A = torch.randn((n, k), dtype=torch.float64, device=engine8.device, generator=gen)
    b = torch.randn(n, dtype=torch.float64, device=engine8.device, generator=gen)
    w = torch.rand(n, dtype=torch.float64, device=engine8.device, generator=gen) + 0.5
    t = (torch.rand(n, device=engine8.device, generator=gen) < 0.1).to(torch.uint8)
    g8 = engine8.gram(A, b, w, t)
    g64 = engine.gram(A, b, w, t)
    assert engine.gram_path(n, k) == "fp64"
    d = g64.diagonal().abs().sqrt()
    rel = float(((g8 - g64).abs() / torch.outer(d, d)).max())
    assert rel < 1e-13, rel
    assert torch.equal(g8, g8.T)


def test_int8_gram_padded_and_odd_lda(engine8):
    rng = np.random.default_rng(8)
    n, k = 700, 150
    for lda in (k, k + 1, k + 6):          # odd pitch takes the scalar-load branch of the conversion
        buf = rng.standard_normal((n, lda))
        b, w = rng.standard_normal(n), 10.0 ** rng.uniform(-1, 1, n)
        Abuf = engine8.to_device(buf)
        g = engine8.gram(Abuf[:, :k], engine8.to_device(b), engine8.to_device(w)).cpu().numpy()
        gram_close(g, buf[:, :k], b, w, None)


def test_int8_gram_empty_masked_and_nonfinite(engine8):
    k = 5
    A = torch.zeros((0, k), dtype=torch.float64, device=engine8.device)
    z = torch.zeros(0, dtype=torch.float64, device=engine8.device)
    assert float(engine8.gram(A, z, z).abs().max()) == 0.0
    rng = np.random.default_rng(0)
    a, b, w = rng.standard_normal((50, k)), rng.standard_normal(50), np.ones(50)
    A, B, W, T = dev(engine8, a, b, w, np.ones(50, dtype=bool))
    assert float(engine8.gram(A, B, W, T).abs().max()) == 0.0
    a[7, 2] = np.inf                       # a non-finite training value poisons the Gram, as in fp64
    A, B, W, _ = dev(engine8, a, b, w)
    assert bool(torch.isnan(engine8.gram(A, B, W)).all())


@pytest.mark.parametrize("name", ["ill", "wide"])
def test_int8_fit_matches_reference_solvers(engine8, name):
    a, b, w, t = synth_system(**SOLVE_CASES[name])
    A, B, W, T = dev(engine8, a, b, w, t)
    x = engine8.fit(A, B, W, T, alpha=0.0, refine=2).coefficients()
    ref = lf.svd_fit(a, b, w, t)
    mr, l2, _ = lf.coeff_rel_err(x, ref)
    assert mr < 1e-10, (mr, l2)
    alpha = 1e-6
    x = engine8.fit(A, B, W, T, alpha=alpha, refine=2).coefficients()
    mr, l2, _ = lf.coeff_rel_err(x, lf.ridge_fit_exact(a, b, w, alpha, t))
    assert mr < 1e-10, (mr, l2)


@pytest.mark.parametrize("n,k", [(3000, 45), (270000, 3)])
def test_int8_gram_bit_exact_vs_integer_oracle(engine8, n, k):
    """Integer work has a bit-exact bar: the device Gram (16 residue GEMMs on the tensor cores + CRT) must equal
    the exact-integer restatement in oracle/int8_gram.py bit for bit -- wide weights, a test mask, and (second
    case) two slabs with their own column scales."""
    from oracle.int8_gram import quantised_gram
    rng = np.random.default_rng(n + k)
    a = rng.standard_normal((n, k)) * 10.0 ** rng.uniform(-4, 1, k)
    b = rng.standard_normal(n)
    w = 10.0 ** rng.uniform(-12, 3.2, n)
    t = rng.random(n) < 0.1
    A, B, W, T = dev(engine8, a, b, w, t)
    g = engine8.gram(A, B, W, T).cpu().numpy()
    assert np.array_equal(g, quantised_gram(a, b, w, t))
