"""CPU: host logic of the out-of-core fit (pipeline.StreamingLinearFit / npy_row_chunks / LinearFitPipeline.fit_stream)
on the test double of tests/fake_engine.py; the kernels it drives are covered by the -m gpu tests."""
import numpy as np
import pytest
import torch

from oracle import linear_fit as lf
from tests.conftest import load_golden
from tests.fake_engine import OracleEngine


def test_streamed_npy_dumps_reproduce_the_golden_ta_fit(tmp_path):
    """Descriptors.npy / Truth-Ref.npy / Weights.npy (the reference's dump format) streamed in 5 chunks give the
    golden Ta coefficients (examples/Ta_Linear_JCP2014/20May21_Standard/Ta_pot.snapcoeff)."""
    from fitsnap_b200.pipeline import StreamingLinearFit, npy_row_chunks
    ta = load_golden("ta_linear.npz")
    a, b, w = ta["a"], ta["b"], ta["w"]
    np.save(tmp_path / "Descriptors.npy", a)
    np.save(tmp_path / "Truth-Ref.npy", b)
    np.save(tmp_path / "Weights.npy", w)
    chunks = npy_row_chunks(tmp_path / "Descriptors.npy", tmp_path / "Truth-Ref.npy", tmp_path / "Weights.npy",
                            chunk_rows=3500)
    res = StreamingLinearFit(alpha=0.0, refine=2, engine=OracleEngine()).fit(chunks)
    assert res.extra["rows_streamed"] == a.shape[0]
    x = res.x.numpy()
    assert lf.coeff_rel_err(x, ta["ref_svd"])[0] < 1e-9
    assert np.max(np.abs(x - ta["snapcoeff"])) < 1e-6          # the reference's own test tolerance (test_examples.py)


def test_streamed_chunks_with_test_mask_equal_one_shot_fit():
    from fitsnap_b200.engine import fit_rows
    from fitsnap_b200.pipeline import StreamingLinearFit
    from tests.synth import SOLVE_CASES, synth_system
    a, b, w, t = synth_system(**SOLVE_CASES["ill"])
    eng = OracleEngine()
    cuts = [0, 1000, 1001, 4200, a.shape[0]]
    chunks = [(a[i:j], b[i:j], w[i:j], t[i:j]) for i, j in zip(cuts[:-1], cuts[1:])]
    res = StreamingLinearFit(alpha=1e-6, refine=2, engine=eng).fit(chunks)
    one = fit_rows(eng, torch.from_numpy(a), torch.from_numpy(b), torch.from_numpy(w),
                   torch.from_numpy(t.astype(np.uint8)), alpha=1e-6, refine=2, diagnostics=False)
    assert np.max(np.abs(res.x.numpy() - one.x.numpy())) < 1e-11 * np.max(np.abs(one.x.numpy()))
    assert lf.coeff_rel_err(res.x.numpy(), lf.ridge_fit_exact(a, b, w, 1e-6, t))[0] < 1e-10
    with pytest.raises(ValueError):
        StreamingLinearFit(engine=eng).fit([])


def test_fit_stream_over_raw_batches_equals_fit_host():
    """Raw LAMMPS blocks streamed batch by batch (never resident as a whole) vs the one-shot host entry."""
    from fitsnap_b200.pipeline import LinearFitPipeline
    rng = np.random.default_rng(5)
    nt, nc, ncfg = 2, 6, 60
    kraw, k = nt * nc, nt * nc + nt
    natoms = rng.integers(1, 7, ncfg).astype(np.int32)
    blocks = [rng.standard_normal((7 + 3 * n, kraw + 1)) for n in natoms]
    vol = rng.uniform(50, 500, ncfg)
    energy = rng.normal(-5, 1, ncfg) * natoms
    forces = [rng.standard_normal((n, 3)) for n in natoms]
    stress = rng.standard_normal((ncfg, 3, 3))
    stress = 0.5 * (stress + stress.transpose(0, 2, 1))
    ew, fw, vw = 10.0 ** rng.uniform(-1, 2, ncfg), 10.0 ** rng.uniform(-1, 1, ncfg), 10.0 ** rng.uniform(-3, -1, ncfg)
    tf = rng.dirichlet(np.ones(nt), ncfg)
    pipe = LinearFitPipeline(nt, nc, False, np.ones(k), alpha=1e-8, refine=2, engine=OracleEngine())
    x_one, _, _ = pipe.fit_host(blocks, natoms, vol, energy, forces, stress, ew, fw, vw, tf, chunks=1)
    cuts = [0, 13, 14, 40, ncfg]
    batches = [(blocks[i:j], natoms[i:j], vol[i:j], energy[i:j], forces[i:j], stress[i:j], ew[i:j], fw[i:j], vw[i:j],
                tf[i:j]) for i, j in zip(cuts[:-1], cuts[1:])]
    calls = []

    def factory():
        calls.append(1)
        return iter(batches)
    res = pipe.fit_stream(factory)
    assert len(calls) == 3                                   # Gram pass + two refinement passes
    assert np.max(np.abs(res.x.numpy() - x_one)) < 1e-11 * np.max(np.abs(x_one))
