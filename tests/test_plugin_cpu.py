"""CPU, build container only (needs /root/reference): the drop-in classes are found by the
reference's own factories under the reference's own names, and the host glue (staging, row
metadata, indices, flush into pt.shared_arrays, device hand-off to the solver) reproduces what the
unmodified reference classes produce.  The device arithmetic is replaced by a test double
(tests/fake_engine.py); the real kernels are checked by the -m gpu tests against the same goldens."""
import numpy as np
import pytest

from oracle import ref_driver as rd

pytestmark = pytest.mark.skipif(not rd.reference_available(), reason="reference tree only exists in the build container")


def _configs(rng, nc, numtypes, n_cfg=9):
    names = ["In", "P"][:numtypes] if numtypes > 1 else ["Ta"]
    cfgs, blocks, vols = [], [], []
    for i in range(n_cfg):
        n = int(rng.integers(1, 9))
        cfgs.append(rd.make_config_dict(n, numtypes, rng, names, group="g%d" % (i % 3), fname="f%d" % i,
                                        eweight=float(10 ** rng.uniform(-2, 2)), fweight=float(10 ** rng.uniform(-2, 2)),
                                        vweight=float(10 ** rng.uniform(-9, -5)), test_bool=bool(i % 4 == 1)))
        blocks.append(rng.standard_normal((1 + 3 * n + 6, nc * numtypes + 1)))
        vols.append(float(rng.uniform(20, 400)))
    return cfgs, blocks, vols


@pytest.fixture()
def registered():
    rd.install_fake_lammps()
    from fitsnap_b200 import plugin
    from tests.fake_engine import OracleEngine
    classes = plugin.register(engine=OracleEngine())
    yield classes
    plugin.unregister()


def test_factories_return_the_dropins(registered):
    from fitsnap3lib.solvers.solver_factory import solver
    from fitsnap3lib.solvers.solver import Solver
    from fitsnap3lib.calculators.calculator_factory import search
    from fitsnap3lib.calculators.lammps_base import LammpsBase
    pt, cfg = rd.make_reference_context(solver="SVD")
    s = solver("SVD", pt, cfg)
    assert type(s) is registered["SVD"] and isinstance(s, Solver) and s.linear
    pt, cfg = rd.make_reference_context(solver="RIDGE", ridge_alpha=1e-6)
    assert type(solver("RIDGE", pt, cfg)) is registered["RIDGE"]
    for name, key in (("LAMMPSSNAP", "LammpsSnap"), ("LAMMPSPACE", "LammpsPace")):
        inst = search(name)
        assert type(inst) is registered[key] and isinstance(inst, LammpsBase)


@pytest.mark.parametrize("bz,efs", [(0, (1, 1, 1)), (1, (1, 1, 0)), (0, (0, 1, 1))])
def test_dropin_calculator_reproduces_reference_rows_and_metadata(registered, bz, efs):
    rng = np.random.default_rng(3)
    kw = dict(numtypes=2, types="In P", twojmax="6 4", bzeroflag=bz, energy=efs[0], force=efs[1], stress=efs[2])
    pt0, cfg0 = rd.make_reference_context(**kw)
    nc = cfg0.sections["BISPECTRUM"].ncoeff
    cfgs, blocks, vols = _configs(rng, nc, 2)
    a_ref, b_ref, w_ref, lists_ref, *_ = rd.ref_scatter(cfgs, blocks, vols, **kw)          # stock classes
    a, b, w, lists, _cfg, pt, calc = rd.ref_scatter(cfgs, blocks, vols, use_factory=True, **kw)   # drop-in via factory
    assert type(calc) is registered["LammpsSnap"]
    assert np.array_equal(a, a_ref) and np.array_equal(b, b_ref) and np.array_equal(w, w_ref)
    for key in ("Row_Type", "Atom_I", "Atom_Type", "Groups", "Configs", "Testing"):
        assert lists[key] == lists_ref[key], key
    assert pt.fitsnap_b200_device["n_rows"] == a_ref.shape[0]


def test_full_plugin_flow_matches_reference_fit(registered):
    """process_configs -> perform_fit -> error_analysis through the factories (fitsnap.py:134-220)."""
    from fitsnap3lib.solvers.solver_factory import solver
    rng = np.random.default_rng(8)
    kw = dict(numtypes=1, types="Ta", twojmax="4", bzeroflag=0)
    pt0, cfg0 = rd.make_reference_context(**kw)
    nc = cfg0.sections["BISPECTRUM"].ncoeff
    cfgs, blocks, vols = _configs(rng, nc, 1, n_cfg=40)
    a_ref, b_ref, w_ref, lists_ref, *_ = rd.ref_scatter(cfgs, blocks, vols, **kw)
    x_ref, _ = rd.ref_fit("SVD", a_ref, b_ref, w_ref, testing=np.array(lists_ref["Testing"]))
    a, b, w, lists, cfg, pt, calc = rd.ref_scatter(cfgs, blocks, vols, use_factory=True, **kw)
    s = solver("SVD", pt, cfg)
    s.refine = 2
    s.perform_fit()                      # reads the device-resident rows left by the calculator
    assert np.max(np.abs(s.fit - x_ref)) < 1e-9 * np.max(np.abs(x_ref))
    s.error_analysis()                   # inherited from the reference's Solver (solver.py:137-435)
    assert len(s.errors) > 0
    assert s.fit.shape[0] == nc + 1      # unchanged by _offset (bzeroflag = 0)


def test_device_error_analysis_matches_reference_table(registered):
    """`error_analysis_device` (ten sums per group from one pass) reproduces the table the reference
    builds through DataFrame(a) + groupby (solver.py:368-429), index and values."""
    from fitsnap3lib.solvers.solver_factory import solver
    rng = np.random.default_rng(11)
    kw = dict(numtypes=1, types="Ta", twojmax="4", bzeroflag=0)
    pt0, cfg0 = rd.make_reference_context(**kw)
    nc = cfg0.sections["BISPECTRUM"].ncoeff
    cfgs, blocks, vols = _configs(rng, nc, 1, n_cfg=30)
    a, b, w, lists, cfg, pt, calc = rd.ref_scatter(cfgs, blocks, vols, use_factory=True, **kw)
    s = solver("SVD", pt, cfg)
    s.refine = 2
    s.perform_fit()
    fit = s.fit.copy()
    s.error_analysis()                       # the reference's implementation (inherited)
    ref = s.errors.copy()
    s.fit = fit                              # bzeroflag = 0: _offset did not touch it
    dev = s.error_analysis_device()
    assert list(dev.index) == list(ref.index) and list(dev.columns) == list(ref.columns)
    assert np.array_equal(dev["ncount"].values, ref["ncount"].values)
    for col in ("mae", "rmse", "rsq"):
        r, d = ref[col].values.astype(float), dev[col].values.astype(float)
        ok = np.isclose(d, r, rtol=1e-9, atol=1e-12) | (np.isnan(d) & np.isnan(r)) | (~np.isfinite(r) & ~np.isfinite(d))
        assert ok.all(), (col, ref[col][~ok], dev[col][~ok])


def test_dropin_anl_matches_reference_mean_and_covariance(registered):
    """[SOLVER] solver = ANL resolves to the drop-in; posterior mean / covariance / samples follow anl.py:40-65
    (host logic on the test double here, the kernels on the GPU in tests/test_gpu_parity.py)."""
    from fitsnap3lib.solvers.solver_factory import solver
    from tests.synth import SOLVE_CASES, synth_system
    a, b, w, t = synth_system(**SOLVE_CASES["well"])
    mean_ref, cov_ref = rd.ref_anl(a, b, w, testing=t, cov_nugget=1e-8)
    pt, cfg = rd.make_reference_context(solver="ANL", extra={"SOLVER": {"cov_nugget": 1e-8, "nsam": 5}})
    s = solver("ANL", pt, cfg)
    assert type(s) is registered["ANL"]
    s.save_files = False
    s.refine = 2
    pt.fitsnap_dict["Testing"] = [bool(v) for v in t]
    s.perform_fit(a=a, b=b, w=w)
    assert np.max(np.abs(s.fit - mean_ref)) < 1e-9 * np.max(np.abs(mean_ref))
    assert np.max(np.abs(s.cov - cov_ref)) < 1e-7 * np.max(np.abs(cov_ref))
    assert s.fit_sam.shape == (5, a.shape[1])


def test_coefficient_file_round_trip(registered):
    """drop-in fit -> reference `_offset` (solver.py:78-102, bzeroflag = 1) -> reference
    `.snapcoeff` text (io/outputs/snap.py:157-188, 18 significant digits) -> parsed back as read_fit does
    (snap.py:90-121): the coefficients survive to the last bit the format keeps."""
    from fitsnap3lib.solvers.solver_factory import solver
    from fitsnap3lib.io.outputs.snap import _to_coeff_string
    rng = np.random.default_rng(21)
    kw = dict(numtypes=2, types="In P", twojmax="4 4", bzeroflag=1)
    pt0, cfg0 = rd.make_reference_context(**kw)
    nc = cfg0.sections["BISPECTRUM"].ncoeff
    cfgs, blocks, vols = _configs(rng, nc, 2, n_cfg=60)
    a, b, w, lists, cfg, pt, calc = rd.ref_scatter(cfgs, blocks, vols, use_factory=True, **kw)
    s = solver("SVD", pt, cfg)
    s.refine = 2
    s.perform_fit()
    raw_fit = s.fit.copy()
    s.error_analysis()                       # applies _offset: one leading 0 per type when bzeroflag = 1
    fit = np.asarray(s.fit, dtype=np.float64).reshape(-1)
    assert fit.shape[0] == 2 * (nc + 1) and fit[0] == 0.0 and fit[nc + 1] == 0.0
    assert np.array_equal(np.delete(fit, [0, nc + 1]), raw_fit)
    text = _to_coeff_string(cfg, fit)
    lines = text.splitlines()
    ntypes, ncoeff1 = (int(v) for v in lines[2].split())
    assert (ntypes, ncoeff1) == (2, nc + 1)
    parsed, pos = [], 3
    for _ in range(ntypes):
        pos += 1                             # element header
        for _j in range(ncoeff1):
            parsed.append(float(lines[pos].split()[0]))
            pos += 1
    parsed = np.array(parsed)
    assert np.max(np.abs(parsed - fit)) <= 1e-17 * np.max(np.abs(fit)) + 1e-300


def test_device_memory_guard_mirrors_the_reference_ram_guard(registered):
    """calculator.py:277-285 aborts when A exceeds half of the RAM unless [MEMORY] override; the drop-in applies the
    same rule to the GPU that will hold the raw blocks and A, b, w."""
    from tests.fake_engine import OracleEngine

    class TinyGpu(OracleEngine):
        def device_memory(self):
            return 1000, 4000           # bytes: anything staged is "too large"

    rng = np.random.default_rng(4)
    kw = dict(numtypes=1, types="Ta", twojmax="4", bzeroflag=0)
    pt0, cfg0 = rd.make_reference_context(**kw)
    nc = cfg0.sections["BISPECTRUM"].ncoeff
    cfgs, blocks, vols = _configs(rng, nc, 1, n_cfg=5)
    ctx = rd.make_reference_context(**kw)
    from fitsnap_b200 import plugin
    plugin.unregister()
    plugin.register(engine=TinyGpu())
    with pytest.raises(MemoryError, match="GPU memory"):
        rd.ref_scatter(cfgs, blocks, vols, use_factory=True, context=ctx, **kw)
    pt, cfg = rd.make_reference_context(**kw)
    cfg.sections["MEMORY"].override = True
    a, *_ = rd.ref_scatter(cfgs, blocks, vols, use_factory=True, context=(pt, cfg), **kw)
    assert a.shape[0] > 0
