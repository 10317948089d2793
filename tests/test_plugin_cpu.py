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


@pytest.mark.parametrize("bz,efs,drop", [(0, (1, 1, 1), False), (1, (1, 1, 0), False), (0, (1, 0, 1), True),
                                         (1, (0, 1, 1), False)])
def test_process_single_after_register_equals_stock(registered, bz, efs, drop):
    """lammps_base.py:101-125: `calculator.process_single(data, i) -> (a, b, w)` through the factory-made drop-in
    equals the stock `_collect_lammps_single` (zero rows of switched-off families, default weights, indices)."""
    rng = np.random.default_rng(5)
    kw = dict(numtypes=2, types="In P", twojmax="6 4", bzeroflag=bz, energy=efs[0], force=efs[1], stress=efs[2])
    pt0, cfg0 = rd.make_reference_context(**kw)
    nc = cfg0.sections["BISPECTRUM"].ncoeff
    cfgs, blocks, vols = _configs(rng, nc, 2, n_cfg=5)
    ref, calc_ref = rd.ref_single(cfgs, blocks, vols, drop_weights=drop, **kw)
    out, calc = rd.ref_single(cfgs, blocks, vols, use_factory=True, drop_weights=drop, **kw)
    assert type(calc) is registered["LammpsSnap"]
    for (a0, b0, w0), (a1, b1, w1) in zip(ref, out):
        assert a0.shape == a1.shape and np.array_equal(a0, a1) and np.array_equal(b0, b1) and np.array_equal(w0, w1)
    assert (calc.shared_index, calc.distributed_index) == (calc_ref.shared_index, calc_ref.distributed_index)


def test_pace_process_single_after_register_equals_stock(registered):
    rng = np.random.default_rng(6)
    nc, nt = 11, 2
    k = nc * nt + nt
    ace = dict(numtypes=nt, ncoeff=nc, bzeroflag=0, bikflag=0, dgradflag=0, blank2J=np.ones(k),
               type_mapping={"In": 1, "P": 2}, rcutfac=[4.0])
    kw = dict(numtypes=nt, types="In P", twojmax="6 6", bzeroflag=0)
    cfgs, blocks, vols = _configs(rng, nc, 2, n_cfg=4)
    ref, _ = rd.ref_single(cfgs, blocks, vols, calculator="LAMMPSPACE", ace=ace, **kw)
    out, calc = rd.ref_single(cfgs, blocks, vols, calculator="LAMMPSPACE", ace=ace, use_factory=True, **kw)
    assert type(calc) is registered["LammpsPace"]
    for r, o in zip(ref, out):
        assert all(np.array_equal(x, y) for x, y in zip(r, o))


def test_bikflag_layout_is_delegated_to_the_stock_collector(registered):
    """with bikflag = 1 LAMMPS emits N energy rows per configuration; the drop-in must not reinterpret that
    block -- it hands the configuration to the reference's own `_collect_lammps` (kept as `_ref_collect_lammps`)."""
    rng = np.random.default_rng(12)
    kw = dict(numtypes=1, types="Ta", twojmax="4", bzeroflag=1,
              extra={"BISPECTRUM": {"bikflag": 1}, "CALCULATOR": {"per_atom_energy": 1}})
    pt0, cfg0 = rd.make_reference_context(**kw)
    nc = cfg0.sections["BISPECTRUM"].ncoeff
    names = ["Ta"]
    cfgs, blocks, vols = [], [], []
    for i in range(4):
        n = int(rng.integers(2, 6))
        cfgs.append(rd.make_config_dict(n, 1, rng, names, group="g", fname="f%d" % i))
        blocks.append(rng.standard_normal((n + 3 * n + 6, nc + 1)))      # bikflag: N energy rows
        vols.append(100.0)
    a_ref, b_ref, w_ref, lists_ref, *_ = rd.ref_scatter(cfgs, blocks, vols, **kw)
    a, b, w, lists, _cfg, pt, calc = rd.ref_scatter(cfgs, blocks, vols, use_factory=True, **kw)
    assert type(calc) is registered["LammpsSnap"] and calc._b200_stock_mode()
    assert np.array_equal(a, a_ref) and np.array_equal(b, b_ref)
    written = np.array(lists_ref["Row_Type"]) != "Energy"     # the reference leaves w of per-atom energy rows unset
    assert np.array_equal(w[written], w_ref[written])
    assert lists["Row_Type"] == lists_ref["Row_Type"]
    assert not hasattr(pt, "fitsnap_b200_device")            # nothing was staged for the device


def test_nonlinear_collectors_survive_register(registered):
    """the network-solver path calls `_collect_lammps_preprocess` / `_collect_lammps_nonlinear`
    (fitsnap.py:161-178); the drop-in classes keep the reference's own methods for them."""
    from fitsnap3lib.calculators.lammps_snap import LammpsSnap as RefSnap
    from fitsnap3lib.calculators.lammps_pace import LammpsPace as RefPace
    for key, ref in (("LammpsSnap", RefSnap), ("LammpsPace", RefPace)):
        cls = registered[key]
        for meth in ("_collect_lammps_nonlinear", "_collect_lammps_preprocess", "_set_computes", "get_width"):
            assert cls.__dict__[meth] is ref.__dict__[meth], (key, meth)
        assert cls.__dict__["_ref_collect_lammps"] is ref.__dict__["_collect_lammps"]


def test_host_mirror_is_lazy_and_in_place_edits_win(registered):
    """The rows assembled on the device reach pt.shared_arrays only when `.array` is read; once it has been read the
    solver must not trust the device copy of that array any more (in-place edits were silently ignored)."""
    from fitsnap3lib.solvers.solver_factory import solver
    from fitsnap_b200.hostmirror import LazyHostMirror
    rng = np.random.default_rng(8)
    kw = dict(numtypes=1, types="Ta", twojmax="4", bzeroflag=0)
    pt0, cfg0 = rd.make_reference_context(**kw)
    nc = cfg0.sections["BISPECTRUM"].ncoeff
    cfgs, blocks, vols = _configs(rng, nc, 1, n_cfg=40)
    a_ref, b_ref, w_ref, lists_ref, *_ = rd.ref_scatter(cfgs, blocks, vols, **kw)
    # drive the drop-in by hand so that nothing reads `.array` behind our back
    from fitsnap3lib.calculators.calculator_factory import calculator as make_calculator
    pt, cfg = rd.make_reference_context(**kw)
    calc = make_calculator("LAMMPSSNAP", pt, cfg)
    calc._prepare_lammps = lambda: calc._set_structure()
    calc.shared_index = calc.distributed_index = 0
    calc.allocate_per_config(cfgs)
    calc.create_a()
    for i, c in enumerate(cfgs):
        rd.FakeLammps.staged_block, rd.FakeLammps.staged_volume = blocks[i], vols[i]
        calc.process_configs(c, i)
    calc.collect_distributed_lists()
    ma, mw = pt.shared_arrays["a"], pt.shared_arrays["w"]
    assert isinstance(ma, LazyHostMirror) and ma.pending and not ma.exposed
    s = solver("SVD", pt, cfg)
    s.refine = 2
    s.perform_fit()
    assert ma.pending and not ma.exposed                      # the fit read the device rows, A never came back
    x0 = s.fit.copy()
    mw.array[:] *= np.where(np.array(lists_ref["Row_Type"]) == "Energy", 7.0, 1.0)   # in-place edit of the weights
    assert mw.exposed and ma.pending
    s.perform_fit()
    x_ref, _ = rd.ref_fit("SVD", a_ref, b_ref, pt.shared_arrays["w"].array, testing=np.array(lists_ref["Testing"]))
    assert np.max(np.abs(s.fit - x_ref)) < 1e-9 * np.max(np.abs(x_ref))
    assert np.max(np.abs(s.fit - x0)) > 1e-6 * np.max(np.abs(x0))
    assert np.array_equal(pt.shared_arrays["a"].array, a_ref) and not ma.pending      # reading A materialises it


def test_error_analysis_override_equals_the_stock_table_and_offsets(registered):
    """`FitSnap.perform_fit` calls `solver.error_analysis()` (fitsnap.py:213-220): the drop-in's override fills
    `solver.errors` like solver.py:368-429 does (device sums instead of DataFrame(a)) and applies `_offset`."""
    from fitsnap3lib.solvers.solver_factory import solver
    from fitsnap3lib.solvers.solver import Solver
    rng = np.random.default_rng(31)
    kw = dict(numtypes=2, types="In P", twojmax="4 4", bzeroflag=1)
    pt0, cfg0 = rd.make_reference_context(**kw)
    nc = cfg0.sections["BISPECTRUM"].ncoeff
    cfgs, blocks, vols = _configs(rng, nc, 2, n_cfg=50)
    a, b, w, lists, cfg, pt, calc = rd.ref_scatter(cfgs, blocks, vols, use_factory=True, **kw)
    s = solver("SVD", pt, cfg)
    s.refine = 2
    s.perform_fit()
    raw = s.fit.copy()
    s.error_analysis()
    dev_err, dev_fit = s.errors.copy(), np.asarray(s.fit).copy()
    assert dev_fit.shape == (2 * (nc + 1), 1) and dev_fit[0, 0] == 0.0        # _offset applied (bzeroflag = 1)
    df = s.df
    assert {"truths", "preds", "weights", "Groups", "Row_Type", "Testing"} <= set(df.columns) and len(df) == a.shape[0]
    s.fit = raw
    Solver.error_analysis(s)                                                  # the reference's implementation
    ref_err = s.errors
    assert list(dev_err.index) == list(ref_err.index) and list(dev_err.columns) == list(ref_err.columns)
    assert np.array_equal(dev_err["ncount"].values, ref_err["ncount"].values)
    for col in ("mae", "rmse", "rsq"):
        r, d = ref_err[col].values.astype(float), dev_err[col].values.astype(float)
        ok = np.isclose(d, r, rtol=1e-9, atol=1e-12) | (np.isnan(d) & np.isnan(r)) | (~np.isfinite(r) & ~np.isfinite(d))
        assert ok.all(), col
    assert np.array_equal(np.asarray(s.fit), dev_fit)
