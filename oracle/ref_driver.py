"""TEST INFRASTRUCTURE: drive the UNMODIFIED reference (`/root/reference`) in the build
container so that (a) `oracle/linear_fit.py` can be pinned to it and (b) golden
fixtures can be generated (`oracle/make_golden.py`).

The reference needs the `lammps` Python module (parallel_tools.py:35) which is not
installed here; `install_fake_lammps()` registers a stand-in whose `extract_compute`
hands back a synthetic `(1+3N+6) x (K_raw+1)` block staged by the caller.  The
reference classes themselves are imported and executed unmodified.

NOT usable on the GPU box (no /root/reference there) -- anything that must run there
uses the committed fixtures under tests/golden/ instead.
"""
from __future__ import annotations

import ctypes
import os
import sys
import types
from types import SimpleNamespace

import numpy as np

REFERENCE_ROOT = os.environ.get("FITSNAP_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "fitsnap3lib"))


class _FakeNumpyView:
    def __init__(self, owner):
        self._o = owner

    def extract_atom(self, name, nelem=None, dim=None, **_kw):
        o = self._o
        if name == "id":
            return np.arange(1, o._natoms + 1, dtype=np.int32)
        if name == "type":
            return np.asarray(o._types, dtype=np.int32)
        if name == "x":
            return np.asarray(o._x, dtype=np.float64).reshape(o._natoms, 3)
        raise KeyError(name)


class FakeLammps:
    """Minimal stand-in for `lammps.lammps` (only what lammps_base.py / lammps_snap.py /
    lammps_pace.py / parallel_tools.py call).  The synthetic compute array and cell
    volume for the NEXT `run 0` are staged on the class by the driver."""
    staged_block = None
    staged_volume = 1.0
    has_exceptions = True
    installed_packages = []

    def __init__(self, *a, **kw):
        self._natoms = 0
        self._types = []
        self._x = []
        self._block = None
        self.numpy = _FakeNumpyView(self)

    def command(self, cmd):
        # lammps_pace.py:47-52 creates atoms one `create_atoms <type> single x y z` command at a time
        parts = str(cmd).split()
        if len(parts) >= 6 and parts[0] == "create_atoms" and parts[2] == "single":
            self._natoms += 1
            self._types.append(int(parts[1]))
            self._x.extend(float(v) for v in parts[3:6])
        elif parts and parts[0] == "clear":
            self._natoms, self._types, self._x = 0, [], []
        return None

    def close(self):
        return None

    def version(self):
        return 20250612

    def create_atoms(self, n, id=None, type=None, x=None, v=None, image=None, shrinkexceed=False,
                     atomid=None, atype=None):
        t = type if type is not None else atype
        self._natoms = int(n)
        self._types = [int(v_) for v_ in t]
        self._x = [float(v_) for v_ in x]

    def get_natoms(self):
        return self._natoms

    def get_thermo(self, what):
        assert what == "vol"
        return float(FakeLammps.staged_volume)

    def extract_compute(self, _name, _style, _rtype):
        # keep a reference so the memory outlives the numpy view the reference builds
        self._block = np.ascontiguousarray(FakeLammps.staged_block, dtype=np.float64)
        self._rowptr = self._block.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        return ctypes.pointer(self._rowptr)


def install_fake_lammps():
    if "lammps" not in sys.modules or not hasattr(sys.modules["lammps"], "_fitsnap_b200_fake"):
        mod = types.ModuleType("lammps")
        mod.lammps = FakeLammps
        mod._fitsnap_b200_fake = True
        sys.modules["lammps"] = mod
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def make_reference_context(solver="SVD", numtypes=1, twojmax="6", bzeroflag=0, types="Ta",
                           energy=1, force=1, stress=1, ridge_alpha=None, ridge_local=0,
                           lasso_alpha=None, lasso_max_iter=2000, apply_transpose=0, extra=None):
    """Build (pt, cfg) for the unmodified reference in stubs (no-MPI) mode
    (SURVEY 8c recipe; io/input.py:141-151 dict mode)."""
    install_fake_lammps()
    from fitsnap3lib.parallel_tools import ParallelTools
    from fitsnap3lib.io.input import Config
    ones = " ".join(["1.0"] * numtypes)
    halves = " ".join(["0.5"] * numtypes)
    d = {
        "BISPECTRUM": {"numTypes": numtypes, "twojmax": twojmax, "rcutfac": 4.67637, "rfac0": 0.99363,
                       "rmin0": 0.0, "wj": ones, "radelem": halves, "type": types,
                       "wselfallflag": 0, "chemflag": 0, "bzeroflag": bzeroflag, "quadraticflag": 0},
        "CALCULATOR": {"calculator": "LAMMPSSNAP", "energy": energy, "force": force, "stress": stress},
        "SOLVER": {"solver": solver},
        "EXTRAS": {"apply_transpose": apply_transpose},
        "REFERENCE": {"units": "metal", "atom_style": "atomic", "pair_style": "zero 10.0",
                      "pair_coeff": "* *"},
    }
    if ridge_alpha is not None:
        d["RIDGE"] = {"alpha": ridge_alpha, "local_solver": ridge_local}
    if lasso_alpha is not None:
        d["LASSO"] = {"alpha": lasso_alpha, "max_iter": lasso_max_iter}
    if extra:
        for sec, kv in extra.items():
            d.setdefault(sec, {}).update(kv)
    pt = ParallelTools()
    cfg = Config(pt, d, arguments_lst=["--overwrite"])
    return pt, cfg


def reference_solver(name, pt, cfg):
    """Instantiate the REFERENCE solver class directly (not through the factory, which
    would return a drop-in class once `fitsnap_b200.plugin` is registered)."""
    install_fake_lammps()
    if name.upper() == "SVD":
        from fitsnap3lib.solvers.svd import SVD as cls
    elif name.upper() == "RIDGE":
        from fitsnap3lib.solvers.ridge import RIDGE as cls
    elif name.upper() == "LASSO":
        from fitsnap3lib.solvers.lasso import LASSO as cls
    elif name.upper() == "ANL":
        from fitsnap3lib.solvers.anl import ANL as cls
    else:
        raise KeyError(name)
    return cls(name, pt, cfg)


def ref_fit(name, a, b, w, testing=None, **ctx):
    """Run the reference `perform_fit` (svd.py:18 / ridge.py:11 / lasso.py:15) on host arrays."""
    pt, cfg = make_reference_context(solver=name, **ctx)
    s = reference_solver(name, pt, cfg)
    if name.upper() == "LASSO" or testing is not None:
        # lasso.py:15 takes no array arguments, and the explicit-array branch of
        # svd.py:46 / ridge.py:39 does not mask `w` (it raises a broadcast error as soon as a
        # test row exists), so train/test splits go through pt.shared_arrays + fitsnap_dict,
        # the route fitsnap.py:190-220 itself uses.
        n, k = a.shape
        pt.create_shared_array("a", n, k)
        pt.create_shared_array("b", n)
        pt.create_shared_array("w", n)
        pt.shared_arrays["a"].array[:] = a.reshape(pt.shared_arrays["a"].array.shape)
        pt.shared_arrays["b"].array[:] = b
        pt.shared_arrays["w"].array[:] = w
        pt.fitsnap_dict["Testing"] = [bool(t) for t in testing] if testing is not None else [False] * n
        s.perform_fit()
    else:
        s.perform_fit(a=a, b=b, w=w, trainall=True)
    return np.array(s.fit, dtype=np.float64), s


def ref_anl(a, b, w, testing=None, cov_nugget=0.0, nsam=0):
    """Run the reference ANL.perform_fit (anl.py:13-67) in a scratch directory (it drops covariance.npy and
    mean.npy into the working directory, anl.py:60-61).  Returns (mean, cov)."""
    import tempfile
    pt, cfg = make_reference_context(solver="ANL", extra={"SOLVER": {"cov_nugget": cov_nugget, "nsam": nsam}})
    s = reference_solver("ANL", pt, cfg)
    # the explicit-array branch (anl.py:28) forgets to mask `w`, like svd.py:46: go through pt.shared_arrays
    n, k = a.shape
    pt.create_shared_array("a", n, k)
    pt.create_shared_array("b", n)
    pt.create_shared_array("w", n)
    pt.shared_arrays["a"].array[:] = a.reshape(pt.shared_arrays["a"].array.shape)
    pt.shared_arrays["b"].array[:] = b
    pt.shared_arrays["w"].array[:] = w
    pt.fitsnap_dict["Testing"] = [bool(t) for t in testing] if testing is not None else [False] * n
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            s.perform_fit()
        finally:
            os.chdir(cwd)
    return np.array(s.fit, dtype=np.float64), np.array(s.cov, dtype=np.float64)


def make_config_dict(natoms, numtypes, rng, type_names, group="G", fname="cfg", eweight=1.0,
                     fweight=1.0, vweight=1.0, test_bool=False):
    """One scraped-configuration dict in the format the reference calculators consume
    (keys read by lammps_base.py:168-217 and lammps_snap.py:393-556)."""
    lat = np.diag(rng.uniform(3.0, 9.0, 3))
    return {
        "NumAtoms": natoms,
        "Energy": float(rng.normal(-5.0 * natoms, 1.0)),
        "AtomTypes": [type_names[int(t)] for t in rng.integers(0, numtypes, natoms)],
        "Positions": rng.uniform(0, 3.0, (natoms, 3)),
        "Forces": rng.normal(0, 1.0, (natoms, 3)),
        "Stress": (lambda s: 0.5 * (s + s.T))(rng.normal(0, 1e4, (3, 3))),
        "Lattice": lat,
        "Group": group, "File": fname,
        "eweight": eweight, "fweight": fweight, "vweight": vweight,
        "test_bool": test_bool,
    }


def ref_scatter(configs, blocks, volumes, calculator="LAMMPSSNAP", ace=None, use_factory=False, context=None, **ctx):
    """Run the unmodified reference calculator (`LammpsSnap`/`LammpsPace`
    allocate_per_config -> create_a -> process_configs -> collect_distributed_lists;
    fitsnap.py:134-188) over synthetic compute blocks.  Returns (A, b, w, fitsnap_dict lists, cfg)."""
    pt, cfg = context if context is not None else make_reference_context(**ctx)
    if calculator == "LAMMPSPACE":
        # SURVEY 8c ACE caveat: the [ACE] section cannot be constructed without mpi4py;
        # inject exactly the attributes lammps_pace.py reads.
        cfg.sections["CALCULATOR"].calculator = "LAMMPSPACE"
        cfg.sections["ACE"] = SimpleNamespace(**ace)
    from fitsnap3lib.calculators.lammps_snap import LammpsSnap
    from fitsnap3lib.calculators.lammps_pace import LammpsPace
    if use_factory:     # through calculators/calculator_factory.py (returns a registered drop-in, if any)
        from fitsnap3lib.calculators.calculator_factory import calculator as make_calculator
        calc = make_calculator(calculator, pt, cfg)
    else:
        cls = LammpsPace if calculator == "LAMMPSPACE" else LammpsSnap
        calc = cls(calculator, pt, cfg)
    calc._prepare_lammps = lambda: calc._set_structure()   # skip compute/pair set-up strings
    if calculator == "LAMMPSPACE":
        calc._set_box = lambda: calc._set_box_helper(numtypes=ace["numtypes"])
    calc.shared_index = 0
    calc.distributed_index = 0
    calc.allocate_per_config(configs)
    calc.create_a()
    for i, c in enumerate(configs):
        FakeLammps.staged_block = blocks[i]
        FakeLammps.staged_volume = volumes[i]
        calc.process_configs(c, i)
    calc.collect_distributed_lists()
    a = np.array(pt.shared_arrays["a"].array, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(len(pt.shared_arrays["b"].array), -1)
    b = np.array(pt.shared_arrays["b"].array, dtype=np.float64)
    w = np.array(pt.shared_arrays["w"].array, dtype=np.float64)
    lists = {k: list(v) for k, v in pt.fitsnap_dict.items() if isinstance(v, list)}
    return a, b, w, lists, cfg, pt, calc


def ref_single(configs, blocks, volumes, calculator="LAMMPSSNAP", ace=None, use_factory=False, context=None,
               drop_weights=False, **ctx):
    """Run `process_single` (lammps_base.py:101-125 -> `_collect_lammps_single`, lammps_snap.py:224-389 /
    lammps_pace.py:197-366) of the unmodified reference calculator -- or, with `use_factory`, of whatever class the
    reference's factory returns -- over synthetic compute blocks.  Returns a list of (a, b, w) per configuration and
    the calculator.  `drop_weights` removes the eweight / fweight / vweight keys (they then default to 1.0)."""
    pt, cfg = context if context is not None else make_reference_context(**ctx)
    if calculator == "LAMMPSPACE":
        cfg.sections["CALCULATOR"].calculator = "LAMMPSPACE"
        cfg.sections["ACE"] = SimpleNamespace(**ace)
    from fitsnap3lib.calculators.lammps_snap import LammpsSnap
    from fitsnap3lib.calculators.lammps_pace import LammpsPace
    if use_factory:
        from fitsnap3lib.calculators.calculator_factory import calculator as make_calculator
        calc = make_calculator(calculator, pt, cfg)
    else:
        cls = LammpsPace if calculator == "LAMMPSPACE" else LammpsSnap
        calc = cls(calculator, pt, cfg)
    calc._prepare_lammps = lambda: calc._set_structure()
    if calculator == "LAMMPSPACE":
        calc._set_box = lambda: calc._set_box_helper(numtypes=ace["numtypes"])
    calc.shared_index = 0
    calc.distributed_index = 0
    out = []
    for i, c in enumerate(configs):
        if drop_weights:
            c = {k_: v for k_, v in c.items() if k_ not in ("eweight", "fweight", "vweight")}
        FakeLammps.staged_block = np.array(blocks[i], dtype=np.float64)    # the reference divides the block in place
        FakeLammps.staged_volume = volumes[i]
        a, b, w = calc.process_single(c, i)
        out.append((np.array(a, dtype=np.float64), np.array(b, dtype=np.float64), np.array(w, dtype=np.float64)))
    return out, calc
