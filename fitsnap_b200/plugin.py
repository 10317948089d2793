"""Register the B200 drop-ins with FitSNAP's own plugin factories.

    import fitsnap_b200.plugin as plugin
    plugin.register()                      # before FitSnap(...) is constructed
    fs = FitSnap(infile_or_dict, comm)     # [SOLVER] solver = SVD | RIDGE | LASSO | ANL, [CALCULATOR] calculator = LAMMPSSNAP | LAMMPSPACE

How discovery works in the reference (and why this is all that is needed):
  * solvers/solver_factory.py:25-34 walks `Solver.__subclasses__()` comparing lower-cased class
    names, WITHOUT a break -- the last matching direct subclass wins.  Defining `class SVD(.., Solver)`
    after the factory module was imported therefore replaces the stock solver for `solver = SVD`.
  * calculators/calculator_factory.py:22-38 does the same two levels deep
    (`Calculator.__subclasses__()` -> their `__subclasses__()`), so the drop-in calculators derive
    from `LammpsBase` directly and borrow the LAMMPS set-up methods of the stock classes (those are
    out of the hot path and stay the reference's code, executed unmodified).
No reference file is edited or monkey-patched; `unregister()` makes the stock classes win again.
"""
from __future__ import annotations

_registered = {}


def register(engine=None):
    """Define the drop-in classes under the reference's bases.  `engine` (optional) is handed to the
    created objects instead of the process-wide default Engine (tests inject a stand-in)."""
    if _registered:
        return dict(_registered)
    import fitsnap3lib.solvers.solver_factory  # noqa: F401  (imports every stock solver first)
    import fitsnap3lib.calculators.calculator_factory  # noqa: F401
    from fitsnap3lib.solvers.solver import Solver
    from fitsnap3lib.calculators.lammps_base import LammpsBase
    from fitsnap3lib.calculators.lammps_snap import LammpsSnap as RefSnap
    from fitsnap3lib.calculators.lammps_pace import LammpsPace as RefPace
    from . import calculators as bc
    from . import solvers as bs

    def solver_class(name, mirror):
        def __init__(self, name_, pt, config):
            Solver.__init__(self, name_, pt, config, linear=True)     # solver.py:19-38 (runs _checks)
            mirror.__init__(self, name_, pt, config)
            self.engine = engine
        return type(name, (mirror, Solver), {"__init__": __init__, "__doc__": mirror.__doc__,
                                             "__module__": __name__, "_offset": Solver._offset})

    def calculator_class(name, mixin, ref):
        def __init__(self, name_, pt, config):
            LammpsBase.__init__(self, name_, pt, config)               # lammps_base.py:8-13
            self._data, self._i, self._lmp, self._row_index = {}, 0, None, 0
            self.pt.check_lammps()
            self._b200_engine = engine
        body = {"__init__": __init__, "__doc__": mixin.__doc__, "__module__": __name__}
        # Everything the stock class defines stays the reference's own code, executed unmodified: the LAMMPS
        # set-up (descriptor evaluation side: get_width, _prepare_lammps, _set_box, _create_atoms, _set_computes,
        # _create_spins, _create_charge, _set_variables) and the nonlinear / preprocessing collectors
        # (_collect_lammps_nonlinear, _collect_lammps_preprocess -- the network solvers' path, fitsnap.py:161-178).
        # Only the two linear collectors are replaced by the mixin; the stock versions are kept under `_ref_*`
        # for the layouts the batched scatter does not cover (bikflag, dgradflag, per_atom_energy).
        replaced = ("__init__", "_collect_lammps", "_collect_lammps_single")
        for meth, fn in ref.__dict__.items():
            if meth.startswith("__") and meth.endswith("__"):
                continue
            if meth in replaced:
                body["_ref_" + meth.lstrip("_")] = fn
            else:
                body[meth] = fn
        return type(name, (mixin, LammpsBase), body)

    _registered["SVD"] = solver_class("SVD", bs.SVD)
    _registered["RIDGE"] = solver_class("RIDGE", bs.RIDGE)
    if hasattr(bs, "LASSO"):
        _registered["LASSO"] = solver_class("LASSO", bs.LASSO)
    if hasattr(bs, "ANL"):
        _registered["ANL"] = solver_class("ANL", bs.ANL)
    _registered["LammpsSnap"] = calculator_class("LammpsSnap", bc.SnapCollectMixin, RefSnap)
    _registered["LammpsPace"] = calculator_class("LammpsPace", bc.PaceCollectMixin, RefPace)
    return dict(_registered)


def unregister():
    """Drop the references so the stock classes are found again (subclass lists are weak)."""
    import gc
    _registered.clear()
    gc.collect()


def is_registered():
    return bool(_registered)
