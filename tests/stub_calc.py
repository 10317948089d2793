"""TEST HELPER: a stand-in for the reference's `LammpsBase` plumbing (`_lmp`, `_extract_atom_*`, `config`,
`pt`) so that the collector mixins of fitsnap_b200.calculators can be driven from committed fixtures on a box
without fitsnap3lib / LAMMPS (the GPU box).  Only what `_CollectMixin` touches is provided."""
import ctypes
from types import SimpleNamespace

import numpy as np

from fitsnap_b200.calculators import PaceCollectMixin, SnapCollectMixin


class _Lmp:
    def __init__(self):
        self.block, self.volume = None, 1.0

    def get_thermo(self, what):
        assert what == "vol"
        return float(self.volume)

    def extract_compute(self, _name, _style, _rtype):
        self._keep = np.ascontiguousarray(self.block, dtype=np.float64)
        self._rowptr = self._keep.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        return ctypes.pointer(self._rowptr)


class _StubBase:
    def __init__(self, engine, section_name, sec, calc, pt=None):
        self.config = SimpleNamespace(sections={section_name: sec, "CALCULATOR": calc,
                                                "MEMORY": SimpleNamespace(override=False)})
        self.pt = pt or SimpleNamespace(fitsnap_dict={}, shared_arrays={}, single_print=lambda *a: None)
        self._lmp = _Lmp()
        self._b200_engine = engine
        self._data, self._i = {}, 0
        self.shared_index = 0
        self.distributed_index = 0
        self._types = None

    def _extract_atom_ids(self, n):
        return 1 + np.arange(n)

    def _extract_atom_types(self, n):
        return np.asarray(self._types[:n])

    def process_single(self, data, block, volume, types):
        """lammps_base.py:101-125 without LAMMPS: the compute array is handed in."""
        self._data, self._types = data, types
        self._lmp.block, self._lmp.volume = block, volume
        return self._collect_lammps_single()


class StubSnap(SnapCollectMixin, _StubBase):
    pass


class StubPace(PaceCollectMixin, _StubBase):
    pass


def fixture_configs(g):
    """Per-configuration (data dict, block, volume, lammps types) of a tests/golden/single_*.npz fixture."""
    nat = g["natoms"]
    roff = np.concatenate([[0], np.cumsum(7 + 3 * nat.astype(np.int64))])
    aoff = np.concatenate([[0], np.cumsum(nat.astype(np.int64))])
    inv = {1: "In", 2: "P"}
    out = []
    for c in range(len(nat)):
        types = g["atom_type_index"][aoff[c]:aoff[c + 1]]
        d = {"NumAtoms": int(nat[c]), "Energy": float(g["energy"][c]),
             "Forces": g["forces"][3 * aoff[c]:3 * aoff[c + 1]].reshape(-1, 3), "Stress": g["stress"][c],
             "AtomTypes": [inv[int(t)] for t in types], "Group": "s", "File": "one%d" % c, "test_bool": False}
        if not bool(g["weights_dropped"]):
            d.update(eweight=float(g["eweight"][c]), fweight=float(g["fweight"][c]), vweight=float(g["vweight"][c]))
        out.append((d, g["raw"][roff[c]:roff[c + 1]], float(g["volume"][c]), types))
    return out


def make_stub(engine, g, pace=False):
    sec = SimpleNamespace(numtypes=int(g["numtypes"]), ncoeff=int(g["ncoeff"]), bzeroflag=int(g["bzeroflag"]),
                          blank2J=np.array(g["blank2j"]), type_mapping={"In": 1, "P": 2}, bikflag=0, dgradflag=0,
                          chemflag=0, wselfallflag=0)
    calc = SimpleNamespace(energy=bool(g["use_energy"]), force=bool(g["use_force"]), stress=bool(g["use_stress"]),
                           per_atom_energy=False, nonlinear=False)
    cls = StubPace if pace else StubSnap
    return cls(engine, "ACE" if pace else "BISPECTRUM", sec, calc)
