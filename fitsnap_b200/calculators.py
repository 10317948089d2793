"""Host-side mirror of FitSNAP's linear Calculator plugins (LammpsSnap / LammpsPace), A-assembly part.

Reference interface being mirrored (paths relative to the FitSNAP tree):
    calculators/lammps_snap.py:391-556  LammpsSnap._collect_lammps   (rows of A, b, w + row metadata)
    calculators/lammps_pace.py:369-509  LammpsPace._collect_lammps
    calculators/lammps_snap.py:224-389  LammpsSnap._collect_lammps_single  (process_single, lammps_base.py:101-125)
    calculators/lammps_pace.py:197-366  LammpsPace._collect_lammps_single
    calculators/calculator.py:261-299   Calculator.create_a (linear branch: a_len, shared arrays)
    calculators/calculator.py:311-326   Calculator.collect_distributed_lists

LAMMPS itself (descriptor evaluation, `run 0`) is out of scope and untouched: the mirror takes the
compute array exactly where the reference takes it (`_extract_compute_np`, lammps_base.py:280-307).
What changes is what happens next.  The reference turns each block into rows with numpy, one
configuration at a time; here `_collect_lammps` only STAGES the block (one memcpy into a pinned
buffer -- the LAMMPS array is a view that dies with the instance) and fills the per-row metadata
lists; `flush()` then ships every staged block H2D once and one kernel writes all rows.

`BlockCollector` is reference-independent (usable and tested without fitsnap3lib);
`SnapCollectMixin` / `PaceCollectMixin` carry the `_collect_lammps` bodies that
`fitsnap_b200.plugin.register()` grafts under the reference's `LammpsBase`.
"""
from __future__ import annotations

import numpy as np
import torch

from .assembly import descriptor_width, pack_configs, rows_per_config, type_fractions
from .hostmirror import LazyHostMirror


def extract_compute_array(lmp, name, shape):
    """View of a LAMMPS global compute array as numpy (what `_extract_compute_np(lmp, name, 0, 2, shape)`
    returns in the reference, lammps_base.py:280-307): `extract_compute(name, 0, 2)` hands back a
    double** whose first row pointer addresses the contiguous block."""
    import ctypes
    ptr = lmp.extract_compute(name, 0, 2)
    first_row = ptr.contents
    n = int(np.prod(shape))
    buf = ctypes.cast(first_row, ctypes.POINTER(ctypes.c_double * n)).contents
    return np.frombuffer(buf, dtype=np.float64).reshape(shape)


class BlockCollector:
    """Accumulates raw compute blocks + per-configuration scalars; `flush()` assembles A, b, w."""

    def __init__(self, engine, numtypes, ncoeff, bzeroflag, blank2j, type_mapping, energy=True, force=True,
                 stress=True, scrub_nonfinite=False, capacity_rows=0):
        self.engine = engine
        self.numtypes, self.ncoeff, self.bzeroflag = int(numtypes), int(ncoeff), bool(bzeroflag)
        self.blank2j = np.ascontiguousarray(blank2j, dtype=np.float64)
        self.type_mapping = type_mapping
        self.rows = (bool(energy), bool(force), bool(stress))
        self.scrub = bool(scrub_nonfinite)
        self.kraw = self.numtypes * self.ncoeff
        self.k = descriptor_width(ncoeff, numtypes, bzeroflag)
        self._cap = 0
        self._raw = None            # pinned (rows, kraw+1) staging
        self._nraw = 0
        self._reserve(max(int(capacity_rows), 1024))
        self.reset()

    def reset(self):
        self._nraw = 0
        self.natoms, self.volume, self.energy, self.forces, self.stress = [], [], [], [], []
        self.ew, self.fw, self.vw, self.tf = [], [], [], []
        self.n_out = 0

    def _reserve(self, rows):
        if rows <= self._cap:
            return
        new_cap = max(rows, 2 * self._cap)
        buf = torch.empty((new_cap, self.kraw + 1), dtype=torch.float64)
        try:
            buf = buf.pin_memory()
        except RuntimeError:
            pass
        if self._raw is not None and self._nraw:
            buf[:self._nraw].copy_(self._raw[:self._nraw])
        self._raw, self._cap = buf, new_cap

    def add(self, block, natoms, volume, energy, forces, stress, eweight, fweight, vweight, atom_types=None):
        """Stage one configuration (lammps_snap.py:393-425 gathers exactly these quantities)."""
        n = int(natoms)
        block = np.asarray(block)
        assert block.shape == (1 + 3 * n + 6, self.kraw + 1), (block.shape, n, self.kraw)
        self._reserve(self._nraw + block.shape[0])
        self._raw[self._nraw:self._nraw + block.shape[0]].numpy()[...] = block
        self._nraw += block.shape[0]
        self.natoms.append(n)
        self.volume.append(float(volume))
        self.energy.append(float(energy))
        self.forces.append(np.asarray(forces, dtype=np.float64).reshape(-1))
        self.stress.append(np.asarray(stress, dtype=np.float64).reshape(3, 3))
        self.ew.append(float(eweight))
        self.fw.append(float(fweight))
        self.vw.append(float(vweight))
        if not self.bzeroflag:
            self.tf.append(type_fractions(atom_types, self.type_mapping, self.numtypes))
        e, f, s = self.rows
        nrows = rows_per_config(n, e, f, s)
        self.n_out += nrows
        return nrows

    def flush(self, first_row=0, out=None):
        """One H2D of everything staged + one scatter launch.  Returns (A, b, w, nonfinite, batch) on
        the device; `out` = (A, b, w) device tensors to write into (rows first_row ...)."""
        e, f, s = self.rows
        ncfg = len(self.natoms)
        batch = pack_configs(self.engine, self._raw[:self._nraw].numpy(), np.asarray(self.natoms, dtype=np.int32),
                             self.volume, self.energy,
                             np.concatenate(self.forces) if ncfg else np.zeros(0),
                             np.stack(self.stress) if ncfg else np.zeros((0, 3, 3)),
                             self.ew, self.fw, self.vw,
                             None if self.bzeroflag else (np.stack(self.tf) if ncfg else np.zeros((0, self.numtypes))),
                             self.blank2j, self.numtypes, self.ncoeff, energy=e, force=f, stress=s,
                             bzeroflag=self.bzeroflag, scrub_nonfinite=self.scrub, first_row=first_row)
        A, b, w, bad = self.engine.scatter(batch, *(out or (None, None, None)))
        return A, b, w, bad, batch


def row_metadata(natoms, lmp_types, energy, force, stress, group, fname, test_bool, with_atom_type=True):
    """The per-row lists `_collect_lammps` appends to pt.fitsnap_dict (lammps_snap.py:478-486,
    513-520, 545-554): Row_Type, Atom_I, Atom_Type, Groups, Configs, Testing."""
    rt, ai, at = [], [], []
    n = int(natoms)
    if energy:
        rt += ["Energy"]
        ai += [0]
        at += [0]
    if force:
        rt += ["Force"] * (3 * n)
        ai += [int(np.floor(i / 3)) for i in range(3 * n)]
        at += [t for t in lmp_types for _ in range(3)]
    if stress:
        rt += ["Stress"] * 6
        ai += [0] * 6
        at += [0] * 6
    m = len(rt)
    meta = {"Row_Type": rt, "Atom_I": ai, "Groups": ["{}".format(group)] * m, "Configs": ["{}".format(fname)] * m,
            "Testing": [bool(test_bool)] * m}
    if with_atom_type:
        meta["Atom_Type"] = at
    return meta


class _CollectMixin:
    """`_collect_lammps` for the linear path: stage the block, fill metadata, advance the indices.
    Subclasses set SECTION ('BISPECTRUM' | 'ACE'), COMPUTE ('snap' | 'pace') and NAN_POLICY."""
    SECTION = "BISPECTRUM"
    COMPUTE = "snap"
    SCRUB = False
    WITH_ATOM_TYPE = True

    def _b200_collector(self):
        col = getattr(self, "_b200_col", None)
        if col is None:
            sec = self.config.sections[self.SECTION]
            calc = self.config.sections["CALCULATOR"]
            from .engine import default_engine
            eng = getattr(self, "_b200_engine", None) or default_engine()
            col = BlockCollector(eng, sec.numtypes, sec.ncoeff, sec.bzeroflag, np.asarray(sec.blank2J, dtype=np.float64),
                                 sec.type_mapping, calc.energy, calc.force, calc.stress, scrub_nonfinite=self.SCRUB)
            self._b200_col = col
            self._b200_first_row = self.shared_index
        return col

    def _b200_stock_mode(self):
        """True for the layouts the batched scatter does not cover: per-atom energy rows (`bikflag`, N energy
        rows per configuration, lammps_snap.py:409-411, 430-433), descriptor-gradient output (`dgradflag`),
        `[CALCULATOR] per_atom_energy` and the nonlinear (network) branch.  Those configurations are handed to
        the reference's own `_collect_lammps*` (kept on the class by plugin.register) -- never reinterpreted."""
        sec = self.config.sections[self.SECTION]
        calc = self.config.sections["CALCULATOR"]
        return bool(getattr(sec, "bikflag", 0) or getattr(sec, "dgradflag", 0) or
                    getattr(calc, "per_atom_energy", False) or getattr(calc, "nonlinear", False))

    def _b200_stock(self, name):
        fn = getattr(self, "_ref_" + name.lstrip("_"), None)
        if fn is None:
            raise NotImplementedError(
                "fitsnap_b200: bikflag / dgradflag / per_atom_energy / nonlinear layouts are assembled by the "
                "reference's own %s, which is only available after fitsnap_b200.plugin.register()" % name)
        return fn()

    def _b200_gather_block(self):
        """What both `_collect_lammps` variants read from LAMMPS (lammps_snap.py:393-428)."""
        d = self._data
        n = d["NumAtoms"]
        sec = self.config.sections[self.SECTION]
        calc = self.config.sections["CALCULATOR"]
        lmp_ids = self._extract_atom_ids(n)
        lmp_types = self._extract_atom_types(n)
        assert np.all(lmp_ids == 1 + np.arange(n)), \
            "LAMMPS seems to have lost atoms\nGroup and configuration: {} {}".format(d["Group"], d["File"])
        volume = self._lmp.get_thermo("vol")
        block = extract_compute_array(self._lmp, self.COMPUTE, (1 + 3 * n + 6, sec.ncoeff * sec.numtypes + 1))
        if not np.isfinite(block).all():
            if not self.SCRUB:      # lammps_snap.py:426-428
                raise ValueError("Nan in computed data of file {} in group {}".format(d["File"], d["Group"]))
            self.pt.single_print("! WARNING! applying np.nan_to_num()")     # lammps_pace.py:399-401
        if calc.energy:
            self._warn_if_no_neighbors(block, n, sec, d)
        return d, n, sec, calc, lmp_types, volume, block

    def _collect_lammps_single(self):
        """`process_single` body (lammps_base.py:101-125 -> lammps_snap.py:224-389 / lammps_pace.py:197-366):
        rows of ONE configuration returned as host `(a, b, w)`, shared arrays untouched.  Layout quirk kept from
        the reference: `a` always has the energy row and the 3N force rows (plus the 6 virial rows iff
        `[CALCULATOR] stress`), and the rows of a switched-off family stay zero (`irow` advances regardless,
        lammps_snap.py:341, 364); missing `eweight`/`fweight`/`vweight` keys default to 1.0 (:337, 360, 383).
        The arithmetic runs in the same scatter kernel as the batched path."""
        if self._b200_stock_mode():
            return self._b200_stock("_collect_lammps_single")
        d, n, sec, calc, lmp_types, volume, block = self._b200_gather_block()
        col = getattr(self, "_b200_single_col", None)
        if col is None:
            from .engine import default_engine
            eng = getattr(self, "_b200_engine", None) or default_engine()
            col = BlockCollector(eng, sec.numtypes, sec.ncoeff, sec.bzeroflag, np.asarray(sec.blank2J, dtype=np.float64),
                                 sec.type_mapping, calc.energy, calc.force, calc.stress, scrub_nonfinite=self.SCRUB,
                                 capacity_rows=1 + 3 * n + 6)
            self._b200_single_col = col
        col.reset()
        e, f, s = col.rows
        na = 1 + 3 * n + (6 if s else 0)
        a = np.zeros((na, col.k))
        b = np.zeros(na)
        w = np.zeros(na)
        nrows = col.add(block, n, volume, d["Energy"], d["Forces"], d["Stress"], d.get("eweight", 1.0),
                        d.get("fweight", 1.0), d.get("vweight", 1.0), d["AtomTypes"])
        if nrows:
            A_d, b_d, w_d, _bad, _batch = col.flush(first_row=0)
            A_h, b_h, w_h = A_d.cpu().numpy(), b_d.cpu().numpy(), w_d.cpu().numpy()
            src = 0
            for on, dst0, cnt in ((e, 0, 1), (f, 1, 3 * n), (s, 1 + 3 * n, 6)):
                if on:
                    a[dst0:dst0 + cnt] = A_h[src:src + cnt]
                    b[dst0:dst0 + cnt] = b_h[src:src + cnt]
                    w[dst0:dst0 + cnt] = w_h[src:src + cnt]
                    src += cnt
        col.reset()
        self.shared_index = nrows                      # lammps_snap.py:386-387 (`index` restarts at 0)
        self.distributed_index += nrows
        return a, b, w

    def _collect_lammps(self):
        if self._b200_stock_mode():
            return self._b200_stock("_collect_lammps")
        d, n, sec, calc, lmp_types, volume, block = self._b200_gather_block()
        col = self._b200_collector()
        nrows = col.add(block, n, volume, d["Energy"], d["Forces"], d["Stress"], d["eweight"], d["fweight"],
                        d["vweight"], d["AtomTypes"])
        meta = row_metadata(n, [int(t) for t in lmp_types], calc.energy, calc.force, calc.stress, d["Group"],
                            d["File"], d["test_bool"], self.WITH_ATOM_TYPE)
        di = self.distributed_index
        for key, vals in meta.items():
            self.pt.fitsnap_dict[key][di:di + nrows] = vals
        self.shared_index += nrows
        self.distributed_index += nrows

    def flush_to_shared_arrays(self):
        """Assemble every staged row on the device, mirror them into pt.shared_arrays (so dumps,
        error analysis and outputs of the reference keep working unchanged) and leave the device
        copies on `pt` for the solver plugin."""
        col = getattr(self, "_b200_col", None)
        if col is None or not col.natoms:
            return
        first = self._b200_first_row
        self._check_device_memory(col)
        A, b, w, bad, batch = col.flush(first_row=0)
        n = batch.n_rows_out
        if int(bad.item()) and not self.SCRUB:
            raise ValueError("Nan in computed data")
        # the host mirror is filled lazily: the copy of A back to the host only happens if somebody reads
        # pt.shared_arrays[...].array (dumps, library users, the stock error analysis); see hostmirror.py
        sa = self.pt.shared_arrays
        eager = bool(getattr(self, "b200_eager_host_mirror", False))
        for name, dev_t in (("a", A), ("b", b), ("w", w)):
            inner = sa[name].unwrap() if isinstance(sa[name], LazyHostMirror) else sa[name]
            sa[name] = LazyHostMirror(inner, dev_t, first, n)
            if eager:
                sa[name].materialize()
        self.pt.fitsnap_b200_device = {"A": A, "b": b, "w": w, "first_row": first, "n_rows": n}
        col.reset()
        self._b200_col = None

    def _check_device_memory(self, col):
        """Device-side twin of the reference's RAM guard (calculator.py:277-285): the raw blocks and A, b, w of this
        rank must not take more than half of the GPU unless `[MEMORY] override = 1`."""
        mem = getattr(col.engine, "device_memory", None)
        if mem is None:
            return
        _free, total = mem()
        need = 8 * (col._nraw * (col.kraw + 1) + col.n_out * (col.k + 2))
        if need / total > 0.5:
            msec = self.config.sections["MEMORY"] if "MEMORY" in getattr(self.config, "sections", {}) else None
            if not getattr(msec, "override", False):
                raise MemoryError("The descriptor matrix and its raw blocks (%.1f GB) are larger than 50%% of the "
                                  "GPU memory (%.1f GB). \n Aborting...!" % (need * 1e-9, total * 1e-9))
            self.pt.single_print("Warning: > 50 % of the GPU memory. I hope you know what you are doing!")

    def collect_distributed_lists(self, allgather=False):
        self.flush_to_shared_arrays()
        return super().collect_distributed_lists(allgather=allgather)


class SnapCollectMixin(_CollectMixin):
    SECTION, COMPUTE, SCRUB, WITH_ATOM_TYPE = "BISPECTRUM", "snap", False, True

    def _warn_if_no_neighbors(self, block, n, sec, d):
        """lammps_snap.py:437-453: B[0,0,0] sums equal to their no-neighbour value => warning."""
        nstride, b000sum0 = sec.ncoeff, 0.0
        if not sec.bzeroflag:
            b000sum0 = 1.0
        if getattr(sec, "chemflag", 0):
            nstride //= sec.numtypes ** 3
            if getattr(sec, "wselfallflag", 0):
                b000sum0 *= sec.numtypes ** 3
        b000sum = float(np.sum(block[0, :sec.ncoeff * sec.numtypes:nstride] / n))
        if abs(b000sum - b000sum0) < 1.0e-10:
            print("! WARNING: Configuration has no SNAP neighbors \nGroup and configuration: {} {}".format(
                d["Group"], d["File"]))


class PaceCollectMixin(_CollectMixin):
    SECTION, COMPUTE, SCRUB, WITH_ATOM_TYPE = "ACE", "pace", True, False

    def _warn_if_no_neighbors(self, block, n, sec, d):
        """lammps_pace.py:415-424."""
        if sec.bzeroflag:
            return
        b000sum = float(np.sum(block[0, :sec.ncoeff * sec.numtypes:sec.ncoeff] / n))
        if abs(b000sum - 1.0) < 1.0e-10:
            self.pt.single_print("! WARNING: Configuration has no PACE neighbors. \nGroup and configuration: {} {}".format(
                d["Group"], d["File"]))
