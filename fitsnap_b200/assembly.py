"""Host-side packing of per-configuration LAMMPS blocks for the batched scatter kernel.

The reference assembles rows one configuration at a time inside
`LammpsSnap._collect_lammps` (fitsnap3lib/calculators/lammps_snap.py:391-556) /
`LammpsPace._collect_lammps` (lammps_pace.py:369-509), writing at `shared_index`.
Here the raw `(1+3N+6) x (K_raw+1)` compute blocks of a batch of configurations are
concatenated into one pinned buffer, shipped H2D once, and a single kernel writes every
row of A, b, w (`Engine.scatter`).  Row order is the reference's: per configuration
energy row, 3N force rows, 6 virial rows (lammps_snap.py:488-549).

No arithmetic of the path happens here: this module only lays bytes out.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _cabi


def rows_per_config(natoms, energy, force, stress):
    """calculator.py:261-272: rows a configuration contributes to A."""
    return int(bool(energy)) + 3 * int(natoms) * int(bool(force)) + 6 * int(bool(stress))


def descriptor_width(ncoeff, numtypes, bzeroflag):
    """lammps_snap.py:15-23 / lammps_pace.py:15-23 `get_width()` (linear branch)."""
    return int(ncoeff) * int(numtypes) + (0 if bzeroflag else int(numtypes))


def type_fractions(atom_types, type_mapping, numtypes):
    """lammps_snap.py:459-462: per-type atom counts divided by the atom count (fp64)."""
    counts = np.zeros(numtypes, dtype=np.float64)
    for a in atom_types:
        counts[type_mapping[a] - 1] += 1
    return counts / len(atom_types)


@dataclass
class ConfigBatch:
    """Device-resident inputs of one `fsb_scatter` call."""
    raw: torch.Tensor
    raw_row_off: torch.Tensor
    out_row_off: torch.Tensor
    natoms: torch.Tensor
    volume: torch.Tensor
    energy: torch.Tensor
    forces: torch.Tensor
    stress: torch.Tensor
    eweight: torch.Tensor
    fweight: torch.Tensor
    vweight: torch.Tensor
    type_fraction: torch.Tensor
    blank2j: torch.Tensor
    ncfg: int
    numtypes: int
    ncoeff: int
    flags: int
    k: int
    row_begin: int
    row_end: int
    row_cfg: torch.Tensor | None = None   # int32 configuration index of every output row
    h2d_bytes: int = 0

    @property
    def n_rows_out(self):
        return self.row_end - self.row_begin


def make_flags(energy, force, stress, bzeroflag, scrub_nonfinite=False):
    f = 0
    if energy:
        f |= _cabi.ROWS_ENERGY
    if force:
        f |= _cabi.ROWS_FORCE
    if stress:
        f |= _cabi.ROWS_STRESS
    if bzeroflag:
        f |= _cabi.BZEROFLAG
    if scrub_nonfinite:
        f |= _cabi.SCRUB_NONFINITE
    return f


def pack_configs(engine, blocks, natoms, volumes, energies, forces, stresses, eweights, fweights, vweights,
                 type_fraction, blank2j, numtypes, ncoeff, energy=True, force=True, stress=True,
                 bzeroflag=False, scrub_nonfinite=False, first_row=0):
    """Concatenate host arrays of a batch of configurations and upload them.

    blocks        list of (1+3N_c+6, ncoeff*numtypes+1) arrays, or one pre-concatenated 2-D array
    forces        list of (N_c, 3) arrays (data["Forces"]) or one concatenated (sum N, 3) array
    stresses      (ncfg, 3, 3)
    type_fraction (ncfg, numtypes) or None when bzeroflag
    """
    natoms = np.asarray(natoms, dtype=np.int32)
    ncfg = int(natoms.shape[0])
    kraw = ncoeff * numtypes
    k = descriptor_width(ncoeff, numtypes, bzeroflag)
    up = engine.to_device
    n_raw = 7 * ncfg + 3 * int(natoms.sum(dtype=np.int64))

    # the descriptor blocks are >99 % of the bytes: put that copy on the wire first and do the
    # bookkeeping below while it is in flight
    if isinstance(blocks, np.ndarray) and blocks.ndim == 2:
        raw = np.ascontiguousarray(blocks, dtype=np.float64)
    else:
        raw_off0 = np.zeros(ncfg + 1, dtype=np.int64)
        np.cumsum(7 + 3 * natoms.astype(np.int64), out=raw_off0[1:])
        raw = np.empty((n_raw, kraw + 1), dtype=np.float64)
        for c, blk in enumerate(blocks):
            raw[raw_off0[c]:raw_off0[c + 1]] = blk
    assert raw.shape == (n_raw, kraw + 1), (raw.shape, n_raw, kraw + 1)
    raw_dev = up(raw)

    raw_rows = 7 + 3 * natoms.astype(np.int64)
    raw_off = np.zeros(ncfg + 1, dtype=np.int64)
    np.cumsum(raw_rows, out=raw_off[1:])
    out_rows = (int(bool(energy)) + 3 * natoms.astype(np.int64) * int(bool(force)) + 6 * int(bool(stress)))
    out_off = np.zeros(ncfg + 1, dtype=np.int64)
    np.cumsum(out_rows, out=out_off[1:])
    out_off += int(first_row)

    if isinstance(forces, np.ndarray):
        fcat = np.ascontiguousarray(forces, dtype=np.float64).reshape(-1)
    else:
        fcat = (np.concatenate([np.asarray(f, dtype=np.float64).reshape(-1) for f in forces])
                if ncfg else np.zeros(0))
    assert fcat.shape[0] == 3 * int(natoms.sum())
    st = np.ascontiguousarray(np.asarray(stresses, dtype=np.float64).reshape(ncfg, 9))
    tf = (np.zeros((ncfg, numtypes)) if type_fraction is None
          else np.ascontiguousarray(type_fraction, dtype=np.float64).reshape(ncfg, numtypes))
    b2j = np.ascontiguousarray(blank2j, dtype=np.float64)
    assert b2j.shape == (k,), (b2j.shape, k)

    host = dict(volume=np.asarray(volumes, dtype=np.float64), energy=np.asarray(energies, dtype=np.float64),
                forces=fcat, stress=st, eweight=np.asarray(eweights, dtype=np.float64),
                fweight=np.asarray(fweights, dtype=np.float64), vweight=np.asarray(vweights, dtype=np.float64),
                type_fraction=tf, blank2j=b2j)
    dev = {name: up(arr) for name, arr in host.items()}
    dev["raw"] = raw_dev
    nbytes = raw.nbytes + sum(a.nbytes for a in host.values()) + raw_off.nbytes + out_off.nbytes + natoms.nbytes
    out_off_dev = up(out_off, dtype=torch.int64)
    # row -> configuration map, built on the device (nothing to upload): it lets the scatter take its TMA-staged
    # kernel and, for narrow matrices, the fused scatter + Gram kernel
    row_cfg = None
    if ncfg and hasattr(engine, "row_map"):
        row_cfg = engine.row_map(out_off_dev, ncfg, int(out_off[-1] - out_off[0]))
    return ConfigBatch(raw_row_off=up(raw_off, dtype=torch.int64), out_row_off=out_off_dev,
                       natoms=up(natoms, dtype=torch.int32), ncfg=ncfg, numtypes=int(numtypes), ncoeff=int(ncoeff),
                       flags=make_flags(energy, force, stress, bzeroflag, scrub_nonfinite), k=k,
                       row_begin=int(out_off[0]), row_end=int(out_off[-1]), row_cfg=row_cfg,
                       h2d_bytes=int(nbytes), **dev)
