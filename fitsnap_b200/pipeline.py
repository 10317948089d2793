"""Public end-to-end entry of the hot path: raw LAMMPS blocks -> (A, b, w) -> coefficients.

This is the call a FitSNAP user reaches through the drop-in plugins
(`fitsnap_b200.calculators.LammpsSnap/LammpsPace` collect the blocks, `fitsnap_b200.solvers.*`
fit them); it mirrors `FitSnap.process_configs` + `FitSnap.perform_fit`
(fitsnap3lib/fitsnap.py:134-220) for the linear solvers.
"""
from __future__ import annotations

import numpy as np
import torch

from .assembly import ConfigBatch, pack_configs
from .engine import Engine, FitResult, default_engine


class LinearFitPipeline:
    def __init__(self, numtypes, ncoeff, bzeroflag, blank2j, energy=True, force=True, stress=True,
                 alpha=0.0, refine=2, group=None, engine: Engine | None = None, scrub_nonfinite=False):
        self.engine = engine or default_engine()
        self.numtypes, self.ncoeff, self.bzeroflag = int(numtypes), int(ncoeff), bool(bzeroflag)
        self.blank2j = np.ascontiguousarray(blank2j, dtype=np.float64)
        self.rows = (bool(energy), bool(force), bool(stress))
        self.alpha, self.refine, self.group = float(alpha), refine, group
        self.scrub = bool(scrub_nonfinite)

    def pack(self, blocks, natoms, volumes, energies, forces, stresses, eweights, fweights, vweights,
             type_fraction=None, first_row=0) -> ConfigBatch:
        e, f, s = self.rows
        return pack_configs(self.engine, blocks, natoms, volumes, energies, forces, stresses, eweights, fweights,
                            vweights, type_fraction, self.blank2j, self.numtypes, self.ncoeff, energy=e, force=f,
                            stress=s, bzeroflag=self.bzeroflag, scrub_nonfinite=self.scrub, first_row=first_row)

    def fit_batch(self, batch: ConfigBatch, testing=None, out=None) -> FitResult:
        """Device-resident step: scatter -> Gram -> (all-reduce) -> factor/solve -> refinement."""
        A, b, w, bad = self.engine.scatter(batch, *(out or (None, None, None)))
        res = self.engine.fit(A, b, w, testing, alpha=self.alpha, refine=self.refine, group=self.group,
                              diagnostics=False)
        res.extra.update(A=A, b=b, w=w, nonfinite=bad)
        return res

    def fit_host(self, blocks, natoms, volumes, energies, forces, stresses, eweights, fweights, vweights,
                 type_fraction=None, testing=None):
        """Host buffers in, host coefficients out (H2D of the blocks and D2H of x included)."""
        batch = self.pack(blocks, natoms, volumes, energies, forces, stresses, eweights, fweights, vweights,
                          type_fraction)
        T = None
        if testing is not None:
            T = self.engine.to_device(np.ascontiguousarray(testing, dtype=np.uint8), dtype=torch.uint8)
        res = self.fit_batch(batch, T)
        x = res.coefficients()          # D2H + sync
        if int(res.extra["nonfinite"].item()) and not self.scrub:
            raise ValueError("Nan in computed data")     # lammps_snap.py:426-428
        return x, res, batch
