"""Lazy host mirror of device-resident rows inside FitSNAP's `pt.shared_arrays`.

The reference's data plane between calculator and solver is `pt.shared_arrays['a'|'b'|'w'].array`
(`SharedArray` / `StubsArray`, fitsnap3lib/parallel_tools.py:944-1077): host numpy arrays that the
calculator fills row by row and every later stage reads.  The drop-in calculator assembles the rows on
the GPU; copying all of A back to the host right away (0.8 GB at 1e6 x 100, 80 GB at 1e7 x 1000) would
put a PCIe round trip in front of a solver that reads the device copy anyway.  `LazyHostMirror`
stands in for the shared-array object inside the `pt.shared_arrays` dict: it forwards every attribute
to the object the reference created and defers the device -> host copy of the assembled rows until
somebody actually asks for `.array` (dumps with `[EXTRAS] dump_descriptors`, library users poking at
the arrays, the stock error analysis).  Once `.array` has been handed out the host copy may be edited
in place (examples/library/bayesian_active_learning.py rescales `w` that way), so the mirror records
that it was `exposed`; the solver then reloads that array from the host instead of trusting the device
copy (a stale device cache must never win over an in-place edit).

Nothing of the reference is patched: only the dict entry is replaced, and `unwrap()` puts the
original object back.
"""
from __future__ import annotations


class LazyHostMirror:
    __slots__ = ("_inner", "_dev", "_first", "_n", "_pending", "exposed")

    def __init__(self, inner, dev, first, n):
        object.__setattr__(self, "_inner", inner)
        object.__setattr__(self, "_dev", dev)          # device tensor holding rows [first, first + n)
        object.__setattr__(self, "_first", int(first))
        object.__setattr__(self, "_n", int(n))
        object.__setattr__(self, "_pending", True)
        object.__setattr__(self, "exposed", False)

    # -- the one attribute that matters -------------------------------------------------
    def materialize(self):
        """Copy the device rows into the host array now (idempotent)."""
        if self._pending:
            host = self._inner.array
            if host.ndim == 1 and self._dev.dim() == 2:      # StubsArray of width 1 (parallel_tools.py:1067)
                host = host.reshape(-1, 1)
            host[self._first:self._first + self._n] = self._dev.detach().cpu().numpy()
            object.__setattr__(self, "_pending", False)

    @property
    def array(self):
        self.materialize()
        object.__setattr__(self, "exposed", True)
        return self._inner.array

    @array.setter
    def array(self, value):
        object.__setattr__(self, "_pending", False)
        object.__setattr__(self, "exposed", True)
        self._inner.array = value

    @property
    def pending(self):
        return self._pending

    def host_shape(self):
        """Shape of the host array WITHOUT triggering the copy."""
        return self._inner.array.shape

    def unwrap(self):
        self.materialize()
        return self._inner

    # -- everything else belongs to the reference's object --------------------------------
    def __getattr__(self, name):
        return getattr(object.__getattribute__(self, "_inner"), name)

    def __setattr__(self, name, value):
        if name == "array" or name in LazyHostMirror.__slots__:
            object.__setattr__(self, name, value)      # `array` is a property: this runs its setter
        else:
            setattr(self._inner, name, value)


def shared_shape(shared):
    """Row count / shape of a shared array without forcing a pending device -> host copy."""
    if isinstance(shared, LazyHostMirror):
        return shared.host_shape()
    return shared.array.shape
