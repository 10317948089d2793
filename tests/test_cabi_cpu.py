"""CPU: the C-ABI library builds, loads, and exports every symbol include/fitsnap_b200.h declares
(no compute call is made -- there is no GPU here), and the product path fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fitsnap_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fsb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_all_bound_and_exported():
    import __graft_entry__
    __graft_entry__.build()
    from fitsnap_b200 import _cabi
    lib = _cabi.load()
    syms = declared_symbols()
    assert len(syms) >= 15
    assert set(syms) == set(_cabi.SIGNATURES), set(syms) ^ set(_cabi.SIGNATURES)
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.fsb_version() >= 100
    assert lib.fsb_status_string(0) == b"ok" and lib.fsb_status_string(3) == b"workspace too small"


def test_argument_validation_without_device():
    """Entry points reject bad arguments before touching the device; fsb_create reports 'no device'."""
    import ctypes
    from fitsnap_b200 import _cabi
    lib = _cabi.load()
    h = ctypes.c_void_p()
    import torch
    if not torch.cuda.is_available():
        assert lib.fsb_create(ctypes.byref(h), 0) == 5          # FSB_ERR_NO_DEVICE
        assert not h.value
    assert lib.fsb_gram(None, None, 0, None, None, None, 0, 1, None, None, 0, None) == 1   # INVALID_ARGUMENT
    assert lib.fsb_factor_bytes(None, 10) == 0


def test_missing_library_fails_loudly(tmp_path):
    from fitsnap_b200 import _cabi
    with pytest.raises(_cabi.NativeLibraryError):
        _cabi.load(str(tmp_path / "nope.so"))


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fitsnap_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine()


def test_solver_mask_resolution_matches_reference_rules():
    """svd.py:35-46: fs_dict['Testing'] wins, then trainall, then pt.fitsnap_dict['Testing']."""
    from types import SimpleNamespace
    from fitsnap_b200.solvers import SVD, RIDGE
    n, k = 6, 2
    a = np.arange(n * k, dtype=float).reshape(n, k)
    pt = SimpleNamespace(_rank=0, shared_arrays={"a": SimpleNamespace(array=a), "b": SimpleNamespace(array=np.ones(n)),
                                                 "w": SimpleNamespace(array=np.ones(n))},
                         fitsnap_dict={"Testing": [False, True, False, False, True, False]})
    s = SVD("SVD", pt, SimpleNamespace(sections={}))
    _, _, _, t = s._resolve_inputs(None, None, None, None, False)
    assert t.tolist() == [False, True, False, False, True, False]
    _, _, _, t = s._resolve_inputs(a, np.ones(n), np.ones(n), None, True)
    assert t is None
    _, _, _, t = s._resolve_inputs(a, np.ones(n), np.ones(n), {"Testing": [True] + [False] * 5}, True)
    assert t.tolist() == [True] + [False] * 5
    with pytest.raises(ValueError):
        s._resolve_inputs(a, np.ones(n), np.ones(n), {"Testing": [True]}, False)
    r = RIDGE("RIDGE", pt, SimpleNamespace(sections={"RIDGE": SimpleNamespace(alpha=3e-5, local_solver=0)}))
    assert r._alpha() == 3e-5
    assert RIDGE("RIDGE", pt, SimpleNamespace(sections={}))._alpha() == 1e-8      # solver_sections/ridge.py:13 default


def test_extract_compute_array_reads_a_lammps_style_double_pointer():
    """lammps_base.py:280-307: extract_compute(name, 0, 2) returns double**; the view must alias it."""
    import ctypes
    from fitsnap_b200.calculators import extract_compute_array, row_metadata
    blk = np.arange(12, dtype=np.float64).reshape(3, 4)

    class Lmp:
        def extract_compute(self, name, style, rtype):
            assert (name, style, rtype) == ("snap", 0, 2)
            self._row = blk.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
            return ctypes.pointer(self._row)

    view = extract_compute_array(Lmp(), "snap", (3, 4))
    assert np.array_equal(view, blk)
    m = row_metadata(2, [1, 2], True, True, True, "grp", "f.json", True)
    assert m["Row_Type"] == ["Energy"] + ["Force"] * 6 + ["Stress"] * 6
    assert m["Atom_I"] == [0, 0, 0, 0, 1, 1, 1] + [0] * 6 and m["Atom_Type"] == [0, 1, 1, 1, 2, 2, 2] + [0] * 6
    assert m["Testing"] == [True] * 13 and m["Groups"] == ["grp"] * 13
