"""ctypes binding of libfitsnap_b200.so (include/fitsnap_b200.h).

The product path has NO CPU fallback: if the shared object is missing or a call fails the
caller gets an exception (`NativeLibraryError` / `FsbError`).  Build it with
`python -m fitsnap_b200.csrc.build` (done by `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint8, c_uint64, c_void_p

LIB_NAME = "libfitsnap_b200.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", LIB_NAME)

# flags / info indices mirrored from the header
ROWS_ENERGY, ROWS_FORCE, ROWS_STRESS, BZEROFLAG, SCRUB_NONFINITE = 1, 2, 4, 8, 16
GRAM_AUTO, GRAM_FP64, GRAM_INT8 = 0, 1, 2
INFO_STATUS, INFO_FIRST_BAD_COLUMN, INFO_NUM_PINNED, INFO_NUM_DEFICIENT, INFO_LEN = 0, 1, 2, 3, 8

# every symbol include/fitsnap_b200.h declares: name -> (restype, argtypes)
_P = c_void_p  # device pointers travel as integers
SIGNATURES = {
    "fsb_version": (c_int, []),
    "fsb_status_string": (c_char_p, [c_int]),
    "fsb_last_cuda_error": (c_char_p, []),
    "fsb_create": (c_int, [POINTER(c_void_p), c_int]),
    "fsb_destroy": (c_int, [c_void_p]),
    "fsb_sm_count": (c_int, [c_void_p, POINTER(c_int)]),
    "fsb_launch_count": (c_int, [c_void_p, POINTER(c_uint64)]),
    "fsb_comm_unique_id": (c_int, [c_void_p, c_size_t]),
    "fsb_comm_init": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_void_p)]),
    "fsb_comm_adopt": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(c_void_p)]),
    "fsb_comm_peer_handle_bytes": (c_size_t, []),
    "fsb_comm_peer_max_bytes": (c_size_t, []),
    "fsb_comm_peer_export": (c_int, [c_void_p, c_void_p, c_size_t]),
    "fsb_comm_peer_attach": (c_int, [c_void_p, c_void_p, c_int]),
    "fsb_comm_peer_disable": (c_int, [c_void_p]),
    "fsb_comm_info": (c_int, [c_void_p, POINTER(c_int64)]),
    "fsb_allreduce": (c_int, [c_void_p, c_void_p, _P, c_int64, c_void_p]),
    "fsb_comm_destroy": (c_int, [c_void_p]),
    "fsb_scatter": (c_int, [c_void_p, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                            c_int32, c_int32, c_int32, c_int32, _P, c_int64, _P, _P, c_int64, _P, _P, c_void_p]),
    "fsb_row_map": (c_int, [c_void_p, _P, c_int32, _P, c_int64, c_void_p]),
    "fsb_scatter_gram": (c_int, [c_void_p, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                 c_int32, c_int32, c_int32, c_int32, _P, c_int64, _P, _P, c_int64, _P, _P, _P, _P, _P,
                                 c_size_t, c_void_p]),
    "fsb_set_gram_path": (c_int, [c_void_p, c_int32]),
    "fsb_get_gram_path": (c_int, [c_void_p, c_int64, c_int32, POINTER(c_int32)]),
    "fsb_gram_workspace_bytes": (c_size_t, [c_void_p, c_int64, c_int32]),
    "fsb_gram": (c_int, [c_void_p, _P, c_int64, _P, _P, _P, c_int64, c_int32, _P, _P, c_size_t, c_void_p]),
    "fsb_factor_bytes": (c_size_t, [c_void_p, c_int32]),
    "fsb_factor": (c_int, [c_void_p, _P, c_int32, c_double, _P, c_size_t, _P, c_void_p]),
    "fsb_factor_solve": (c_int, [c_void_p, _P, c_int32, _P, c_int64, c_double, _P, _P, c_void_p]),
    "fsb_pinv_bytes": (c_size_t, [c_void_p, c_int32]),
    "fsb_pinv_factor": (c_int, [c_void_p, _P, c_int32, c_double, _P, c_size_t, _P, c_void_p]),
    "fsb_pinv_factor_shifted": (c_int, [c_void_p, _P, c_int32, c_double, c_double, _P, c_size_t, _P, c_void_p]),
    "fsb_pinv_apply": (c_int, [c_void_p, _P, c_int32, _P, c_int64, _P, _P, c_void_p]),
    "fsb_lasso": (c_int, [c_void_p, _P, c_int32, c_int64, c_double, c_int32, c_double, _P, _P, c_void_p]),
    "fsb_residual_workspace_bytes": (c_size_t, [c_void_p, c_int64, c_int32]),
    "fsb_residual": (c_int, [c_void_p, _P, c_int64, _P, _P, _P, c_int64, c_int32, _P, _P, _P, c_size_t, c_void_p]),
    "fsb_group_stats": (c_int, [c_void_p, _P, c_int64, _P, _P, _P, c_int64, c_int32, _P, c_int32, _P, c_void_p]),
    "fsb_predict": (c_int, [c_void_p, _P, c_int64, c_int64, c_int32, _P, _P, c_void_p]),
}


class NativeLibraryError(RuntimeError):
    """libfitsnap_b200.so is missing or lacks a declared symbol."""


UNSUPPORTED = 4


class FsbError(RuntimeError):
    """A C-ABI call returned a non-zero fsb_status."""

    def __init__(self, fn, status, text, cuda_text=""):
        self.fn, self.status = fn, status
        msg = "%s failed: %s (status %d)" % (fn, text, status)
        if cuda_text:
            msg += " -- " + cuda_text
        super().__init__(msg)


_lib = None


def load(path=None):
    """Load the shared object and bind every declared symbol (no compute is triggered)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("FITSNAP_B200_LIB") or LIB_PATH
    if not os.path.exists(p):
        raise NativeLibraryError(
            "%s not found at %s; build it with `python -m fitsnap_b200.csrc.build` "
            "(there is no CPU fallback for this path)" % (LIB_NAME, p))
    try:
        lib = ctypes.CDLL(p)
    except OSError as e:
        raise NativeLibraryError("cannot load %s: %s" % (p, e)) from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryError("%s does not export %s" % (p, name)) from e
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(fn_name, status):
    if status != 0:
        lib = load()
        text = lib.fsb_status_string(status).decode()
        cuda_text = lib.fsb_last_cuda_error().decode() if status in (2, 4) else ""
        raise FsbError(fn_name, status, text, cuda_text)
