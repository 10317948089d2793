"""Device engine of the linear-fit hot path: thin Python over the C-ABI.

PyTorch is used ONLY as the device container (allocation, pinned staging, the stream
handle) and for the one collective (`torch.distributed.all_reduce`, NCCL over NVLink on a
B200 box; gloo in CPU tests of the host logic).  All arithmetic of the path runs in the
hand-written sm_100a kernels behind `libfitsnap_b200.so`.  There is no CPU fallback.

Path: a5 mask+weight -> a6/a7/a8 Gram + factor + solve (+ refinement against A)
-> a9 predictions; a3 scatter builds A, b, w from raw LAMMPS blocks.
"""
from __future__ import annotations

import ctypes
import os
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _cabi


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


@dataclass
class Factor:
    buf: torch.Tensor
    info: torch.Tensor
    k: int
    alpha: float


@dataclass
class PinvFactor:
    buf: torch.Tensor
    info: torch.Tensor
    k: int


@dataclass
class FitResult:
    x: torch.Tensor                 # (k,) device fp64
    gaug: torch.Tensor              # (k+1, k+1) device fp64, already all-reduced
    info: torch.Tensor              # int32[8] device
    last_correction: torch.Tensor | None = None   # |dx|_inf / |x|_inf of the last refinement step (device scalar)
    launches: int = 0               # kernels of this library launched for the fit
    extra: dict = field(default_factory=dict)

    def coefficients(self):
        return self.x.detach().cpu().numpy().astype(np.float64, copy=True)

    def info_host(self):
        return self.info.detach().cpu().numpy()


_NP_OF_TORCH = {torch.float64: np.float64, torch.int64: np.int64, torch.int32: np.int32, torch.uint8: np.uint8}


class HostStager:
    """Persistent pinned staging ring for pageable host arrays.

    A pageable array cannot be read by the copy engine; `tensor.pin_memory()` per call costs a `cudaHostAlloc` of the
    whole array (hundreds of ms per GB) plus a single-threaded memcpy.  Here NSLOT pinned slots are allocated once per
    engine; an upload walks the array in slot-sized chunks: a few host threads copy chunk c into a free slot (numpy
    releases the GIL for plain copies) while the DMA of chunk c-1 is still on the wire, then the H2D of the slot is
    queued on the CURRENT stream (so the kernels that follow need no extra synchronisation) and an event marks the
    slot reusable."""
    SMALL = 1 << 20            # below this a plain copy is cheaper than the ring
    SLOT_BYTES = 32 << 20
    NSLOT = 4

    def __init__(self, device):
        self.device = device
        self.slots = [torch.empty(self.SLOT_BYTES, dtype=torch.uint8).pin_memory() for _ in range(self.NSLOT)]
        self.events = [None] * self.NSLOT
        self.next = 0
        # 8 copy threads: measured on the 16-core bench host, 14 threads LOWER the upload rate of a pageable 816 MB
        # matrix from 34 to 28 GB/s (they compete with the DMA reads of the pinned slots for memory bandwidth)
        nthr = max(1, min(8, (os.cpu_count() or 2) // 2))
        self.pool = ThreadPoolExecutor(max_workers=nthr)
        self.nthr = nthr

    def _fill(self, dst_np, src_np):
        n = src_np.shape[0]
        if n < (4 << 20) or self.nthr == 1:
            np.copyto(dst_np[:n], src_np)
            return
        step = -(-n // self.nthr)
        futs = [self.pool.submit(np.copyto, dst_np[o:min(n, o + step)], src_np[o:min(n, o + step)])
                for o in range(0, n, step)]
        for f in futs:
            f.result()

    def upload(self, a, dtype, out=None):
        if out is None:
            out = torch.empty(a.shape, dtype=dtype, device=self.device)
        src = a.reshape(-1).view(np.uint8)
        dst = out.view(-1).view(torch.uint8)
        stream = torch.cuda.current_stream(self.device)
        for off in range(0, src.shape[0], self.SLOT_BYTES):
            n = min(self.SLOT_BYTES, src.shape[0] - off)
            i = self.next
            self.next = (i + 1) % self.NSLOT
            if self.events[i] is not None:
                self.events[i].synchronize()
            self._fill(self.slots[i].numpy(), src[off:off + n])
            dst[off:off + n].copy_(self.slots[i][:n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            self.events[i] = ev
        return out


class Engine:
    """One engine per process / GPU (one process per GPU under torchrun)."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("fitsnap_b200.Engine needs a CUDA device (B200, sm_100a); no CPU fallback exists")
        self.lib = _cabi.load()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        h = ctypes.c_void_p()
        _cabi.check("fsb_create", self.lib.fsb_create(ctypes.byref(h), self.device.index))
        self._h = h
        n = ctypes.c_int()
        _cabi.check("fsb_sm_count", self.lib.fsb_sm_count(self._h, ctypes.byref(n)))
        self.sm_count = n.value
        self._ws = {}
        self._stager = None
        self._comms = {}

    @property
    def launch_count(self):
        """Kernels this library has launched so far (counted inside the library at every launch site)."""
        n = ctypes.c_uint64()
        _cabi.check("fsb_launch_count", self.lib.fsb_launch_count(self._h, ctypes.byref(n)))
        return int(n.value)

    def comm_for(self, group):
        """The device communicator (`DeviceComm`: NCCL + NVLink peer windows behind `fsb_allreduce`) bound to a
        torch.distributed group; created collectively on first use."""
        c = self._comms.get(group)
        if c is None:
            c = self._comms[group] = DeviceComm(self, group)
        return c

    def __del__(self):
        try:
            for c in getattr(self, "_comms", {}).values():
                c.close()
            self._comms = {}
        except Exception:
            pass
        try:
            if getattr(self, "_h", None):
                self.lib.fsb_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _workspace(self, tag, nbytes):
        t = self._ws.get(tag)
        if t is None or t.numel() < nbytes:
            t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._ws[tag] = t
        return t

    def device_memory(self):
        """(free, total) bytes of this engine's GPU."""
        return torch.cuda.mem_get_info(self.device)

    def to_device(self, arr, dtype=torch.float64, non_blocking=True):
        """Host numpy -> device tensor (or pass a device tensor through).  Pinned sources go straight onto the copy
        engine; pageable sources (the ordinary numpy arrays a FitSNAP user passes to `perform_fit(a=, b=, w=)`) are
        staged through this engine's persistent pinned ring (`HostStager`) -- no per-call `cudaHostAlloc`."""
        if isinstance(arr, torch.Tensor):
            if arr.device.type == "cuda":
                return arr.to(device=self.device, dtype=dtype)
            arr = arr.numpy()
        a = np.ascontiguousarray(arr)
        want = _NP_OF_TORCH[dtype]
        if a.dtype != want:
            a = a.astype(want)
        t = torch.from_numpy(a)
        if a.nbytes == 0:
            return torch.empty(a.shape, dtype=dtype, device=self.device)
        if t.is_pinned():
            return t.to(self.device, non_blocking=non_blocking)
        if a.nbytes < HostStager.SMALL:
            return t.to(self.device)
        if self._stager is None:
            self._stager = HostStager(self.device)
        return self._stager.upload(a, dtype)

    def upload_into(self, dst, arr):
        """Host numpy (same dtype, C-contiguous) -> an existing contiguous device tensor, on the current stream:
        straight onto the copy engine when the source is pinned, through the pinned ring otherwise."""
        a = np.ascontiguousarray(arr)
        assert dst.is_contiguous() and a.dtype == _NP_OF_TORCH[dst.dtype] and a.size == dst.numel()
        if a.nbytes == 0:
            return dst
        t = torch.from_numpy(a)
        if t.is_pinned() or a.nbytes < HostStager.SMALL:
            dst.copy_(t.view(dst.shape), non_blocking=True)
            return dst
        if self._stager is None:
            self._stager = HostStager(self.device)
        return self._stager.upload(a, dst.dtype, out=dst)

    @staticmethod
    def _check_matrix(A, b, w, testing):
        assert A.dtype == torch.float64 and A.dim() == 2 and A.stride(1) == 1, "A must be fp64 row-major"
        n, k = A.shape
        assert b.dtype == torch.float64 and w.dtype == torch.float64
        assert b.numel() == n and w.numel() == n and b.is_contiguous() and w.is_contiguous()
        if testing is not None:
            assert testing.dtype == torch.uint8 and testing.numel() == n and testing.is_contiguous()
        lda = A.stride(0) if n > 1 else max(k, A.stride(0))
        return n, k, lda

    def set_gram_path(self, path):
        """'auto' | 'fp64' (DMMA) | 'int8' (tcgen05 exact-integer Gram); see include/fitsnap_b200.h."""
        code = {"auto": _cabi.GRAM_AUTO, "fp64": _cabi.GRAM_FP64, "int8": _cabi.GRAM_INT8}[path]
        _cabi.check("fsb_set_gram_path", self.lib.fsb_set_gram_path(self._h, code))

    def gram_path(self, n_rows, k):
        out = ctypes.c_int32()
        _cabi.check("fsb_get_gram_path", self.lib.fsb_get_gram_path(self._h, n_rows, k, ctypes.byref(out)))
        return {_cabi.GRAM_FP64: "fp64", _cabi.GRAM_INT8: "int8"}[out.value]

    # ------------------------------------------------------------------ kernels
    def gram(self, A, b, w, testing=None):
        n, k, lda = self._check_matrix(A, b, w, testing)
        gaug = torch.empty((k + 1, k + 1), dtype=torch.float64, device=self.device)
        nbytes = self.lib.fsb_gram_workspace_bytes(self._h, n, k)
        ws = self._workspace("gram", nbytes)
        _cabi.check("fsb_gram", self.lib.fsb_gram(self._h, _ptr(A), lda, _ptr(b), _ptr(w), _ptr(testing), n, k,
                                                   _ptr(gaug), _ptr(ws), ws.numel(), self._stream()))
        return gaug

    def factor(self, gaug, alpha=0.0):
        k = gaug.shape[0] - 1
        assert gaug.is_contiguous() and gaug.dtype == torch.float64
        nbytes = self.lib.fsb_factor_bytes(self._h, k)
        buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        info = torch.empty(_cabi.INFO_LEN, dtype=torch.int32, device=self.device)
        _cabi.check("fsb_factor", self.lib.fsb_factor(self._h, _ptr(gaug), k, float(alpha), _ptr(buf), nbytes,
                                                       _ptr(info), self._stream()))
        return Factor(buf, info, k, float(alpha))

    def solve(self, factor, rhs, rhs_stride=1, x_in=None, out=None):
        k = factor.k
        x = out if out is not None else torch.empty(k, dtype=torch.float64, device=self.device)
        _cabi.check("fsb_factor_solve",
                    self.lib.fsb_factor_solve(self._h, _ptr(factor.buf), k, _ptr(rhs), int(rhs_stride),
                                              factor.alpha, _ptr(x_in), _ptr(x), self._stream()))
        return x

    def pinv_factor(self, gaug, rcond=None, alpha=0.0):
        """G^+ (alpha = 0) or the range-restricted ridge inverse (alpha > 0) through a Jacobi eigendecomposition
        (rank-deficient fallback, see csrc/pinv.cu)."""
        k = gaug.shape[0] - 1
        if rcond is None:
            # resolution of a Gram formed and diagonalised in fp64: eigenvalues below ~k eps lambda_max are rounding
            # noise of the Jacobi sweeps as much as of the Gram (a cut at exactly k eps let a null direction through)
            rcond = 32 * k * 2.220446049250313e-16
        nbytes = self.lib.fsb_pinv_bytes(self._h, k)
        buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        info = torch.zeros(2, dtype=torch.int32, device=self.device)
        _cabi.check("fsb_pinv_factor_shifted",
                    self.lib.fsb_pinv_factor_shifted(self._h, _ptr(gaug), k, float(rcond), float(alpha), _ptr(buf),
                                                     nbytes, _ptr(info), self._stream()))
        return PinvFactor(buf, info, k)

    def pinv_apply(self, pf, rhs, rhs_stride=1, x_in=None):
        x = torch.empty(pf.k, dtype=torch.float64, device=self.device)
        _cabi.check("fsb_pinv_apply", self.lib.fsb_pinv_apply(self._h, _ptr(pf.buf), pf.k, _ptr(rhs), int(rhs_stride),
                                                               _ptr(x_in), _ptr(x), self._stream()))
        return x

    def lasso(self, gaug, n_train, alpha, max_iter=2000, tol=1e-12):
        """Coordinate descent on the reduced problem (lasso.py:25-29 objective)."""
        k = gaug.shape[0] - 1
        x = torch.empty(k, dtype=torch.float64, device=self.device)
        info = torch.empty(2, dtype=torch.int32, device=self.device)
        _cabi.check("fsb_lasso", self.lib.fsb_lasso(self._h, _ptr(gaug), k, int(n_train), float(alpha), int(max_iter),
                                                     float(tol), _ptr(x), _ptr(info), self._stream()))
        return x, info

    def residual(self, A, b, w, testing, x):
        """g = aw^T (bw - aw x) over this rank's rows."""
        n, k, lda = self._check_matrix(A, b, w, testing)
        g = torch.empty(k, dtype=torch.float64, device=self.device)
        nbytes = self.lib.fsb_residual_workspace_bytes(self._h, n, k)
        ws = self._workspace("residual", nbytes)
        _cabi.check("fsb_residual", self.lib.fsb_residual(self._h, _ptr(A), lda, _ptr(b), _ptr(w), _ptr(testing),
                                                           n, k, _ptr(x), _ptr(g), _ptr(ws), ws.numel(),
                                                           self._stream()))
        return g

    def predict(self, A, x):
        assert A.dtype == torch.float64 and A.dim() == 2 and A.stride(1) == 1
        n, k = A.shape
        lda = A.stride(0) if n > 1 else max(k, A.stride(0))
        y = torch.empty(n, dtype=torch.float64, device=self.device)
        _cabi.check("fsb_predict", self.lib.fsb_predict(self._h, _ptr(A), lda, n, k, _ptr(x), _ptr(y), self._stream()))
        return y

    def group_stats(self, A, b, w, group_id, x, n_groups):
        """Per-group error sums (n_groups x 10) for the linear error analysis, one pass over A."""
        n, k, lda = self._check_matrix(A, b, w, None)
        assert group_id.dtype == torch.int32 and group_id.numel() == n and group_id.is_contiguous()
        stats = torch.zeros((int(n_groups), 10), dtype=torch.float64, device=self.device)
        _cabi.check("fsb_group_stats", self.lib.fsb_group_stats(self._h, _ptr(A), lda, _ptr(b), _ptr(w), _ptr(group_id),
                                                                 n, k, _ptr(x), int(n_groups), _ptr(stats),
                                                                 self._stream()))
        return stats

    def scatter(self, batch, A=None, b=None, w=None, lda=None):
        """Assemble rows of (A, b, w) from a `ConfigBatch` already on the device."""
        k = batch.k
        n_out = batch.n_rows_out
        lda = int(lda or k)
        if A is None:
            A = torch.empty((batch.row_end, lda), dtype=torch.float64, device=self.device)[:, :k]
            b = torch.empty(batch.row_end, dtype=torch.float64, device=self.device)
            w = torch.empty(batch.row_end, dtype=torch.float64, device=self.device)
        nonfinite = torch.zeros(1, dtype=torch.int32, device=self.device)
        _cabi.check("fsb_scatter", self.lib.fsb_scatter(
            self._h, _ptr(batch.raw), _ptr(batch.raw_row_off), _ptr(batch.out_row_off), _ptr(batch.natoms),
            _ptr(batch.volume), _ptr(batch.energy), _ptr(batch.forces), _ptr(batch.stress),
            _ptr(batch.eweight), _ptr(batch.fweight), _ptr(batch.vweight), _ptr(batch.type_fraction),
            _ptr(batch.blank2j), batch.ncfg, batch.numtypes, batch.ncoeff, batch.flags,
            _ptr(A), A.stride(0) if A.shape[0] > 1 else lda, _ptr(b), _ptr(w), n_out, _ptr(batch.row_cfg), _ptr(nonfinite), self._stream()))
        return A, b, w, nonfinite

    def row_map(self, out_row_off, ncfg, n_rows):
        """int32 row -> configuration map of a batch (`fsb_row_map`)."""
        m = torch.empty(int(n_rows), dtype=torch.int32, device=self.device)
        _cabi.check("fsb_row_map", self.lib.fsb_row_map(self._h, _ptr(out_row_off), int(ncfg), _ptr(m), int(n_rows),
                                                         self._stream()))
        return m

    def scatter_gram(self, batch, A=None, b=None, w=None, testing=None, lda=None, store_a=True):
        """Fused K1 + K2..K4 (`fsb_scatter_gram`): assemble the rows of `batch` AND form their augmented Gram in one
        pass over the raw blocks.  Returns (A, b, w, nonfinite, gaug) -- A is None with store_a=False (streaming mode)
        -- or None when the layout is not covered by the fused kernel (the caller then runs scatter + gram)."""
        if batch.row_cfg is None or batch.ncfg == 0 or batch.n_rows_out == 0:
            return None
        all_rows = _cabi.ROWS_ENERGY | _cabi.ROWS_FORCE | _cabi.ROWS_STRESS
        k = batch.k
        if (batch.flags & all_rows) != all_rows or (k + 1 + 7) // 8 > 13 or self.gram_path(batch.n_rows_out, k) != "fp64":
            return None
        lda = int(lda or k)
        if b is None:
            if store_a:
                A = torch.empty((batch.row_end, lda), dtype=torch.float64, device=self.device)[:, :k]
            b = torch.empty(batch.row_end, dtype=torch.float64, device=self.device)
            w = torch.empty(batch.row_end, dtype=torch.float64, device=self.device)
        if not store_a:
            A = None
        if testing is not None:
            assert testing.dtype == torch.uint8 and testing.numel() == batch.n_rows_out and testing.is_contiguous()
        n_out = batch.n_rows_out
        gaug = torch.empty((k + 1, k + 1), dtype=torch.float64, device=self.device)
        ws = self._workspace("gram", self.lib.fsb_gram_workspace_bytes(self._h, n_out, k))
        nonfinite = torch.zeros(1, dtype=torch.int32, device=self.device)
        a_lda = (A.stride(0) if A.shape[0] > 1 else lda) if A is not None else k
        st = self.lib.fsb_scatter_gram(
            self._h, _ptr(batch.raw), _ptr(batch.raw_row_off), _ptr(batch.out_row_off), _ptr(batch.natoms),
            _ptr(batch.volume), _ptr(batch.energy), _ptr(batch.forces), _ptr(batch.stress),
            _ptr(batch.eweight), _ptr(batch.fweight), _ptr(batch.vweight), _ptr(batch.type_fraction),
            _ptr(batch.blank2j), batch.ncfg, batch.numtypes, batch.ncoeff, batch.flags,
            _ptr(A), a_lda, _ptr(b), _ptr(w), n_out, _ptr(batch.row_cfg), _ptr(nonfinite), _ptr(testing), _ptr(gaug),
            _ptr(ws), ws.numel(), self._stream())
        if st == _cabi.UNSUPPORTED:
            return None
        _cabi.check("fsb_scatter_gram", st)
        return A, b, w, nonfinite, gaug

    # ------------------------------------------------------------------ the fit
    def fit(self, A, b, w, testing=None, alpha=0.0, refine=2, group=None, diagnostics=True):
        return fit_rows(self, A, b, w, testing, alpha=alpha, refine=refine, group=group, diagnostics=diagnostics)

    def refine_once(self, A, b, w, testing, res, group=None):
        return refine_rows(self, A, b, w, testing, res, group=group)


class DeviceComm:
    """`fsb_comm_t` bound to the ranks of a torch.distributed group: torch.distributed is only the OUT-OF-BAND channel
    that ships the NCCL unique id and the CUDA IPC handles of the peer windows at set-up; every all-reduce of the fit
    afterwards is one `fsb_allreduce` on the caller's stream (NVLink peer-window kernel for the small Gram / k-vector
    messages, ncclAllReduce for large ones) -- see csrc/comm.cu.  `FSB_NO_PEER=1` keeps everything on NCCL."""

    def __init__(self, engine, group):
        import socket
        import torch.distributed as dist
        self.engine, self.group = engine, group
        lib = engine.lib
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._c = None
        idbuf = ctypes.create_string_buffer(128)
        if self.rank == 0:
            _cabi.check("fsb_comm_unique_id", lib.fsb_comm_unique_id(idbuf, 128))
        box = [bytes(idbuf.raw) if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0), group=group)
        c = ctypes.c_void_p()
        _cabi.check("fsb_comm_init", lib.fsb_comm_init(engine._h, box[0], self.world, self.rank, ctypes.byref(c)))
        self._c = c
        # peer windows: ranks of one box only, and only if every rank could map every window
        self.peer = False
        if self.world > 1 and not os.environ.get("FSB_NO_PEER"):
            hb = int(lib.fsb_comm_peer_handle_bytes())
            hbuf = ctypes.create_string_buffer(hb)
            st = lib.fsb_comm_peer_export(self._c, hbuf, hb)
            mine = (socket.gethostname(), bytes(hbuf.raw) if st == 0 else None)
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            ok = all(h is not None and host == everyone[0][0] for host, h in everyone)
            if ok:
                ok = lib.fsb_comm_peer_attach(self._c, b"".join(h for _host, h in everyone), self.world) == 0
            votes = [None] * self.world
            dist.all_gather_object(votes, bool(ok), group=group)
            self.peer = all(votes)
            if not self.peer:
                lib.fsb_comm_peer_disable(self._c)
        self.peer_max_bytes = int(lib.fsb_comm_peer_max_bytes())

    def all_reduce(self, t):
        assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        _cabi.check("fsb_allreduce", self.engine.lib.fsb_allreduce(self.engine._h, self._c, _ptr(t), t.numel(),
                                                                   self.engine._stream()))
        return t

    def uses_peer(self, numel):
        return self.peer and numel * 8 <= self.peer_max_bytes

    def info(self):
        out = (ctypes.c_int64 * 4)()
        _cabi.check("fsb_comm_info", self.engine.lib.fsb_comm_info(self._c, out))
        return {"peer_windows": bool(out[0]), "peer_calls": int(out[1]), "nccl_calls": int(out[2]), "world": int(out[3])}

    def close(self):
        if self._c is not None:
            self.engine.lib.fsb_comm_destroy(self._c)
            self._c = None


def _all_reduce(t, group, engine=None):
    """In-place sum of `t` over `group`.  Device tensors of a real Engine go through the library's own collective
    (`fsb_allreduce`); host tensors (the gloo tests of the host logic) through torch.distributed."""
    import torch.distributed as dist
    if group is None or not dist.is_initialized() or dist.get_world_size(group) <= 1:
        return
    if t.is_cuda and hasattr(engine, "comm_for") and t.dtype == torch.float64:
        if not t.is_contiguous():
            raise ValueError("all-reduce of a non-contiguous device tensor")
        engine.comm_for(group).all_reduce(t)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def fit_rows(engine, A, b, w, testing=None, alpha=0.0, refine=2, group=None, diagnostics=True, gaug=None):
    """Weighted least squares / ridge on this rank's row shard (`engine` = `Engine`, or any object
    with gram/factor/solve/residual -- the CPU gloo tests drive this function with a stand-in).

    G~ = [aw|bw]^T[aw|bw] (one pass over A) -> ONE all-reduce of (k+1)^2 doubles over `group`
    (mirrors comm.Allreduce of examples/library/transpose_trick/example.py:241-242) -> equilibrated
    Cholesky of G + alpha I -> x; then `refine` rounds of
        x += (G + alpha I)^-1 (aw^T (bw - aw x) - alpha x)
    with the residual streamed from A (one pass + one k-vector all-reduce per round).
    Every rank ends with the same x (replicated solve, no broadcast).
    """
    start = getattr(engine, "launch_count", 0)
    if gaug is None:                 # a caller that streamed the rows in has accumulated it already
        gaug = engine.gram(A, b, w, testing)
    _all_reduce(gaug, group, engine)
    k = gaug.shape[0] - 1
    f = engine.factor(gaug, alpha)
    x = engine.solve(f, gaug[:, k], rhs_stride=k + 1)
    last = None
    for _ in range(int(refine)):
        g = engine.residual(A, b, w, testing, x)
        _all_reduce(g, group, engine)
        x_new = engine.solve(f, g, x_in=x)
        if diagnostics:
            last = (x_new - x).abs().max() / x_new.abs().max().clamp_min(1e-300)
        x = x_new
    return FitResult(x=x, gaug=gaug, info=f.info, last_correction=last,
                     launches=getattr(engine, "launch_count", 0) - start, extra={"factor": f})


def fit_rows_min_norm(engine, A, b, w, testing, gaug, refine=3, group=None, rcond=None, alpha=0.0):
    """Eigenvalue-truncated solve for a numerically rank-deficient system.  alpha = 0: the minimum-norm least-squares
    solution (what gelsd returns, svd.py:54): x0 = G^+ c, then x += G^+ aw^T (bw - aw x).  alpha > 0: the ridge
    solution through V diag(1/(lambda + alpha)) V^T over the numerical range of G (sklearn's Ridge falls back to an
    SVD-based solve when its Cholesky breaks down), refinement rhs aw^T (bw - aw x) - alpha x.  The residual is streamed from A; `gaug` is the (already
    all-reduced) augmented Gram of `fit_rows`."""
    start = getattr(engine, "launch_count", 0)
    k = gaug.shape[0] - 1
    pf = engine.pinv_factor(gaug, rcond, alpha=alpha)
    x = engine.pinv_apply(pf, gaug[:, k], rhs_stride=k + 1)
    last = None
    for _ in range(int(refine)):
        g = engine.residual(A, b, w, testing, x)
        _all_reduce(g, group, engine)
        if alpha:
            g = g - float(alpha) * x
        x_new = engine.pinv_apply(pf, g, x_in=x)
        last = (x_new - x).abs().max() / x_new.abs().max().clamp_min(1e-300)
        x = x_new
    return FitResult(x=x, gaug=gaug, info=pf.info, last_correction=last,
                     launches=getattr(engine, "launch_count", 0) - start, extra={"pinv": pf, "min_norm": True})


def refine_rows(engine, A, b, w, testing, res, group=None):
    """One more refinement round on top of a FitResult (adaptive tail for ill-conditioned fits)."""
    start = getattr(engine, "launch_count", 0)
    f = res.extra["factor"]
    g = engine.residual(A, b, w, testing, res.x)
    _all_reduce(g, group, engine)
    x_new = engine.solve(f, g, x_in=res.x)
    last = (x_new - res.x).abs().max() / x_new.abs().max().clamp_min(1e-300)
    return FitResult(x=x_new, gaug=res.gaug, info=res.info, last_correction=last,
                     launches=res.launches + getattr(engine, "launch_count", 0) - start, extra=res.extra)


_default_engine = None


def local_device_index():
    """GPU of this process: the node-local rank the launcher exported (torchrun LOCAL_RANK, Open MPI / MVAPICH /
    Intel MPI / Slurm equivalents) modulo the device count -- so the ranks of one node spread over its GPUs the way the
    reference's ranks spread over a node's cores (parallel_tools.py:262-300) -- else the current device."""
    n = torch.cuda.device_count()
    for var in ("LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID",
                "SLURM_LOCALID"):
        v = os.environ.get(var)
        if v is not None and n > 0:
            try:
                return int(v) % n
            except ValueError:
                pass
    return torch.cuda.current_device()


def default_engine():
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(local_device_index())
    return _default_engine
