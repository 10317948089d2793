"""Public end-to-end entry of the hot path: raw LAMMPS blocks -> (A, b, w) -> coefficients.

This is the call a FitSNAP user reaches through the drop-in plugins
(`fitsnap_b200.calculators.LammpsSnap/LammpsPace` collect the blocks, `fitsnap_b200.solvers.*`
fit them); it mirrors `FitSnap.process_configs` + `FitSnap.perform_fit`
(fitsnap3lib/fitsnap.py:134-220) for the linear solvers.
"""
from __future__ import annotations

import numpy as np
import torch

from types import SimpleNamespace

from .assembly import ConfigBatch, make_flags, pack_configs
from .engine import Engine, FitResult, default_engine, fit_rows


class LinearFitPipeline:
    def __init__(self, numtypes, ncoeff, bzeroflag, blank2j, energy=True, force=True, stress=True,
                 alpha=0.0, refine=2, group=None, engine: Engine | None = None, scrub_nonfinite=False):
        self.engine = engine or default_engine()
        self.numtypes, self.ncoeff, self.bzeroflag = int(numtypes), int(ncoeff), bool(bzeroflag)
        self.blank2j = np.ascontiguousarray(blank2j, dtype=np.float64)
        self.rows = (bool(energy), bool(force), bool(stress))
        self.alpha, self.refine, self.group = float(alpha), refine, group
        self.scrub = bool(scrub_nonfinite)

    #: narrow layouts (k + 1 <= 104, E + F + S rows, row map given) go through the fused scatter + Gram kernel
    fuse_scatter_gram = True

    def pack(self, blocks, natoms, volumes, energies, forces, stresses, eweights, fweights, vweights,
             type_fraction=None, first_row=0) -> ConfigBatch:
        e, f, s = self.rows
        return pack_configs(self.engine, blocks, natoms, volumes, energies, forces, stresses, eweights, fweights,
                            vweights, type_fraction, self.blank2j, self.numtypes, self.ncoeff, energy=e, force=f,
                            stress=s, bzeroflag=self.bzeroflag, scrub_nonfinite=self.scrub, first_row=first_row)

    def fit_batch(self, batch: ConfigBatch, testing=None, out=None) -> FitResult:
        """Device-resident step: scatter -> Gram -> (all-reduce) -> factor/solve -> refinement."""
        fused = self.engine.scatter_gram(batch, *(out or (None, None, None)), testing=testing) \
            if (self.fuse_scatter_gram and hasattr(self.engine, "scatter_gram")) else None
        if fused is not None:       # rows assembled and contracted in one pass (fsb_scatter_gram)
            A, b, w, bad, gaug = fused
            res = fit_rows(self.engine, A, b, w, testing, alpha=self.alpha, refine=self.refine, group=self.group,
                           diagnostics=False, gaug=gaug)
        else:
            A, b, w, bad = self.engine.scatter(batch, *(out or (None, None, None)))
            res = self.engine.fit(A, b, w, testing, alpha=self.alpha, refine=self.refine, group=self.group,
                                  diagnostics=False)
        res.extra.update(A=A, b=b, w=w, nonfinite=bad)
        return res

    def fit_stream(self, batches, refine=None) -> FitResult:
        """Out-of-core variant of `fit_host`: `batches` is a re-iterable (list, or callable returning an iterator) of
        argument tuples of `pack` (blocks, natoms, volumes, energies, forces, stresses, eweights, fweights, vweights[,
        type_fraction]); every batch is uploaded, scattered into rows and folded into the running Gram / residual
        (see StreamingLinearFit), so neither the raw blocks nor A are ever resident as a whole."""
        eng = self.engine

        def rows():
            for args in StreamingLinearFit._iterate(batches):
                batch = self.pack(*args)
                A, b, w, bad = eng.scatter(batch)
                if int(bad.item()) and not self.scrub:
                    raise ValueError("Nan in computed data")     # lammps_snap.py:426-428
                yield A, b, w
        sf = StreamingLinearFit(self.alpha, self.refine if refine is None else refine, self.group, eng)
        return sf.fit(rows)

    def capture(self, batch: ConfigBatch, testing=None, out=None, warmup=2) -> "CapturedStep":
        """CUDA graph of one device-resident step over fixed buffers (see CapturedStep)."""
        return CapturedStep(self, batch, testing, out, warmup)

    #: raw bytes above which fit_host streams the blocks in chunks (copy of chunk c+1 overlaps the
    #: scatter + partial Gram of chunk c); below it one copy + one launch of each kernel is faster
    pipeline_min_bytes = 64 << 20
    pipeline_chunks = 8

    def fit_host(self, blocks, natoms, volumes, energies, forces, stresses, eweights, fweights, vweights,
                 type_fraction=None, testing=None, chunks=None):
        """Host buffers in, host coefficients out (H2D of the blocks and D2H of x included)."""
        natoms = np.asarray(natoms, dtype=np.int32)
        if chunks is None:
            raw_bytes = (7 * len(natoms) + 3 * int(natoms.sum(dtype=np.int64))) * (self.ncoeff * self.numtypes + 1) * 8
            chunks = self.pipeline_chunks if raw_bytes >= self.pipeline_min_bytes else 1
        chunks = max(1, min(int(chunks), len(natoms)))
        if chunks > 1 and isinstance(blocks, np.ndarray) and isinstance(forces, np.ndarray):
            return self._fit_host_streamed(blocks, natoms, volumes, energies, forces, stresses, eweights, fweights,
                                           vweights, type_fraction, testing, chunks)
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

    def _fit_host_streamed(self, raw, natoms, volumes, energies, forces, stresses, eweights, fweights, vweights,
                           type_fraction, testing, chunks):
        """Chunked upload on a side stream; each chunk is scattered into its rows of (A, b, w) and its
        rows' Gram is added to the running sum while the next chunk is on the wire.  After the last
        chunk only 1/chunks of the scatter + Gram, the factorisation and the refinement passes remain."""
        eng = self.engine
        ncfg = len(natoms)
        e, f, s = self.rows
        raw_rows = 7 + 3 * natoms.astype(np.int64)
        raw_off = np.concatenate([[0], np.cumsum(raw_rows)])
        out_rows = int(e) + 3 * natoms.astype(np.int64) * int(f) + 6 * int(s)
        out_off = np.concatenate([[0], np.cumsum(out_rows)])
        atom_off = np.concatenate([[0], np.cumsum(natoms.astype(np.int64))])
        # chunk boundaries: equal shares of raw rows, cut at configuration boundaries
        targets = raw_off[-1] * np.arange(1, chunks) / chunks
        cuts = np.unique(np.concatenate([[0], np.searchsorted(raw_off, targets, side="left"), [ncfg]]))
        k = self.blank2j.shape[0]
        n_out = int(out_off[-1])
        dev = eng.device
        A = torch.empty((n_out, k), dtype=torch.float64, device=dev)
        b = torch.empty(n_out, dtype=torch.float64, device=dev)
        w = torch.empty(n_out, dtype=torch.float64, device=dev)
        T = None
        if testing is not None:
            T = eng.to_device(np.ascontiguousarray(testing, dtype=np.uint8), dtype=torch.uint8)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        copy_stream = self._copy_stream
        main = torch.cuda.current_stream(dev)
        copy_stream.wait_stream(main)                 # A/b/w allocations and earlier work are ordered
        forces = np.ascontiguousarray(forces, dtype=np.float64).reshape(-1)
        stresses = np.ascontiguousarray(np.asarray(stresses, dtype=np.float64).reshape(ncfg, 9))
        tf = (np.zeros((ncfg, self.numtypes)) if type_fraction is None
              else np.ascontiguousarray(type_fraction, dtype=np.float64).reshape(ncfg, self.numtypes))
        f64 = lambda arr: np.ascontiguousarray(arr, dtype=np.float64)
        kraw = self.ncoeff * self.numtypes
        raw = np.ascontiguousarray(raw, dtype=np.float64)
        assert raw.shape == (int(raw_off[-1]), kraw + 1), (raw.shape, int(raw_off[-1]), kraw + 1)
        # The copy engine must never idle: the first raw chunk goes on the wire at once, the per-configuration scalars of
        # ALL chunks (1 % of the bytes) follow in one go, then the other raw chunks back to back into one device buffer -- round 2 measured 17.5 ms per
        # step against 14.4 ms of pure PCIe time when every chunk re-uploaded its own 13 small arrays in between.
        with torch.cuda.stream(copy_stream):
            up = eng.to_device
            raw_dev = torch.empty((int(raw_off[-1]), kraw + 1), dtype=torch.float64, device=dev)
            chunk_events = []

            def send_chunk(c0, c1):
                r0, r1 = int(raw_off[int(c0)]), int(raw_off[int(c1)])
                eng.upload_into(raw_dev[r0:r1], raw[r0:r1])
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                chunk_events.append(ev)

            send_chunk(cuts[0], cuts[1])              # the wire is busy while the host prepares the small arrays
            meta = dict(volume=up(f64(volumes)), energy=up(f64(energies)), forces=up(forces), stress=up(stresses),
                        eweight=up(f64(eweights)), fweight=up(f64(fweights)), vweight=up(f64(vweights)),
                        type_fraction=up(tf), blank2j=up(self.blank2j))
            raw_off_dev = up(raw_off.astype(np.int64), dtype=torch.int64)
            out_off_dev = up(out_off.astype(np.int64), dtype=torch.int64)
            natoms_dev = up(natoms, dtype=torch.int32)
            ev_meta = torch.cuda.Event()
            ev_meta.record(copy_stream)
            for c0, c1 in zip(cuts[1:-1], cuts[2:]):
                send_chunk(c0, c1)
        h2d = raw.nbytes + sum(int(t.numel()) * t.element_size() for t in meta.values()) + \
            raw_off_dev.numel() * 8 + out_off_dev.numel() * 8 + natoms_dev.numel() * 4
        main.wait_event(ev_meta)
        flags = make_flags(e, f, s, self.bzeroflag, self.scrub)
        gaug = None
        bad = None
        batches = []
        for ev, c0, c1 in zip(chunk_events, cuts[:-1], cuts[1:]):
            c0, c1 = int(c0), int(c1)
            r0, r1 = int(out_off[c0]), int(out_off[c1])
            off_c = out_off_dev[c0:c1 + 1]
            # row -> configuration map of the chunk: a device kernel on the MAIN stream (nothing to upload); it runs
            # while the chunk's blocks are still on the wire
            row_cfg = eng.row_map(off_c, c1 - c0, r1 - r0)
            batch = ConfigBatch(raw=raw_dev, raw_row_off=raw_off_dev[c0:c1 + 1], out_row_off=off_c,
                                natoms=natoms_dev[c0:c1], volume=meta["volume"][c0:c1], energy=meta["energy"][c0:c1],
                                forces=meta["forces"][3 * int(atom_off[c0]):3 * int(atom_off[c1])],
                                stress=meta["stress"][c0:c1], eweight=meta["eweight"][c0:c1],
                                fweight=meta["fweight"][c0:c1], vweight=meta["vweight"][c0:c1],
                                type_fraction=meta["type_fraction"][c0:c1], blank2j=meta["blank2j"], ncfg=c1 - c0,
                                numtypes=self.numtypes, ncoeff=self.ncoeff, flags=flags, k=k, row_begin=r0, row_end=r1,
                                row_cfg=row_cfg, h2d_bytes=0)
            main.wait_event(ev)
            batches.append(batch)
            fused = eng.scatter_gram(batch, A, b, w, testing=None if T is None else T[r0:r1], lda=k) \
                if self.fuse_scatter_gram else None
            if fused is not None:
                bad_c, g_c = fused[3], fused[4]
            else:
                _, _, _, bad_c = eng.scatter(batch, A, b, w, lda=k)
                g_c = eng.gram(A[r0:r1], b[r0:r1], w[r0:r1], None if T is None else T[r0:r1])
            gaug = g_c if gaug is None else gaug.add_(g_c)
            bad = bad_c if bad is None else bad.add_(bad_c)
        for t in list(meta.values()) + [raw_off_dev, out_off_dev, natoms_dev, raw_dev]:
            t.record_stream(main)                    # allocated on the copy stream, consumed on the main one
        res = fit_rows(eng, A, b, w, T, alpha=self.alpha, refine=self.refine, group=self.group, diagnostics=False,
                       gaug=gaug)
        res.extra.update(A=A, b=b, w=w, nonfinite=bad, batches=batches)
        x = res.coefficients()          # D2H + sync (the chunk buffers are alive until here)
        if int(bad.item()) and not self.scrub:
            raise ValueError("Nan in computed data")     # lammps_snap.py:426-428
        summary = SimpleNamespace(ncfg=ncfg, k=k, row_begin=0, row_end=n_out, n_rows_out=n_out, h2d_bytes=int(h2d),
                                  chunks=len(batches))
        return x, res, summary


class StreamingLinearFit:
    """Out-of-core fit: the design matrix is never resident as a whole (the mode of
    examples/library/transpose_trick/example.py:226-246, where C += a^T a, d += a^T b per configuration).

    `chunks` is a RE-ITERABLE of row chunks -- a list, or a zero-argument callable returning a fresh iterator --
    each `(a, b, w)` or `(a, b, w, testing)`, host numpy or device tensors.  Pass 0 accumulates the augmented Gram
    chunk by chunk (one `fsb_gram` per chunk, summed on the device), all-reduces it once, factors and solves; each
    refinement round streams the chunks again and accumulates `aw^T (bw - aw x)` (one `fsb_residual` per chunk, one
    k-vector all-reduce per round).  Device memory: one chunk + O(k^2).  The result equals `Engine.fit` on the
    stacked rows up to the summation order of the chunk Grams."""

    def __init__(self, alpha=0.0, refine=2, group=None, engine: Engine | None = None):
        self.engine = engine or default_engine()
        self.alpha, self.refine, self.group = float(alpha), int(refine), group

    @staticmethod
    def _iterate(chunks):
        return iter(chunks() if callable(chunks) else chunks)

    def _device_chunk(self, chunk):
        eng = self.engine
        a, b, w = chunk[0], chunk[1], chunk[2]
        t = chunk[3] if len(chunk) > 3 else None
        A = eng.to_device(a)
        B = eng.to_device(b).reshape(-1)
        W = eng.to_device(w).reshape(-1)
        T = None
        if t is not None:
            T = t.to(eng.device, torch.uint8) if isinstance(t, torch.Tensor) else \
                eng.to_device(np.ascontiguousarray(t, dtype=np.uint8), dtype=torch.uint8)
        return A, B, W, T

    def fit(self, chunks) -> FitResult:
        from .engine import _all_reduce
        eng = self.engine
        start = getattr(eng, "launch_count", 0)
        gaug, n_rows = None, 0
        for chunk in self._iterate(chunks):
            A, B, W, T = self._device_chunk(chunk)
            g = eng.gram(A, B, W, T)
            gaug = g if gaug is None else gaug.add_(g)
            n_rows += int(A.shape[0])
        if gaug is None:
            raise ValueError("StreamingLinearFit.fit: no chunks")
        _all_reduce(gaug, self.group, eng)
        k = gaug.shape[0] - 1
        f = eng.factor(gaug, self.alpha)
        x = eng.solve(f, gaug[:, k], rhs_stride=k + 1)
        for _ in range(self.refine):
            gsum = None
            for chunk in self._iterate(chunks):
                A, B, W, T = self._device_chunk(chunk)
                g = eng.residual(A, B, W, T, x)
                gsum = g if gsum is None else gsum.add_(g)
            _all_reduce(gsum, self.group, eng)
            x = eng.solve(f, gsum, x_in=x)
        return FitResult(x=x, gaug=gaug, info=f.info, launches=getattr(eng, "launch_count", 0) - start,
                         extra={"factor": f, "rows_streamed": n_rows})


def npy_row_chunks(descriptors, truth, weights, chunk_rows=1 << 20, testing=None):
    """Re-iterable over the `.npy` dumps FitSNAP writes with `[EXTRAS] dump_descriptors / dump_truth /
    dump_weights` (calculators/calculator.py:329-337; default names Descriptors.npy, Truth-Ref.npy, Weights.npy,
    io/sections/extras.py:32-37), memory-mapped: feed it to `StreamingLinearFit.fit` to refit dumped matrices
    that do not fit in device (or host) memory.  `testing`: optional bool array / list over all rows."""
    def factory():
        a = np.load(descriptors, mmap_mode="r")
        b = np.load(truth, mmap_mode="r")
        w = np.load(weights, mmap_mode="r")
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        n = a.shape[0]
        if b.shape[0] != n or w.shape[0] != n:
            raise ValueError("row counts differ: %s %s %s" % (a.shape, b.shape, w.shape))
        t = None if testing is None else np.asarray(testing, dtype=bool)
        for r0 in range(0, n, int(chunk_rows)):
            r1 = min(n, r0 + int(chunk_rows))
            # copies out of the (read-only) memory maps: the upload pins them anyway
            out = (np.array(a[r0:r1], dtype=np.float64), np.array(b[r0:r1], dtype=np.float64),
                   np.array(w[r0:r1], dtype=np.float64))
            yield out if t is None else out + (t[r0:r1],)
    return factory


class CapturedStep:
    """One device-resident step (scatter -> Gram -> [all-reduce] -> factor/solve -> refinement) captured into a
    CUDA graph: the ~11 launches of a narrow fit (k ~ 100; many more for blocked factorisations) become one
    `cudaGraphLaunch`, which is what a caller re-fitting the same buffers wants (hyper-parameter scans over
    group weights: rewrite `batch.eweight/fweight/vweight` in place, replay).  The buffers of `batch`, `testing`
    and `out` are baked into the graph by address; results land in the same `FitResult` tensors on every replay.
    All launches go through the C-ABI on the capturing stream; the library allocates nothing and never
    synchronises, so the path is capturable as is -- including the row-sharded step, whose all-reduces are the
    library's own peer-window kernel (`fsb_allreduce`, csrc/comm.cu): an ordinary kernel node, no communicator
    stream (every rank must capture and replay the same sequence)."""

    def __init__(self, pipe: LinearFitPipeline, batch: ConfigBatch, testing=None, out=None, warmup=2):
        self.pipe, self.batch = pipe, batch
        dev = pipe.engine.device
        if pipe.group is not None:
            import torch.distributed as dist
            if dist.is_initialized() and dist.get_world_size(pipe.group) > 1:
                comm = pipe.engine.comm_for(pipe.group)
                if not comm.uses_peer((batch.k + 1) ** 2):
                    raise RuntimeError("CapturedStep: the Gram all-reduce of this shape goes through NCCL; only the "
                                       "peer-window collective (fsb_allreduce <= %d bytes) is captured into a graph"
                                       % comm.peer_max_bytes)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):            # warm-up: workspaces, function attributes, NCCL channels
            for _ in range(max(1, int(warmup))):
                res = pipe.fit_batch(batch, testing, out)
            if out is None:                       # keep the outputs of the scatter at a fixed address
                out = (res.extra["A"], res.extra["b"], res.extra["w"])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = pipe.engine.launch_count
        with torch.cuda.graph(self.graph):
            self.result = pipe.fit_batch(batch, testing, out)
        self.launches = pipe.engine.launch_count - n0      # kernels inside one replay (counted by the library)

    def replay(self) -> FitResult:
        self.graph.replay()
        return self.result
