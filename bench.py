#!/usr/bin/env python
"""bench.py -- FitSNAP linear-fit hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c5|c4s|c4_shard|c4] [--impl reference]

One STEP = one pass of the hot path over one batch of synthetic configurations:
    scatter (raw LAMMPS blocks -> A, b, w)  ->  fused mask/weight/Gram  ->  [all-reduce over the row shards]
    ->  equilibrated Cholesky solve  ->  2 rounds of refinement streamed from A.

Headline workload = BASELINE.json configs[1]: synthetic A 1e6 x 100 fp64, ridge alpha 1e-6, PER GPU (weak scaling:
every rank owns a 1e6-row shard; one all-reduce of the 101x101 Gram + one 100-vector all-reduce per refinement round,
through the library's own collective `fsb_allreduce`).  The same JSON line carries, under "workloads", the TARGET
SHAPE of the north star -- 1.25e6 x 1000 per GPU, i.e. at 8 GPUs exactly BASELINE configs[3] (1e7 x 1000) -- which runs
the int8 tcgen05 Gram; --no-secondary skips it.

value    : rows/s, inputs resident in HBM (raw blocks on device), whole job over N GPUs.  The K timed steps are K
           replays of the step captured into a CUDA graph ("launch_mode": "cuda_graph_replay") when the replay
           reproduces the eagerly launched coefficients bit for bit on every rank; the eagerly launched K steps
           ("eager": {...}) are timed first and carry the per-phase CUDA events ("phases_ms_rank0").
e2e      : same metric through the public host API (`LinearFitPipeline.fit_host`): pinned HOST raw blocks -> H2D ->
           same device path -> D2H of the coefficients, all inside the timed region.
e2e_plugin : the reference-facing SOLVER call on ordinary (pageable) numpy arrays: `RIDGE.perform_fit(a=A, b=b, w=w,
           trainall=True)` (ridge.py:11 signature) -- upload of A through the pinned ring + fit + D2H of the
           coefficients -- and the calculator -> solver hand-off (`BlockCollector.add` per configuration, flush, fit).
roofline : the dominant kernel (fused Gram) timed with CUDA events on its stream inside the timed steps.
coeff_max_rel_err : at EVERY N -- coefficients of the timed (sharded) fit against a host solution built from the
           oracle's Gram of every shard (numpy, gathered out of band), plus bit-equality of x across ranks.
cpu_baseline : the oracle (numpy restatement of the reference + scipy/sklearn, kind "port") on the host cores, on a
           bounded sample of the same workload (rank 0, N=1 only).
--impl reference : the reference's CPU path, step for step (oracle/linear_fit.py *_as_reference: per-configuration
           numpy assembly with the dense diag(blank2J) product, Python-list training mask, two copies of A, sklearn
           Ridge, residual product) on the FULL per-GPU workload of the headline config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: configs per rank, atoms per config, numtypes, ncoeff  (rows/config = 1 + 3N + 6, K = nt*nc + nt)
    "c2": dict(ncfg=10000, natoms=31, numtypes=2, ncoeff=49, desc="synthetic A 1e6x100 fp64, ridge 1e-6 (BASELINE configs[1])"),
    "c3": dict(ncfg=1841, natoms=64, numtypes=2, ncoeff=239, desc="InP-like 367k x 480 (BASELINE configs[2] shape)"),
    "c5": dict(ncfg=41230, natoms=12, numtypes=2, ncoeff=54, desc="WBe-like 1.77M x 110 (BASELINE configs[4] shape)"),
    "c4s": dict(ncfg=10000, natoms=31, numtypes=2, ncoeff=499, desc="ACE-like 1e6 x 1000 (BASELINE configs[3] shape, 1/10 rows per GPU)"),
    # one eighth of BASELINE configs[3] per GPU: at 8 GPUs this IS the 1e7 x 1000 matrix
    "c4_shard": dict(ncfg=12500, natoms=31, numtypes=2, ncoeff=499, desc="ACE-like 1.25e6 x 1000 per GPU = BASELINE configs[3] (1e7 x 1000) at 8 GPUs"),
    # the whole BASELINE configs[3] matrix on ONE GPU: 80 GB of raw blocks + 80 GB of A (run with --no-e2e)
    "c4": dict(ncfg=100000, natoms=31, numtypes=2, ncoeff=499, desc="ACE-like 1e7 x 1000, the full BASELINE configs[3] matrix on one GPU"),
}
ALPHA = 1.0e-6
REFINE = 2
I8_MIN_COLS, I8_MIN_ROWS = 384, 65536     # FSB_GRAM_AUTO rule (csrc/fsb_common.cuh), for the config of the reference arm


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def make_config(name, world, gram_path=None):
    """The `config` object of the JSON line: identical for the CUDA arm and the reference arm."""
    wl = WORKLOADS[name]
    n, nt, nc = wl["natoms"], wl["numtypes"], wl["ncoeff"]
    kraw, k = nt * nc, nt * nc + nt
    n_rows = wl["ncfg"] * (7 + 3 * n)
    if gram_path is None:
        gram_path = "int8" if (k + 1 >= I8_MIN_COLS and n_rows >= I8_MIN_ROWS) else "fp64"
    return {"workload": name + ": " + wl["desc"], "rows_per_gpu": n_rows, "k": k, "configs_per_gpu": wl["ncfg"],
            "atoms_per_config": n, "alpha": ALPHA, "refine_rounds": REFINE, "gram_path": gram_path,
            "parallelism": "row-shard x%d, 1 all-reduce of (k+1)^2 + %d of k doubles" % (world, REFINE),
            "l2": "inputs (A %.0f MB + raw %.0f MB per GPU) larger than the 126 MB L2; no flush" %
                  (n_rows * k * 8 / 1e6, n_rows * (kraw + 1) * 8 / 1e6)}


# ------------------------------------------------------------------------------------------------
def synth_host(wl, seed, ncfg=None):
    """Host-side synthetic configurations (numpy) for the reference arm."""
    rng = np.random.default_rng(seed)
    ncfg = ncfg or wl["ncfg"]
    n, nt, nc = wl["natoms"], wl["numtypes"], wl["ncoeff"]
    kraw = nt * nc
    k = kraw + nt
    rows_raw = 7 + 3 * n
    colscale = 10.0 ** rng.uniform(-3, 0, kraw)
    raw = rng.standard_normal((ncfg * rows_raw, kraw + 1))
    raw[:, :kraw] *= colscale
    vol = rng.uniform(200.0, 2000.0, ncfg)
    r3 = raw.reshape(ncfg, rows_raw, kraw + 1)
    r3[:, 0, :kraw] *= n                                      # energy rows are divided by N
    r3[:, 1 + 3 * n:, :kraw] *= (vol / 1.6021765e6)[:, None, None]   # virial rows are scaled by 1.6e6/V
    cls = rng.choice(3, ncfg, p=[0.05, 0.85, 0.10])
    wtab = np.array([1e-2, 1.0, 100.0])
    tf = rng.dirichlet(np.ones(nt), ncfg)
    # truths: any finite numbers do for timing; b = noise around 0 keeps the solve well posed
    return dict(raw=raw, natoms=np.full(ncfg, n, dtype=np.int32), volume=vol,
                eweight=wtab[cls], fweight=wtab[(cls + 1) % 3], vweight=wtab[(cls + 2) % 3] * 1e-3,
                type_fraction=tf, blank2j=np.ones(k), k=k, ncfg=ncfg,
                energy=rng.standard_normal(ncfg) * n, forces=rng.standard_normal(3 * n * ncfg),
                stress=(lambda s: 0.5 * (s + s.transpose(0, 2, 1)))(rng.standard_normal((ncfg, 3, 3))))


def oracle_configs(h, wl, lo, hi):
    n = wl["natoms"]
    rr = 7 + 3 * n
    out = []
    for c in range(lo, hi):
        out.append(dict(block=h["raw"][c * rr:(c + 1) * rr], natoms=n, volume=h["volume"][c], energy=h["energy"][c],
                        forces=h["forces"][3 * n * c:3 * n * (c + 1)], stress=h["stress"][c],
                        eweight=h["eweight"][c], fweight=h["fweight"][c], vweight=h["vweight"][c],
                        type_fraction=h["type_fraction"][c]))
    return out


def cpu_reference_step(h, wl, ncfg_sample, faithful=True):
    """The reference's CPU path over `ncfg_sample` configurations: per-configuration row assembly
    (lammps_snap.py:391-556) + RIDGE.perform_fit (ridge.py:11-60), restated step for step in oracle/linear_fit.py
    (`assemble_as_reference`, `ridge_perform_fit_as_reference`)."""
    from oracle import linear_fit as lf
    cfgs = oracle_configs(h, wl, 0, ncfg_sample)
    t0 = time.perf_counter()
    if faithful:
        a, b, w = lf.assemble_as_reference(cfgs, wl["numtypes"], wl["ncoeff"], 0, h["blank2j"])
    else:
        a, b, w = lf.assemble(cfgs, wl["numtypes"], wl["ncoeff"], 0, h["blank2j"])
    t1 = time.perf_counter()
    if faithful:
        x, _res = lf.ridge_perform_fit_as_reference(a, b, w, ALPHA, None)
    else:
        x = lf.ridge_fit(a, b, w, ALPHA)
    t2 = time.perf_counter()
    return a.shape[0], t1 - t0, t2 - t1, x, (a, b, w)


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [v.strip() for v in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import threadpoolctl
    wl = WORKLOADS[args.workload]
    cores = os.cpu_count()
    budget_s = float(os.environ.get("FSB_REF_BUDGET_S", "420"))
    ncfg = wl["ncfg"]                     # the FULL per-GPU workload ...
    if args.ref_sample_configs:
        ncfg = min(ncfg, args.ref_sample_configs)
    h = synth_host(wl, seed=2024, ncfg=ncfg)
    t0 = time.perf_counter()
    rows, t_asm, t_fit, _, _ = cpu_reference_step(h, wl, ncfg)
    first = time.perf_counter() - t0
    total_steps = max(args.warmup, 1) + args.steps
    note = "full per-GPU workload"
    if first * total_steps > budget_s and ncfg > 100:
        # ... unless K + W steps of it would not end within the budget: bounded sample of the same workload
        ncfg = max(100, int(ncfg * budget_s / (first * total_steps)))
        note = "bounded sample (a full-size step took %.1f s; %d steps would exceed the %.0f s budget)" % (
            first, total_steps, budget_s)
    for _ in range(max(args.warmup, 1) - 1):
        cpu_reference_step(h, wl, ncfg)
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        rows, t_asm, t_fit, _, _ = cpu_reference_step(h, wl, ncfg)
        ts.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(ts))
    value = rows / (ms / 1e3)
    line = {"impl": "reference", "metric": "design_matrix_rows_per_s", "value": value, "unit": "rows/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(args.workload, args.gpus),
            "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "port",
                             "sample": "%s: %d configs = %d rows per step; reference-style row assembly "
                                       "(lammps_snap.py:430-549 incl. the dense diag(blank2J) matmul) %.2fs + "
                                       "RIDGE.perform_fit as ridge.py:24-60 (list mask, 2 copies of A, sklearn Ridge, "
                                       "residual product) %.2fs; BLAS threads %s" %
                                       (note, ncfg, rows, t_asm, t_fit,
                                        [p.get("num_threads") for p in threadpoolctl.threadpool_info()])},
            "rows_per_step": rows,
            "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class Shard:
    """Synthetic per-rank shard of a workload, generated on the device (seeded per rank)."""

    def __init__(self, eng, name, rank):
        import torch
        from fitsnap_b200.assembly import ConfigBatch, make_flags
        wl = WORKLOADS[name]
        self.name, self.wl = name, wl
        dev = eng.device
        ncfg, n, nt, nc = wl["ncfg"], wl["natoms"], wl["numtypes"], wl["ncoeff"]
        kraw, k = nt * nc, nt * nc + nt
        rows_raw = 7 + 3 * n
        n_rows = ncfg * rows_raw
        self.ncfg, self.n, self.nt, self.nc, self.kraw, self.k, self.rows_raw, self.n_rows = ncfg, n, nt, nc, kraw, k, rows_raw, n_rows
        gen = torch.Generator(device=dev).manual_seed(2024 + rank)
        colscale = 10.0 ** (torch.rand(kraw, dtype=torch.float64, device=dev, generator=gen) * -3.0)
        raw = torch.randn((ncfg * rows_raw, kraw + 1), dtype=torch.float64, device=dev, generator=gen)
        raw[:, :kraw] *= colscale
        vol = torch.rand(ncfg, dtype=torch.float64, device=dev, generator=gen) * 1800.0 + 200.0
        r3 = raw.view(ncfg, rows_raw, kraw + 1)
        r3[:, 0, :kraw] *= n
        r3[:, 1 + 3 * n:, :kraw] *= (vol / 1.6021765e6)[:, None, None]
        cls = torch.multinomial(torch.tensor([0.05, 0.85, 0.10], device=dev), ncfg, replacement=True, generator=gen)
        wtab = torch.tensor([1e-2, 1.0, 100.0], dtype=torch.float64, device=dev)
        tf = torch.rand((ncfg, nt), dtype=torch.float64, device=dev, generator=gen)
        tf = tf / tf.sum(1, keepdim=True)
        x_true = torch.randn(k, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
        raw_off = torch.arange(ncfg + 1, dtype=torch.int64, device=dev) * rows_raw
        self.batch = batch = ConfigBatch(
            raw=raw, raw_row_off=raw_off, out_row_off=raw_off.clone(),
            natoms=torch.full((ncfg,), n, dtype=torch.int32, device=dev), volume=vol,
            energy=torch.zeros(ncfg, dtype=torch.float64, device=dev),
            forces=torch.zeros(3 * n * ncfg, dtype=torch.float64, device=dev),
            stress=torch.zeros((ncfg, 9), dtype=torch.float64, device=dev),
            eweight=wtab[cls], fweight=wtab[(cls + 1) % 3], vweight=wtab[(cls + 2) % 3] * 1e-3,
            type_fraction=tf, blank2j=torch.ones(k, dtype=torch.float64, device=dev),
            ncfg=ncfg, numtypes=nt, ncoeff=nc, flags=make_flags(True, True, True, False), k=k,
            row_begin=0, row_end=n_rows,
            row_cfg=torch.arange(ncfg, dtype=torch.int32, device=dev).repeat_interleave(rows_raw))
        self.A = torch.empty((n_rows, k), dtype=torch.float64, device=dev)
        self.b = torch.empty(n_rows, dtype=torch.float64, device=dev)
        self.w = torch.empty(n_rows, dtype=torch.float64, device=dev)
        eng.scatter(batch, self.A, self.b, self.w)
        y = eng.predict(self.A, x_true) + 1e-3 * torch.randn(n_rows, dtype=torch.float64, device=dev, generator=gen)
        y2 = y.view(ncfg, rows_raw)
        ref = r3[:, :, kraw]
        batch.energy = (y2[:, 0] * n + ref[:, 0]).contiguous()
        batch.forces = (y2[:, 1:1 + 3 * n] + ref[:, 1:1 + 3 * n]).reshape(-1).contiguous()
        sv = y2[:, 1 + 3 * n:] + ref[:, 1 + 3 * n:]
        st = torch.zeros((ncfg, 3, 3), dtype=torch.float64, device=dev)
        for q, (i_, j_) in enumerate(zip([0, 1, 2, 1, 0, 0], [0, 1, 2, 2, 2, 1])):
            st[:, i_, j_] = sv[:, q]
            st[:, j_, i_] = sv[:, q]
        batch.stress = st.reshape(ncfg, 9).contiguous()
        self.out = (self.A, self.b, self.w)


def gather_objects(obj, world, group):
    import torch.distributed as dist
    if world == 1:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj, group=group)
    return out


def host_reference_solution(sh, world, group, x_dev_host, alpha):
    """Coefficients of the WHOLE sharded system from the oracle, independent of the device Gram and of the device
    collective: every rank forms [aw|bw]^T[aw|bw] of ITS shard with numpy (oracle.linear_fit.gram), the 101 x 101
    blocks travel out of band (pickled objects), rank 0 adds them, solves the equilibrated ridge system and refines
    once against the oracle's residual of every shard."""
    from oracle import linear_fit as lf
    import scipy.linalg as sl
    a = sh.A.cpu().numpy()
    b = sh.b.cpu().numpy()
    w = sh.w.cpu().numpy()
    G, c, _btb, _n = lf.gram(a, b, w, None)
    parts = gather_objects((G, c), world, group)
    k = G.shape[0]
    Gs = np.sum([p[0] for p in parts], axis=0) + alpha * np.eye(k)
    cs = np.sum([p[1] for p in parts], axis=0)
    d = 1.0 / np.sqrt(np.diag(Gs))
    cf = sl.cho_factor(Gs * d[:, None] * d[None, :], lower=True)
    solve = lambda r: d * sl.cho_solve(cf, d * r)
    x = solve(cs)
    aw, bw = lf.weighted_system(a, b, w, None)
    for _ in range(2):
        x = np.array(gather_objects(x, world, group)[0])          # rank 0's iterate everywhere
        g = aw.T @ (bw - aw @ x)
        gs = np.sum(gather_objects(g, world, group), axis=0) - alpha * x
        x = x + solve(gs)
    x = np.array(gather_objects(x, world, group)[0])
    mr, l2, _ = lf.coeff_rel_err(x_dev_host, x)
    return {"max_rel_vs_host_ridge_of_all_shards": mr, "l2_rel": l2, "rows_total": int(a.shape[0]) * world,
            "how": "oracle Gram of every shard (numpy), summed on the host, equilibrated Cholesky + 2 refinement "
                   "rounds against the oracle residual; independent of the device Gram and collective"}


def sample_parity(eng, sh, alpha, nrows):
    """Device fit of the first `nrows` rows of this rank's shard (same kernels and Gram path as the timed step)
    against the oracle's exact ridge statement on exactly those rows."""
    from oracle import linear_fit as lf
    res = eng.fit(sh.A[:nrows], sh.b[:nrows], sh.w[:nrows], None, alpha=alpha, refine=REFINE, diagnostics=False)
    a, b, w = sh.A[:nrows].cpu().numpy(), sh.b[:nrows].cpu().numpy(), sh.w[:nrows].cpu().numpy()
    t0 = time.perf_counter()
    ref = lf.ridge_fit_exact(a, b, w, alpha)
    dt = time.perf_counter() - t0
    mr, l2, _ = lf.coeff_rel_err(res.coefficients(), ref)
    return {"max_rel_vs_exact_ridge": mr, "l2_rel": l2, "rows": int(nrows), "gram_path": eng.gram_path(nrows, sh.k),
            "oracle_seconds": round(dt, 2)}


def run_workload(eng, name, rank, world, group, steps, warmup, full, args, peaks):
    """Time one workload; returns the dict that becomes the JSON line (full=True) or an entry of "workloads"."""
    import torch
    import torch.distributed as dist
    from fitsnap_b200.pipeline import LinearFitPipeline
    hbm_peak, bf16_peak, bf16_sus, peak_kind = peaks
    dev = eng.device
    sh = Shard(eng, name, rank)
    k, n_rows, kraw = sh.k, sh.n_rows, sh.kraw
    batch, out = sh.batch, sh.out
    pipe = LinearFitPipeline(sh.nt, sh.nc, False, np.ones(k), alpha=ALPHA, refine=REFINE, group=group, engine=eng)
    comm = eng.comm_for(group) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()

    def allreduce(t):
        if comm is not None:
            comm.all_reduce(t)

    for _ in range(warmup):
        res = pipe.fit_batch(batch, None, out)
    fused = (not args.no_fused) and eng.scatter_gram(batch, *out) is not None     # narrow layouts: K1 + K2..K4 in one kernel
    pipe.fuse_scatter_gram = fused
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(dev.index) if (rank == 0 and full) else None
    launches0 = eng.launch_count
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    phase_ev = []
    for _ in range(steps):
        # the step, with CUDA events between its phases on the stream the kernels are launched on
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        if fused:
            Ad, bd, wd, _bad, gaug = eng.scatter_gram(batch, *out)
            ev[1].record()
        else:
            Ad, bd, wd, _bad = eng.scatter(batch, *out)
            ev[1].record()
            gaug = eng.gram(Ad, bd, wd, None)
        ev[2].record()
        allreduce(gaug)
        f = eng.factor(gaug, ALPHA)
        x = eng.solve(f, gaug[:, k], rhs_stride=k + 1)
        ev[3].record()
        for _r in range(REFINE):
            g = eng.residual(Ad, bd, wd, None, x)
            allreduce(g)
            x = eng.solve(f, g, x_in=x)
        ev[4].record()
        phase_ev.append(ev)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    launches = (eng.launch_count - launches0) // steps
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    ms_step = float(t.item())
    mean_ms = lambda i, j: float(np.mean([ev[i].elapsed_time(ev[j]) for ev in phase_ev]))
    if fused:
        gram_ms = mean_ms(0, 2)
        phases_ms = {"scatter+gram (one fused kernel + split-K reduction)": gram_ms}
    else:
        gram_ms = mean_ms(1, 2)
        phases_ms = {"scatter": mean_ms(0, 1), "gram": gram_ms}
    phases_ms["allreduce_factor_solve"] = mean_ms(2, 3)
    phases_ms["refine_%dx(residual+allreduce+solve)" % REFINE] = mean_ms(3, 4)
    value = world * n_rows / (ms_step / 1e3)
    x_dev = x.detach().cpu().numpy()
    gram_path = eng.gram_path(n_rows, k)

    # ---- bit-equality of the replicated solve across ranks ----------------------------------------
    same_x = None
    if world > 1:
        xs = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(xs, x.contiguous(), group=group)
        same_x = bool(all(torch.equal(xs[0], v) for v in xs))

    # ---- the same step replayed from a CUDA graph ------------------------------------------------------
    graph_info = None
    if full or args.graph_all:
        if world > 1 and not (comm.uses_peer((k + 1) * (k + 1))):
            graph_info = {"skipped": "the all-reduce of this shape goes through NCCL; only the peer-window collective "
                                     "is captured"}
        else:
            try:
                cap = pipe.capture(batch, None, out)
                for _ in range(3):
                    cap.replay()
                torch.cuda.synchronize()
                barrier()
                q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                q0.record()
                for _ in range(steps):
                    rg = cap.replay()
                q1.record()
                torch.cuda.synchronize()
                barrier()
                tg = torch.tensor([q0.elapsed_time(q1) / steps], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(tg, op=dist.ReduceOp.MAX, group=group)
                graph_info = {"ms_per_step": float(tg.item()), "rows_per_s": world * n_rows / (float(tg.item()) / 1e3),
                              "kernels_per_replay": int(cap.launches), "same_x_as_eager": bool(torch.equal(rg.x, x))}
                del cap
            except Exception as exc:                  # report, never hide: the eager numbers above stand on their own
                graph_info = {"error": repr(exc)[:200]}

    clocks = sampler.stop() if sampler else None
    # The headline step is the graph replay when it exists and reproduces the eager coefficients bit for bit: the same
    # kernels on the same buffers, launched by one cudaGraphLaunch instead of 12-15 calls through Python / ctypes
    # (the guide's own advice for launch-bound inner loops).  The eager numbers and the per-phase events stay beside it.
    launch_mode, eager = "eager", None
    ok_here = bool(graph_info and graph_info.get("same_x_as_eager") and graph_info.get("ms_per_step"))
    if full and not args.eager_headline:
        okt = torch.tensor([1.0 if ok_here else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(okt, op=dist.ReduceOp.MIN, group=group)
        if float(okt.item()) == 1.0:
            eager = {"ms_per_step": ms_step, "rows_per_s": value}
            ms_step = graph_info["ms_per_step"]
            value = graph_info["rows_per_s"]
            launch_mode = "cuda_graph_replay"

    # ---- parity of the timed path ---------------------------------------------------------------------
    coeff_err = None
    if not args.no_parity:
        if n_rows * k <= 2.0e8:         # host Gram of every shard is seconds
            coeff_err = host_reference_solution(sh, world, group, x_dev, ALPHA)
        else:                           # wide shapes: a sample of rank 0's shard through the same Gram path
            ns = min(n_rows, 81920)
            coeff_err = sample_parity(eng, sh, ALPHA, ns) if rank == 0 else None
            # optimality of the timed (sharded) solution itself: the gradient of the ridge objective at x, formed by
            # the streaming residual kernel (independent of the Gram) and all-reduced
            gopt = eng.residual(sh.A, sh.b, sh.w, None, x)
            cnorm = gaug[:k, k].clone()
            allreduce(gopt)
            gopt = gopt - ALPHA * x
            if rank == 0:
                coeff_err["timed_solution_gradient_over_rhs"] = float(gopt.abs().max() / cnorm.abs().max())
        if coeff_err is not None and same_x is not None:
            coeff_err["x_bit_identical_on_all_ranks"] = same_x

    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    flops = (2.0 * k * k + 2.0 * k) * n_rows
    achieved = flops / (gram_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s",
                "frac": achieved / bf16_peak, "traffic": None,
                "peak_kind": "dense bf16 cuBLAS, %s (MEASURED_PEAKS.json)" % peak_kind,
                "achieved_kind": "algorithmic fp64 flops (2k^2+2k per row, full Gram convention) / Gram time",
                "hbm_gbs_during_gram": (16.0 * k + 24.0 if fused else 8.0 * (k + 2)) * n_rows / (gram_ms * 1e-3) / 1e9,
                "hbm_peak_gbs": hbm_peak}
    if fused:
        roofline["fused_note"] = ("scatter_gram_kernel: the raw blocks are scattered into A, b, w AND contracted in the "
                                  "same kernel; its time covers both (algorithmic bytes 16k+24 per row)")
    try:    # ncu-measured DRAM bytes per launch of the dominant kernel, when a capture of this shape is committed
        tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("%s:%s" % (name, "fused" if fused else gram_path))
        if tr:
            roofline["traffic"] = tr["bytes_per_launch"]
            roofline["traffic_note"] = "%s, %d rows x %d: %.3g B from ncu vs %.3g B algorithmic (%s)" % (
                tr["kernel"], tr["rows"], tr["k"], tr["bytes_per_launch"], tr["algorithmic_bytes"], tr["source"])
    except (OSError, ValueError):
        pass
    if gram_path == "int8":
        n_i = -(-(k + 1) // 128)
        ops = sum(2.0 * 128 * (256 if (2 * jj + 1 < n_i and 2 * jj + 1 <= i_) else 128) * n_rows * 16
                  for i_ in range(n_i) for jj in range(i_ // 2 + 1))
        top = ops / (gram_ms * 1e-3) / 1e12
        roofline.update({
            "kernel": "int8 tcgen05 Gram (tcgen05.mma kind::i8, exact integer Gram through 16 CRT moduli)",
            "int8_top_s_executed_over_whole_gram": top,
            "int8_peak_top_s": 2.0 * bf16_peak,
            "int8_peak_kind": "2 x the measured dense bf16 peak (kind::i8 issues at twice the bf16 rate; no int8 "
                              "figure in MEASURED_PEAKS.json)",
            "frac_of_int8_peak": top / (2.0 * bf16_peak),
            "frac_of_int8_peak_sustained": top / (2.0 * bf16_sus),
            "fp64_equivalent_tflops": achieved})
    else:
        roofline.update({
            "kernel": ("scatter_gram_kernel" if fused else "gram_rowsplit_kernel" if k + 1 <= 104 else
                       ("gram_dmma_kernel" if k + 1 <= 128 else "preweight_kernel + gram_tma_kernel")) +
                      " (+gram_reduce_kernel)",
            "frac_of_fp64_dmma_peak": achieved / 37.1,
            "fp64_note": "tcgen05 has no f64 kind; this kernel runs on DMMA.8x8x4 whose measured peak on this pool is "
                         "37.1 TFLOP/s (tools/ubench/fp64_rates.cu); algorithmic flops count the full K x K Gram, "
                         "the kernel executes the lower triangle only"})
    entry = {
        "ms_per_step": ms_step, "rows_per_s": value, "steps": steps, "warmup": warmup,
        "launch_mode": launch_mode, "eager": eager,
        "config": make_config(name, world, gram_path),
        "gram_ms": gram_ms, "gram_tflops_algorithmic": achieved,
        "phases_ms_rank0": phases_ms,
        "scatter_GBps": None if fused else (16.0 * k + 24.0) * n_rows / (phases_ms["scatter"] * 1e-3) / 1e9,
        "scatter_frac_of_hbm_peak": None if fused else
        (16.0 * k + 24.0) * n_rows / (phases_ms["scatter"] * 1e-3) / 1e9 / hbm_peak,
        "fused_scatter_gram": bool(fused),
        "roofline": roofline, "coeff_max_rel_err": coeff_err, "gpu_launches": int(launches),
        "cuda_graph_replay": graph_info,
        "collective": comm.info() if comm is not None else None,
    }
    return entry, sh, pipe, clocks, x_dev


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--secondary", default="c4_shard", choices=sorted(WORKLOADS))
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gram-path", default="auto", choices=["auto", "fp64", "int8"],
                    help="Gram arithmetic: fp64 DMMA, int8 tcgen05 (exact integer, CRT), or the library's choice")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-fused", action="store_true", help="scatter and Gram as two kernels even for narrow layouts")
    ap.add_argument("--graph-all", action="store_true", help="CUDA-graph replay for the secondary workload too")
    ap.add_argument("--eager-headline", action="store_true",
                    help="report the eagerly launched step as the headline even when the graph replay is available")
    ap.add_argument("--ref-sample-configs", type=int, default=0, help="reference arm: cap the configurations per step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # pin this rank's host threads (and therefore its pinned staging pages) to the NUMA node of its GPU
    from fitsnap_b200.distributed import bind_to_gpu_numa
    numa = bind_to_gpu_numa(local)

    import torch
    import torch.distributed as dist
    from fitsnap_b200.engine import Engine

    torch.cuda.set_device(local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        group = dist.group.WORLD
    eng = Engine(local)
    eng.set_gram_path(args.gram_path)
    dev = eng.device
    peaks = load_peaks()
    wl = WORKLOADS[args.workload]

    entry, sh, pipe, clocks, x_dev = run_workload(eng, args.workload, rank, world, group, args.steps, args.warmup, True,
                                                  args, peaks)
    n_rows, k, ncfg, n = sh.n_rows, sh.k, sh.ncfg, sh.n
    batch = sh.batch

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- CPU baseline: the oracle port on a bounded sample (rank 0, N = 1) ---------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import linear_fit as lf
        import threadpoolctl
        sample_cfg = max(1, min(ncfg, int(250_000 // sh.rows_raw)))
        rr = sh.rows_raw
        hs = dict(raw=batch.raw[:sample_cfg * rr].cpu().numpy(), volume=batch.volume[:sample_cfg].cpu().numpy(),
                  energy=batch.energy[:sample_cfg].cpu().numpy(), forces=batch.forces[:3 * n * sample_cfg].cpu().numpy(),
                  stress=batch.stress[:sample_cfg].cpu().numpy().reshape(sample_cfg, 3, 3),
                  eweight=batch.eweight[:sample_cfg].cpu().numpy(), fweight=batch.fweight[:sample_cfg].cpu().numpy(),
                  vweight=batch.vweight[:sample_cfg].cpu().numpy(),
                  type_fraction=batch.type_fraction[:sample_cfg].cpu().numpy(), blank2j=np.ones(k))
        cpu_reference_step(hs, wl, min(sample_cfg, 50))        # warm-up (imports, BLAS threads)
        rows_s, t_asm, t_fit, x_cpu, (a_s, b_s, w_s) = cpu_reference_step(hs, wl, sample_cfg)
        cpu_baseline = {"value": rows_s / (t_asm + t_fit), "unit": "rows/s", "cores": os.cpu_count(), "kind": "port",
                        "sample": "first %d configs = %d rows of this workload: reference-style row assembly %.2fs + "
                                  "RIDGE.perform_fit as ridge.py:24-60 %.2fs (BLAS threads %s)" %
                                  (sample_cfg, rows_s, t_asm, t_fit,
                                   [p.get("num_threads") for p in threadpoolctl.threadpool_info()])}
        if entry["coeff_max_rel_err"] is not None:
            # the sample also pins the scatter (bit-exact) and the device fit of exactly these rows
            assert np.array_equal(sh.A[:rows_s].cpu().numpy(), a_s), "device scatter differs from the oracle"
            res_s = eng.fit(sh.A[:rows_s], sh.b[:rows_s], sh.w[:rows_s], None, alpha=ALPHA, refine=REFINE,
                            diagnostics=False)
            mr, l2, _ = lf.coeff_rel_err(res_s.coefficients(), lf.ridge_fit_exact(a_s, b_s, w_s, ALPHA))
            entry["coeff_max_rel_err"].update({"sample_rows": int(rows_s), "sample_max_rel_vs_exact_ridge": mr,
                                               "sample_max_rel_vs_sklearn_ridge":
                                                   lf.coeff_rel_err(res_s.coefficients(), x_cpu)[0],
                                               "scatter_bit_exact": True})

    # ---- e2e: host buffers through the public API -------------------------------------------------
    e2e = None
    e2e_plugin = None
    if not args.no_e2e:
        host = {}
        for name in ("raw", "volume", "energy", "forces", "stress", "eweight", "fweight", "vweight", "type_fraction"):
            tdev = getattr(batch, name)
            th = torch.empty(tdev.shape, dtype=tdev.dtype, pin_memory=True)
            th.copy_(tdev)
            host[name] = th.numpy()
        natoms_h = np.full(ncfg, n, dtype=np.int32)
        torch.cuda.synchronize()

        def e2e_step():
            xh, r_, b_ = pipe.fit_host(host["raw"], natoms_h, host["volume"], host["energy"], host["forces"],
                                       host["stress"].reshape(ncfg, 3, 3), host["eweight"], host["fweight"],
                                       host["vweight"], host["type_fraction"])
            return xh, b_

        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            xh, b_ = e2e_step()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            xh, b_ = e2e_step()
        torch.cuda.synchronize()
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX, group=group)
        dt = float(tt.item())
        # context for the e2e number: the raw PCIe rate of this box (one pinned copy of the blocks, device-timed)
        raw_h = torch.from_numpy(host["raw"])
        dst = torch.empty_like(batch.raw)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dst.copy_(raw_h, non_blocking=True)
        c0.record()
        dst.copy_(raw_h, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_gbps = raw_h.numel() * 8 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        del dst
        e2e = {"value": world * n_rows / dt, "unit": "rows/s", "h2d_bytes_per_step": int(b_.h2d_bytes),
               "chunks": int(getattr(b_, "chunks", 1)), "pcie_h2d_GBps": round(h2d_gbps, 2),
               "pcie_bound_ms": round(int(b_.h2d_bytes) / (h2d_gbps * 1e9) * 1e3, 3),
               "d2h_bytes_per_step": int(8 * k + 4), "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "api": "fitsnap_b200.pipeline.LinearFitPipeline.fit_host (pinned host raw blocks -> H2D -> scatter -> "
                      "fit -> D2H coefficients)", "host_numa_binding": numa,
               "max_abs_diff_vs_device_resident_x": float(np.max(np.abs(xh - x_dev)))}

        # ---- the reference-facing solver call on ordinary numpy arrays (rank-local rows, sharded over `group`) ----
        from types import SimpleNamespace
        from fitsnap_b200.solvers import RIDGE
        a_np = np.empty((n_rows, k))              # pageable, as a FitSNAP user's pt.shared_arrays['a'].array is
        a_np[...] = sh.A.cpu().numpy()
        b_np, w_np = sh.b.cpu().numpy().copy(), sh.w.cpu().numpy().copy()
        cfg = SimpleNamespace(sections={"RIDGE": SimpleNamespace(alpha=ALPHA, local_solver=0),
                                        "EXTRAS": SimpleNamespace(apply_transpose=0)})
        pt = SimpleNamespace(_rank=rank, shared_arrays={}, fitsnap_dict={})
        solver = RIDGE("RIDGE", pt, cfg)
        solver.engine, solver.process_group, solver.refine = eng, group, REFINE
        for _ in range(2):
            solver.perform_fit(a=a_np, b=b_np, w=w_np, trainall=True)
        torch.cuda.synchronize()
        barrier()
        p_steps = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(p_steps):
            solver.perform_fit(a=a_np, b=b_np, w=w_np, trainall=True)
        barrier()
        dtp = (time.perf_counter() - t0) / p_steps
        tt = torch.tensor([dtp], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX, group=group)
        dtp = float(tt.item())
        e2e_plugin = {"value": world * n_rows / dtp, "unit": "rows/s", "ms_per_step": dtp * 1e3, "steps": p_steps,
                      "h2d_bytes_per_step": int(a_np.nbytes + b_np.nbytes + w_np.nbytes), "d2h_bytes_per_step": 8 * k,
                      "api": "fitsnap_b200.solvers.RIDGE.perform_fit(a=A, b=b, w=w, trainall=True) on pageable numpy "
                             "arrays (ridge.py:11 signature): pinned-ring upload + Gram + solve + refinement + D2H; "
                             "scatter not included (A is given)",
                      "upload_GBps": round((a_np.nbytes + b_np.nbytes + w_np.nbytes) / dtp / 1e9, 2),
                      "max_abs_diff_vs_device_resident_x": float(np.max(np.abs(solver.fit - x_dev)))}
        del a_np

        # ---- calculator -> solver hand-off: one BlockCollector.add per configuration (what _collect_lammps does),
        #      one flush (H2D + scatter), fit from the rows left on the device
        from fitsnap_b200.calculators import BlockCollector
        col = BlockCollector(eng, sh.nt, sh.nc, False, np.ones(k), {"A": 1, "B": 2}, capacity_rows=ncfg * sh.rows_raw)
        types = ["A"] * (n // 2) + ["B"] * (n - n // 2)
        raw3 = host["raw"].reshape(ncfg, sh.rows_raw, kraw_of(sh) + 1)
        f3 = host["forces"].reshape(ncfg, n, 3)
        s3 = host["stress"].reshape(ncfg, 3, 3)

        def handoff():
            col.reset()
            for c in range(ncfg):
                col.add(raw3[c], n, host["volume"][c], host["energy"][c], f3[c], s3[c], host["eweight"][c],
                        host["fweight"][c], host["vweight"][c], types)
            A_, b__, w__, _bad, _batch = col.flush()
            r = eng.fit(A_, b__, w__, None, alpha=ALPHA, refine=REFINE, group=group, diagnostics=False)
            return r.coefficients()

        handoff()
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            handoff()
        barrier()
        dth = (time.perf_counter() - t0) / 2
        e2e_plugin["calculator_handoff"] = {
            "ms_per_step": dth * 1e3, "rows_per_s_per_gpu": n_rows / dth,
            "what": "BlockCollector.add x %d configurations (host staging, as the drop-in _collect_lammps) + flush "
                    "(H2D + scatter) + fit from device-resident rows + D2H" % ncfg}

    # ---- the target shape (north star): 1.25e6 x 1000 per GPU -> int8 tcgen05 Gram -------------------------------
    workloads = {}
    if not args.no_secondary and args.secondary != args.workload:
        del sh, pipe, batch
        torch.cuda.empty_cache()
        try:
            e2, sh2, _p2, _c2, _x2 = run_workload(eng, args.secondary, rank, world, group, max(3, min(args.steps, 5)), 3,
                                                  False, args, peaks)
            workloads[args.secondary] = e2
            del sh2, _p2
        except Exception as exc:
            workloads[args.secondary] = {"error": repr(exc)[:300]}
        torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": "design_matrix_rows_per_s", "value": entry["rows_per_s"], "unit": "rows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": entry["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": entry["config"],
            "gram_tflops_algorithmic": entry["gram_tflops_algorithmic"],
            "gram_ms": entry["gram_ms"],
            "phases_ms_rank0": entry["phases_ms_rank0"],
            "scatter_GBps": entry["scatter_GBps"], "scatter_frac_of_hbm_peak": entry["scatter_frac_of_hbm_peak"],
            "roofline": entry["roofline"],
            "coeff_max_rel_err": entry["coeff_max_rel_err"],
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "e2e_plugin": e2e_plugin,
            "gpu_launches": entry["gpu_launches"],
            "launch_mode": entry["launch_mode"], "eager": entry["eager"],
            "cuda_graph_replay": entry["cuda_graph_replay"],
            "collective": entry["collective"],
            "workloads": workloads,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def kraw_of(sh):
    return sh.kraw


if __name__ == "__main__":
    main()
