#!/usr/bin/env python
"""bench.py -- FitSNAP linear-fit hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c5|c4s] [--impl reference]

One STEP = one pass of the hot path over one batch of synthetic configurations:
    scatter (raw LAMMPS blocks -> A, b, w)  ->  fused mask/weight/Gram  ->  [NCCL all-reduce]
    ->  equilibrated Cholesky solve  ->  2 rounds of refinement streamed from A.
Default workload = BASELINE.json configs[1]: synthetic A 1e6 x 100 fp64, ridge alpha 1e-6, per GPU
(weak scaling: every rank owns a 1e6-row shard; one all-reduce of the 101x101 Gram + one
100-vector all-reduce per refinement round).

value  : rows/s, inputs resident in HBM (raw blocks on device), whole job over N GPUs.
e2e    : same metric through the public host API (`LinearFitPipeline.fit_host`): pinned HOST
         raw blocks -> H2D -> same device path -> D2H of the coefficients, all inside the timed region.
roofline : the dominant kernel (fused Gram) timed with CUDA events on its stream inside the timed steps.
cpu_baseline : the oracle (numpy restatement of the reference + scipy/sklearn, kind "port") on the
         host cores, on a bounded sample of the same workload (rank 0, N=1 only).
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
    # the whole BASELINE configs[3] matrix on ONE GPU: 80 GB of raw blocks + 80 GB of A (run with --no-e2e)
    "c4": dict(ncfg=100000, natoms=31, numtypes=2, ncoeff=499, desc="ACE-like 1e7 x 1000, the full BASELINE configs[3] matrix on one GPU"),
}
ALPHA = 1.0e-6
REFINE = 2


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


# ------------------------------------------------------------------------------------------------
def synth_host(wl, seed, ncfg=None):
    """Host-side synthetic configurations (numpy), used by the CPU baseline / reference arm and,
    for the e2e leg, as the pinned host buffers."""
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
    return dict(raw=raw, natoms=np.full(ncfg, n, dtype=np.int32), volume=vol,
                eweight=wtab[cls], fweight=wtab[(cls + 1) % 3], vweight=wtab[(cls + 2) % 3] * 1e-3,
                type_fraction=tf, blank2j=np.ones(k), x_true=rng.standard_normal(k), k=k, ncfg=ncfg,
                noise_seed=seed + 1)


def finish_truths(h, a_rows_times_x, wl):
    """Given y = A x_true per output row, set Energy/Forces/Stress so that b = y + 1e-3 N(0,1)."""
    rng = np.random.default_rng(h["noise_seed"])
    ncfg, n = h["ncfg"], wl["natoms"]
    kraw = wl["numtypes"] * wl["ncoeff"]
    y = a_rows_times_x.reshape(ncfg, 7 + 3 * n) + 1e-3 * rng.standard_normal((ncfg, 7 + 3 * n))
    r3 = h["raw"].reshape(ncfg, 7 + 3 * n, kraw + 1)
    ref = r3[:, :, kraw]
    h["energy"] = y[:, 0] * n + ref[:, 0]
    h["forces"] = (y[:, 1:1 + 3 * n] + ref[:, 1:1 + 3 * n]).reshape(-1)
    st = np.zeros((ncfg, 3, 3))
    vi, vj = [0, 1, 2, 1, 0, 0], [0, 1, 2, 2, 2, 1]
    sv = y[:, 1 + 3 * n:] + ref[:, 1 + 3 * n:]
    for q in range(6):
        st[:, vi[q], vj[q]] = sv[:, q]
        st[:, vj[q], vi[q]] = sv[:, q]
    h["stress"] = st
    return h


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


def cpu_reference_step(h, wl, ncfg_sample):
    """The reference's CPU path on a bounded sample: per-configuration row assembly
    (lammps_snap.py:391-556 restated in oracle/linear_fit.py) + RIDGE.perform_fit (ridge.py:11-60)."""
    from oracle import linear_fit as lf
    t0 = time.perf_counter()
    a, b, w = lf.assemble(oracle_configs(h, wl, 0, ncfg_sample), wl["numtypes"], wl["ncoeff"], 0, h["blank2j"])
    t1 = time.perf_counter()
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
def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import threadpoolctl
    cores = os.cpu_count()
    # bounded sample: ~100k rows per step keeps K steps within minutes
    sample_cfg = max(1, min(wl["ncfg"], int(100_000 // (7 + 3 * wl["natoms"]))))
    h = synth_host(wl, seed=2024, ncfg=sample_cfg)
    from oracle import linear_fit as lf
    a0, _, _ = lf.assemble(oracle_configs(dict(h, energy=np.zeros(sample_cfg), forces=np.zeros(3 * wl["natoms"] * sample_cfg),
                                               stress=np.zeros((sample_cfg, 3, 3))), wl, 0, sample_cfg),
                           wl["numtypes"], wl["ncoeff"], 0, h["blank2j"])
    finish_truths(h, a0 @ h["x_true"], wl)
    for _ in range(max(args.warmup, 1)):
        cpu_reference_step(h, wl, sample_cfg)
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        rows, t_asm, t_fit, _, _ = cpu_reference_step(h, wl, sample_cfg)
        ts.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(ts))
    value = rows / (ms / 1e3)
    line = {"impl": "reference", "metric": "design_matrix_rows_per_s", "value": value, "unit": "rows/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "k": h["k"], "alpha": ALPHA,
                       "sample_rows_per_step": rows},
            "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "port",
                             "sample": "%d configs = %d rows of the same workload per step; oracle assemble "
                                       "(%.2fs) + sklearn Ridge (%.2fs); BLAS threads %s" %
                                       (sample_cfg, rows, t_asm, t_fit,
                                        [p.get("num_threads") for p in threadpoolctl.threadpool_info()])},
            "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gram-path", default="auto", choices=["auto", "fp64", "int8"],
                    help="Gram arithmetic: fp64 DMMA, int8 tcgen05 (exact integer, CRT), or the library's choice")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist
    from fitsnap_b200.engine import Engine
    from fitsnap_b200.pipeline import LinearFitPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        group = dist.group.WORLD
    eng = Engine(local)
    eng.set_gram_path(args.gram_path)
    dev = eng.device
    hbm_peak, bf16_peak, peak_kind = load_peaks()

    # ---- synthetic shard, generated on the device (seeded per rank) -------------------------------
    ncfg, n, nt, nc = wl["ncfg"], wl["natoms"], wl["numtypes"], wl["ncoeff"]
    kraw, k = nt * nc, nt * nc + nt
    rows_raw = 7 + 3 * n
    n_rows = ncfg * rows_raw
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

    from fitsnap_b200.assembly import ConfigBatch, make_flags
    raw_off = torch.arange(ncfg + 1, dtype=torch.int64, device=dev) * rows_raw
    batch = ConfigBatch(raw=raw, raw_row_off=raw_off, out_row_off=raw_off.clone(),
                        natoms=torch.full((ncfg,), n, dtype=torch.int32, device=dev), volume=vol,
                        energy=torch.zeros(ncfg, dtype=torch.float64, device=dev),
                        forces=torch.zeros(3 * n * ncfg, dtype=torch.float64, device=dev),
                        stress=torch.zeros((ncfg, 9), dtype=torch.float64, device=dev),
                        eweight=wtab[cls], fweight=wtab[(cls + 1) % 3], vweight=wtab[(cls + 2) % 3] * 1e-3,
                        type_fraction=tf, blank2j=torch.ones(k, dtype=torch.float64, device=dev),
                        ncfg=ncfg, numtypes=nt, ncoeff=nc, flags=make_flags(True, True, True, False), k=k,
                        row_begin=0, row_end=n_rows,
                        row_cfg=torch.arange(ncfg, dtype=torch.int32, device=dev).repeat_interleave(rows_raw))
    pipe = LinearFitPipeline(nt, nc, False, np.ones(k), alpha=ALPHA, refine=REFINE, group=group, engine=eng)
    A = torch.empty((n_rows, k), dtype=torch.float64, device=dev)
    bvec = torch.empty(n_rows, dtype=torch.float64, device=dev)
    wvec = torch.empty(n_rows, dtype=torch.float64, device=dev)
    eng.scatter(batch, A, bvec, wvec)
    y = eng.predict(A, x_true) + 1e-3 * torch.randn(n_rows, dtype=torch.float64, device=dev, generator=gen)
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
    del y, y2, sv, st
    out = (A, bvec, wvec)

    def step():
        return pipe.fit_batch(batch, None, out)

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- value: device-resident timed region ------------------------------------------------------
    for _ in range(args.warmup):
        res = step()
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    gram_ev = []
    launches0 = eng.launch_count
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    phase_ev = []
    for _ in range(args.steps):
        # the step, with CUDA events between its phases on the stream the kernels are launched on
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        Ad, bd, wd, _bad = eng.scatter(batch, *out)
        ev[1].record()
        gaug = eng.gram(Ad, bd, wd, None)
        ev[2].record()
        gram_ev.append((ev[1], ev[2]))
        if world > 1:
            dist.all_reduce(gaug, group=group)
        f = eng.factor(gaug, ALPHA)
        x = eng.solve(f, gaug[:, k], rhs_stride=k + 1)
        ev[3].record()
        for _r in range(REFINE):
            g = eng.residual(Ad, bd, wd, None, x)
            if world > 1:
                dist.all_reduce(g, group=group)
            x = eng.solve(f, g, x_in=x)
        ev[4].record()
        phase_ev.append(ev)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = (eng.launch_count - launches0) // args.steps
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    t = torch.tensor([ms_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    ms_step = float(t.item())
    gram_ms = float(np.mean([a_.elapsed_time(b_) for a_, b_ in gram_ev]))
    names = ("scatter", "gram", "allreduce_factor_solve", "refine_%dx(residual+allreduce+solve)" % REFINE)
    phases_ms = {nm: float(np.mean([ev[i].elapsed_time(ev[i + 1]) for ev in phase_ev])) for i, nm in enumerate(names)}
    phase_rows_per_s = {nm: n_rows / (ms * 1e-3) for nm, ms in phases_ms.items() if ms > 0}
    value = world * n_rows / (ms_step / 1e3)

    # ---- the same step replayed from a CUDA graph (one cudaGraphLaunch instead of `launches` launches) ----
    # (single GPU only: capturing the NCCL all-reduces of the sharded step hung the 2-GPU run of this round)
    graph_info = {"skipped": "single-GPU only (NCCL all-reduce inside a capture hung at 2 GPUs)"} if world > 1 else None
    try:
        if world > 1:
            raise StopIteration
        cap = pipe.capture(batch, None, out)
        for _ in range(3):
            cap.replay()
        torch.cuda.synchronize()
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(args.steps):
            rg = cap.replay()
        q1.record()
        torch.cuda.synchronize()
        barrier()
        tg = torch.tensor([q0.elapsed_time(q1) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tg, op=dist.ReduceOp.MAX, group=group)
        graph_info = {"ms_per_step": float(tg.item()), "rows_per_s": world * n_rows / (float(tg.item()) / 1e3),
                      "kernels_per_replay": int(cap.launches),
                      "same_x_as_eager": bool(torch.equal(rg.x, x))}
    except StopIteration:
        pass
    except Exception as exc:                      # report, never hide: the eager numbers above stand on their own
        graph_info = {"error": repr(exc)[:200]}

    # ---- parity of the timed path: coefficients vs the oracle on the SAME (A, b, w) (rank 0, N=1) --
    coeff_err = None
    cpu_baseline = None
    x_dev = x.detach().cpu().numpy()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import linear_fit as lf
        import threadpoolctl
        # bounded CPU sample: first `sample_cfg` configurations of this very workload
        sample_cfg = max(1, min(ncfg, int(250_000 // rows_raw)))
        hs = dict(raw=raw[:sample_cfg * rows_raw].cpu().numpy(), volume=vol[:sample_cfg].cpu().numpy(),
                  energy=batch.energy[:sample_cfg].cpu().numpy(), forces=batch.forces[:3 * n * sample_cfg].cpu().numpy(),
                  stress=batch.stress[:sample_cfg].cpu().numpy().reshape(sample_cfg, 3, 3),
                  eweight=batch.eweight[:sample_cfg].cpu().numpy(), fweight=batch.fweight[:sample_cfg].cpu().numpy(),
                  vweight=batch.vweight[:sample_cfg].cpu().numpy(), type_fraction=tf[:sample_cfg].cpu().numpy(),
                  blank2j=np.ones(k))
        cpu_reference_step(hs, wl, min(sample_cfg, 50))        # warm-up (imports, BLAS threads)
        rows_s, t_asm, t_fit, x_cpu, (a_s, b_s, w_s) = cpu_reference_step(hs, wl, sample_cfg)
        cpu_baseline = {"value": rows_s / (t_asm + t_fit), "unit": "rows/s", "cores": os.cpu_count(), "kind": "port",
                        "sample": "first %d configs = %d rows of this workload: oracle row assembly %.2fs + "
                                  "sklearn Ridge %.2fs (BLAS threads %s)" %
                                  (sample_cfg, rows_s, t_asm, t_fit,
                                   [p.get("num_threads") for p in threadpoolctl.threadpool_info()])}
        # parity: device fit of exactly that sample vs the oracle's exact ridge statement
        ns = rows_s
        res_s = eng.fit(A[:ns], bvec[:ns], wvec[:ns], None, alpha=ALPHA, refine=REFINE, diagnostics=False)
        assert np.array_equal(A[:ns].cpu().numpy(), a_s), "device scatter differs from the oracle"
        mr, l2, _ = lf.coeff_rel_err(res_s.coefficients(), lf.ridge_fit_exact(a_s, b_s, w_s, ALPHA))
        coeff_err = {"max_rel_vs_exact_ridge": mr, "l2_rel_vs_exact_ridge": l2,
                     "max_rel_vs_sklearn_ridge": lf.coeff_rel_err(res_s.coefficients(), x_cpu)[0],
                     "rows": int(ns), "scatter_bit_exact": True}

    # ---- e2e: host buffers through the public API -------------------------------------------------
    e2e = None
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
                      "fit -> D2H coefficients)",
               "max_abs_diff_vs_device_resident_x": float(np.max(np.abs(xh - x_dev)))}

    if rank == 0:
        flops = (2.0 * k * k + 2.0 * k) * n_rows
        achieved = flops / (gram_ms * 1e-3) / 1e12
        gram_path = eng.gram_path(n_rows, k)
        roofline = {"bound": "tensor", "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s",
                    "frac": achieved / bf16_peak, "traffic": None,
                    "peak_kind": "dense bf16 cuBLAS, %s (MEASURED_PEAKS.json)" % peak_kind,
                    "achieved_kind": "algorithmic fp64 flops (2k^2+2k per row, full Gram convention) / Gram time",
                    "hbm_gbs_during_gram": 8.0 * (k + 2) * n_rows / (gram_ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm_peak}
        try:    # ncu-measured DRAM bytes per launch of the dominant kernel, when a capture of this shape is committed
            tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(
                "%s:%s" % (args.workload, gram_path))
            if tr:
                roofline["traffic"] = tr["bytes_per_launch"]
                roofline["traffic_note"] = "%s, %d rows x %d: %.3g B from ncu vs %.3g B algorithmic (%s)" % (
                    tr["kernel"], tr["rows"], tr["k"], tr["bytes_per_launch"], tr["algorithmic_bytes"], tr["source"])
        except (OSError, ValueError):
            pass
        if gram_path == "int8":
            n_i = -(-(k + 1) // 128)
            ops = sum(2.0 * 128 * (256 if 2 * jj + 1 < n_i else 128) * n_rows * 16
                      for i_ in range(n_i) for jj in range(i_ // 2 + 1))
            roofline.update({
                "kernel": "i8_gemm_kernel (tcgen05.mma kind::i8) + i8_colmax/i8_convert/i8_crt_kernel",
                "int8_top_s_executed_over_whole_gram": ops / (gram_ms * 1e-3) / 1e12,
                "note": "fp64 Gram recast as 16 exact int8 GEMMs modulo coprime moduli (CRT); the tcgen05 kernel "
                        "alone keeps the tensor pipe ~84 % busy (profiles/r01_i8_*.txt); the Gram time also holds "
                        "the column-maximum, residue-conversion and CRT passes"})
        else:
            roofline.update({
                "kernel": ("gram_rowsplit_kernel" if k + 1 <= 104 else
                           ("gram_dmma_kernel" if k + 1 <= 128 else "preweight_kernel + gram_tma_kernel")) +
                          " (+gram_reduce_kernel)",
                "fp64_note": "tcgen05 has no f64 kind; this kernel runs on DMMA.8x8x4 whose measured peak on this "
                             "pool is 37.1 TFLOP/s (tools/ubench/fp64_rates.cu): frac_of_fp64_peak = %.3f "
                             "(algorithmic flops count the full K x K Gram, the kernel executes the lower "
                             "triangle only)" % (achieved / 37.1)})
        line = {
            "metric": "design_matrix_rows_per_s", "value": value, "unit": "rows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "rows_per_gpu": n_rows, "k": k,
                       "configs_per_gpu": ncfg, "atoms_per_config": n, "alpha": ALPHA, "refine_rounds": REFINE,
                       "gram_path": gram_path,
                       "parallelism": "row-shard x%d, 1 all-reduce of (k+1)^2 + %d of k doubles" % (world, REFINE),
                       "l2": "inputs (A %.0f MB + raw %.0f MB per GPU) larger than the 126 MB L2; no flush" %
                             (n_rows * k * 8 / 1e6, n_rows * (kraw + 1) * 8 / 1e6)},
            "gram_tflops_algorithmic": achieved,
            "gram_ms": gram_ms,
            "phases_ms_rank0": phases_ms,
            "phase_rows_per_s_per_gpu": phase_rows_per_s,
            "roofline": roofline,
            "coeff_max_rel_err": coeff_err,
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "cuda_graph_replay": graph_info,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
