"""Row-sharded fit parity under torchrun (one process per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_parity.py
Every rank takes its contiguous row shard of the seeded synthetic systems of tests/synth.py, the
Gram is all-reduced, the solve is replicated; rank 0 checks the coefficients against the oracle on
the FULL matrix and that all ranks hold bit-identical coefficients."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from fitsnap_b200.distributed import shard_rows
from fitsnap_b200.engine import Engine
from oracle import linear_fit as lf
from tests.synth import SOLVE_CASES, synth_system


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local)
    ok = True
    cases = [("well", 0.0, "auto"), ("ill", 0.0, "auto"), ("wide", 1e-6, "auto"), ("zerocol", 0.0, "auto"),
             ("ill", 0.0, "int8"), ("wide", 1e-6, "int8")]     # int8: the tcgen05 exact-integer Gram per shard
    for case, alpha, path in cases:
        eng.set_gram_path(path)
        a, b, w, t = synth_system(**SOLVE_CASES[case])
        lo, hi = shard_rows(a.shape[0], world, rank)
        A, B, W = eng.to_device(a[lo:hi]), eng.to_device(b[lo:hi]), eng.to_device(w[lo:hi])
        T = eng.to_device(t[lo:hi].astype(np.uint8), dtype=torch.uint8)
        res = eng.fit(A, B, W, T, alpha=alpha, refine=3, group=dist.group.WORLD)
        x = res.x.clone()
        gathered = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(gathered, x)
        if rank == 0:
            same = all(torch.equal(gathered[0], g) for g in gathered)
            ref = lf.svd_fit(a, b, w, t) if alpha == 0.0 else lf.ridge_fit_exact(a, b, w, alpha, t)
            mr, l2, _ = lf.coeff_rel_err(x.cpu().numpy(), ref)
            good = same and mr < 1e-10
            ok = ok and good
            print("dist_parity world=%d case=%s alpha=%g gram=%s: max_rel=%.2e l2=%.2e identical_on_all_ranks=%s %s"
                  % (world, case, alpha, path, mr, l2, same, "OK" if good else "FAIL"), flush=True)
    # ---- the mirror solver classes made sharded by distributed.attach (perform_fit + error analysis) ----
    from types import SimpleNamespace
    from fitsnap_b200 import distributed
    from fitsnap_b200.solvers import RIDGE, SVD
    a, b, w, t = synth_system(**SOLVE_CASES["ill"])
    lo, hi = shard_rows(a.shape[0], world, rank)
    rt = np.where(np.arange(a.shape[0]) % 5 == 0, "Energy", "Force")
    for cls, alpha in ((SVD, 0.0), (RIDGE, 1e-6)):
        pt = SimpleNamespace(_rank=rank, shared_arrays={},
                             fitsnap_dict={"Testing": [bool(v) for v in t[lo:hi]], "Groups": ["g%d" % (i % 3) for i in range(lo, hi)],
                                           "Row_Type": list(rt[lo:hi])})
        cfg = SimpleNamespace(sections={"RIDGE": SimpleNamespace(alpha=alpha, local_solver=0)})
        s = cls(cls.__name__, pt, cfg)
        distributed.attach(s, engine=eng)
        s.perform_fit(a=a[lo:hi], b=b[lo:hi], w=w[lo:hi], fs_dict=pt.fitsnap_dict)
        errs = s.error_analysis_device(a=a[lo:hi], b=b[lo:hi], w=w[lo:hi], fs_dict=pt.fitsnap_dict)
        xs = [None] * world
        dist.all_gather_object(xs, s.fit)
        ns = [None] * world
        dist.all_gather_object(ns, int(errs["ncount"].sum()))
        if rank == 0:
            ref = lf.svd_fit(a, b, w, t) if alpha == 0.0 else lf.ridge_fit_exact(a, b, w, alpha, t)
            mr = lf.coeff_rel_err(s.fit, ref)[0]
            same = all(np.array_equal(xs[0], v) for v in xs)
            # every row is counted once in the '*ALL' block and once in its group block, weighted and unweighted
            rows_ok = all(v == ns[0] for v in ns) and int(errs.loc["*ALL"].loc["Unweighted"]["ncount"].sum()) == a.shape[0]
            good = same and mr < 1e-10 and rows_ok
            ok = ok and good
            print("dist_parity world=%d solver=%s attach: max_rel=%.2e identical_on_all_ranks=%s error_rows_ok=%s %s"
                  % (world, cls.__name__, mr, same, rows_ok, "OK" if good else "FAIL"), flush=True)

    # ---- CUDA-graph replay of the sharded device-resident step (peer-window collective inside the graph) ----
    comm = eng.comm_for(dist.group.WORLD)
    info = comm.info()
    if info["peer_windows"]:
        from fitsnap_b200.pipeline import LinearFitPipeline
        eng.set_gram_path("auto")
        rng = np.random.default_rng(100 + rank)
        nt, nc, ncfg, n = 2, 14, 300, 9
        kraw, k = nt * nc, nt * nc + nt
        raw = rng.standard_normal((ncfg * (7 + 3 * n), kraw + 1))
        pipe = LinearFitPipeline(nt, nc, False, np.ones(k), alpha=1e-8, refine=2, group=dist.group.WORLD, engine=eng)
        st = rng.standard_normal((ncfg, 3, 3))
        batch = pipe.pack(raw, np.full(ncfg, n, dtype=np.int32), rng.uniform(50, 500, ncfg), rng.standard_normal(ncfg),
                          rng.standard_normal(ncfg * n * 3), 0.5 * (st + st.transpose(0, 2, 1)), np.ones(ncfg), np.ones(ncfg),
                          np.full(ncfg, 1e-4), rng.dirichlet(np.ones(nt), ncfg))
        eager = pipe.fit_batch(batch).x.clone()
        cap = pipe.capture(batch)
        for _ in range(3):
            xg = cap.replay().x.clone()
        torch.cuda.synchronize()
        xs = [torch.empty_like(xg) for _ in range(world)]
        dist.all_gather(xs, xg)
        if rank == 0:
            good = bool(torch.equal(eager, xg)) and all(torch.equal(xs[0], v) for v in xs)
            ok = ok and good
            print("dist_parity world=%d cuda_graph_of_sharded_step: replay == eager and identical on all ranks: %s %s"
                  % (world, good, "OK" if good else "FAIL"), flush=True)
    if rank == 0:
        print("dist_parity collective: peer_windows=%s peer_calls=%d nccl_calls=%d OK"
              % (info["peer_windows"], comm.info()["peer_calls"], comm.info()["nccl_calls"]), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
