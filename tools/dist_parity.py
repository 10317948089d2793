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
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
