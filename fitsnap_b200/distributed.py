"""Row-sharding of the design matrix across the GPUs of one box (one process per GPU).

Mirrors the reference's data distribution: every MPI rank owns the rows of a contiguous block
of the node-shared A (`ParallelTools.new_slice_a`, fitsnap3lib/parallel_tools.py:594-651), all
rows of a configuration stay on one rank (`split_within_node`, parallel_tools.py:509-511), and
the multi-node ScaLAPACK solver uses a 1-D row-block grid (lib/scalapack_solver/scalapack.py:37-54).
Here the per-rank blocks never meet in one address space: each GPU forms the Gram of its shard
and a single all-reduce (NCCL over NVLink) sums the (k+1)^2 doubles, after which every rank
solves the same k x k system.
"""
from __future__ import annotations

import numpy as np


def shard_rows(n_rows, world, rank):
    """Contiguous, balanced [lo, hi) row range of `rank` (first n % world ranks get one extra row)."""
    base, rem = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_configs_by_rows(rows_per_config, world):
    """Split configurations into `world` CONTIGUOUS groups with near-equal row counts (a
    configuration is never split).  Returns an int array of world+1 config boundaries.
    The reference deals configs round-robin by count (parallel_tools.py:466, 511); balancing by
    rows instead keeps the per-GPU Gram work equal when atom counts vary (2..257 in the examples)."""
    rows = np.asarray(rows_per_config, dtype=np.int64)
    csum = np.concatenate([[0], np.cumsum(rows)])
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        j = int(np.searchsorted(csum, target, side="left"))
        if j > 0 and abs(csum[j - 1] - target) <= abs(csum[min(j, len(rows))] - target):
            j -= 1
        bounds.append(min(max(j, bounds[-1]), len(rows)))
    bounds.append(len(rows))
    return np.asarray(bounds, dtype=np.int64)


def attach(target, group=None, engine=None):
    """Make the drop-in solver (and calculator) of `target` row-sharded over `group`.

    `target` is a `FitSnap` instance (its `.solver` / `.calculator` are used) or a solver object.  One process per
    GPU: every rank runs the reference's flow over ITS OWN configurations (the way every MPI rank of the reference
    assembles its own row block, parallel_tools.py:594-651), the rows stay on that rank's GPU, and
    `solver.perform_fit()` -- which the reference already calls on every rank (fitsnap.py:198-200) -- forms the local
    Gram, all-reduces the (k+1)^2 doubles once and solves the replicated k x k system; `solver.fit` ends up identical
    on every rank.  `error_analysis()` all-reduces the per-group sums the same way.

        torch.distributed.init_process_group("nccl")           # torchrun exports RANK / LOCAL_RANK / WORLD_SIZE
        plugin.register()
        fs = FitSnap(settings, comm=None, arglist=["--overwrite"])
        distributed.attach(fs)                                  # binds LOCAL_RANK -> GPU, sets solver.process_group
        fs.process_configs(data=my_share_of_the_configurations)
        fs.perform_fit()
    """
    import torch.distributed as dist
    from .engine import default_engine
    if group is None:
        if not dist.is_initialized():
            raise RuntimeError("attach(): torch.distributed is not initialised and no process group was given")
        group = dist.group.WORLD
    eng = engine or default_engine()
    solver = getattr(target, "solver", target)
    solver.process_group = group
    solver.engine = eng
    calc = getattr(target, "calculator", None)
    if calc is not None:
        calc._b200_engine = eng
    return solver


def bind_to_gpu_numa(local_gpu):
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off (sysfs `numa_node` of the GPU's PCI
    function), so that the pinned staging buffers it allocates afterwards are first-touched on that node and its H2D
    copies do not cross the socket interconnect.  With 8 ranks staging at once from one node the per-GPU H2D rate
    halved in round 1 (55.6 -> 27.4 GB/s).  Returns a short description, or None when the topology cannot be read
    (single-node boxes, containers without sysfs) -- never raises."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_gpu))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:            # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return None
        os.sched_setaffinity(0, use)
        return "gpu %d (%s) -> numa node %d, %d cpus" % (int(local_gpu), bus, node, len(use))
    except Exception:
        return None
