"""Profiling driver: the fused scatter + Gram kernel on the c2 workload (run under ncu or plain)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from fitsnap_b200.engine import Engine

eng = Engine(0)
sh = bench.Shard(eng, sys.argv[1] if len(sys.argv) > 1 else "c2", 0)


def timeit(fn, n=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("scatter_gram (A stored) ms", timeit(lambda: eng.scatter_gram(sh.batch, *sh.out)))
print("scatter_gram (A not stored) ms", timeit(lambda: eng.scatter_gram(sh.batch, None, sh.b, sh.w, store_a=False)))
print("scatter ms", timeit(lambda: eng.scatter(sh.batch, *sh.out)))
print("gram ms", timeit(lambda: eng.gram(sh.A, sh.b, sh.w, None)))
