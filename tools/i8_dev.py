"""Development driver of the int8 tcgen05 Gram: accuracy against an exact (integer / long double) Gram and
timing against the DMMA path.  python tools/i8_dev.py [n_rows k] ..."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from fitsnap_b200.engine import Engine


def exact_gram(A, b, w):
    aug = np.concatenate([A * w[:, None], (w * b)[:, None]], axis=1)
    L = aug.astype(np.longdouble)
    return np.asarray(L.T @ L, dtype=np.float64)


def run(n, k, eng, check=True, reps=3):
    rng = np.random.default_rng(n + k)
    dev = eng.device
    g = torch.Generator(device=dev); g.manual_seed(n * 7 + k)
    A = torch.randn((n, k), dtype=torch.float64, device=dev, generator=g)
    A *= 10.0 ** (torch.rand(k, dtype=torch.float64, device=dev, generator=g) * 3 - 3)
    b = torch.randn(n, dtype=torch.float64, device=dev, generator=g)
    w = 10.0 ** (torch.randint(-2, 3, (n,), device=dev, generator=g).double())
    out = {}
    for path in ("fp64", "int8"):
        eng.set_gram_path(path)
        G = eng.gram(A, b, w)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); G = eng.gram(A, b, w); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out[path] = (G.cpu().numpy(), min(ts))
    G64, t64 = out["fp64"]; G8, t8 = out["int8"]
    d = np.sqrt(np.abs(np.diag(G64))); d[d == 0] = 1
    scale = np.outer(d, d)
    msg = "n=%d k=%d  fp64 %.3f ms  int8 %.3f ms (x%.2f)  |G8-G64|/sqrt(GiiGjj) max %.2e" % (
        n, k, t64, t8, t64 / t8, np.max(np.abs(G8 - G64) / scale))
    if check:
        Gx = exact_gram(A.cpu().numpy(), b.cpu().numpy(), w.cpu().numpy())
        msg += "  vs exact: int8 %.2e fp64 %.2e  sym %s" % (np.max(np.abs(G8 - Gx) / scale), np.max(np.abs(G64 - Gx) / scale),
                                                  np.array_equal(G8, G8.T))
    print(msg, flush=True)


if __name__ == "__main__":
    eng = Engine(0)
    args = [int(a) for a in sys.argv[1:]]
    shapes = list(zip(args[0::2], args[1::2])) or [(5000, 40), (20000, 300), (70000, 1000)]
    for n, k in shapes:
        run(n, k, eng, check=(n * k <= 3e7))
