"""Ad-hoc per-kernel timing on the GPU box (CUDA events, warm-up, L2-flush between iterations)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fitsnap_b200.engine import Engine

def timeit(fn, iters=5, warm=2, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.fill_(1.0)
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts), sorted(ts)[len(ts) // 2]

def main():
    eng = Engine(0)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    shapes = [(1_000_000, 100), (356_536, 480), (1_000_000, 1000)]
    if len(sys.argv) > 1:
        shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
    for n, k in shapes:
        gen = torch.Generator(device="cuda").manual_seed(1)
        A = torch.randn((n, k), dtype=torch.float64, device="cuda", generator=gen)
        b = torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
        w = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) + 0.5
        out = {"n": n, "k": k}
        tmin, tmed = timeit(lambda: eng.gram(A, b, w), flush=flush)
        out["gram_ms"] = tmin; out["gram_tflops_alg"] = (2.0 * k * k + 2 * k) * n / tmin / 1e9
        out["gram_GBs"] = 8.0 * (k + 2) * n / tmin / 1e6
        gaug = eng.gram(A, b, w)
        tmin, _ = timeit(lambda: eng.factor(gaug, 1e-6), flush=flush); out["factor_ms"] = tmin
        f = eng.factor(gaug, 1e-6)
        tmin, _ = timeit(lambda: eng.solve(f, gaug[:, k], rhs_stride=k + 1), flush=flush); out["solve_ms"] = tmin
        x = eng.solve(f, gaug[:, k], rhs_stride=k + 1)
        tmin, _ = timeit(lambda: eng.residual(A, b, w, None, x), flush=flush); out["residual_ms"] = tmin
        out["residual_GBs"] = 8.0 * (k + 2) * n / tmin / 1e6
        tmin, _ = timeit(lambda: eng.predict(A, x), flush=flush); out["predict_ms"] = tmin
        out["predict_GBs"] = 8.0 * (k + 1) * n / tmin / 1e6
        tmin, _ = timeit(lambda: eng.fit(A, b, w, None, alpha=1e-6, refine=2, diagnostics=False), flush=flush)
        out["fit_ms"] = tmin; out["fit_rows_per_s"] = n / tmin * 1e3
        print(json.dumps(out), flush=True)
        del A, b, w

if __name__ == "__main__":
    main()
