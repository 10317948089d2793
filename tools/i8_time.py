"""Time the int8 Gram alone: python tools/i8_time.py n k   (FSB_I8_CONVERTERS=0|256|512 picks the overlap mode)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fitsnap_b200.engine import Engine
n, k = int(sys.argv[1]), int(sys.argv[2])
eng = Engine(0)
g = torch.Generator(device=eng.device); g.manual_seed(1)
A = torch.randn((n, k), dtype=torch.float64, device=eng.device, generator=g)
b = torch.randn(n, dtype=torch.float64, device=eng.device, generator=g)
w = torch.rand(n, dtype=torch.float64, device=eng.device, generator=g) + 0.5
eng.set_gram_path("int8")
for _ in range(2):
    G = eng.gram(A, b, w)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); G = eng.gram(A, b, w); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
eng.set_gram_path("fp64")
ns = min(n, 200000)
G64 = eng.gram(A[:ns], b[:ns], w[:ns])
eng.set_gram_path("int8")
G8 = eng.gram(A[:ns], b[:ns], w[:ns])
err = float((G8 - G64).abs().max() / G64.abs().max())
print(json.dumps({"n": n, "k": k, "converters": os.environ.get("FSB_I8_CONVERTERS", "default"), "gram_ms_min": min(ts),
                  "gram_ms_med": sorted(ts)[2], "fp64_equiv_tflops": (2.0 * k * k + 2 * k) * n / min(ts) / 1e9,
                  "rel_diff_vs_dmma_gram_on_%d_rows" % ns: err}))
