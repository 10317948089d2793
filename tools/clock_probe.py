"""Sustained-load probe: run one kernel back to back for ~2 s while sampling nvidia-smi clocks,
power and throttle reasons every 20 ms.  Answers "what SM clock does this kernel actually run at?"
    python tools/clock_probe.py gram 1000000x100
    python tools/clock_probe.py residual 1000000x100
"""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fitsnap_b200.engine import Engine

Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu"


def main():
    what, shape = sys.argv[1], sys.argv[2]
    n, k = (int(v) for v in shape.split("x"))
    secs = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
    eng = Engine(0)
    gen = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn((n, k), dtype=torch.float64, device="cuda", generator=gen)
    b = torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
    w = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    x = torch.randn(k, dtype=torch.float64, device="cuda", generator=gen)
    fn = {"gram": lambda: eng.gram(A, b, w), "residual": lambda: eng.residual(A, b, w, None, x),
          "predict": lambda: eng.predict(A, x)}[what]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); fn(); e.record(); torch.cuda.synchronize()
    one = s.elapsed_time(e)
    iters = max(10, int(secs * 1e3 / one))
    p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + Q, "--format=csv,noheader,nounits", "-lms", "20"],
                         stdout=subprocess.PIPE, text=True)
    time.sleep(0.2)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    time.sleep(0.05)
    p.terminate()
    rows = [[v.strip() for v in l.split(",")] for l in p.communicate()[0].strip().splitlines()]
    rows = [r for r in rows if len(r) >= 7]
    clk = np.array([float(r[0]) for r in rows]); pw = np.array([float(r[2]) for r in rows])
    busy = pw > (pw.min() + 0.5 * (pw.max() - pw.min()))
    out = {"what": what, "n": n, "k": k, "iters": iters, "ms_first": one, "ms_sustained": ms,
           "sm_mhz_median_under_load": float(np.median(clk[busy])) if busy.any() else None,
           "sm_mhz_min": float(clk.min()), "sm_mhz_max": float(clk.max()), "power_w_max": float(pw.max()),
           "power_w_median_under_load": float(np.median(pw[busy])) if busy.any() else None,
           "sw_power_cap_active_samples": sum(r[3].lower().startswith("active") for r in rows),
           "hw_slowdown_samples": sum(r[4].lower().startswith("active") for r in rows),
           "samples": len(rows), "temp_max": max(float(r[6]) for r in rows)}
    if what == "gram":
        out["tflops_alg_sustained"] = (2.0 * k * k + 2 * k) * n / ms / 1e9
    else:
        out["GBs_sustained"] = 8.0 * (k + 2) * n / ms / 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
