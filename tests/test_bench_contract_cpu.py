"""CPU: the reference arm of bench.py (the oracle port timed on the host cores) prints ONE JSON line with the
keys the driver reads; the CUDA arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, **kw):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          cwd=ROOT, timeout=600, **kw)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-sample-configs", "300"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "design_matrix_rows_per_s" and d["unit"] == "rows/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["config"]["workload"].startswith("c2")
    # the reference arm reports the SAME config object the CUDA arm prints for this workload and GPU count
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.make_config("c2", 1)
    assert d["rows_per_step"] == 300 * 100 and "diag(blank2J)" in d["cpu_baseline"]["sample"]


def test_reference_arm_runs_on_rank_zero_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cuda_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "1", "--no-e2e", "--no-cpu-baseline"])
    assert r.returncode != 0 and r.stdout.strip() == ""
