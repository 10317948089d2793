"""2-GPU parity (skipped with fewer than 2 devices): spawns two ranks (one per GPU, NCCL) and asserts what
tools/dist_parity.py prints -- row-sharded fits through the library's own collective (peer windows and NCCL),
coefficients against the oracle on the FULL matrix, bit-identical on both ranks -- plus the mirror solver classes
made sharded by `distributed.attach` and the CUDA-graph replay of the sharded step."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "dist_parity.py")]
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=600, env=env)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("no_peer", [False, True])
def test_two_rank_sharded_fit_parity(no_peer):
    r = _run({"FSB_NO_PEER": "1"} if no_peer else None)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("dist_parity")]
    assert len(lines) >= 8 and all(ln.rstrip().endswith("OK") for ln in lines), "\n".join(lines)
    assert any("peer_windows=%s" % (not no_peer) in ln for ln in lines), "\n".join(lines)
