"""The gradient all-reduce inside the gradient kernels over NVLink peer memory: two ranks, one per GPU,
from the same rollout buffer and optimizer state -- the peer-memory update must leave both ranks bit-identical to each other
and (two-term sums are order independent) bit-identical to the NCCL all-reduce path; the single-block optimizer steps agree
with it to reduction-order noise.  Needs two GPUs in one box: skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("batch", [2048, 128])
def test_peer_memory_allreduce_matches_nccl_on_two_gpus(batch):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + (os.getpid() + batch) % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "p2p_check.py"), str(batch)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    out = r.stdout + r.stderr
    line = [l for l in out.splitlines() if l.startswith("world 2 batch")]
    assert r.returncode == 0 and line, out[-2000:]
    assert "ranks in sync True" in line[0]
    if batch > 256:
        assert "bitwise True" in line[0]
