"""GPU (>= 2 devices): numerical check of the data-parallel path on hardware -- the NVLink exchange kernel against NCCL,
and the head's exchanged gradients against the mean of the shard-local gradients (tests/dp_worker.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one node")
def test_data_parallel_exchange_and_gradient_mean():
    n = 2
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DP_CHECK_OK" in r.stdout, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
