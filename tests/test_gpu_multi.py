"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): tests/mgpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,no_replica", [(2, "0"), (2, "1")])
def test_row_sharded_results_identical_to_single_gpu(world, no_replica):
    """no_replica=1 forces the cooperative schedule (every rank scans its row shard of every PAM problem, one all-reduce
    per step); the default gathers the shards and spreads bootstrap replicates / sweep k values over the ranks."""
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ)
    if no_replica == "1":
        env["RID_NO_REPLICA"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "mgpu_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "MGPU OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
