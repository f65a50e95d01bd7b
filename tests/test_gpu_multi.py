"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): tests/mgpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,mode", [(2, "replica"), (2, "cooperative"), (2, "replica_upper"), (2, "sweep_replica"),
                                        (8, "replica_upper"), (8, "cooperative")])
def test_row_sharded_results_identical_to_single_gpu(world, mode):
    """cooperative: RID_NO_REPLICA=1, every rank scans its row shard of every PAM problem, one all-reduce per step.
    replica (default): the copy engines pull the peers' row blocks into a full copy per rank and the bootstrap
    replicates are spread over the ranks; replica_upper makes the first pull partial at this small size
    (RID_REPLICA_MARGIN=64: only the columns the symmetric compaction needs), so the fetch-the-rest path runs too;
    sweep_replica also spreads the k values of the sweep over the replicas."""
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ)
    if mode == "cooperative":
        env["RID_NO_REPLICA"] = "1"
    if mode == "replica_upper":
        env["RID_REPLICA_MARGIN"] = "64"
    if world > 2:
        env["MGPU_CELLS"] = "6000"
    if mode == "sweep_replica":
        env["RID_SWEEP_REPLICA"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "mgpu_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "MGPU OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
