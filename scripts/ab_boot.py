"""A/B of the bootstrap phase at cfg4: look-ahead depth (RID_BOOT_AHEAD), per-kernel brackets on/off.  Prints the
device time of the replicate phase of rid_clusterboot per variant (the sweep runs first so that c1 comes from it)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import raceid3_stemid2_package_b200 as rid
from raceid3_stemid2_package_b200.synth import synth_features_torch
N, G, K, B = 100000, 3000, 30, 50
ctx = rid.Context(device=0); rid.set_context(ctx)
x, _ = synth_features_torch(N, G, K, 1004, torch.device("cuda", 0))
torch.cuda.synchronize()
samples = rid.bootstrap_samples(N, B, 17000, packed=True)
dm = ctx.dist((x.data_ptr(), G, N), "pearson")
dm.gap_sweep(K)
ref = None
variants = [("warm", {}), ("default", {}),
            ("mb1", {"RID_COLSUM_MB": "1"}), ("default", {}), ("mb1", {"RID_COLSUM_MB": "1"}), ("imad", {"RID_COLSUM_IMAD": "1"})] + \
           [(a, b) for a, b in (os.environ.get("AB_EXTRA") and eval(os.environ["AB_EXTRA"]) or [])]
for name, env in variants:
    for k_, v_ in env.items(): os.environ[k_] = v_
    t0 = time.perf_counter(); r = dm.clusterboot(K, samples); wall = time.perf_counter() - t0
    for k_ in env: os.environ.pop(k_, None)
    st = ctx.stats()
    print(f"{name:10s} wall {wall*1e3:7.1f} dev {ctx.last_ms:7.1f} first {st['first_picks_and_replica_wait_ms']:6.1f} replicates {st['replicates_ms']:7.1f} "
          f"host_enqueue {st['replicates_host_enqueue_ms']:6.1f} redone {st['boot_redone']:.0f} ahead {st['boot_ahead']:.0f}", flush=True)
    if ref is None: ref = r
    assert np.array_equal(ref["partition"], r["partition"]) and np.array_equal(ref["bootresult"], r["bootresult"])
print("identical results")
