"""Breakdown of rid_gap_sweep at cfg4 (per-kind CUDA-event totals, swaps) and of one clusterboot, same box."""
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
KINDS = ("scan_swap", "scan_delta", "scan_build", "scan_group", "compact")
def show(tag):
    p = {k: ctx.profile_get(k) for k in KINDS}
    print(tag, f"dev {ctx.last_ms:.1f} ms |", " ".join(f"{k}={v['ms']:.1f}ms/{v['launches']}/{v['work']/1e9:.0f}GB" for k, v in p.items()),
          "| sum", f"{sum(v['ms'] for v in p.values()):.1f}", flush=True)
for rep in range(2):
    s0 = ctx.swaps
    ctx.profile(True); lw = dm.gap_sweep(K); show(f"sweep   swaps {ctx.swaps - s0:4d}"); ctx.profile(False)
    t0 = time.perf_counter(); lw = dm.gap_sweep(K); print(f"sweep unprofiled dev {ctx.last_ms:.1f} wall {(time.perf_counter()-t0)*1e3:.1f}")
    os.environ["RID_SWEEP_BUILD_FIRST"] = "1"
    t0 = time.perf_counter(); lw2 = dm.gap_sweep(K); print(f"sweep BUILD-first unprofiled dev {ctx.last_ms:.1f} wall {(time.perf_counter()-t0)*1e3:.1f}", np.array_equal(lw, lw2))
    os.environ.pop("RID_SWEEP_BUILD_FIRST")
    s0 = ctx.swaps
    ctx.profile(True); r = dm.clusterboot(K, samples); show(f"boot    swaps {ctx.swaps - s0:4d}"); ctx.profile(False)
    r = dm.clusterboot(K, samples); print(f"boot unprofiled dev {ctx.last_ms:.1f}", ctx.stats())
