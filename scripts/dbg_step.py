"""Diagnostic: per-call wall / device ms of dist -> gap_sweep -> clusterboot -> free over a few steps (RID_TRACE=1 shows
pool misses)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import raceid3_stemid2_package_b200 as rid
from raceid3_stemid2_package_b200.synth import synth_features_torch
N, G, K, B = 100000, 3000, 30, int(sys.argv[1]) if len(sys.argv) > 1 else 6
ctx = rid.Context(device=0); rid.set_context(ctx)
x, _ = synth_features_torch(N, G, K, 1004, torch.device("cuda", 0)); torch.cuda.synchronize()
samples = rid.bootstrap_samples(N, B, 17000)
for it in range(5):
    print(f"--- step {it}", file=sys.stderr, flush=True)
    w0 = time.perf_counter(); dm = ctx.dist((x.data_ptr(), G, N), "pearson"); w1 = time.perf_counter(); d = ctx.last_ms
    lw = dm.gap_sweep(30); w2 = time.perf_counter(); s_ms = ctx.last_ms
    r = dm.clusterboot(K, samples); w3 = time.perf_counter(); b = ctx.last_ms
    dm.free(); ctx.sync(); w4 = time.perf_counter()
    print(f"step {it}: dist {1e3*(w1-w0):.0f}/{d:.0f} sweep {1e3*(w2-w1):.0f}/{s_ms:.0f} boot {1e3*(w3-w2):.0f}/{b:.0f} free {1e3*(w4-w3):.1f} total {1e3*(w4-w0):.0f}", flush=True)
for it in range(3):
    dm = ctx.dist((x.data_ptr(), G, N), "pearson")
    t = []
    for j in range(3):
        w1 = time.perf_counter(); lw = dm.gap_sweep(30); t.append((time.perf_counter() - w1) * 1e3)
    dm.free()
    print("sweeps on one matrix:", " ".join(f"{v:.0f}" for v in t), flush=True)
