"""A/B: bootstrap replicates with the symmetric (upper triangle + mirror) compaction against whole-row compaction."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import raceid3_stemid2_package_b200 as rid
from raceid3_stemid2_package_b200.synth import synth_features_torch
N, G, K, B = 100000, 3000, 30, 50
ctx = rid.Context(device=0); rid.set_context(ctx)
x, _ = synth_features_torch(N, G, K, 1004, torch.device("cuda", 0))
torch.cuda.synchronize()
samples = rid.bootstrap_samples(N, B, 17000)
dm = ctx.dist((x.data_ptr(), G, N), "pearson")
res = {}
for mode in ("rows", "triangle", "triangle", "sync", "triangle", "triangle", "sync"):
    os.environ.pop("RID_COMPACT_FULL", None); os.environ.pop("RID_BOOT_SYNC", None)
    if mode == "rows": os.environ["RID_COMPACT_FULL"] = "1"
    if mode == "sync": os.environ["RID_BOOT_SYNC"] = "1"
    prof_all = len(res) < 2
    ctx.profile(True if prof_all else ["compact"])
    t0 = time.perf_counter(); r = dm.clusterboot(K, samples); wall = time.perf_counter() - t0
    p = ctx.profile_get("compact")
    allk = {k: ctx.profile_get(k) for k in ("scan_swap", "scan_delta", "scan_build", "compact")}
    ctx.profile(False)
    print("   ", " ".join(f"{k}={v['ms']:.1f}ms/{v['launches']}" for k, v in allk.items()), "sum", f"{sum(v['ms'] for v in allk.values()):.1f}")
    print(mode, f"boot wall {wall*1e3:.0f} ms dev {ctx.last_ms:.0f} | compact {p['launches']} launches {p['ms']:.1f} ms "
          f"alg {p['work']/p['ms']/1e6:.0f} GB/s req {p['traffic']/p['ms']/1e6:.0f} GB/s", flush=True)
    if mode in res:
        assert np.array_equal(res[mode]["partition"], r["partition"]) and np.array_equal(res[mode]["bootresult"], r["bootresult"])
    res[mode] = r
for a in ("triangle", "sync"):
    assert np.array_equal(res["rows"]["partition"], res[a]["partition"])
    assert np.array_equal(res["rows"]["bootresult"], res[a]["bootresult"])
print("identical results")
