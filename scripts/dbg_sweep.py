import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import raceid3_stemid2_package_b200 as rid
from raceid3_stemid2_package_b200.synth import synth_features_torch
N, G = int(sys.argv[1]) if len(sys.argv) > 1 else 100000, 3000
ctx = rid.Context(device=0); rid.set_context(ctx)
x, _ = synth_features_torch(N, G, 30, 1004, torch.device("cuda", 0)); torch.cuda.synchronize()
dm = ctx.dist((x.data_ptr(), G, N), "pearson")
res = {}
for mode in ("full", "chain", "chain", "full"):
    os.environ.pop("RID_SWEEP_FULL", None)
    if mode == "full": os.environ["RID_SWEEP_FULL"] = "1"
    try:
        ctx.profile(True)
        t0 = time.perf_counter(); lw = dm.gap_sweep(30); wall = time.perf_counter() - t0
        pk = {k: ctx.profile_get(k) for k in ("scan_swap", "scan_delta", "scan_build", "scan_group")}
        ctx.profile(False)
        print(mode, f"wall {wall*1e3:.0f} ms dev {ctx.last_ms:.0f}", " ".join(f"{k}={v['ms']:.0f}ms/{v['launches']}" for k, v in pk.items()), flush=True)
        if mode in res: assert np.array_equal(res[mode], lw)
        res[mode] = lw
    except Exception as e:
        print(mode, "FAILED:", e, flush=True)
        ctx.profile(False)
if len(res) == 2:
    print("identical" if np.array_equal(res["full"], res["chain"]) else "DIFFERENT", res["full"][:4], res["chain"][:4])
