"""cmdscale(d, k = 2) at cfg4 size: iterations, wall time, eigenvalues."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import raceid3_stemid2_package_b200 as rid
from raceid3_stemid2_package_b200.synth import synth_features_torch
N, G = int(sys.argv[1]) if len(sys.argv) > 1 else 100000, 3000
ctx = rid.Context(device=0); rid.set_context(ctx)
x, _ = synth_features_torch(N, G, 30, 1004, torch.device("cuda", 0)); torch.cuda.synchronize()
dm = ctx.dist((x.data_ptr(), G, N), "pearson")
for tol in (1e-6, 1e-9):
    t0 = time.perf_counter(); pts, ev, it = dm.cmdscale(2, tol=tol, with_eig=True); wall = time.perf_counter() - t0
    print(f"N={N} tol={tol:g}: {it} iterations, wall {wall*1e3:.0f} ms ({wall*1e3/it:.1f} ms/iteration), eig {ev}, points range {pts.min(0)} {pts.max(0)}", flush=True)
