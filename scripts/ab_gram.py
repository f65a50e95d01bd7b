"""A/B: CTA-pair (cta_group::2) int8 distance contraction against the single-CTA kernel; results must be bit-identical."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import raceid3_stemid2_package_b200 as rid
from raceid3_stemid2_package_b200.synth import synth_features_torch
ctx = rid.Context(device=0); rid.set_context(ctx)
dev = torch.device("cuda", 0)
os.environ["RID_GRAM"] = "i8"  # a fallback to the FP64 kernel would be an error, not a silent pass

def run(x, G, N, kind):
    os.environ["RID_I8_PAIR"] = "1" if kind == "pair" else "0"
    os.environ["RID_I8_PERSIST"] = "1" if kind == "persist" else "0"
    ctx.profile(["gram_i8"])
    dm = ctx.dist((x.data_ptr(), G, N), "pearson")
    p = ctx.profile_get("gram_i8"); ctx.profile(False)
    return dm, p

for (N, G) in ((1000, 200), (5000, 700), (4999, 3000)):
    x, _ = synth_features_torch(N, G, 7, 99, dev); torch.cuda.synchronize()
    a, _ = run(x, G, N, "single"); A = a.as_matrix(); a.free()
    b, _ = run(x, G, N, "pair"); B = b.as_matrix(); b.free()
    c, _ = run(x, G, N, "persist"); Cm = c.as_matrix(); c.free()
    print(N, G, "pair identical" if np.array_equal(A, B) else f"pair DIFFERENT max {np.nanmax(np.abs(A-B))}",
          "| persist (integer epilogue): max |diff|", np.nanmax(np.abs(A - Cm)), "cells differing", int((A != Cm).sum()),
          "symmetric", bool(np.array_equal(Cm, Cm.T)), flush=True)
    assert np.array_equal(A, B) and np.nanmax(np.abs(A - Cm)) <= 2.0 ** -27
N, G = 100000, 3000
x, _ = synth_features_torch(N, G, 30, 1004, dev); torch.cuda.synchronize()
rng = np.random.default_rng(1)
rows = np.sort(rng.choice(N, 600, replace=False)).astype(np.int32); cols = np.sort(rng.choice(N, 700, replace=False)).astype(np.int32)
ref = None
for kind in ("single", "persist", "pair", "persist", "single", "persist"):
    dm, p = run(x, G, N, kind)
    blk = dm.sub(rows, cols); tail = dm.sub(np.arange(N - 300, N, dtype=np.int32), np.arange(N - 300, N, dtype=np.int32))
    dm.free()
    print(kind, f"{p['ms']:.2f} ms  {p['work']/p['ms']/1e9:.0f} TOP/s", flush=True)
    if ref is None: ref = (blk, tail)
    assert max(np.abs(ref[0] - blk).max(), np.abs(ref[1] - tail).max()) <= 2.0 ** -27
print("consistent at 100k")
