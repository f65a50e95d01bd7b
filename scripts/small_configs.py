"""Times configs 1-2 of BASELINE.json (the reference's bundled data, golden fixtures) end to end through the public API
and prints one JSON line per config; also checks the result against the golden partition."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import raceid3_stemid2_package_b200 as rid  # noqa: E402

ctx = rid.Context(0)
rid.set_context(ctx)
for name in ("cfg1_intestinalDataSmall", "cfg2_intestinalData"):
    g = np.load(f"tests/golden/{name}.npz")
    x = g["x"]
    G, N = x.shape
    sc = rid.SCseq(np.ones((G, N)))
    sc.cell_names = [f"c{i}" for i in range(N)]
    sc.genes = [f"g{i}" for i in range(G)]
    sc.genes_all = list(sc.genes)
    sc.counts = np.ones(N)
    sc.ndata = x - 0.1
    sc.cluster["features"] = list(sc.genes)
    best = None
    for rep in range(4):
        t0 = time.perf_counter()
        sc = rid.compdist(sc, metric="pearson")
        t1 = time.perf_counter()
        sc = rid.clustexp(sc, verbose=False)  # sat=TRUE, clustnr=30, bootnr=50, rseed=17000
        t2 = time.perf_counter()
        if rep and (best is None or t2 - t0 < best[0]):
            best = (t2 - t0, t1 - t0, t2 - t1)
    k = int(sc.cluster["kpart"].max())
    print(json.dumps({"config": name, "cells": N, "features": G, "compdist_s": best[1], "clustexp_s": best[2],
                      "total_s": best[0], "clusters": k, "cln_matches_oracle": bool(k == int(g["cln_sat"])),
                      "launches_total": ctx.launches}))
