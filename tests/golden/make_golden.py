"""Generates tests/golden/*.npz.  Run HERE (the container that has /root/reference); the GPU box never reads
/root/reference.  Inputs come from the reference's bundled fixtures (data/intestinalDataSmall.RData,
data/intestinalData.RData) through the host-side filterdata restatement; expected outputs come from the CPU oracle
(oracle/oracle.c) — PARITY UNPINNED against R itself (no R in this image), see oracle.c header.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from raceid3_stemid2_package_b200 import scseq  # noqa: E402
from raceid3_stemid2_package_b200.rdata import load_dgcmatrix  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/data"


def make(name, rdata, mintotal, kmax, k_boot, bootnr):
    M, genes, cells = load_dgcmatrix(os.path.join(REF, rdata))
    sc = scseq.SCseq(M, genes_all=genes, cells_all=cells)
    sc = scseq.filterdata(sc, mintotal=mintotal)
    x = scseq.getfdata(sc, g=sc.cluster["features"])  # genes x cells, what compdist hands to dist.gen
    out = {"x": x, "n_cells": np.int64(x.shape[1]), "n_features": np.int64(x.shape[0])}
    for metric in ("pearson", "spearman", "logpearson", "euclidean"):
        D, nz = orc.dist(x, metric)
        out[f"D_{metric}"] = D
        out[f"nz_{metric}"] = np.int64(nz)
    D = out["D_pearson"]
    n = D.shape[0]
    kmax = min(kmax, n - 1)
    out["logW"] = orc.gap_sweep(D, kmax)
    out["cln_sat"] = np.int64(orc.saturation(out["logW"]))
    for k in (2, 3, 5, k_boot):
        r = orc.pam(D, k, want_sil=True)
        out[f"pam{k}_id_med"] = r["id_med"]
        out[f"pam{k}_clustering"] = r["clustering"]
        out[f"pam{k}_objective"] = r["objective"]
        out[f"pam{k}_avg_sil"] = np.float64(r["avg_sil"])
        out[f"pam{k}_n_swaps"] = np.int64(r["n_swaps"])
    samples = scseq.bootstrap_samples(n, bootnr, 17000)
    cb = orc.clusterboot(D, k_boot, samples)
    out["boot_k"] = np.int64(k_boot)
    out["boot_partition"] = cb["partition"]
    out["boot_result"] = cb["bootresult"]
    out["boot_mean"] = cb["bootmean"]
    out["boot_sample_lens"] = np.asarray([len(s) for s in samples], dtype=np.int64)
    out["boot_sample0"] = samples[0]
    med, ids = orc.medoids(D, cb["partition"])
    out["medoids"] = med
    out["refresh"] = orc.refresh(D, med)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, "cells", n, "features", x.shape[0], "cln_sat", int(out["cln_sat"]), "nz", int(out["nz_pearson"]))


if __name__ == "__main__":
    make("cfg1_intestinalDataSmall.npz", "intestinalDataSmall.RData", 3000, 30, 4, 20)
    make("cfg2_intestinalData.npz", "intestinalData.RData", 3000, 30, 7, 50)
