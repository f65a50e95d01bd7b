"""Synthetic UMI-count feature matrices of the shapes BASELINE.json names.

C true clusters; gene base rates ~ LogNormal(-8.5, 1.5) normalised to sum 1; per cluster 10 % of the genes get a
fold change 2^N(0,1); library size ~ LogNormal(log 8000, 0.35) clipped to >= 3000; counts ~ Poisson(L * p).
The matrix is then transformed exactly like the reference does before dist.gen (R/RaceID.R:129,727):
x = c / colSums(c) * min(colSums(c)) + 0.1.  Generated on whatever device torch is given (plumbing), cell-major
(N x G, each cell's G values contiguous) — the layout rid_dist() takes.
"""
from __future__ import annotations

import numpy as np


def synth_counts_numpy(n_cells: int, n_genes: int, n_clusters: int, seed: int, structure: str = "clusters"):
    """structure "clusters": every cell follows the profile of its cluster.  "ring": every cell is a random mixture of
    the profiles of cluster c and c+1 (a closed differentiation trajectory): a continuum without gaps, on which PAM's
    BUILD is not already the local optimum and SWAP really has to work (7-9 exchanges per run at k = 30)."""
    rng = np.random.default_rng(seed)
    p = rng.lognormal(-8.5, 1.5, size=n_genes)
    p /= p.sum()
    prof = np.tile(p, (n_clusters, 1))
    for c in range(n_clusters):
        sel = rng.random(n_genes) < 0.10
        prof[c, sel] *= 2.0 ** rng.normal(0.0, 1.0, size=int(sel.sum()))
        prof[c] /= prof[c].sum()
    lab = rng.integers(0, n_clusters, size=n_cells)
    L = np.maximum(rng.lognormal(np.log(8000.0), 0.35, size=n_cells), 3000.0)
    rate = prof[lab]
    if structure == "ring":
        a = rng.random(n_cells)[:, None]
        rate = a * rate + (1.0 - a) * prof[(lab + 1) % n_clusters]
    counts = rng.poisson(L[:, None] * rate)  # N x G
    return counts.astype(np.float64), lab


def transform_like_getfdata(counts_ng: np.ndarray) -> np.ndarray:
    tot = counts_ng.sum(axis=1, keepdims=True)
    return counts_ng / tot * tot.min() + 0.1


def synth_features_numpy(n_cells, n_genes, n_clusters, seed, structure: str = "clusters"):
    """N x G float64 feature matrix (cell-major) + true labels; zero-variance cells are impossible here because
    every cell has >= 3000 counts spread over the genes."""
    c, lab = synth_counts_numpy(n_cells, n_genes, n_clusters, seed, structure)
    return transform_like_getfdata(c), lab


def synth_features_torch(n_cells, n_genes, n_clusters, seed, device, structure: str = "clusters", want_counts: bool = False):
    """Same model, generated with torch on `device`; returns (x [N,G] float64 tensor, labels), or with
    want_counts=True (counts [N,G] float64 tensor, labels) — the raw UMI counts before the getfdata transform."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    p = torch.exp(torch.randn(n_genes, generator=g, device=device, dtype=torch.float64) * 1.5 - 8.5)
    p = p / p.sum()
    prof = p.repeat(n_clusters, 1)
    sel = torch.rand(n_clusters, n_genes, generator=g, device=device) < 0.10
    fc = torch.exp2(torch.randn(n_clusters, n_genes, generator=g, device=device, dtype=torch.float64))
    prof = torch.where(sel, prof * fc, prof)
    prof = prof / prof.sum(dim=1, keepdim=True)
    lab = torch.randint(0, n_clusters, (n_cells,), generator=g, device=device)
    L = torch.exp(torch.randn(n_cells, generator=g, device=device, dtype=torch.float64) * 0.35 + np.log(8000.0))
    L = torch.clamp(L, min=3000.0)
    x = torch.empty((n_cells, n_genes), dtype=torch.float64, device=device)
    step = 8192
    mix = torch.rand(n_cells, generator=g, device=device, dtype=torch.float64) if structure == "ring" else None
    for s in range(0, n_cells, step):
        e = min(s + step, n_cells)
        pr = prof[lab[s:e]]
        if mix is not None:
            a = mix[s:e, None]
            pr = a * pr + (1.0 - a) * prof[(lab[s:e] + 1) % n_clusters]
        rate = (L[s:e, None] * pr).to(torch.float32)
        x[s:e] = torch.poisson(rate, generator=g).to(torch.float64)
    if want_counts:
        return x, lab
    tot = x.sum(dim=1, keepdim=True)
    x = x / tot * tot.min() + 0.1
    return x, lab
