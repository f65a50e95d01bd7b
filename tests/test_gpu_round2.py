"""GPU parity tests added in round 2: value-level checks at benchmarked sizes (VERDICT r1 #1c), the numeric corners
(euclidean range, adopted matrices, the 31-bit grid's top value), the `samp` arm end to end, the findoutliers distance
consumers, cancellation.  All through the C ABI, against the CPU oracle / FP64 numpy on the same seeded inputs."""
import threading

import numpy as np
import pytest

from conftest import ari

pytestmark = pytest.mark.gpu

DIST_TOL = 1e-6


def _zrows(x_ng, metric):
    """FP64 numpy restatement of cor()'s operands (bench.py's transform_rows_fp64; pinned against the oracle on CPU)."""
    import bench

    return bench.transform_rows_fp64(x_ng, metric)


def _rows_ref(x_ng, metric, ridx):
    Z = _zrows(x_ng, metric)
    if metric == "euclidean":
        return np.stack([np.sqrt(((Z - Z[r]) ** 2).sum(1)) for r in ridx])
    ref = 1.0 - np.clip(Z[ridx] @ Z.T, -1.0, 1.0)
    ref[np.arange(len(ridx)), ridx] = 0.0
    return ref


def test_persistent_gemm_many_tiles_per_cta(ctx, monkeypatch):
    """The persistent int8 tcgen05 kernel at a shape where every CTA walks >= 20 tiles with 24 K-blocks each
    (N = 9000, G = 3000: 3 337 tiles of 128 x 96 over 148 CTAs; cfg4 runs ~2 760 per CTA with the same 24 K-blocks),
    ragged in N and G.  Sampled rows against FP64 numpy, against the FP64 DMMA kernel, and the mirror property."""
    from raceid3_stemid2_package_b200.synth import synth_features_numpy

    N, G = 9000, 3000
    x, _ = synth_features_numpy(N, G, 12, 77)
    x[4321] = x[17]
    monkeypatch.setenv("RID_GRAM", "i8")
    dm = ctx.dist_cellmajor(x, "pearson")
    rng = np.random.default_rng(3)
    ridx = np.sort(rng.choice(N, 96, replace=False)).astype(np.int32)
    ridx[:4] = [0, 17, 127, 128]
    ridx[-3:] = [4321, 8959, 8999]
    ridx = np.unique(ridx).astype(np.int32)
    got = dm.sub(rows=ridx)
    ref = _rows_ref(x, "pearson", ridx)
    assert np.max(np.abs(got - ref)) <= DIST_TOL
    assert np.array_equal(got, dm.sub(cols=ridx).T)  # the mirror tiles
    i17 = int(np.nonzero(ridx == 17)[0][0])
    assert got[i17, 4321] <= 4e-7
    dm.free()
    monkeypatch.setenv("RID_GRAM", "f64")
    dm = ctx.dist_cellmajor(x, "pearson")
    got64 = dm.sub(rows=ridx)
    dm.free()
    assert np.max(np.abs(got - got64)) <= 4e-7


def test_cfg3_size_spearman_rows_and_subset_pam(ctx, oracle):
    """BASELINE configs[2] at full size (20 000 cells x 2 000 genes, spearman, cln = 20): sampled rows of D against FP64
    numpy (scipy average-tie ranks), a 3 000-cell subset PAM identical to the oracle, the full-size PAM's properties."""
    from raceid3_stemid2_package_b200.synth import synth_features_numpy

    N, G, k = 20000, 2000, 20
    x, truth = synth_features_numpy(N, G, k, 1003)
    dm = ctx.dist_cellmajor(x, "spearman")
    rng = np.random.default_rng(8)
    ridx = np.sort(rng.choice(N, 48, replace=False)).astype(np.int32)
    got = dm.sub(rows=ridx)
    assert np.max(np.abs(got - _rows_ref(x, "spearman", ridx))) <= DIST_TOL
    S = rng.permutation(N)[:3000].astype(np.int32)
    r = dm.pam(k, subset=S)
    Dsub = dm.sub(S, S)
    ro = oracle.pam(Dsub, k)
    assert np.array_equal(r["id_med"], ro["id_med"]) and np.array_equal(r["clustering"], ro["clustering"])
    assert r["n_swaps"] == ro["n_swaps"] and np.allclose(r["objective"], ro["objective"], atol=1e-13)
    rf = dm.pam(k)
    cols = dm.sub(cols=rf["id_med"].astype(np.int32))
    assert np.array_equal(np.argmin(cols, axis=1) + 1, rf["clustering"])  # every cell with its nearest medoid
    assert abs(cols.min(axis=1).mean() - rf["objective"][1]) < 1e-12
    assert rf["objective"][1] <= rf["objective"][0] + 1e-15
    assert ari(rf["clustering"], truth) > 0.5  # (spearman on sparse counts: 0.78 at this size)
    dm.free()


def test_euclidean_large_range_keeps_1e6(ctx, oracle):
    """getfdata values scale with min(counts): with distances beyond 4096 the 31-bit grid alone is coarser than the 1e-6
    budget (step > 2e-6).  Exported distances are then recomputed from the retained rows in FP64 (1e-6 holds for any
    range); PAM on the grid still matches the oracle run on the exact distances (ARI = 1, same medoids)."""
    from raceid3_stemid2_package_b200.synth import synth_features_numpy

    x, _ = synth_features_numpy(600, 120, 5, 21)
    x = np.ascontiguousarray(x.T) * 4000.0  # genes x cells
    x[:, 33] = x[:, 7]
    Do, _ = oracle.dist(x, "euclidean")
    assert Do.max() > 8192.0
    dm = ctx.dist(x, "euclidean")
    assert dm.step > 2e-6
    Dg = dm.as_matrix()
    assert np.max(np.abs(Dg - Do)) <= DIST_TOL
    assert Dg[33, 7] == 0.0 and np.array_equal(Dg, Dg.T)
    blk = dm.sub(rows=np.array([5, 33, 599], dtype=np.int32), cols=np.array([7, 0, 123, 598], dtype=np.int32))
    assert np.max(np.abs(blk - Do[np.ix_([5, 33, 599], [7, 0, 123, 598])])) <= DIST_TOL
    r = dm.pam(5)
    ro = oracle.pam(Do, 5)
    assert np.array_equal(r["id_med"], ro["id_med"]) and ari(r["clustering"], ro["clustering"]) == 1.0
    dm.free()


def test_adopted_matrix_small_range_and_grid_top(ctx, oracle):
    """rid_dmat_from_host picks its step from the range (a fixed 2^-27 would reduce a matrix with maximum 1e-4 to a few
    thousand levels and invent ties); a maximum that rounds up to 2^31 steps is clamped instead of wrapping."""
    rng = np.random.default_rng(2)
    P = rng.normal(size=(260, 3))
    D = np.sqrt(((P[:, None, :] - P[None, :, :]) ** 2).sum(-1))
    small = D * (1e-4 / D.max())
    dm = ctx.from_matrix(small)
    assert dm.step < 1e-13
    Dg = dm.as_matrix()
    assert np.max(np.abs(Dg - small)) <= dm.step
    assert len(np.unique(Dg)) > 30000  # ~33 670 distinct pairs survive the quantisation
    for k in (3, 7):
        r, ro = dm.pam(k), oracle.pam(small, k)  # against the oracle on the REAL values: no invented ties
        assert np.array_equal(r["id_med"], ro["id_med"]) and np.array_equal(r["clustering"], ro["clustering"])
    dm.free()
    top = D * ((4096.0 - 1e-10) / D.max())  # frexp gives e = 12; rint(max * 2^19) == 2^31 exactly
    dm = ctx.from_matrix(top)
    Dg = dm.as_matrix()
    assert Dg.max() < 4096.0 and np.max(np.abs(Dg - top)) <= dm.step
    r, ro = dm.pam(4), oracle.pam(Dg, 4)
    assert np.array_equal(r["id_med"], ro["id_med"]) and np.array_equal(r["clustering"], ro["clustering"])
    assert np.allclose(r["objective"], ro["objective"], rtol=1e-13)
    lw = dm.gap_sweep(5)
    assert np.max(np.abs(lw - oracle.gap_sweep(Dg, 5))) < 1e-12
    dm.free()


def test_clustexp_samp_arm_end_to_end(ctx, oracle):
    """clustexp(samp=...) (R/RaceID_utils.R:112,133): the sweep runs on `sample(N, samp)` drawn after set.seed(rseed),
    clusterboot uses bootmethod="subset" (resamples without replacement) and `jaccard` is `subsetmean`."""
    import raceid3_stemid2_package_b200 as rid
    from raceid3_stemid2_package_b200.rrng import RRng
    from raceid3_stemid2_package_b200.synth import synth_features_numpy

    rid.set_context(ctx)
    N, G = 900, 150
    x, _ = synth_features_numpy(N, G, 5, 4)
    sc = rid.SCseq(np.ones((G, N)))
    sc.cell_names = [f"c{i}" for i in range(N)]
    sc.genes = sc.genes_all = [f"g{i}" for i in range(G)]
    sc.counts = np.ones(N)
    sc.ndata = x.T - 0.1
    sc.cluster["features"] = list(sc.genes)
    sc = rid.compdist(sc, metric="pearson")
    samp, K, B = 300, 9, 5
    sc = rid.clustexp(sc, samp=samp, clustnr=K, bootnr=B, rseed=4711, verbose=False)
    Dg = sc.distances.as_matrix()
    sub = (RRng(4711).sample_int(N, samp, replace=False) - 1).astype(np.int32)
    lwo = oracle.gap_sweep(Dg, K, subset=sub)
    assert np.max(np.abs(sc.cluster["gap"]["Tab"][:, 0] - lwo)) < 1e-12
    assert sc.cluster["gap"]["n"] == samp
    cln = oracle.saturation(lwo)
    smp = rid.bootstrap_samples(N, B, 4711, samp=samp)
    assert all(len(s) == samp and len(np.unique(s)) == samp for s in smp)
    cb = oracle.clusterboot(Dg, cln, smp)
    assert np.array_equal(sc.cluster["kpart"], cb["partition"])
    assert np.allclose(sc.cluster["jaccard"], cb["bootmean"], atol=1e-14)  # subsetmean
    clb = sc.cluster["clb"]
    assert clb["bootmethod"] == "subset" and np.array_equal(clb["subsetresult"], cb["bootresult"]) and clb["bootmean"] is None
    # an explicit cln with samp: the sweep still runs (for the plot) and cln is kept
    sc = rid.clustexp(sc, samp=samp, cln=3, clustnr=K, bootnr=2, rseed=4711, verbose=False)
    assert int(sc.cluster["kpart"].max()) == 3 and len(sc.cluster["jaccard"]) == 3
    sc.distances.free()
    rid.set_context(None)


def _findoutliers_merge_literal(D, names, kpart, out, rseed, q):
    """findoutliers' "cluster outliers" block (R/RaceID.R:983-1031) on a plain full matrix, line by line."""
    from raceid3_stemid2_package_b200.rrng import RRng

    pos = {c: i for i, c in enumerate(names)}
    rng = RRng(rseed)
    clp = []
    cpart = kpart.copy()
    for i in range(1, cpart.max() + 1):
        cells = [names[j] for j in np.nonzero(cpart == i)[0]]
        tcol = [cells[j - 1] for j in rng.sample_int(len(cells), min(500, len(cells)), replace=False)]
        t = [c for c in tcol if c not in out]
        if len(t) > 1:
            ii = [pos[c] for c in t]
            clp.extend(D[np.ix_(ii, ii)].reshape(-1).tolist())
    clp = np.asarray(clp)
    clp = clp[clp > 0]
    cadd = []
    if len(out) == 1:
        cadd = [list(out)]
    elif len(out) > 1:
        n = list(out)
        ii = [pos[c] for c in n]
        m = D[np.ix_(ii, ii)]
        thr = np.quantile(clp, q)
        for _ in range(len(out)):
            if len(n) > 1:
                key = [min(m[r, c] for c in range(len(n)) if c != r) for r in range(len(n))]
                o = sorted(range(len(n)), key=lambda r: (key[r], r))
                m = m[np.ix_(o, o)]
                n = [n[r] for r in o]
                f = [(m[r, 0] < thr) or (m[r, 0] == clp.min()) for r in range(len(n))]
                fi = [r for r in range(len(n)) if f[r]]
                ind = [0]
                if len(fi) > 1:
                    ind = [a for a, r in enumerate(fi) if sum(m[r, c] <= thr for c in fi) > len(fi) / 2]
                grp = [n[fi[a]] for a in ind]
                cadd.append(grp)
                g = [c not in grp for c in n]
                n = [c for c, kk in zip(n, g) if kk]
                m = m[np.ix_(g, g)]
                if not any(g):
                    break
            elif len(n) == 1:
                cadd.append(n)
                break
    for grp in cadd:
        new = cpart.max() + 1
        for c in grp:
            cpart[pos[c]] = new
    return clp, cadd, cpart


def test_findoutliers_distance_consumers_by_name(ctx, oracle):
    """the distance reads of findoutliers — `D[tcol, tcol]` on per-cluster samples of <= 500 cells
    (R/RaceID.R:990-993) and `D[out, out]` (:1001-1003), indexed by cell NAME — served from the device-resident matrix,
    then the final refresh (:1034-1060).  Checked against the same block on a plain host matrix."""
    import raceid3_stemid2_package_b200 as rid
    from raceid3_stemid2_package_b200.synth import synth_features_numpy

    rid.set_context(ctx)
    N, G = 1600, 120
    x, truth = synth_features_numpy(N, G, 3, 9)  # three big clusters: two of them exceed the 500-cell sample cap
    sc = rid.SCseq(np.ones((G, N)))
    rng = np.random.default_rng(1)
    sc.cell_names = [f"cell_{j}" for j in rng.permutation(N)]  # names unrelated to positions
    sc.genes = sc.genes_all = [f"g{i}" for i in range(G)]
    sc.counts = np.ones(N)
    sc.ndata = x.T - 0.1
    sc.cluster["features"] = list(sc.genes)
    sc = rid.compdist(sc, metric="pearson")
    sc = rid.clustexp(sc, sat=False, cln=3, bootnr=1, rseed=17000, verbose=False)
    kpart = np.asarray(sc.cluster["kpart"]).copy()
    assert np.bincount(kpart)[1:].max() > 500
    D = sc.distances.as_matrix()
    # pretend outliers: one cell of cluster 2 with its six nearest cells (they merge) and a far-away cell of cluster 3
    c2 = np.nonzero(kpart == 2)[0]
    seed_cell = c2[0]
    near = np.argsort(D[seed_cell])[1:7]
    out = [sc.cell_names[j] for j in [seed_cell, *near, np.nonzero(kpart == 3)[0][5]]]
    clp, cadd, cpart = _findoutliers_merge_literal(D, sc.cell_names, kpart, out, 17000, 0.95)
    got_clp = rid.outlier_distance_sample(sc, out)
    assert np.array_equal(got_clp, clp)
    sc = rid.findoutliers_merge(sc, out, outdistquant=0.95)
    assert sc.out["cadd"] == cadd and len(cadd) >= 1
    medo, _ = oracle.medoids(D, cpart.astype(np.int32))
    ref = oracle.refresh(D, medo)
    assert np.array_equal(sc.cpart, ref)
    medo2, _ = oracle.medoids(D, ref)
    assert sc.medoids == [sc.cell_names[j] for j in medo2]
    with pytest.raises(ValueError):
        rid.findoutliers_merge(sc, out, outdistquant=1.5)
    with pytest.raises(ValueError):
        rid.findoutliers_merge(sc, ["no_such_cell"])
    sc.distances.free()
    rid.set_context(None)


def test_cancellation_poll_and_flag(ctx, oracle):
    """long calls are interruptible.  The poll callback (what the R stub wires to R_CheckUserInterrupt) and
    the asynchronous flag both end the running call with RID_ERR_INTERRUPTED; the context and the matrix stay usable
    and the next call gives the oracle's answer."""
    from raceid3_stemid2_package_b200 import lib
    from raceid3_stemid2_package_b200.scseq import bootstrap_samples
    from raceid3_stemid2_package_b200.synth import synth_features_numpy

    N = 2500
    x, _ = synth_features_numpy(N, 200, 6, 5)
    dm = ctx.dist_cellmajor(x, "pearson")
    samples = bootstrap_samples(N, 40, 17000)
    calls = {"n": 0}

    def poll():
        calls["n"] += 1
        return calls["n"] >= 2

    ctx.set_interrupt_poll(poll)
    with pytest.raises(lib.RaceIDError) as ei:
        dm.clusterboot(6, samples)
    assert ei.value.code == lib.INTERRUPTED and calls["n"] >= 2
    calls["n"] = 1  # the next poll cancels
    with pytest.raises(lib.RaceIDError) as ei:
        dm.gap_sweep(12)
    assert ei.value.code == lib.INTERRUPTED
    ctx.set_interrupt_poll(None)
    # the flag, raised from another thread while the call runs (or just before it: it is sticky until seen)
    t = threading.Timer(0.005, ctx.request_abort)
    t.start()
    try:
        dm.clusterboot(6, samples)
        interrupted = False
    except lib.RaceIDError as e:
        interrupted = e.code == lib.INTERRUPTED
    t.join()
    ctx.clear_abort()
    assert interrupted in (True, False)
    # nothing is left behind: same answers as the oracle afterwards
    Dg = dm.as_matrix()
    r = dm.clusterboot(6, samples[:3])
    ro = oracle.clusterboot(Dg, 6, samples[:3])
    assert np.array_equal(r["partition"], ro["partition"]) and np.array_equal(r["bootresult"], ro["bootresult"])
    lw = dm.gap_sweep(6)
    assert np.max(np.abs(lw - oracle.gap_sweep(Dg, 6))) < 1e-12
    dm.free()


def test_sparse_scseq_compdist_equals_dense(ctx):
    """An SCseq built from a sparse count matrix (upstream: dgCMatrix) runs compdist through rid_dist_csc: bit-identical
    to the dense getfdata -> rid_dist route when both see the same column totals."""
    import raceid3_stemid2_package_b200 as rid
    from scipy.sparse import csc_matrix

    rid.set_context(ctx)
    rng = np.random.default_rng(15)
    counts = rng.poisson(0.5, size=(300, 220)).astype(np.float64)  # integer counts: column sums are exact either way
    counts[:, counts.sum(0) == 0] = 1.0
    a = rid.filterdata(rid.SCseq(counts), mintotal=1, minexpr=1, minnumber=3)
    b = rid.filterdata(rid.SCseq(csc_matrix(counts)), mintotal=1, minexpr=1, minnumber=3)
    assert a.genes == b.genes and a.cluster["features"] == b.cluster["features"]
    assert np.array_equal(rid.getfdata(a), rid.getfdata(b))
    for metric in ("pearson", "logpearson"):
        a = rid.compdist(a, metric=metric)
        b = rid.compdist(b, metric=metric)
        assert np.array_equal(a.distances.as_matrix(), b.distances.as_matrix())
    # the state filterdata leaves behind when nothing is filtered (what bench.py's e2e leg builds), all genes as features
    c = rid.SCseq.from_filtered_counts(csc_matrix(counts))
    c = rid.compdist(c, metric="pearson")
    tot = counts.sum(0)
    ref = ctx.dist(counts / tot * tot.min() + 0.1, "pearson")
    assert np.array_equal(c.distances.as_matrix(), ref.as_matrix())
    ref.free()
    for o in (a, b, c):
        o.distances.free()
    rid.set_context(None)


def test_int8_peak_measurement_is_sane(ctx):
    """rid_measure_i8_peak (bench.py's tensor ceiling): between half and 1.3x the nominal 4.5 POP/s dense int8 rate."""
    v = ctx.measure_i8_peak()
    assert 1500.0 < v < 6000.0, v


@pytest.mark.parametrize("case", ["clusters", "ties", "curve", "subset_compact", "euclid31"])
def test_build_lazy_equals_full_rescan(ctx, oracle, case, monkeypatch):
    """BUILD's opt-in lazy evaluation (RID_BUILD_LAZY=1: exact gains only for the candidates whose upper bound can still
    win, from the third pick on) against the default incremental rescans and the oracle: identical picks, including
    exact ties (largest index wins) and the 31-bit grid."""
    from raceid3_stemid2_package_b200.synth import synth_features_numpy

    rng = np.random.default_rng(11)
    subset = None
    if case == "clusters":
        x, _ = synth_features_numpy(1500, 200, 7, 3)
        dm = ctx.dist_cellmajor(x, "pearson")
        ks = (3, 8, 25)
    elif case == "ties":
        D = rng.integers(1, 4, size=(300, 300)).astype(np.float64)  # three distinct values: ties everywhere
        D = np.triu(D, 1)
        D = D + D.T
        dm = ctx.from_matrix(D)
        ks = (3, 9, 40)
    elif case == "curve":
        t = np.sort(rng.uniform(0, 1, 900))
        P = np.stack([t, np.sin(7 * t), np.cos(3 * t) * t]) + rng.normal(0, 0.01, (3, 900))
        dm = ctx.dist(P, "euclidean")
        ks = (4, 12, 30)
    elif case == "subset_compact":
        x, _ = synth_features_numpy(1800, 150, 5, 8)
        x[100] = x[7]
        dm = ctx.dist_cellmajor(x, "spearman")
        subset = rng.permutation(1800)[:1100].astype(np.int32)  # 1100^2 words > 1 MB: the compact symmetric copy
        ks = (3, 10)
    else:
        P = rng.normal(size=(700, 4)) * 900.0
        D = np.sqrt(((P[:, None, :] - P[None, :, :]) ** 2).sum(-1))
        dm = ctx.from_matrix(D)
        ks = (5, 17)
    Dg = dm.as_matrix()
    for k in ks:
        monkeypatch.setenv("RID_BUILD_LAZY", "1")
        a = dm.pam(k, subset=subset)
        monkeypatch.delenv("RID_BUILD_LAZY", raising=False)
        b = dm.pam(k, subset=subset)
        o = oracle.pam(Dg, k, subset=subset)
        for r in (a, b):
            assert np.array_equal(r["id_med"], o["id_med"]) and np.array_equal(r["clustering"], o["clustering"])
            assert r["n_swaps"] == o["n_swaps"] and np.allclose(r["objective"], o["objective"], rtol=1e-13, atol=1e-13)
    lw = dm.gap_sweep(12, subset=subset)
    assert np.max(np.abs(lw - oracle.gap_sweep(Dg, 12, subset=subset))) < 1e-12
    dm.free()


@pytest.mark.parametrize("metric", ["pearson", "euclidean"])
def test_clusterboot_first_pick_batched_and_fused(ctx, oracle, metric, monkeypatch):
    """rid_clusterboot takes BUILD's step-0 column sums of all replicates in batched passes over the resident matrix
    (10 replicates per pass), so every replicate's first medoid is known before its compaction, which then accumulates
    the second step's sums directly (no full rescan for that step).  Identical to the oracle and to the plain flow
    (RID_NO_FIRST=1); 36 replicates = two batches (or more); resamples large enough for the compact symmetric copy; the euclidean
    case runs on the 31-bit grid (all four byte planes of the tensor-core column sums in use)."""
    from raceid3_stemid2_package_b200.scseq import bootstrap_samples
    from raceid3_stemid2_package_b200.synth import synth_features_numpy

    N, B, k = 1500, 36, 6
    if metric == "pearson":
        x, _ = synth_features_numpy(N, 180, k, 33)
        x[1200] = x[3]  # duplicated cells: tied column sums
        dm = ctx.dist_cellmajor(x, metric)
    else:
        rng = np.random.default_rng(4)
        t = np.sort(rng.uniform(0, 1, N))
        P = np.stack([t, np.sin(7 * t), np.cos(3 * t) * t]) + rng.normal(0, 0.01, (3, N))
        dm = ctx.dist(P * 50.0, metric)
    Dg = dm.as_matrix()
    samples = bootstrap_samples(N, B, 17000)
    assert min(len(s) for s in samples) ** 2 * 4 >= 1 << 20
    ro = oracle.clusterboot(Dg, k, samples)
    # default: the column sums of 32 replicates per pass on the tensor cores (k_colsum_mma<2>: u8 byte planes, exact);
    # RID_COLSUM_MB=1: 16 per pass; RID_COLSUM_IMAD=1: the 10-per-pass multiply-add kernel; RID_NO_FIRST=1: no batched
    # first picks at all
    for env in (None, "RID_COLSUM_MB", "RID_COLSUM_IMAD", "RID_NO_FIRST"):
        if env:
            monkeypatch.setenv(env, "1")
        r = dm.clusterboot(k, samples)
        if env:
            monkeypatch.delenv(env, raising=False)
        assert np.array_equal(r["partition"], ro["partition"]) and np.array_equal(r["id_med"], ro["id_med"]), env
        assert np.array_equal(r["bootresult"], ro["bootresult"]), env
    # "subset" resamples (no replacement, arbitrary order: the first pick's tie rule follows the POSITIONS)
    sub = bootstrap_samples(N, 3, 99, samp=900)
    r = dm.clusterboot(4, sub)
    ro = oracle.clusterboot(Dg, 4, sub)
    assert np.array_equal(r["partition"], ro["partition"]) and np.array_equal(r["bootresult"], ro["bootresult"])
    dm.free()


def test_spearman_rank_by_sort_equals_counting_kernel(ctx, oracle, monkeypatch):
    """The sort-based rank kernel (bitonic sort in shared memory + tie groups by binary search, O(G log^2 G) per cell)
    against the O(G^2) counting kernel (RID_RANK_COUNT=1): bit-identical matrices; heavy ties (sparse counts), G not a
    power of two, G = 2, a constant cell (zero variance -> NA)."""
    from raceid3_stemid2_package_b200.synth import synth_features_numpy

    for (N, G, seed) in ((400, 333, 1), (257, 2, 2), (300, 1025, 3), (150, 4096, 4)):
        x, _ = synth_features_numpy(N, max(G, 8), 4, seed)
        x = np.ascontiguousarray(x[:, :G])
        x[5] = 0.1  # all values tied: zero variance
        if G > 2:
            x[7, : G // 2] = x[7, 0]  # one huge tie group
        monkeypatch.delenv("RID_RANK_COUNT", raising=False)
        a = ctx.dist_cellmajor(x, "spearman")
        Da = a.as_matrix()
        a.free()
        monkeypatch.setenv("RID_RANK_COUNT", "1")
        b = ctx.dist_cellmajor(x, "spearman")
        Db = b.as_matrix()
        b.free()
        monkeypatch.delenv("RID_RANK_COUNT", raising=False)
        assert np.array_equal(np.isnan(Da), np.isnan(Db)) and np.isnan(Da[5, 0]) and a.n_zero_var == b.n_zero_var == 1
        ok = ~np.isnan(Da)
        assert np.array_equal(Da[ok], Db[ok])
        Do, nz = oracle.dist(np.ascontiguousarray(x.T), "spearman")
        assert nz == 1 and np.max(np.abs(Da[ok] - Do[ok])) <= DIST_TOL


def test_clusterboot_takes_c1_from_the_sweep(ctx, oracle, monkeypatch):
    """clustexp(sat=TRUE) runs clusGapExt's pam(as.dist(D), k) for k = 1..clustnr and then fpc::clusterboot, whose c1
    is the very same pam(as.dist(D), cln) call (R/RaceID_utils.R:113,132).  rid_gap_sweep leaves its full-set results
    with the matrix and rid_clusterboot takes c1 from there: identical to the oracle and to RID_NO_MEMO=1, for a k
    inside the sweep; a k outside the sweep, a sweep on a sample of the cells (`samp`) and a fresh matrix all run c1."""
    from raceid3_stemid2_package_b200.scseq import bootstrap_samples
    from raceid3_stemid2_package_b200.synth import synth_features_numpy

    N, B = 1400, 4
    x, _ = synth_features_numpy(N, 150, 7, 5, "ring")  # a continuum: SWAP really iterates
    dm = ctx.dist_cellmajor(x, "pearson")
    Dg = dm.as_matrix()
    samples = bootstrap_samples(N, B, 17000)

    def hits():
        return ctx.stats()["c1_taken_from_sweep"]

    h0 = hits()
    r_plain = dm.clusterboot(5, samples)  # nothing memoised yet
    assert hits() == h0
    sub = np.random.default_rng(1).permutation(N)[:700].astype(np.int32)
    dm.gap_sweep(6, subset=sub)  # a sweep on a sample of the cells says nothing about the full set
    r_sub = dm.clusterboot(5, samples)
    assert hits() == h0
    lw = dm.gap_sweep(6)
    assert np.max(np.abs(lw - oracle.gap_sweep(Dg, 6))) < 1e-12
    r_memo = dm.clusterboot(5, samples)
    assert hits() == h0 + 1
    r_out = dm.clusterboot(8, samples)  # k beyond the sweep
    assert hits() == h0 + 1
    monkeypatch.setenv("RID_NO_MEMO", "1")
    r_off = dm.clusterboot(5, samples)
    monkeypatch.delenv("RID_NO_MEMO", raising=False)
    assert hits() == h0 + 1
    ro = oracle.clusterboot(Dg, 5, samples)
    for r in (r_plain, r_sub, r_memo, r_off):
        assert np.array_equal(r["partition"], ro["partition"]) and np.array_equal(r["id_med"], ro["id_med"])
        assert np.array_equal(r["bootresult"], ro["bootresult"])
    ro8 = oracle.clusterboot(Dg, 8, samples)
    assert np.array_equal(r_out["partition"], ro8["partition"]) and np.array_equal(r_out["bootresult"], ro8["bootresult"])
    dm.free()
    dm2 = ctx.dist_cellmajor(x, "pearson")  # the memo lives and dies with the matrix
    dm2.clusterboot(5, samples)
    assert hits() == h0 + 1
    dm2.free()


@pytest.mark.parametrize("metric", ["pearson", "euclidean"])
def test_sweep_wk_from_atoms_equals_batched_passes(ctx, oracle, metric, monkeypatch):
    """W.k of all partitions of a sweep from ONE grouped-column-sum pass over the common refinement of the partitions
    (the atoms), against the 6-partitions-per-pass kernel (RID_WK_BATCH=1) and the oracle: identical logW (the
    within-cluster sums are exact integers either way).  Full set and a `samp`-style subset (compact copy, positions in
    arbitrary order); a cap of 4 classes forces the overflow fallback; both fixed-point grids."""
    from raceid3_stemid2_package_b200.synth import synth_features_numpy

    N, K = 1300, 9
    if metric == "pearson":
        x, _ = synth_features_numpy(N, 160, 6, 11, "ring")
        dm = ctx.dist_cellmajor(x, metric)
    else:
        rng = np.random.default_rng(8)
        P = rng.normal(0, 1, (4, N)) + 3.0 * rng.integers(0, 5, N)
        dm = ctx.dist(P * 20.0, metric)
    Dg = dm.as_matrix()
    sub = np.random.default_rng(2).permutation(N)[:1100].astype(np.int32)
    for subset in (None, sub):
        ref = oracle.gap_sweep(Dg, K, subset=subset)
        lw_atoms = dm.gap_sweep(K, subset=subset)
        A = ctx.stats()["wk_atoms"]
        assert A >= K  # the atoms path ran (at least as many classes as the finest partition has clusters)
        monkeypatch.setenv("RID_WK_BATCH", "1")
        lw_batch = dm.gap_sweep(K, subset=subset)
        monkeypatch.delenv("RID_WK_BATCH", raising=False)
        monkeypatch.setenv("RID_WK_ATOM_CAP", "4")
        lw_over = dm.gap_sweep(K, subset=subset)
        assert ctx.stats()["wk_atoms"] == -1  # more classes than the cap: the batched passes did it
        monkeypatch.delenv("RID_WK_ATOM_CAP", raising=False)
        # BUILD's picks 3..K come from the chain's sums by default; RID_SWEEP_BUILD_FIRST=1: BUILD(K.max) up front
        monkeypatch.setenv("RID_SWEEP_BUILD_FIRST", "1")
        lw_first = dm.gap_sweep(K, subset=subset)
        monkeypatch.delenv("RID_SWEEP_BUILD_FIRST", raising=False)
        assert np.array_equal(lw_atoms, lw_first)
        assert np.array_equal(lw_atoms, lw_batch) and np.array_equal(lw_atoms, lw_over)
        assert np.max(np.abs(lw_atoms - ref)) < 1e-12
    dm.free()


def test_many_clusters_keep_the_incremental_paths(ctx, oracle, monkeypatch):
    """k = 70 and k = 100 (> 64: round 1 fell back to full rescans there): incremental SWAP keyed by (old, new) group
    pairs (k^2 <= 16 384 keys) and the enqueue-only bootstrap against the oracle on data without cluster structure (SWAP
    iterates); the chained sweep with the atoms pass against the plain one (full scans, batched W.k) up to K.max = 72."""
    from raceid3_stemid2_package_b200.scseq import bootstrap_samples

    rng = np.random.default_rng(21)
    N = 1000
    x = rng.gamma(2.0, 1.0, (N, 40)) + 0.1
    dm = ctx.dist_cellmajor(x, "pearson")
    Dg = dm.as_matrix()
    r = dm.pam(70)
    o = oracle.pam(Dg, 70)
    assert np.array_equal(r["id_med"], o["id_med"]) and np.array_equal(r["clustering"], o["clustering"])
    assert r["n_swaps"] == o["n_swaps"] and o["n_swaps"] >= 3
    sub = rng.permutation(N)[:600].astype(np.int32)
    r = dm.pam(100, subset=sub)
    o = oracle.pam(Dg[np.ix_(sub, sub)], 100)
    assert np.array_equal(r["id_med"], o["id_med"]) and np.array_equal(r["clustering"], o["clustering"])
    assert r["n_swaps"] == o["n_swaps"]
    samples = bootstrap_samples(N, 2, 17000)
    r = dm.clusterboot(70, samples)
    o = oracle.clusterboot(Dg, 70, samples)
    assert np.array_equal(r["partition"], o["partition"]) and np.array_equal(r["bootresult"], o["bootresult"])
    lw = dm.gap_sweep(72)
    for v in ("RID_SWEEP_FULL", "RID_SWAP_FULL", "RID_WK_BATCH"):
        monkeypatch.setenv(v, "1")
    lw_plain = dm.gap_sweep(72)
    assert np.array_equal(lw, lw_plain)
    dm.free()
