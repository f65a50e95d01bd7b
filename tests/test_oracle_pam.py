"""Pins the oracle's PAM / W.k / clusterboot / compmedoids restatements with brute-force and independent numpy
re-derivations on small inputs (the reference holds no golden vectors for this path)."""
import itertools

import numpy as np
import pytest

from conftest import load_golden


def _dist(n=40, seed=0, dup=False):
    rng = np.random.default_rng(seed)
    cent = rng.normal(size=(4, 6)) * 3
    pts = cent[rng.integers(0, 4, n)] + rng.normal(size=(n, 6))
    if dup:
        pts[5] = pts[2]
        pts[9] = pts[2]
    D = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1))
    D = np.round(D * 64) / 64  # coarse grid => plenty of exact ties
    np.fill_diagonal(D, 0.0)
    return D


def _td(D, med):
    return D[:, med].min(axis=1).sum()


def _py_build(D, k):
    n = len(D)
    dn = np.full(n, D.max() * 1.1 + 1)
    med = []
    for _ in range(k):
        best, arg = 0.0, -1
        for i in range(n):
            if i in med:
                continue
            gain = np.maximum(dn - D[i], 0).sum()
            if best <= gain:
                best, arg = gain, i
        med.append(arg)
        dn = np.minimum(dn, D[arg])
    return med


@pytest.mark.parametrize("seed,dup", [(0, False), (1, True), (2, False)])
def test_build_order_is_greedy_with_largest_index_ties(oracle, seed, dup):
    D = _dist(seed=seed, dup=dup)
    r = oracle.pam(D, 5)
    assert list(r["build_order"]) == _py_build(D, 5)
    # prefix property used by the GPU sweep: BUILD(k) is the first k picks of BUILD(K)
    for k in (1, 2, 3, 4):
        assert list(oracle.pam(D, k)["build_order"]) == _py_build(D, 5)[:k]


@pytest.mark.parametrize("seed,dup,k", [(0, False, 3), (1, True, 4), (3, False, 6), (4, True, 2)])
def test_swap_sequence_is_steepest_descent_first_tie(oracle, seed, dup, k):
    D = _dist(seed=seed, dup=dup)
    n = len(D)
    r = oracle.pam(D, k)
    med = list(r["build_order"])
    td = _td(D, med)
    assert abs(r["objective"][0] - td / n) < 1e-12
    for h, i, dz in r["swaps"]:
        h, i = int(h), int(i)
        # brute force: every (h', i') swap cost by recomputing TD from scratch
        best = (1.0, -1, -1)
        for hh in range(n):
            if hh in med:
                continue
            for ii in sorted(med):
                m2 = [hh if m == ii else m for m in med]
                d = _td(D, m2) - td
                if d < best[0] - 1e-12:
                    best = (d, hh, ii)
        assert (best[1], best[2]) == (h, i)
        assert abs(best[0] - dz) < 1e-9
        med[med.index(i)] = h
        new = _td(D, med)
        assert new < td
        td = new
    assert abs(r["objective"][1] - td / n) < 1e-12
    # fixed point: no single swap improves
    for hh in range(n):
        if hh in med:
            continue
        for ii in med:
            m2 = [hh if m == ii else m for m in med]
            assert _td(D, m2) >= td - 1e-12
    assert sorted(r["id_med"]) == sorted(med)


def test_colform_deltas_equal_literal_loop(oracle):
    """The column formulation the GPU uses (TD'(h,c) = sum_j min(d_jh, cap_c(j))) gives the literal pam.c deltas."""
    D = _dist(seed=5, dup=True)
    n = len(D)
    med = list(oracle.pam(D, 4)["build_order"])
    dz = oracle.swap_deltas_colform(D, med)
    td = _td(D, med)
    for h in range(n):
        for c, i in enumerate(med):
            if h in med:
                assert np.isinf(dz[h, c])
                continue
            m2 = [h if m == i else m for m in med]
            assert abs(dz[h, c] - (_td(D, m2) - td)) < 1e-10


def test_labels_numbered_by_first_appearance(oracle):
    D = _dist(seed=6)
    r = oracle.pam(D, 4)
    lab = r["clustering"]
    seen = []
    for v in lab:
        if v not in seen:
            seen.append(v)
    assert seen == [1, 2, 3, 4]
    for c in range(4):
        assert lab[r["id_med"][c]] == c + 1
    near = np.argmin(D[:, r["id_med"]], axis=1) + 1
    assert np.array_equal(near, lab)


def test_subset_equals_materialised_submatrix(oracle):
    D = _dist(seed=7, n=50)
    sub = np.array([40, 3, 17, 22, 9, 31, 8, 0, 49, 12, 13, 27, 5, 44, 30, 2], dtype=np.int32)
    a = oracle.pam(D, 3, subset=sub)
    b = oracle.pam(D[np.ix_(sub, sub)], 3)
    assert np.array_equal(a["id_med"], b["id_med"]) and np.array_equal(a["clustering"], b["clustering"])
    assert np.array_equal(a["objective"], b["objective"])


def test_k_bounds_and_na(oracle):
    D = _dist(seed=8)
    with pytest.raises(ValueError):
        oracle.pam(D, 0)
    with pytest.raises(ValueError):
        oracle.pam(D, len(D))
    D2 = D.copy()
    D2[3, 1] = D2[1, 3] = np.nan
    with pytest.raises(ValueError):
        oracle.pam(D2, 2)
    r = oracle.pam(D, 1)
    assert r["n_swaps"] == 0 and set(r["clustering"]) == {1}


def test_silhouette_matches_numpy(oracle):
    D = _dist(seed=9)
    r = oracle.pam(D, 3, want_sil=True)
    lab = r["clustering"]
    sw = []
    for i in range(len(D)):
        own = lab == lab[i]
        if own.sum() == 1:
            sw.append(0.0)
            continue
        a = D[i, own].sum() / (own.sum() - 1)
        b = min(D[i, lab == c].mean() for c in set(lab) if c != lab[i])
        sw.append((b - a) / max(a, b))
    assert abs(np.mean(sw) - r["avg_sil"]) < 1e-12


def test_wk_and_sweep(oracle):
    D = _dist(seed=10)
    lab = oracle.pam(D, 3)["clustering"]
    ref = 0.5 * sum(D[np.ix_(lab == c, lab == c)].sum() / (lab == c).sum() for c in (1, 2, 3))
    assert abs(oracle.wk(D, lab, 3) - ref) < 1e-10
    lw = oracle.gap_sweep(D, 6)
    assert abs(lw[0] - np.log(0.5 * D.sum() / len(D))) < 1e-12
    assert abs(lw[2] - np.log(ref)) < 1e-12


def test_saturation_rule(oracle):
    from raceid3_stemid2_package_b200.scseq import saturation_cln

    rng = np.random.default_rng(0)
    for _ in range(20):
        g = np.cumsum(-np.abs(rng.normal(size=12)) * np.linspace(1.0, 0.05, 12)) + 5
        assert oracle.saturation(g) == saturation_cln(g)


def test_clusterboot_jaccard(oracle):
    D = _dist(seed=11, n=60)
    rng = np.random.default_rng(1)
    samples = []
    for _ in range(5):
        s = rng.integers(0, 60, 60)
        _, first = np.unique(s, return_index=True)
        samples.append(s[np.sort(first)].astype(np.int32))
    cb = oracle.clusterboot(D, 3, samples)
    part = cb["partition"]
    assert np.array_equal(part, oracle.pam(D, 3)["clustering"])
    for b, s in enumerate(samples):
        bl = oracle.pam(D, 3, subset=s)["clustering"]
        for j in range(3):
            a = part[s] == j + 1
            best = max((a & (bl == c + 1)).sum() / max((a | (bl == c + 1)).sum(), 1) for c in range(3))
            assert abs(cb["bootresult"][j, b] - best) < 1e-15
    assert np.allclose(cb["bootmean"], cb["bootresult"].mean(axis=1))


def test_medoids_and_refresh(oracle):
    D = _dist(seed=12, dup=True)
    lab = oracle.pam(D, 4)["clustering"].copy()
    lab[0] = 0  # a cell outside `part`
    med, ids = oracle.medoids(D, lab)
    assert list(ids) == [1, 2, 3, 4]
    for c, m in zip(ids, med):
        f = np.nonzero(lab == c)[0]
        y = D[np.ix_(f, f)].mean(axis=0)
        assert m == f[np.nonzero(y == y.min())[0][0]]
    new = oracle.refresh(D, med)
    assert np.array_equal(new, np.argmin(D[:, med], axis=1) + 1)


@pytest.mark.parametrize("name", ["cfg1_intestinalDataSmall.npz", "cfg2_intestinalData.npz"])
def test_golden_pam_and_boot(oracle, name):
    """The committed golden vectors are reproducible from the oracle (they are what the GPU path is held to)."""
    g = load_golden(name)
    D = g["D_pearson"]
    for k in (2, 3, 5):
        r = oracle.pam(D, k, want_sil=True)
        assert np.array_equal(r["id_med"], g[f"pam{k}_id_med"])
        assert np.array_equal(r["clustering"], g[f"pam{k}_clustering"])
        assert np.array_equal(r["objective"], g[f"pam{k}_objective"])
    lw = oracle.gap_sweep(D, len(g["logW"]))
    assert np.array_equal(lw, g["logW"])
    assert oracle.saturation(lw) == int(g["cln_sat"])


def test_silhouette_matches_sklearn_and_tie_rules(oracle):
    """orc_silhouette restates cluster::silhouette(part, dist) (plotsilhouette, reference R/RaceID.R:1283-1284).  The
    widths are pinned by scikit-learn's independent implementation; the neighbour rule (first minimum in increasing
    cluster number) and the singleton rule (s = 0) by construction."""
    from sklearn.metrics import silhouette_samples

    rng = np.random.default_rng(3)
    pts = np.concatenate([rng.normal(c, 0.7, (40, 3)) for c in ((0, 0, 0), (4, 0, 0), (0, 5, 0), (3, 3, 3))])
    D = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1))
    lab = np.repeat([2, 5, 7, 9], 40).astype(np.int32)  # arbitrary positive cluster numbers
    lab[[3, 50, 90]] = [5, 9, 2]                          # a few misassigned cells -> negative widths
    nb, sw = oracle.silhouette(D, lab)
    assert np.allclose(sw, silhouette_samples(D, lab, metric="precomputed"), atol=1e-12)
    means = np.stack([D[:, lab == c].mean(1) for c in (2, 5, 7, 9)], 1)
    for i in range(len(lab)):
        other = [c for c in (2, 5, 7, 9) if c != lab[i]]
        assert nb[i] == other[int(np.argmin([means[i, (2, 5, 7, 9).index(c)] for c in other]))]
    # singleton cluster -> width 0; unlabelled cell -> NaN / neighbour 0 and it is ignored by everybody else
    lab2 = lab.copy()
    lab2[0] = 11
    lab2[1] = 0
    nb2, sw2 = oracle.silhouette(D, lab2)
    assert sw2[0] == 0.0 and nb2[0] in (2, 5, 7, 9) and np.isnan(sw2[1]) and nb2[1] == 0
    keep = lab2 > 0
    ref = silhouette_samples(D[np.ix_(keep, keep)], lab2[keep], metric="precomputed")
    assert np.allclose(sw2[keep], ref, atol=1e-12)
    # exact tie between two neighbour clusters: the lower cluster number wins
    Dt = np.array([[0, 1, 2, 2], [1, 0, 2, 2], [2, 2, 0, 9], [2, 2, 9, 0]], dtype=np.float64)
    nbt, _ = oracle.silhouette(Dt, np.array([1, 1, 3, 2], dtype=np.int32))
    assert nbt[0] == 2 and nbt[1] == 2
    with pytest.raises(ValueError):
        oracle.silhouette(D, np.ones(len(lab), dtype=np.int32))
