"""Pins the oracle's dist.gen restatement (oracle.c orc_dist; reference R/RaceID_utils.R:9-15) against independent
double-precision implementations available here: numpy.corrcoef, scipy.stats.rankdata, scipy pdist."""
import numpy as np
import pytest
from scipy.spatial.distance import pdist, squareform
from scipy.stats import rankdata

from conftest import load_golden


def _x(seed=0, G=60, N=40, ties=True):
    rng = np.random.default_rng(seed)
    c = rng.poisson(1.5, size=(G, N)).astype(np.float64)
    if not ties:
        c += rng.random((G, N))
    return c / c.sum(0, keepdims=True) * c.sum(0).min() + 0.1


def test_pearson_matches_numpy(oracle):
    x = _x()
    D, nz = oracle.dist(x, "pearson")
    assert nz == 0
    ref = 1.0 - np.corrcoef(x.T)
    np.fill_diagonal(ref, 0.0)
    assert np.max(np.abs(D - ref)) < 1e-13
    assert np.all(np.diag(D) == 0.0)
    assert np.array_equal(D, D.T)
    assert D.min() >= 0.0 and D.max() <= 2.0


def test_logpearson_matches_numpy(oracle):
    x = _x(1)
    D, _ = oracle.dist(x, "logpearson")
    ref = 1.0 - np.corrcoef(np.log2(x).T)
    np.fill_diagonal(ref, 0.0)
    assert np.max(np.abs(D - ref)) < 1e-13


def test_spearman_average_ties_matches_scipy(oracle):
    x = _x(2)  # Poisson counts: a huge tie group at 0.1
    D, _ = oracle.dist(x, "spearman")
    r = np.apply_along_axis(lambda v: rankdata(v, method="average"), 0, x)
    ref = 1.0 - np.corrcoef(r.T)
    np.fill_diagonal(ref, 0.0)
    assert np.max(np.abs(D - ref)) < 1e-13


def test_euclidean_matches_scipy(oracle):
    x = _x(3)
    D, _ = oracle.dist(x, "euclidean")
    ref = squareform(pdist(x.T, "euclidean"))
    assert np.max(np.abs(D - ref)) < 1e-11


def test_zero_variance_cell_gives_na(oracle):
    x = _x(4)
    x[:, 5] = 0.7
    D, nz = oracle.dist(x, "pearson")
    assert nz == 1
    assert np.all(np.isnan(D[5, :])) and np.all(np.isnan(D[:, 5]))
    assert not np.isnan(np.delete(np.delete(D, 5, 0), 5, 1)).any()


def test_duplicate_cells_distance_zero(oracle):
    x = _x(5)
    x[:, 7] = x[:, 3]
    for metric in ("pearson", "spearman", "euclidean"):
        D, _ = oracle.dist(x, metric)
        assert abs(D[7, 3]) < 1e-15


@pytest.mark.parametrize("name", ["cfg1_intestinalDataSmall.npz", "cfg2_intestinalData.npz"])
def test_golden_distances(oracle, name):
    g = load_golden(name)
    x = g["x"]
    for metric in ("pearson", "spearman", "logpearson", "euclidean"):
        D, nz = oracle.dist(x, metric)
        assert nz == int(g[f"nz_{metric}"])
        assert np.array_equal(D, g[f"D_{metric}"])
    ref = 1.0 - np.corrcoef(x.T)
    np.fill_diagonal(ref, 0.0)
    assert np.max(np.abs(g["D_pearson"] - ref)) < 1e-12


def test_cmdscale_recovers_euclidean_configurations(oracle):
    """oracle.cmdscale restates stats::cmdscale (reference R/RaceID.R:1093).  Classical MDS of Euclidean distances
    recovers the configuration up to a rigid motion: with k = d the pairwise distances of the returned points equal
    the input, and the eigenvalues are those of the centred Gram matrix."""
    rng = np.random.default_rng(11)
    X = rng.normal(size=(60, 3)) * np.array([5.0, 2.0, 0.5])
    D = np.sqrt(((X[:, None, :] - X[None, :, :]) ** 2).sum(-1))
    P, ev = oracle.cmdscale(D, 3)
    D2 = np.sqrt(((P[:, None, :] - P[None, :, :]) ** 2).sum(-1))
    assert np.max(np.abs(D2 - D)) < 1e-10
    Xc = X - X.mean(0)
    assert np.allclose(ev, np.sort(np.linalg.eigvalsh(Xc @ Xc.T))[::-1][:3], rtol=1e-10)
    assert np.allclose(P.mean(0), 0.0, atol=1e-12)  # double centring: coordinates are centred
    # a non-Euclidean "distance" (1 - correlation) still gives the k largest eigenvalues, in decreasing order
    Z = rng.normal(size=(40, 25))
    Dc = 1.0 - np.corrcoef(Z)
    _, ev2 = oracle.cmdscale(Dc, 2)
    assert ev2[0] >= ev2[1] > 0
