"""ctypes front end of the CPU oracle (oracle.c).  TEST INFRASTRUCTURE ONLY — see oracle.c header.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
PARITY UNPINNED (no R in this image, no golden vectors in the reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[bool, C.CDLL] = {}

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    need = force or not all(os.path.exists(os.path.join(_HERE, f)) for f in ("liboracle.so", "liboracle_omp.so"))
    if need:
        subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib(omp: bool = False) -> C.CDLL:
    if omp not in _LIBS:
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle_omp.so" if omp else "liboracle.so"))
        L.orc_num_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_set_threads.restype = None
        L.orc_dist.restype = C.c_int
        L.orc_dist.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _dp]
        L.orc_pam.restype = C.c_int
        L.orc_pam.argtypes = [_dp, C.c_int, C.c_void_p, C.c_int, C.c_int, _ip, _ip, _dp, C.c_void_p,
                              C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_wk.restype = C.c_double
        L.orc_wk.argtypes = [_dp, C.c_int, C.c_void_p, C.c_int, _ip, C.c_int]
        L.orc_gap_sweep.restype = C.c_int
        L.orc_gap_sweep.argtypes = [_dp, C.c_int, C.c_void_p, C.c_int, C.c_int, _dp]
        L.orc_saturation.restype = C.c_int
        L.orc_saturation.argtypes = [_dp, C.c_int]
        L.orc_clusterboot.restype = C.c_int
        L.orc_clusterboot.argtypes = [_dp, C.c_int, C.c_int, _ip, _ip, C.c_int, _ip, _ip, _dp, _dp]
        L.orc_medoids.restype = C.c_int
        L.orc_medoids.argtypes = [_dp, C.c_int, _ip, _ip, _ip]
        L.orc_refresh.restype = None
        L.orc_refresh.argtypes = [_dp, C.c_int, _ip, C.c_int, _ip]
        L.orc_silhouette.argtypes = [_dp, C.c_int, _ip, _ip, _dp]
        L.orc_silhouette.restype = C.c_int
        L.orc_swap_deltas_colform.restype = None
        L.orc_swap_deltas_colform.argtypes = [_dp, C.c_int, C.c_void_p, C.c_int, _ip, C.c_int, _dp]
        _LIBS[omp] = L
    return _LIBS[omp]


METRICS = {"pearson": 0, "spearman": 1, "logpearson": 2, "euclidean": 3}


def _colmajor(D: np.ndarray) -> np.ndarray:
    """N x N numpy array -> buffer whose memory is column-major (R layout).  We pass the transpose of a
    C-contiguous array so that buf[c*n + r] == D[r, c]."""
    return np.ascontiguousarray(np.asarray(D, dtype=np.float64).T)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def dist(x_gn: np.ndarray, metric: str = "pearson", omp: bool = False):
    """x_gn: G x N (genes x cells) like getfdata()'s result.  Returns (D [N,N] float64, n_zero_var)."""
    x = np.asarray(x_gn, dtype=np.float64)
    G, N = x.shape
    buf = np.ascontiguousarray(x.T)  # cell-major: each cell's G values contiguous == G x N column-major
    D = np.empty((N, N), dtype=np.float64)
    nz = lib(omp).orc_dist(buf, G, N, METRICS[metric], D)
    return D.T.copy(), nz  # symmetric anyway; .T maps column-major buffer to numpy [row, col]


def pam(D: np.ndarray, k: int, subset=None, want_sil: bool = False, init_med=None, max_swaps: int = 100000):
    Dc = _colmajor(D)
    n_full = Dc.shape[0]
    sub = None if subset is None else np.ascontiguousarray(subset, dtype=np.int32)
    m = n_full if sub is None else len(sub)
    med = np.zeros(k, dtype=np.int32)
    labels = np.zeros(m, dtype=np.int32)
    obj = np.zeros(2)
    sil = C.c_double(0.0)
    swaps = np.zeros(3 * max_swaps)
    nsw = C.c_int(0)
    border = np.zeros(k, dtype=np.int32)
    im = None if init_med is None else np.ascontiguousarray(init_med, dtype=np.int32)
    rc = lib().orc_pam(Dc, n_full, _ptr(sub), m, k, med, labels, obj, C.byref(sil) if want_sil else None,
                       _ptr(swaps), max_swaps, C.byref(nsw), _ptr(border), _ptr(im))
    if rc == -1:
        raise ValueError("Number of clusters 'k' must be in {1,2, .., n-1}")
    if rc == -2:
        raise ValueError("No clustering performed, NAs in the computed dissimilarity matrix.")
    return {"id_med": med, "clustering": labels, "objective": obj, "avg_sil": sil.value if want_sil else None,
            "swaps": swaps[: 3 * min(nsw.value, max_swaps)].reshape(-1, 3), "n_swaps": nsw.value,
            "build_order": border}


def wk(D, labels, k, subset=None) -> float:
    Dc = _colmajor(D)
    sub = None if subset is None else np.ascontiguousarray(subset, dtype=np.int32)
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    return lib().orc_wk(Dc, Dc.shape[0], _ptr(sub), len(lab), lab, k)


def gap_sweep(D, kmax, subset=None) -> np.ndarray:
    Dc = _colmajor(D)
    sub = None if subset is None else np.ascontiguousarray(subset, dtype=np.int32)
    m = Dc.shape[0] if sub is None else len(sub)
    out = np.zeros(kmax)
    rc = lib().orc_gap_sweep(Dc, Dc.shape[0], _ptr(sub), m, kmax, out)
    if rc:
        raise ValueError(f"orc_gap_sweep rc={rc}")
    return out


def saturation(logW) -> int:
    g = np.ascontiguousarray(logW, dtype=np.float64)
    return lib().orc_saturation(g, len(g))


def clusterboot(D, k, samples):
    """samples: list of 0-based index arrays (already unique()'d in first-occurrence order)."""
    Dc = _colmajor(D)
    n = Dc.shape[0]
    B = len(samples)
    cat = np.ascontiguousarray(np.concatenate(samples) if B else np.zeros(0), dtype=np.int32)
    lens = np.ascontiguousarray([len(s) for s in samples], dtype=np.int32)
    part = np.zeros(n, dtype=np.int32)
    med = np.zeros(k, dtype=np.int32)
    br = np.zeros((B, k))
    bm = np.zeros(k)
    rc = lib().orc_clusterboot(Dc, n, k, cat, lens, B, part, med, br, bm)
    if rc:
        raise ValueError(f"orc_clusterboot rc={rc}")
    return {"partition": part, "id_med": med, "bootresult": br.T.copy(), "bootmean": bm}


def medoids(D, labels):
    Dc = _colmajor(D)
    n = Dc.shape[0]
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    out = np.zeros(n, dtype=np.int32)
    ids = np.zeros(n, dtype=np.int32)
    nc = lib().orc_medoids(Dc, n, lab, out, ids)
    return out[:nc].copy(), ids[:nc].copy()


def cmdscale(D, k: int = 2):
    """stats::cmdscale(d, k) restated with numpy (upstream: R src/library/stats/R/cmdscale.R + C_DoubleCentre): x = D^2,
    double centring, e = eigen(-x/2, symmetric = TRUE), points = e$vectors[, 1:k] * sqrt(e$values[1:k]).  Returns
    (points n x k, eigenvalues k).  Eigenvector signs are LAPACK's: compare up to a sign per axis."""
    D = np.asarray(D, dtype=np.float64)
    x = D * D
    x = x - x.mean(axis=0, keepdims=True)
    x = x - x.mean(axis=1, keepdims=True)
    w, v = np.linalg.eigh(-x / 2.0)
    idx = np.argsort(-w)[:k]
    ev = w[idx]
    return v[:, idx] * np.sqrt(np.where(ev > 0, ev, np.nan)), ev


def silhouette(D, labels):
    """cluster::silhouette(part, dist): (neighbor, sil_width) per cell; labels 0 = not in the partition."""
    Dc = _colmajor(D)
    n = Dc.shape[0]
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    nb = np.zeros(n, dtype=np.int32)
    sw = np.zeros(n)
    if lib().orc_silhouette(Dc, n, lab, nb, sw) < 0:
        raise ValueError("silhouette needs 2 <= number of clusters <= n - 1")
    return nb, sw


def refresh(D, medoid_idx):
    Dc = _colmajor(D)
    n = Dc.shape[0]
    md = np.ascontiguousarray(medoid_idx, dtype=np.int32)
    out = np.zeros(n, dtype=np.int32)
    lib().orc_refresh(Dc, n, md, len(md), out)
    return out


def swap_deltas_colform(D, med, subset=None):
    Dc = _colmajor(D)
    sub = None if subset is None else np.ascontiguousarray(subset, dtype=np.int32)
    m = Dc.shape[0] if sub is None else len(sub)
    md = np.ascontiguousarray(med, dtype=np.int32)
    out = np.zeros((m, len(md)))
    lib().orc_swap_deltas_colform(Dc, Dc.shape[0], _ptr(sub), m, md, len(md), out)
    return out
