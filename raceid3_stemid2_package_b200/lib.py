"""ctypes binding of libraceid_b200.so (C ABI declared in include/raceid_b200.h).

This is the only way the Python host layer reaches compute.  There is NO CPU fallback: if the shared library is
missing, or no sm_100 device is usable, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libraceid_b200.so")

METRIC_CODES = {"pearson": 0, "spearman": 1, "logpearson": 2, "euclidean": 3}

# every symbol include/raceid_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "rid_last_error", "rid_version", "rid_nccl_unique_id", "rid_ctx_create", "rid_ctx_destroy", "rid_ctx_sync",
    "rid_ctx_last_ms", "rid_ctx_launches", "rid_dist", "rid_dist_csc", "rid_dmat_from_host", "rid_dmat_free", "rid_dmat_info",
    "rid_shard_rows", "rid_ctx_profile", "rid_ctx_profile_get", "rid_ctx_event_record", "rid_ctx_event_elapsed_ms", "rid_measure_dmma_peak", "rid_measure_read_peak",
    "rid_dmat_sub", "rid_pam", "rid_gap_sweep", "rid_clusterboot", "rid_medoids", "rid_refresh", "rid_within_disp", "rid_knn", "rid_silhouette", "rid_cmdscale",
    "rid_ctx_request_abort", "rid_ctx_clear_abort", "rid_ctx_set_interrupt_poll", "rid_measure_i8_peak", "rid_r_bootstrap_samples", "rid_ctx_swaps", "rid_ctx_trim", "rid_ctx_stat",
]

INTERRUPTED = -6  # RID_ERR_INTERRUPTED
POLL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)


class RaceIDError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libraceid_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C raceid3_stemid2_package_b200/csrc).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, ip, dp = C.c_void_p, C.c_void_p, C.c_void_p
    L.rid_last_error.restype = C.c_char_p
    L.rid_version.restype = C.c_int
    L.rid_nccl_unique_id.argtypes = [vp]
    L.rid_ctx_create.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.rid_ctx_destroy.argtypes = [vp]
    L.rid_ctx_sync.argtypes = [vp]
    L.rid_ctx_swaps.argtypes = [vp]
    L.rid_ctx_swaps.restype = C.c_longlong
    L.rid_ctx_trim.argtypes = [vp]
    L.rid_ctx_stat.argtypes = [vp, C.c_int]
    L.rid_ctx_stat.restype = C.c_double
    L.rid_ctx_request_abort.argtypes = [vp]
    L.rid_ctx_clear_abort.argtypes = [vp]
    L.rid_ctx_set_interrupt_poll.argtypes = [vp, vp, vp]
    L.rid_measure_i8_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.rid_ctx_last_ms.argtypes = [vp]
    L.rid_ctx_last_ms.restype = C.c_double
    L.rid_ctx_launches.argtypes = [vp]
    L.rid_ctx_launches.restype = C.c_longlong
    L.rid_shard_rows.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.rid_ctx_profile.argtypes = [vp, C.c_int]
    L.rid_ctx_profile_get.argtypes = [vp, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_double),
                                      C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.rid_measure_dmma_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.rid_measure_read_peak.argtypes = [vp, C.c_double, C.POINTER(C.c_double)]
    L.rid_ctx_event_record.argtypes = [vp, C.c_int]
    L.rid_ctx_event_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.rid_dist.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_int)]
    L.rid_dist_csc.argtypes = [vp, ip, ip, dp, C.c_int, C.c_int, ip, C.c_int, dp, C.c_int, C.POINTER(vp), C.POINTER(C.c_int)]
    L.rid_dmat_from_host.argtypes = [vp, dp, C.c_int, C.POINTER(vp)]
    L.rid_dmat_free.argtypes = [vp]
    L.rid_dmat_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.rid_dmat_sub.argtypes = [vp, ip, C.c_int, ip, C.c_int, dp]
    L.rid_pam.argtypes = [vp, ip, C.c_int, C.c_int, ip, ip, dp, dp, ip]
    L.rid_gap_sweep.argtypes = [vp, ip, C.c_int, C.c_int, dp]
    L.rid_clusterboot.argtypes = [vp, C.c_int, ip, ip, C.c_int, ip, ip, dp]
    L.rid_medoids.argtypes = [vp, ip, C.c_int, ip, ip, ip]
    L.rid_refresh.argtypes = [vp, ip, C.c_int, ip]
    L.rid_within_disp.argtypes = [vp, ip, C.c_int, ip, C.c_int, dp]
    L.rid_silhouette.argtypes = [vp, ip, C.c_int, ip, dp]
    L.rid_cmdscale.argtypes = [vp, C.c_int, C.c_double, C.c_int, dp, dp, ip]
    L.rid_knn.argtypes = [vp, C.c_int, ip, dp]
    L.rid_r_bootstrap_samples.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_int, ip, ip]
    for name in SYMBOLS:
        getattr(L, name)  # AttributeError here = the library does not export what include/raceid_b200.h declares
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise RaceIDError(rc, load().rid_last_error().decode("utf-8", "replace"))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def shard_rows(n: int, rank: int, world: int) -> tuple[int, int]:
    """Row block [row0, row1) owned by `rank` (host-only helper of the library)."""
    a, b = C.c_int(), C.c_int()
    _check(load().rid_shard_rows(n, rank, world, C.byref(a), C.byref(b)))
    return a.value, b.value


def r_bootstrap_samples(n: int, bootnr: int, seed: int, subset_size: int = 0):
    """The resamples clusterboot draws from R's generator (host-only C helper, no GPU): (concatenated int32 indices,
    int32 lengths)."""
    cap = n if subset_size <= 0 else min(n, int(subset_size))
    cat = np.empty(max(bootnr * cap, 1), dtype=np.int32)
    lens = np.zeros(max(bootnr, 1), dtype=np.int32)
    _check(load().rid_r_bootstrap_samples(int(n), int(bootnr), int(seed) & 0xFFFFFFFF, int(subset_size), _p(cat), _p(lens)))
    lens = lens[:bootnr]
    return cat[:int(lens.sum())], lens


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(load().rid_nccl_unique_id(buf))
    return buf.raw


class Context:
    """One GPU (one rank of a row-sharded job).  Mirrors rid_ctx."""

    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, nccl_uid: bytes | None = None):
        self._h = C.c_void_p()
        self.rank, self.world, self.device = rank, world, device
        uid = C.create_string_buffer(nccl_uid, 128) if nccl_uid is not None else None
        _check(load().rid_ctx_create(device, rank, world, uid, C.byref(self._h)))

    def close(self):
        if self._h:
            load().rid_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _check(load().rid_ctx_sync(self._h))

    def request_abort(self):
        """Ask the compute call running on this context (from another thread, or a signal handler) to stop: it returns
        RID_ERR_INTERRUPTED (RaceIDError with code -6) at its next poll point."""
        _check(load().rid_ctx_request_abort(self._h))

    def clear_abort(self):
        _check(load().rid_ctx_clear_abort(self._h))

    def set_interrupt_poll(self, fn=None):
        """fn() -> truthy to cancel; called on the calling thread between waves of GPU work (what the R stub does with
        R_CheckUserInterrupt).  None removes the callback."""
        if fn is None:
            self._poll_keep = None
            _check(load().rid_ctx_set_interrupt_poll(self._h, None, None))
            return
        cb = POLL_FN(lambda _arg: 1 if fn() else 0)
        self._poll_keep = cb  # keep the thunk alive
        _check(load().rid_ctx_set_interrupt_poll(self._h, C.cast(cb, C.c_void_p), None))

    def measure_i8_peak(self) -> float:
        """tcgen05.mma kind::i8 issue-rate ceiling measured on this GPU right now, TOP/s."""
        v = C.c_double()
        _check(load().rid_measure_i8_peak(self._h, C.byref(v)))
        return v.value

    @property
    def last_ms(self) -> float:
        return load().rid_ctx_last_ms(self._h)

    @property
    def launches(self) -> int:
        return int(load().rid_ctx_launches(self._h))

    @property
    def swaps(self) -> int:
        return int(load().rid_ctx_swaps(self._h))

    def stats(self) -> dict:
        names = ("swaps", "boot_redone", "c1_ms", "first_picks_and_replica_wait_ms", "replicates_ms", "replica_pull_bytes", "boot_ahead",
                 "replicates_host_enqueue_ms", "replica_pull_ms", "c1_taken_from_sweep", "wk_atoms")
        return {n: load().rid_ctx_stat(self._h, i) for i, n in enumerate(names)}

    def trim(self):
        _check(load().rid_ctx_trim(self._h))

    PROF_KINDS = {"gram": 0, "scan_swap": 1, "scan_build": 2, "scan_group": 3, "prep": 4, "compact": 5, "gram_i8": 6, "scan_delta": 7}

    def profile(self, kinds=True):
        """kinds: True = every kernel kind, False/None = off, or a list of kind names (see PROF_KINDS) — recording only
        the kernels a roofline needs keeps the timed region free of tens of thousands of CUDA events."""
        if kinds is True:
            mask = -1
        elif not kinds:
            mask = 0
        else:
            mask = 0
            for k in kinds:
                mask |= 1 << self.PROF_KINDS[k]
        _check(load().rid_ctx_profile(self._h, mask))

    def profile_get(self, kind: str) -> dict:
        n, ms, work, tr = C.c_longlong(), C.c_double(), C.c_double(), C.c_double()
        _check(load().rid_ctx_profile_get(self._h, self.PROF_KINDS[kind], C.byref(n), C.byref(ms), C.byref(work),
                                          C.byref(tr)))
        return {"launches": n.value, "ms": ms.value, "work": work.value, "traffic": tr.value}

    def measure_dmma_peak(self) -> float:
        v = C.c_double()
        _check(load().rid_measure_dmma_peak(self._h, C.byref(v)))
        return v.value

    def measure_read_peak(self, gib: float = 8.0) -> float:
        v = C.c_double()
        _check(load().rid_measure_read_peak(self._h, gib, C.byref(v)))
        return v.value

    def event_record(self, slot: int):
        _check(load().rid_ctx_event_record(self._h, slot))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_double()
        _check(load().rid_ctx_event_elapsed_ms(self._h, a, b, C.byref(ms)))
        return ms.value

    # ---- distance matrices ----
    def dist(self, x_gn, metric: str = "pearson") -> "DistMatrix":
        """x_gn: G x N numpy array (genes x cells), or a (device_ptr, G, N) tuple for device-resident input laid
        out cell-major (each cell's G doubles contiguous)."""
        if metric not in METRIC_CODES:
            raise ValueError("metric has to be one of the following: spearman, pearson, logpearson, euclidean")
        h = C.c_void_p()
        nz = C.c_int(0)
        if isinstance(x_gn, tuple):
            ptr, G, N = x_gn
            _check(load().rid_dist(self._h, C.c_void_p(ptr), G, N, METRIC_CODES[metric], 1, C.byref(h), C.byref(nz)))
        else:
            x = np.asarray(x_gn, dtype=np.float64)
            G, N = x.shape
            buf = np.ascontiguousarray(x.T)  # == G x N column-major
            _check(load().rid_dist(self._h, _p(buf), G, N, METRIC_CODES[metric], 0, C.byref(h), C.byref(nz)))
        return DistMatrix(self, h, nz.value)

    def dist_cellmajor(self, x_ng: np.ndarray, metric: str = "pearson") -> "DistMatrix":
        """x_ng: N x G C-contiguous float64 host array (one row per cell) — no transpose copy."""
        x = np.ascontiguousarray(x_ng, dtype=np.float64)
        N, G = x.shape
        h = C.c_void_p()
        nz = C.c_int(0)
        _check(load().rid_dist(self._h, _p(x), G, N, METRIC_CODES[metric], 0, C.byref(h), C.byref(nz)))
        return DistMatrix(self, h, nz.value)

    def dist_csc(self, indices, indptr, data, n_rows_all: int, feature_rows, counts, metric: str = "pearson") -> "DistMatrix":
        """compdist straight from the filtered dgCMatrix (CSC over the kept cells): getfdata is fused on the device."""
        i, p_, fr = _i32(indices), _i32(indptr), _i32(feature_rows)
        v = np.ascontiguousarray(data, dtype=np.float64)
        c = np.ascontiguousarray(counts, dtype=np.float64)
        h = C.c_void_p()
        nz = C.c_int(0)
        _check(load().rid_dist_csc(self._h, _p(i), _p(p_), _p(v), n_rows_all, len(c), _p(fr), len(fr), _p(c),
                                   METRIC_CODES[metric], C.byref(h), C.byref(nz)))
        return DistMatrix(self, h, nz.value)

    def from_matrix(self, D: np.ndarray) -> "DistMatrix":
        D = np.asarray(D, dtype=np.float64)
        if D.ndim != 2 or D.shape[0] != D.shape[1]:
            raise ValueError("distance matrix must be square")
        buf = np.ascontiguousarray(D.T)  # column-major
        h = C.c_void_p()
        _check(load().rid_dmat_from_host(self._h, _p(buf), D.shape[0], C.byref(h)))
        return DistMatrix(self, h, 0)


class DistMatrix:
    """Device-resident, row-block-sharded distance matrix (rid_dmat).  Plays the role of `sc@distances`."""

    def __init__(self, ctx: Context, handle: C.c_void_p, n_zero_var: int):
        self.ctx = ctx
        self._h = handle
        self.n_zero_var = n_zero_var
        n, r0, r1, has_na = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        step = C.c_double()
        _check(load().rid_dmat_info(self._h, C.byref(n), C.byref(r0), C.byref(r1), C.byref(step), C.byref(has_na)))
        self.N, self.row0, self.row1, self.step, self.has_na = n.value, r0.value, r1.value, step.value, bool(has_na.value)

    def free(self):
        if self._h:
            load().rid_dmat_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    @property
    def shape(self):
        return (self.N, self.N)

    def sub(self, rows=None, cols=None) -> np.ndarray:
        r, c = _i32(rows), _i32(cols)
        m = self.N if r is None else len(r)
        n = self.N if c is None else len(c)
        out = np.empty((n, m), dtype=np.float64)  # column-major m x n
        _check(load().rid_dmat_sub(self._h, _p(r), m, _p(c), n, _p(out)))
        return out.T

    def as_matrix(self) -> np.ndarray:
        return np.ascontiguousarray(self.sub())

    def pam(self, k: int, subset=None, want_sil: bool = False):
        s = _i32(subset)
        m = self.N if s is None else len(s)
        med = np.zeros(max(k, 1), dtype=np.int32)
        lab = np.zeros(m, dtype=np.int32)
        obj = np.zeros(2)
        sil = np.zeros(1)
        nsw = np.zeros(1, dtype=np.int32)
        _check(load().rid_pam(self._h, _p(s), m, k, _p(med), _p(lab), _p(obj), _p(sil) if want_sil else None, _p(nsw)))
        return {"id_med": med[:k], "clustering": lab, "objective": obj, "avg_sil": float(sil[0]) if want_sil else None,
                "n_swaps": int(nsw[0])}

    def gap_sweep(self, kmax: int, subset=None) -> np.ndarray:
        s = _i32(subset)
        m = self.N if s is None else len(s)
        out = np.zeros(kmax)
        _check(load().rid_gap_sweep(self._h, _p(s), m, kmax, _p(out)))
        return out

    def clusterboot(self, k: int, samples):
        """samples: list of index vectors, or the (concatenated, lengths) pair r_bootstrap_samples returns."""
        if isinstance(samples, tuple):
            cat, lens = _i32(samples[0]), _i32(samples[1])
            B = len(lens)
            if B == 0:
                cat, lens = np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32)
        else:
            B = len(samples)
            cat = _i32(np.concatenate(samples)) if B else np.zeros(1, dtype=np.int32)
            lens = _i32([len(s) for s in samples]) if B else np.zeros(1, dtype=np.int32)
        part = np.zeros(self.N, dtype=np.int32)
        med = np.zeros(k, dtype=np.int32)
        br = np.zeros((max(B, 1), k))
        _check(load().rid_clusterboot(self._h, k, _p(cat), _p(lens), B, _p(part), _p(med), _p(br)))
        br = br[:B]
        return {"partition": part, "id_med": med, "bootresult": br.T.copy(),
                "bootmean": br.mean(axis=0) if B else np.ones(k)}

    def medoids(self, labels):
        lab = _i32(labels)
        out = np.zeros(self.N, dtype=np.int32)
        ids = np.zeros(self.N, dtype=np.int32)
        nc = C.c_int(0)
        _check(load().rid_medoids(self._h, _p(lab), len(lab), _p(out), _p(ids), C.byref(nc)))
        return out[:nc.value].copy(), ids[:nc.value].copy()

    def refresh(self, medoid_idx) -> np.ndarray:
        md = _i32(medoid_idx)
        lab = np.zeros(self.N, dtype=np.int32)
        _check(load().rid_refresh(self._h, _p(md), len(md), _p(lab)))
        return lab

    def knn(self, K: int, with_dist: bool = False):
        """apply(D, 1, function(x) head(order(x), K)): returns idx [N, K] (0-based) and optionally the distances."""
        idx = np.zeros((self.N, K), dtype=np.int32)
        dist = np.zeros((self.N, K)) if with_dist else None
        _check(load().rid_knn(self._h, K, _p(idx), _p(dist)))
        return (idx, dist) if with_dist else idx

    def cmdscale(self, k: int = 2, tol: float = 1e-9, max_iter: int = 1000, with_eig: bool = False):
        """stats::cmdscale(d, k) (R/RaceID.R:1093): N x k coordinates (axis signs: largest coordinate positive)."""
        pts = np.zeros((k, self.N))  # column-major N x k
        ev = np.zeros(k)
        it = np.zeros(1, dtype=np.int32)
        _check(load().rid_cmdscale(self._h, k, float(tol), int(max_iter), _p(pts), _p(ev), _p(it)))
        out = np.ascontiguousarray(pts.T)
        return (out, ev, int(it[0])) if with_eig else out

    def silhouette(self, labels):
        """cluster::silhouette(part, dist) (R/RaceID.R:1283-1284): neighbour cluster and silhouette width per cell."""
        lab = _i32(labels)
        if len(lab) != self.N:
            raise ValueError("labels must have one entry per cell")
        nb = np.zeros(self.N, dtype=np.int32)
        sw = np.zeros(self.N)
        _check(load().rid_silhouette(self._h, _p(lab), self.N, _p(nb), _p(sw)))
        return nb, sw

    def within_disp(self, labels, k: int, subset=None) -> float:
        s = _i32(subset)
        lab = _i32(labels)
        W = np.zeros(1)
        _check(load().rid_within_disp(self._h, _p(s), len(lab), _p(lab), k, _p(W)))
        return float(W[0])
