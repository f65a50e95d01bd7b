"""Host-side mirror of the reference's R API for the clustering hot path.

Same names, argument meaning and error behaviour as the R functions they stand in for, so that code (and tests)
written against RaceID read the same:

    sc = SCseq(expdata)                   # R/RaceID.R:47-83
    sc = filterdata(sc, mintotal=3000)    # R/RaceID.R:109-242   (host-side producer of the hot path's input)
    sc = compdist(sc, metric="pearson")   # R/RaceID.R:752-842   -> libraceid_b200 rid_dist
    sc = clustexp(sc, cln=..., ...)       # R/RaceID.R:894-912   -> rid_gap_sweep / rid_clusterboot
    m  = compmedoids(sc, part)            # R/RaceID.R:1298-1316 -> rid_medoids

Everything numeric on the hot path runs in libraceid_b200.so (CUDA, sm_100a) through raceid3_stemid2_package_b200.lib;
this module only validates arguments, draws R-compatible bootstrap index vectors, and packs results into the
same fields the R object exposes (`distances`, `cluster$kpart|jaccard|gap|clb|features`, `clusterpar`, `fcol`).
There is no CPU fallback.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import lib as _lib
from .rrng import RRng, r_unique

_default_ctx: _lib.Context | None = None


def get_context() -> _lib.Context:
    """The process-wide GPU context (device from LOCAL_RANK, rank/world from torch.distributed if initialised)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = _lib.Context(device=0)
    return _default_ctx


def set_context(ctx: _lib.Context | None) -> None:
    global _default_ctx
    _default_ctx = ctx


def init_distributed(device: int | None = None) -> _lib.Context:
    """One process per GPU under torchrun: rank 0 makes the NCCL id, torch.distributed ships it (plumbing only),
    the library then owns its own communicator for the row-sharded kernels."""
    import os

    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", rank))
    box = [_lib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx = _lib.Context(device=device, rank=rank, world=world, nccl_uid=box[0])
    set_context(ctx)
    return ctx


# ------------------------------------------------------------------------------------------------------------
# SCseq object (the slots the hot path reads and writes)
# ------------------------------------------------------------------------------------------------------------
@dataclass
class SCseq:
    expdata: np.ndarray  # genes x cells
    genes_all: list = field(default_factory=list)  # rownames(expdata)
    cells_all: list = field(default_factory=list)  # colnames(expdata)
    ndata: np.ndarray | None = None
    cell_names: list = field(default_factory=list)  # colnames(ndata) == names(counts)
    counts: np.ndarray | None = None
    genes: list = field(default_factory=list)
    dimRed: dict = field(default_factory=dict)
    distances: object = None  # DistMatrix handle (slot type "ANY" upstream) or a plain numpy matrix
    cluster: dict = field(default_factory=dict)
    background: dict = field(default_factory=dict)
    cpart: np.ndarray | None = None
    medoids: list = field(default_factory=list)
    fcol: list = field(default_factory=list)
    filterpar: dict = field(default_factory=dict)
    clusterpar: dict = field(default_factory=dict)
    out: dict = field(default_factory=dict)
    fcounts: object = None  # sparse input only: expdata[, kept cells] as CSC (what rid_dist_csc reads)

    @classmethod
    def from_filtered_counts(cls, counts_csc, genes=None, cells=None, features=None) -> "SCseq":
        """An object in the state filterdata leaves it in when every cell passes `mintotal` and every gene is kept:
        counts_csc is the genes x cells count matrix (scipy CSC; its arrays are used as they are, e.g. page-locked).
        Used where the feature matrix is generated directly; `features` defaults to all genes."""
        obj = cls.__new__(cls)
        G, N = counts_csc.shape
        obj.expdata = counts_csc
        obj.genes_all = list(genes) if genes is not None else [f"g{i + 1}" for i in range(G)]
        obj.cells_all = list(cells) if cells is not None else [f"c{i + 1}" for i in range(N)]
        obj.cell_names = obj.cells_all
        obj.counts = np.asarray(counts_csc.sum(axis=0)).ravel()
        obj.fcounts = counts_csc
        obj.ndata = None  # expdata / counts: built lazily by getfdata when a host consumer asks for it
        obj.genes = obj.genes_all
        obj.dimRed, obj.background, obj.filterpar, obj.clusterpar, obj.out = {}, {}, {}, {}, {}
        obj.cluster = {"features": list(features) if features is not None else obj.genes_all}
        obj.distances, obj.cpart, obj.medoids, obj.fcol = None, None, [], []
        return obj

    def __post_init__(self):
        # expdata may be dense or a scipy.sparse matrix: upstream it is a dgCMatrix (R/RaceID.R:72-83,
        # Matrix(as.matrix(expdata), sparse=TRUE)); a sparse input stays sparse all the way to the device
        if _is_sparse(self.expdata):
            x = self.expdata.tocsc()
            if x.dtype != np.float64:
                x = x.astype(np.float64)
            vals = x.data
        else:
            x = np.asarray(self.expdata, dtype=np.float64)
            vals = x
        # validity (R/RaceID.R:54-69)
        if x.ndim != 2 or x.shape[0] < 2:
            raise ValueError("input data must have more than one row")
        if x.shape[1] < 2:
            raise ValueError("input data must have more than one column")
        if vals.size and np.isnan(vals.min()):
            raise ValueError("NAs are not allowed in input data")
        if vals.size and vals.min() < 0:
            raise ValueError("negative values are not allowed in input data")
        self.expdata = x
        if not self.genes_all:
            self.genes_all = [f"g{i + 1}" for i in range(x.shape[0])]
        if not self.cells_all:
            self.cells_all = [f"c{i + 1}" for i in range(x.shape[1])]


def _is_sparse(x) -> bool:
    return hasattr(x, "tocsc") and hasattr(x, "indptr") or (hasattr(x, "tocsc") and hasattr(x, "nnz"))


def _csc_div_cols(m, colvals):
    """expdata / counts column-wise on a CSC matrix (true divisions, bit-identical to the dense `t(t(x) / counts)`)."""
    from scipy.sparse import csc_matrix

    return csc_matrix((m.data / np.repeat(colvals, np.diff(m.indptr)), m.indices, m.indptr), shape=m.shape)


def _fitbackground(x: np.ndarray, names: list, mthr: float = -1.0):
    """fitbackground (R/RaceID_utils.R:150-175): quadratic fit of log2 variance on log2 mean; genes whose variance
    exceeds the fit plus its standard errors are the features."""
    m = x.mean(axis=1)
    v = x.var(axis=1, ddof=1)
    with np.errstate(divide="ignore"):
        ml, vl = np.log2(m), np.log2(v)
    f = np.isfinite(ml) & np.isfinite(vl) & (ml > -np.inf) & (vl > -np.inf)
    ml_f, vl_f = ml[f], vl[f]
    mm = -8.0
    while True:
        X = np.stack([np.ones_like(ml_f), ml_f, ml_f ** 2], axis=1)
        coef, *_ = np.linalg.lstsq(X, vl_f, rcond=None)
        if coef[2] >= 0 or mm >= mthr:
            break
        mm += 0.5
        keep = ml_f > mm
        ml_f, vl_f = ml_f[keep], vl_f[keep]
    resid = vl_f - X @ coef
    dof = max(len(vl_f) - 3, 1)
    sigma2 = float(resid @ resid) / dof
    err = np.sqrt(np.diag(sigma2 * np.linalg.inv(X.T @ X)))
    with np.errstate(divide="ignore", invalid="ignore"):
        lm = np.log2(m)
        uvar = 2.0 ** (coef[0] + err[0] + lm * (coef[1] + err[1]) + (coef[2] + err[2]) * lm ** 2)
        vln = np.log2(v) - np.log2(uvar)
    sel = np.nan_to_num(vln, nan=-1.0) > 0
    return {"coef": coef, "err": err}, [n for n, s in zip(names, sel) if s]


def filterdata(object: SCseq, mintotal=3000, minexpr=5, minnumber=5, LBatch=None, knn=10, CGenes=None, FGenes=None,
               ccor=0.4, bmode="RaceID", verbose=True) -> SCseq:
    """filterdata (R/RaceID.R:109-242), default arm (no LBatch / CGenes).  Host-side: it produces the hot path's
    input and is O(G*N)."""
    if not isinstance(mintotal, (int, float)) or mintotal <= 0:
        raise ValueError("mintotal has to be a positive number")
    if not isinstance(minexpr, (int, float)) or minexpr < 0:
        raise ValueError("minexpr has to be a non-negative number")
    if not isinstance(minnumber, (int, float)) or round(minnumber) != minnumber or minnumber < 0:
        raise ValueError("minnumber has to be a non-negative integer number")
    if bmode not in ("RaceID", "mnnCorrect"):
        raise ValueError("bmode has to be one of RaceID, mnnCorrect")
    if LBatch is not None or CGenes is not None:
        raise NotImplementedError("batch-effect and CGenes arms of filterdata are outside the accelerated hot path")
    object.dimRed = {}
    sparse = _is_sparse(object.expdata)
    counts = np.asarray(object.expdata.sum(axis=0)).ravel()
    f = counts >= mintotal
    object.counts = counts[f]
    object.cell_names = [c for c, k in zip(object.cells_all, f) if k]
    sub = object.expdata[:, f] if not f.all() else object.expdata
    if sparse:
        sub = sub.tocsc()
        if minexpr > 0:
            g = np.asarray((sub >= minexpr).sum(axis=1)).ravel() >= minnumber
        else:
            g = np.full(sub.shape[0], sub.shape[1] >= minnumber)
        object.fcounts = sub                                  # raw counts of the kept cells: the device path's input
        object.ndata = _csc_div_cols(sub, counts[f])
        sub = None
    else:
        g = (sub >= minexpr).sum(axis=1) >= minnumber
        object.ndata = sub / counts[f][None, :]
    genes = [n for n, k in zip(object.genes_all, g) if k]
    if FGenes:
        genes = [n for n in genes if n not in set(FGenes)]
    object.genes = genes
    object.filterpar = dict(mintotal=mintotal, minexpr=minexpr, minnumber=minnumber, CGenes=CGenes, FGenes=FGenes,
                            BGenes=None, bmode=bmode)
    gi = _gene_index(object, genes)
    fit, _ = _fitbackground(sub[gi, :] if sub is not None else object.fcounts[gi, :].toarray(), genes)
    object.background["vfit"] = fit
    _, feats = _fitbackground(getfdata(object), genes)
    object.cluster["features"] = feats
    return object


def _gene_index(object: SCseq, names) -> np.ndarray:
    pos = {n: i for i, n in enumerate(object.genes_all)}
    return np.asarray([pos[n] for n in names], dtype=np.int64)


def getfdata(object: SCseq, g=None, n=None) -> np.ndarray:
    """getfdata (R/RaceID.R:723-728): (ndata * min(counts[n]))[fgenes, n] + 0.1, genes x cells."""
    if object.ndata is None and object.fcounts is not None:
        object.ndata = _csc_div_cols(object.fcounts, object.counts)
    if g is None:
        fgenes = object.genes
    else:
        gs = set(g)
        fgenes = [x for x in object.genes_all if x in gs]  # rownames(ndata) order
    if n is None:
        cols = np.arange(len(object.cell_names))
    else:
        ns = set(n)
        cols = np.asarray([i for i, c in enumerate(object.cell_names) if c in ns], dtype=np.int64)
    gi = _gene_index(object, fgenes)
    if _is_sparse(object.ndata):
        return object.ndata[gi, :][:, cols].toarray() * object.counts[cols].min() + 0.1
    return object.ndata[np.ix_(gi, cols)] * object.counts[cols].min() + 0.1


_METRICS = ("spearman", "pearson", "logpearson", "euclidean", "kendall")


def compdist(object: SCseq, metric="pearson", FSelect=True, knn=None, alpha=1, no_cores=1) -> SCseq:
    """compdist (R/RaceID.R:752-842), knn=NULL branch.  `object.distances` becomes a device-resident DistMatrix."""
    if metric not in _METRICS:
        raise ValueError("metric has to be one of the following: spearman, pearson, logpearson, euclidean, kendall")
    if not isinstance(FSelect, (bool, int, float)):
        raise ValueError("FSelect has to be logical (TRUE/FALSE)")
    if knn is not None and not isinstance(knn, (int, float)):
        raise ValueError("knn has to be NULL or integer > 0")
    if knn is not None:
        raise NotImplementedError("the knn-imputation branch (R/RaceID.R:773-836) is outside the accelerated hot path")
    if metric == "kendall":
        raise NotImplementedError("kendall is not part of the accelerated hot path")
    n = object.cluster.get("features") if FSelect else object.genes
    x = None
    if object.dimRed.get("x") is not None:
        x = np.asarray(object.dimRed["x"], dtype=np.float64)
        if metric == "logpearson" and x.min() <= 0:
            metric = "pearson"
    elif object.fcounts is None:
        x = getfdata(object, g=n)
    old = object.distances
    if isinstance(old, _lib.DistMatrix):
        old.free()
    if x is not None:
        dm = get_context().dist(x, metric)
    else:
        # sparse object: the dgCMatrix triplet goes to the device as it is and getfdata (R/RaceID.R:723-728) is fused
        # into the prep kernel's producer (rid_dist_csc) — no dense host matrix, no host transposes
        fc = object.fcounts
        if n is object.genes_all or len(n) == len(object.genes_all):
            frows = np.arange(len(object.genes_all), dtype=np.int32)
        else:
            ns = set(n)
            frows = np.asarray([i for i, gname in enumerate(object.genes_all) if gname in ns], dtype=np.int32)
        dm = get_context().dist_csc(fc.indices, fc.indptr, fc.data, fc.shape[0], frows, object.counts, metric)
    if dm.n_zero_var:
        import warnings

        warnings.warn("the standard deviation is zero")
    object.distances = dm
    object.clusterpar["metric"] = metric
    object.clusterpar["FSelect"] = FSelect
    object.cluster["features"] = n
    return object


def _as_dmat(d) -> _lib.DistMatrix:
    if isinstance(d, _lib.DistMatrix):
        return d
    return get_context().from_matrix(np.asarray(d, dtype=np.float64))


def saturation_cln(logW: np.ndarray) -> int:
    """clustfun's saturation rule (R/RaceID_utils.R:116-124)."""
    g = np.asarray(logW, dtype=np.float64)
    y = g[:-1] - g[1:]
    L = len(y)
    hit = []
    for i in range(L):
        tail = y[i:]
        mm = tail.mean()
        nn = math.sqrt(tail.var(ddof=1)) if len(tail) > 1 else float("nan")
        v = y[i] - (mm + nn)
        if not math.isnan(v) and v < 0:
            hit.append(i + 1)
    if not hit:
        raise ValueError("saturation rule found no cluster number (min(which(...)) is empty)")
    return max(min(hit), 1)


def bootstrap_samples(n: int, bootnr: int, rseed: int, samp=None, packed: bool = False):
    """The resamples fpc::clusterboot draws after set.seed(seed): sample(n,n,replace=TRUE) then unique() in
    first-occurrence order ("boot"), or sample(n, subtuning, replace=FALSE) ("subset").  0-based.
    Drawn by the library's host-only C restatement of R's generator (rid_r_bootstrap_samples; 13 M Mersenne-Twister
    words at 100k cells x 50 resamples); `bootstrap_samples_py` is the Python mirror it is checked against.
    packed=True returns the (concatenated, lengths) pair DistMatrix.clusterboot takes without re-packing."""
    cat, lens = _lib.r_bootstrap_samples(int(n), int(bootnr), int(rseed), 0 if samp is None else int(samp))
    if packed:
        return cat, lens
    return np.split(cat, np.cumsum(lens)[:-1]) if bootnr else []


def bootstrap_samples_py(n: int, bootnr: int, rseed: int, samp=None):
    """bootstrap_samples through the Python mirror of R's generator (rrng.py)."""
    rng = RRng(rseed)
    out = []
    for _ in range(bootnr):
        if samp is None:
            out.append((r_unique(rng.sample_int(n, n, replace=True)) - 1).astype(np.int32))
        else:
            out.append((rng.sample_int(n, min(n, int(samp)), replace=False) - 1).astype(np.int32))
    return out


def clustexp(object: SCseq, sat=True, samp=None, cln=None, clustnr=30, bootnr=50, rseed=17000, FUNcluster="kmedoids",
             verbose=True) -> SCseq:
    """clustexp (R/RaceID.R:894-912) + clustfun's kmedoids arms (R/RaceID_utils.R:102-148)."""
    if not isinstance(clustnr, (int, float)) or isinstance(clustnr, bool):
        raise ValueError("clustnr has to be a positive integer")
    if round(clustnr) != clustnr or clustnr <= 0:
        raise ValueError("clustnr has to be a positive integer")
    if not isinstance(bootnr, (int, float)) or isinstance(bootnr, bool):
        raise ValueError("bootnr has to be a positive integer")
    if round(bootnr) != bootnr or bootnr <= 0:
        raise ValueError("bootnr has to be a positive integer")
    if not isinstance(sat, (bool, int, float)):
        raise ValueError("sat has to be logical (TRUE/FALSE)")
    if cln is not None:
        if not isinstance(cln, (int, float)) or isinstance(cln, bool):
            raise ValueError("cln has to be a non-negative integer")
        if round(cln) != cln or cln < 0:
            raise ValueError("cln has to be a non-negative integer")
    if not isinstance(rseed, (int, float)):
        raise ValueError("rseed has to be numeric")
    if not sat and cln is None:
        raise ValueError("cln has to be a positive integer or sat has to be TRUE")
    if samp is not None:
        if not isinstance(samp, (int, float)):
            raise ValueError("samp needs to be an integer >1")
        if samp <= 1:
            raise ValueError("samp needs to be an integer >1")
    if FUNcluster not in ("kmedoids", "kmeans", "hclust"):
        raise ValueError("FUNcluster needs to be one of kmedoids, kmeans, hclust")
    if object.distances is None:
        raise ValueError("Run compdist before clustexp")
    if FUNcluster != "kmedoids":
        raise NotImplementedError("only FUNcluster='kmedoids' is part of the accelerated hot path")
    if cln is None:
        sat = True
    clustnr, bootnr = int(clustnr), int(bootnr)
    object.clusterpar = dict(clustnr=clustnr, bootnr=bootnr, samp=samp, metric=object.clusterpar.get("metric"), sat=sat,
                             cln=cln, rseed=rseed, FSelect=object.clusterpar.get("FSelect"), FUNcluster=FUNcluster)
    y = clustfun(object.distances, clustnr, bootnr, samp, sat, cln, rseed, FUNcluster, verbose)
    clb = y["clb"]
    # (the single-cluster result carries no subsetmean: R's y$clb$subsetmean is NULL there, R/RaceID_utils.R:126-130)
    object.cluster = dict(kpart=clb["result"]["partition"],
                          jaccard=clb.get("bootmean") if samp is None else clb.get("subsetmean"), gap=y["gpr"], clb=clb,
                          features=object.cluster.get("features"))
    k = int(np.max(clb["result"]["partition"]))
    object.fcol = list(RRng(111111).sample_int(k) - 1)  # set.seed(111111); sample(rainbow(k)) as a permutation
    return object


def clustfun(diM, clustnr=20, bootnr=50, samp=None, sat=True, cln=None, rseed=17000, FUNcluster="kmedoids",
             verbose=True) -> dict:
    """clustfun (R/RaceID_utils.R:102-148), kmedoids arms."""
    if clustnr < 2:
        raise ValueError("Choose clustnr > 1")
    dm = _as_dmat(diM)
    N = dm.N
    if cln is None:
        cln = 0
    cln = int(cln)
    if N - 1 < clustnr:
        clustnr = N - 1
    gpr = None
    if not (sat or cln > 0):
        return {"clb": None, "gpr": None}
    f = cln == 0
    # the resamples depend on (N, bootnr, rseed, samp) only: draw them on a helper thread (the C helper releases the GIL)
    # while the GPU runs the sweep
    import threading

    box = {}

    def _draw():
        try:
            box["samples"] = bootstrap_samples(N, bootnr, int(rseed), samp, packed=True)
        except BaseException as e:  # surfaced by the join below
            box["error"] = e

    drawer = threading.Thread(target=_draw, daemon=True)
    drawer.start()
    if sat:
        rng = RRng(int(rseed))
        sub = None
        if samp is not None:
            sub = (rng.sample_int(N, min(int(samp), N), replace=False) - 1).astype(np.int32)
        if verbose:
            print(f"Clustering k = 1,2,..., K.max (= {clustnr}): .. ")
        logW = dm.gap_sweep(clustnr, subset=sub)
        if verbose:
            print("done.")
        nanv = np.full(clustnr, np.nan)
        gpr = {"Tab": np.stack([logW, nanv, nanv, nanv], axis=1), "n": N if sub is None else len(sub), "B": 100}
        if f:
            cln = saturation_cln(logW)
    if cln <= 1:
        clb = {"result": {"partition": np.ones(N, dtype=np.int32)}, "bootmean": np.array([1.0])}
        return {"clb": clb, "gpr": gpr}
    drawer.join()
    if "error" in box:
        raise box["error"]
    r = dm.clusterboot(cln, box["samples"])
    key = "boot" if samp is None else "subset"
    clb = {"result": {"partition": r["partition"], "nc": cln,
                      "result": {"pamobject": {"id.med": r["id_med"] + 1, "clustering": r["partition"]}, "nc": cln},
                      "clusterlist": [r["partition"] == (j + 1) for j in range(cln)]},
           "partition": r["partition"], "nc": cln, "B": bootnr, "bootmethod": key,
           f"{key}result": r["bootresult"], f"{key}mean": r["bootmean"],
           f"{key}brd": (r["bootresult"] <= 0.5).sum(axis=1), f"{key}recover": (r["bootresult"] >= 0.75).sum(axis=1)}
    if samp is not None:
        clb["bootmean"] = None
    return {"clb": clb, "gpr": gpr}


def compmedoids(object: SCseq, part) -> list:
    """compmedoids (R/RaceID.R:1298-1316), distance-matrix arm.  part: cluster numbers aligned with the cells of
    `ndata` (0 / missing = cell not in part).  Returns medoid cell names ordered by increasing cluster number."""
    if object.distances is None:
        raise NotImplementedError("the dimRed fallback of compmedoids (R/RaceID.R:1310) is outside the hot path")
    dm = _as_dmat(object.distances)
    idx, _ = dm.medoids(np.asarray(part, dtype=np.int32))
    names = object.cell_names if object.cell_names else [str(i) for i in range(dm.N)]
    return [names[i] for i in idx]


def silhouette(object: SCseq, final: bool = False) -> np.ndarray:
    """The data behind plotsilhouette (R/RaceID.R:1270-1285): `silhouette(kpart, as.dist(object@distances))`, an n x 3
    matrix with columns cluster, neighbor, sil_width (cells in `ndata` order).  final=TRUE uses `cpart`."""
    if object.distances is None:
        raise ValueError("run compdist before plotsilhouette")
    part = object.cpart if final else (object.cluster or {}).get("kpart")
    if part is None or len(part) == 0:
        raise ValueError("run findoutliers before plotsilhouette" if final else "run clustexp before plotsilhouette")
    part = np.asarray(part, dtype=np.int32)
    if part.max() < 2:
        raise ValueError("only a single cluster: no silhouette plot")
    nb, sw = _as_dmat(object.distances).silhouette(part)
    return np.column_stack([part.astype(np.float64), nb.astype(np.float64), sw])


def tsne_initial_config(object: SCseq, k: int = 2) -> np.ndarray:
    """`cmdscale(as.dist(object@distances), k = 2)`, the `initial_config` comptsne hands to Rtsne when
    `initial_cmd = TRUE` (R/RaceID.R:1086-1093).  Rtsne itself is outside this package's scope."""
    if object.distances is None:
        raise ValueError("run compdist before comptsne (the dimRed arm of comptsne does not use a distance matrix)")
    return _as_dmat(object.distances).cmdscale(k)


def outlier_distance_sample(object: SCseq, out=()) -> np.ndarray:
    """`clp2p.cl` of findoutliers (R/RaceID.R:986-994): after `set.seed(clusterpar$rseed)`, for every cluster of `kpart`
    `tcol <- sample(cells of the cluster, min(500, size))`, and the distances among the sampled cells that are not
    outliers, `as.vector(t(object@distances[tcol', tcol']))`, are appended when more than one cell is left; zeros are
    dropped.  The blocks are read from the device-resident matrix by cell NAME (rid_dmat_sub): D never leaves HBM."""
    kpart = np.asarray(object.cluster["kpart"], dtype=np.int64)
    names = object.cell_names if object.cell_names else [str(i) for i in range(len(kpart))]
    dm = _as_dmat(object.distances)
    pos = {c: i for i, c in enumerate(names)}
    outset = set(out)
    rng = RRng(int(object.clusterpar.get("rseed", 17000)))
    chunks = []
    for i in range(1, int(kpart.max()) + 1):
        members = np.nonzero(kpart == i)[0]
        if len(members) == 0:
            # sample(character(0), 0) draws nothing
            continue
        pick = rng.sample_int(len(members), min(500, len(members)), replace=False) - 1
        tcol = [names[members[j]] for j in pick]
        keep = [c for c in tcol if c not in outset]
        if len(keep) > 1:
            idx = np.asarray([pos[c] for c in keep], dtype=np.int32)
            chunks.append(dm.sub(idx, idx).reshape(-1))  # as.vector(t(M)): row-major
    v = np.concatenate(chunks) if chunks else np.zeros(0)
    return v[v > 0]


def findoutliers_merge(object: SCseq, out, outdistquant=.95, verbose=False) -> SCseq:
    """The distance-consuming half of findoutliers (R/RaceID.R:983-1060): `out` are the outlier cell names the negative
    binomial model produced (R/RaceID.R:943-981; a statistical model outside the accelerated path).  Outliers are merged
    into new clusters by their mutual distances `object@distances[out, out]` against the `outdistquant` quantile of the
    within-cluster distance sample, then the final partition is refreshed (medoids by mean within-cluster distance,
    every cell to its nearest medoid).  Both distance reads are sub-block gathers by cell name on the device."""
    if not isinstance(outdistquant, (int, float)) or isinstance(outdistquant, bool):
        raise ValueError("outdistquant has to be a number between 0 and 1")
    if outdistquant < 0 or outdistquant > 1:
        raise ValueError("outdistquant has to be a number between 0 and 1")
    if object.cluster.get("kpart") is None or len(object.cluster["kpart"]) == 0:
        raise ValueError("run clustexp before findoutliers")
    kpart = np.asarray(object.cluster["kpart"], dtype=np.int64)
    names = object.cell_names if object.cell_names else [str(i) for i in range(len(kpart))]
    pos = {c: i for i, c in enumerate(names)}
    out = list(out)
    for c in out:
        if c not in pos:
            raise ValueError(f"subscript out of bounds: unknown cell {c!r}")
    dm = _as_dmat(object.distances)
    clp = outlier_distance_sample(object, out)
    cpart = kpart.copy()
    cadd = []
    if len(out) == 1:
        cadd = [out]
    elif len(out) > 1:
        if len(clp) == 0:
            raise ValueError("missing value where TRUE/FALSE needed")  # quantile() / min() of an empty sample in R
        q = float(np.quantile(clp, outdistquant))  # type 7, R's default
        cmin = float(clp.min())
        n = list(out)
        idx = np.asarray([pos[c] for c in n], dtype=np.int32)
        m = dm.sub(idx, idx)
        for _ in range(len(out)):
            if len(n) > 1:
                mm = m.copy()
                np.fill_diagonal(mm, np.inf)  # x[1:(length(x)-1)][-x[length(x)]]: the row without its own column
                o = np.argsort(mm.min(axis=1), kind="stable")
                m = m[np.ix_(o, o)]
                n = [n[j] for j in o]
                f = (m[:, 0] < q) | (m[:, 0] == cmin)
                fi = np.nonzero(f)[0]
                ind = np.asarray([0]) if len(fi) else np.zeros(0, dtype=np.int64)
                if len(fi) > 1:
                    ind = np.nonzero((m[np.ix_(fi, fi)] <= q).sum(axis=1) > len(fi) / 2)[0]
                if len(fi) == 0:
                    # n[f][1] is NA in R; nothing can be merged
                    raise ValueError("no outlier lies below the distance quantile (n[f][ind] is NA in R)")
                grp = [n[fi[j]] for j in ind]
                cadd.append(grp)
                g = np.asarray([c not in set(grp) for c in n])
                n = [c for c, k in zip(n, g) if k]
                m = m[np.ix_(g, g)]
                if not g.any():
                    break
            elif len(n) == 1:
                cadd.append(n)
                break
    for grp in cadd:
        sel = np.asarray([pos[c] for c in grp], dtype=np.int64)
        new = int(cpart.max()) + 1
        if len(sel):
            cpart[sel] = new
    object.out = {"out": out, "cadd": cadd, "clp2p.cl": clp}
    return refresh_partition(object, cpart)


def refresh_partition(object: SCseq, cpart):
    """The final-cluster block of findoutliers (R/RaceID.R:1034-1060): per-cluster medoid by mean within-cluster
    distance, every cell to its nearest medoid (first minimum), empty clusters compacted, medoids recomputed."""
    dm = _as_dmat(object.distances)
    md, _ = dm.medoids(np.asarray(cpart, dtype=np.int32))
    new = dm.refresh(md)
    object.cpart = new
    object.fcol = list(RRng(111111).sample_int(int(new.max())) - 1)
    object.medoids = compmedoids(object, new)
    return object
