/*
 * raceid_b200.h — C ABI of libraceid_b200.so: the B200-native replacement for RaceID's clustering hot path.
 *
 * This is the drop-in boundary.  The reference reaches native code through Rcpp-generated
 * `.Call` stubs (reference src/RcppExports.cpp:15-24, registration table :154-172, R stubs
 * R/RcppExports.R:4-46); the functions below are what an equivalent `.Call` shim binds when the bodies of
 * `dist.gen` (R/RaceID_utils.R:9-15), `clustfun` (R/RaceID_utils.R:102-148) and `compmedoids`
 * (R/RaceID.R:1298-1316) are replaced.  INTEGRATION.md shows the Rcpp stub for each entry point.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, <0 on error (rid_last_error() gives the text,
 *     the R stub turns it into Rcpp::stop => an R condition, like BEGIN_RCPP/END_RCPP does upstream);
 *   - inputs are borrowed and read-only; outputs are caller-allocated host buffers;
 *   - matrices follow R: column-major; cell/medoid indices crossing this ABI are 0-based, cluster labels 1-based;
 *   - the library draws no random numbers: bootstrap index vectors come from the caller (R's sample()).
 *   - there is NO CPU fallback: every compute entry point fails if no sm_100 device is usable.
 */
#ifndef RACEID_B200_H
#define RACEID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rid_ctx rid_ctx;   /* one GPU (one rank of a row-sharded job)            */
typedef struct rid_dmat rid_dmat; /* device-resident distance matrix, row-block sharded */

enum { RID_PEARSON = 0, RID_SPEARMAN = 1, RID_LOGPEARSON = 2, RID_EUCLIDEAN = 3 };

enum {
    RID_OK = 0,
    RID_ERR_ARG = -1,     /* bad argument (k outside {1..n-1}, unknown metric, ...)                     */
    RID_ERR_CUDA = -2,    /* CUDA / NCCL runtime failure, or no usable sm_100 device                    */
    RID_ERR_NA = -3,      /* NAs in the dissimilarity matrix (pam() refuses them upstream)               */
    RID_ERR_NOMEM = -4,
    RID_ERR_UNSUPPORTED = -5,
    RID_ERR_INTERRUPTED = -6 /* the caller asked for cancellation (rid_ctx_request_abort / the poll callback)   */
};

const char *rid_last_error(void);
int rid_version(void);

/* ---- contexts -------------------------------------------------------------------------------------- */
/* One context per GPU.  world==1: single-GPU.  world>1: this process is rank `rank` of a job that shards the
 * distance matrix by row blocks; `nccl_uid` is the 128-byte id made by rid_nccl_unique_id()
 * on rank 0 and shipped to the other ranks by the caller (torch.distributed store, MPI, a file ...). */
int rid_nccl_unique_id(void *uid128);
int rid_ctx_create(int device, int rank, int world, const void *nccl_uid, rid_ctx **out);
int rid_ctx_destroy(rid_ctx *ctx);
int rid_ctx_sync(rid_ctx *ctx);
/* Cancellation (upstream cluster::pam polls R_CheckUserInterrupt once per candidate row).
 * rid_ctx_request_abort may be called from any thread or from a signal handler while a compute call is running on
 * `ctx`; rid_ctx_set_interrupt_poll installs a callback the library runs ON THE CALLING THREAD between enqueued waves
 * of GPU work (per k of the sweep, per SWAP batch, per wave of bootstrap replicates) — the R stub passes a function
 * that wraps R_CheckUserInterrupt in R_ToplevelExec and returns non-zero when the user pressed Ctrl-C.  Either way the
 * running call drains the GPU, releases its work buffers and returns RID_ERR_INTERRUPTED; the context and the
 * distance matrix stay usable.  Multi-rank jobs agree on the answer (one small all-reduce per poll point). */
int rid_ctx_request_abort(rid_ctx *ctx);
int rid_ctx_clear_abort(rid_ctx *ctx);
int rid_ctx_set_interrupt_poll(rid_ctx *ctx, int (*poll)(void *), void *arg);
/* Milliseconds the GPU spent in the last compute call of this ctx, from CUDA events on its stream. */
double rid_ctx_last_ms(const rid_ctx *ctx);
/* Number of kernels this ctx has launched since creation (bench.py's gpu_launches). */
long long rid_ctx_launches(const rid_ctx *ctx);
/* Accepted SWAP exchanges summed over every PAM run this ctx has finished (c1, sweep, bootstrap replicates). */
long long rid_ctx_swaps(const rid_ctx *ctx);
/* Diagnostics (bench.py): 0 swaps, 1 bootstrap replicates redone (more SWAP iterations needed than were enqueued ahead),
 * 2-4 milliseconds of the last rid_clusterboot spent in c1 / in the batched first picks + waiting for the replica
 * pulls / in the replicates, 5 bytes the last replica pull moved over NVLink, 6 look-ahead iterations enqueued per
 * replicate, 7 host milliseconds spent inside the enqueue calls of the replicates, 8 milliseconds the last replica pull
 * took on the copy stream, 9 rid_clusterboot calls that took c1 from the sweep, 10 classes of the common refinement of
 * the last sweep's partitions (the "atoms" its W.k pass grouped the rows by; -1: too many, batched passes ran). */
double rid_ctx_stat(const rid_ctx *ctx, int which);
/* Give the cached device buffers back to the driver (collective when world > 1: call it on every rank). */
int rid_ctx_trim(rid_ctx *ctx);

/* The row block [row0,row1) of an N x N matrix that rank `rank` of `world` owns (pure host function). */
int rid_shard_rows(int N, int rank, int world, int *row0, int *row1);

/* Measurement hooks (bench.py): per-kernel CUDA-event timing on the library's own stream.
 * kind: 0 Gram/DMMA distance kernel (work = flops), 1 SWAP scan, 2 BUILD scan, 3 grouped column sums (W.k, medoids,
 * silhouette), 4 prep, 5 subset compaction, 6 int8 tcgen05 distance kernel (work = int8 ops), 7 incremental SWAP pass (work = algorithmic bytes; traffic = bytes requested from memory). */
int rid_ctx_profile(rid_ctx *ctx, int kind_mask /* bit k = record kind k; -1 = all; 0 = stop */);
int rid_ctx_profile_get(rid_ctx *ctx, int kind, long long *launches, double *ms, double *work, double *traffic);
/* Ceilings measured on the spot: issue rate of the FP64 tensor pipe (DMMA m8n8k4), TFLOP/s, and a read-only
 * 128-bit streaming loop over `gib` GiB of HBM, GB/s (MEASURED_PEAKS.json holds a copy figure and a bf16 figure only). */
int rid_measure_dmma_peak(rid_ctx *ctx, double *tflops);
int rid_measure_read_peak(rid_ctx *ctx, double gib, double *gbs);
/* Issue-rate ceiling of the int8 tensor pipe (tcgen05.mma kind::i8, M=128 N=192 K=32 from shared memory), TOP/s. */
int rid_measure_i8_peak(rid_ctx *ctx, double *tops);
/* user events on the library's stream, slots 0..7 */
int rid_ctx_event_record(rid_ctx *ctx, int slot);
int rid_ctx_event_elapsed_ms(rid_ctx *ctx, int slot_a, int slot_b, double *ms);

/* ---- distance matrix: replaces dist.gen (R/RaceID_utils.R:9-15) as called at R/RaceID.R:769 ---------- */
/* x: G x N column-major doubles (column c = the G feature values of cell c), i.e. what cor()/dist() see.
 * x_on_device != 0: x is a device pointer on ctx's GPU (benchmarks with resident inputs).
 * Every rank passes the full x; rank r keeps rows [row0,row1) of the N x N result.
 * n_zero_var: number of zero-variance cells (cor() returns NA for them and warns; their rows/cols are NA). */
int rid_dist(rid_ctx *ctx, const double *x, int G, int N, int metric, int x_on_device, rid_dmat **out,
             int *n_zero_var);
/* Same, fused with getfdata (R/RaceID.R:723-728, "next" row of SURVEY §8f-2): the filtered count matrix in the package's
 * native dgCMatrix storage (CSC: row_idx/col_ptr/val over the kept cells, n_rows_all genes), counts[c] = the cell's
 * total, feature_rows[g] = matrix row of feature g.  x = (val / counts[c]) * min(counts) + 0.1 is built on the device. */
int rid_dist_csc(rid_ctx *ctx, const int *row_idx, const int *col_ptr, const double *val, int n_rows_all, int N,
                 const int *feature_rows, int G, const double *counts, int metric, rid_dmat **out, int *n_zero_var);
/* Adopt a distance matrix computed elsewhere (clustexp() only needs @distances, R/RaceID.R:903,907).
 * D: N x N column-major.  Like as.dist() the LOWER triangle is authoritative and is mirrored. */
int rid_dmat_from_host(rid_ctx *ctx, const double *D, int N, rid_dmat **out);
int rid_dmat_free(rid_dmat *dm);
/* N, owned row range, the value of one quantisation step (distances are stored as 32-bit fixed point,
 * see DESIGN.md), 1 if the matrix holds NAs. */
int rid_dmat_info(const rid_dmat *dm, int *N, int *row0, int *row1, double *step, int *has_na);
/* as.matrix() on demand: out is m x n column-major, rows ridx (NULL = 0..m-1), columns cidx (NULL = 0..n-1).
 * Collective when world>1 (every rank calls with the same arguments; every rank receives the block).
 * Values are q * step (what the clustering kernels see), within step / 2 <= 1e-6 of R's doubles.  A euclidean matrix
 * whose range makes step larger than 2e-6 (largest distance >= 4096) keeps its feature rows, and this call then
 * returns sqrt(sum (x_i - x_j)^2) recomputed in FP64 like stats::dist, so the 1e-6 budget holds for any range. */
int rid_dmat_sub(rid_dmat *dm, const int *ridx, int m, const int *cidx, int n, double *out);

/* ---- cluster::pam(as.dist(x[subset,subset]), k) as called at R/RaceID_utils.R:113,221 ------------------ */
/* subset: m cell indices (order significant: position = object index inside pam); NULL = all cells.
 * medoids[k]: id.med, as POSITIONS into subset, in cluster-number order.  labels[m]: `clustering`.
 * objective[2]: build, swap.  avg_sil: silinfo$avg.width (NULL = skip).  n_swaps: accepted swaps (nullable). */
int rid_pam(rid_dmat *dm, const int *subset, int m, int k, int *medoids, int *labels, double *objective,
            double *avg_sil, int *n_swaps);

/* ---- clusGapExt(..., FUNcluster=pam, random=FALSE, diss=TRUE) k-loop, R/RaceID_utils.R:58-61 ----------- */
/* logW[Kmax].  With subset == NULL the PAM(k) results of k = 2..Kmax stay with the (immutable) matrix: clustexp follows
 * the sweep with fpc::clusterboot, whose first step is the same pam(as.dist(D), cln) call (R/RaceID_utils.R:113,132), and
 * rid_clusterboot takes c1 from there instead of repeating it (identical by construction; RID_NO_MEMO=1 disables). */
int rid_gap_sweep(rid_dmat *dm, const int *subset, int m, int Kmax, double *logW);

/* ---- fpc::clusterboot(..., clustermethod=pamkdCBI, k) as called at R/RaceID_utils.R:132-133 ------------ */
/* samples: the B resamples concatenated (already unique()'d, first-occurrence order); sample_len[B].
 * partition[N], medoids[k] (cell indices of c1's id.med), bootresult k x B column-major. */
int rid_clusterboot(rid_dmat *dm, int k, const int *samples, const int *sample_len, int B, int *partition,
                    int *medoids, double *bootresult);

/* ---- compmedoids (R/RaceID.R:1298-1316) and the final-partition refresh (R/RaceID.R:1044-1054) -------- */
/* labels[N]: positive cluster numbers, 0 = cell not in `part`.  medoid_idx / cluster_ids sized >= #clusters. */
int rid_medoids(rid_dmat *dm, const int *labels, int N, int *medoid_idx, int *cluster_ids, int *n_clusters);
int rid_refresh(rid_dmat *dm, const int *medoid_idx, int k, int *labels);

/* ---- row-wise nearest neighbours: apply(D, 1, function(x) head(order(x), K)) — R/RaceID.R:774,1123,
 * R/VarID_functions.R:549, R/StemID.R:192 ("next" row of SURVEY §8f).  idx: K x N column-major (cell c's K neighbours
 * contiguous, 0-based, ties to the lower index like order()); dist (nullable): the matching distances. */
int rid_knn(rid_dmat *dm, int K, int *idx, double *dist);

/* ---- cluster::silhouette(part, dist) as plotsilhouette calls it (R/RaceID.R:1283-1284; "next" row of SURVEY §8f) --- */
/* labels[N]: positive cluster numbers, 0 = cell not in the partition.  neighbor[N]: the other cluster with the smallest
 * mean distance (first one in sorted cluster-id order on ties; 0 for unlabelled cells); width[N]: (b - a) / max(a, b),
 * 0 for singleton clusters, NaN for unlabelled cells.  Needs 2 <= #clusters <= #cells - 1. */
int rid_silhouette(rid_dmat *dm, const int *labels, int N, int *neighbor, double *width);

/* ---- stats::cmdscale(d, k): classical MDS, the t-SNE start configuration of comptsne (R/RaceID.R:1093; "next" row of
 * SURVEY §8f).  points: N x k column-major = E_k diag(sqrt(lambda_k)) of B = -1/2 J D^2 J for the k algebraically largest
 * eigenvalues; eig[k]: those eigenvalues (a non-positive one gives a NaN column, cmdscale warns and drops it).  The sign
 * of every axis is fixed so that its largest-magnitude coordinate is positive (LAPACK's is arbitrary).  tol: relative
 * residual |B v - lambda v| / |lambda| of the k pairs (<= 0: 1e-9); max_iter (<= 0: 1000); iters (nullable): used.
 * k <= 4 on this path (block of 8 vectors); NA in the matrix is an error like in cmdscale. */
int rid_cmdscale(rid_dmat *dm, int k, double tol, int max_iter, double *points, double *eig, int *iters);

/* ---- host-only helper: the resamples fpc::clusterboot draws after set.seed(seed) (R/RaceID_utils.R:132-133), from R's
 * default generator (Mersenne-Twister, "Rejection" sample.kind).  subset_size <= 0: bootmethod "boot" — sample(n, n,
 * replace=TRUE) then unique() in first-occurrence order; > 0: bootmethod "subset" — sample(n, min(n, subset_size)).
 * out_cat holds up to B * n 0-based indices (the resamples concatenated), out_len[B] their lengths.  No GPU involved;
 * an R host passes the vectors its own sample() drew instead. */
int rid_r_bootstrap_samples(int n, int B, unsigned int seed, int subset_size, int *out_cat, int *out_len);

/* ---- W.k (R/RaceID_utils.R:39-53) for a given partition; used by tests and by callers of clusGapExt ---- */
int rid_within_disp(rid_dmat *dm, const int *subset, int m, const int *labels, int k, double *W);

#ifdef __cplusplus
}
#endif
#endif /* RACEID_B200_H */
