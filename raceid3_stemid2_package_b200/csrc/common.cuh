// common.cuh — shared declarations for libraceid_b200.so (context, distance-matrix handle, error plumbing).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/raceid_b200.h"

typedef unsigned long long u64;
typedef uint32_t u32;

// NaN sentinel inside the fixed-point matrix (zero-variance cells: cor() gives NA).
#define RID_Q_NA 0xFFFFFFFFu

void rid_set_error(const char *fmt, ...);

#define RID_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            rid_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));   \
            return RID_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

#define RID_TRY(call)                \
    do {                             \
        int rc__ = (call);           \
        if (rc__ != RID_OK) return rc__; \
    } while (0)

// ---- NCCL, bound at run time (dlopen) so that single-GPU use needs no NCCL at all -------------------------
struct NcclApi;

struct rid_ctx {
    int device = 0;
    int rank = 0, world = 1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr; // copy-engine work that overlaps kernels on `stream` (peer pulls of matrix shards)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // cooperative cancellation (long calls must be interruptible): a flag any thread / signal handler may
    // set, and an optional poll callback run on the calling thread between enqueued waves of work
    volatile int abort_flag = 0;
    int (*poll_fn)(void *) = nullptr;
    void *poll_arg = nullptr;
    double last_ms = 0.0;
    long long launches = 0;
    long long swaps = 0; // accepted SWAP exchanges of every PAM run read back so far (bench.py's n_swaps)
    // rid_ctx_stat: [0] swaps (mirror), [1] bootstrap replicates redone because they needed more SWAP iterations than were
    // enqueued ahead, [2..4] ms of the last rid_clusterboot: c1 / wait for the replica pulls / replicates,
    // [5] bytes the last replica pull moved over NVLink, [6] look-ahead iterations the last rid_clusterboot enqueued
    double stat[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; // [8] ms the last replica pull took on the copy stream
    cudaEvent_t phase_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    void *comm = nullptr; // ncclComm_t
    NcclApi *nccl = nullptr;
    // optional per-kernel profiling (rid_ctx_profile): CUDA-event pairs around the hot kernels, on `stream`
    int prof_on = 0;
    struct ProfRec { int kind; cudaEvent_t a, b; double work; double traffic; };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> prof_pool;
    // finished records folded into per-kind totals (their events go back to the pool): a long timed region never has to
    // create events on the fly
    struct ProfAcc { long long n = 0; double ms = 0, work = 0, traffic = 0; };
    ProfAcc prof_acc[16];
    // device-side work counters of the passes whose row count is only known on the device (never read back while a
    // clustering call is in flight): [0] rows x subset size, [1] rows x pitch of the incremental BUILD passes,
    // [2], [3] the same for the incremental SWAP passes
    u64 *prof_dev = nullptr;
    cudaEvent_t user_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // freed large device buffers kept for reuse (a 40 GB cudaMalloc/cudaFree pair per compdist() costs ~0.3 s)
    struct PoolBuf { void *p; size_t bytes; };
    std::vector<PoolBuf> pool;
    // peer distance-matrix shards mapped through CUDA IPC (multi-rank symmetric Gram kernel)
    struct IpcMap { unsigned char handle[64]; void *ptr; };
    std::vector<IpcMap> ipc;
    // allocations of this rank whose IPC handle went out to the peers: a peer may hold them mapped, so they are never
    // given back to the driver (pool eviction / trim) before rid_ipc_release_all has closed the mappings everywhere
    std::vector<void *> exported;
    void *scratch = nullptr; // small device scratch (barriers, handle exchange)
    // tile list of the persistent distance kernel (device), cached per problem shape
    void *i8_tiles = nullptr;
    long long i8_tiles_key[6] = {-1, -1, -1, -1, -1, -1};
    long long i8_ntiles = 0;
    size_t i8_tiles_bytes = 0;
    // page-locked host staging (grow-only), e.g. the subset tables and result slots of the bootstrap pipeline
    void *pin = nullptr;
    size_t pin_bytes = 0;
    int live_sharded = 0; // multi-rank matrices alive (their peers' mappings must stay open)
};

// page-locked host memory owned by the context (contents are not preserved when it grows)
void *rid_pinned(rid_ctx *ctx, size_t bytes);

// Collect the device pointers of every rank's buffer (peers mapped with CUDA IPC).  peers[rank] = mine.
// `mine` may point inside a larger allocation: the offset from the allocation base travels with the handle.
// Returns RID_OK with *all_ok = 1 only if every rank could map every peer.  Collective.
int rid_peer_pointers(rid_ctx *ctx, void *mine, void **peers, int *all_ok);
// Collective: every rank closes its mappings of the peers' buffers, then (after a barrier) forgets which of its own
// buffers were exported, so that they may be freed again.
int rid_ipc_release_all(rid_ctx *ctx);
// Collective: *all_ok = 1 iff every rank passed ok != 0.
int rid_agree_all(rid_ctx *ctx, int ok, int *all_ok);
// dist.cu: start / finish pulling the other ranks' row blocks into dm->full (no-ops without a replica buffer)
int rid_replica_start(rid_dmat *dm);
int rid_replica_complete(rid_dmat *dm);
// RID_ERR_INTERRUPTED if the caller asked for cancellation (flag or poll callback); clears the request.  With
// collective = true the answer is agreed over all ranks (an all-reduce: call it at the same point on every rank).
int rid_check_abort(rid_ctx *ctx, bool collective);
int rid_barrier(rid_ctx *ctx);
int rid_allgather(rid_ctx *ctx, const void *send, void *recv, size_t bytes_per_rank);
int rid_bcast_ranges(rid_ctx *ctx, void *buf, const size_t *off /* [world + 1] byte offsets */);

// cudaMalloc/cudaFree through the context's reuse pool (exact-size match)
cudaError_t rid_pool_alloc(rid_ctx *ctx, void **p, size_t bytes);
void rid_pool_free(rid_ctx *ctx, void *p, size_t bytes);
void rid_pool_trim(rid_ctx *ctx);
// Short-lived device buffers of the clustering calls (label tables, contingency tables, ...) go through the same pool
// with their size rounded up to a power of two >= 256 KB, so that a steady-state call makes NO cudaMalloc / cudaFree:
// those go through the driver's resource manager, where they queue behind whatever else is talking to it (a monitoring
// process polling the GPUs is enough) -- measured as random 0.1-1.2 s stalls of single sweep / clusterboot calls.
static inline size_t rid_tmp_size(size_t bytes)
{
    size_t s = (size_t)256 << 10;
    while (s < bytes) s <<= 1;
    return s;
}
template <class T> static inline cudaError_t rid_tmp_alloc(rid_ctx *ctx, T **p, size_t bytes)
{
    void *q = nullptr;
    cudaError_t e = rid_pool_alloc(ctx, &q, rid_tmp_size(bytes));
    *p = e == cudaSuccess ? reinterpret_cast<T *>(q) : nullptr;
    return e;
}
static inline void rid_tmp_free(rid_ctx *ctx, void *p, size_t bytes) { if (p) rid_pool_free(ctx, p, rid_tmp_size(bytes)); }

// kinds for rid_ctx_profile_get
enum { RID_PROF_GRAM = 0, RID_PROF_SCAN_SWAP = 1, RID_PROF_SCAN_BUILD = 2, RID_PROF_SCAN_GROUP = 3, RID_PROF_PREP = 4,
       RID_PROF_COMPACT = 5, RID_PROF_GRAM_I8 = 6, RID_PROF_SCAN_DELTA = 7, RID_PROF_NKINDS = 8 };

// RAII-less helpers: call prof_begin right before the launch and prof_end right after it.
cudaEvent_t rid_prof_begin(rid_ctx *ctx, int kind);
void rid_prof_end(rid_ctx *ctx, cudaEvent_t a, int kind, double work, double traffic);

int rid_allreduce_u64(rid_ctx *ctx, u64 *buf, size_t count);
int rid_allreduce_i32(rid_ctx *ctx, int *buf, size_t count);
int rid_allreduce_f64(rid_ctx *ctx, double *buf, size_t count);

struct PamWs; // pam.cu

struct rid_dmat {
    rid_ctx *ctx = nullptr;
    int N = 0;
    int row0 = 0, row1 = 0; // owned rows [row0,row1)
    size_t ld = 0;          // pitch in elements, multiple of 128
    u32 *D = nullptr;       // (row1-row0) x ld, fixed point: value = q * step
    int bits = 28;          // q <= 2^bits  (28: correlation distances, 31: general)
    double step = 0.0;      // value of one quantisation step
    size_t alloc_rows = 0;  // rows allocated (== the shard size of rank 0 when world > 1, so that shards all-gather)
    int has_na = 0;
    std::vector<unsigned char> na_cell; // host copy: 1 if the cell's row/col is NA
    PamWs *ws = nullptr;
    // ---- multi-rank replica ----
    // When N x ld words fit next to the work buffers, every rank allocates the WHOLE matrix; `D` then points at the
    // rank's own row block inside `full`.  The other ranks' blocks are pulled over NVLink by the copy engines
    // (cudaMemcpy2DAsync from the peers' IPC-mapped buffers on ctx->copy_stream), overlapping the kernels of the first
    // clustering phase.  replica_state: 0 nothing pulled, 1 columns >= replica_lo[s] of every peer block s (all the
    // symmetric compaction of a bootstrap replicate reads), 2 complete.
    int world = 1, rank = 0; // the job geometry the matrix was created under (ctx->world is switched to 1 in local phases)
    u32 *full = nullptr;
    size_t full_bytes = 0;
    u32 *peer_full[16] = {nullptr};
    u32 *peer_shard[16] = {nullptr}; // every rank's row block (CUDA IPC); valid when peers_ok
    int peers_ok = 0;
    rid_dmat *parent = nullptr;      // set in a replica view: the sharded matrix it was made from
    int replica_state = 0;
    int replica_lo[16] = {0};
    cudaEvent_t replica_ev = nullptr;
    cudaEvent_t pull_t0 = nullptr, pull_t1 = nullptr; // timing of the first pull (diagnostics)
    rid_dmat *replica_view = nullptr; // the "one GPU owns every row" view of `full` the bootstrap PAM runs use
    // euclidean only: the raw feature matrix (N x Gx doubles, one row per cell) kept so that exported distances are
    // recomputed in FP64 difference form when the fixed-point step alone would exceed the 1e-6 budget
    double *xraw = nullptr;
    size_t xraw_bytes = 0;
    int xraw_G = 0, xraw_ld = 0;
    // PAM results on ALL cells that rid_gap_sweep produced on this (immutable) matrix, by k.  clustexp(sat=TRUE) runs
    // the sweep and then fpc::clusterboot, whose first step is the very same pam(as.dist(D), cln) call
    // (R/RaceID_utils.R:113 and :132 -> pamkdCBI): rid_clusterboot takes c1 from here instead of repeating it.
    struct PamMemo {
        int k = 0, nswaps = 0;
        std::vector<int> labels, medoids; // cstat numbering (1-based labels by cell; medoid cell of cluster c)
    };
    std::vector<PamMemo> pam_memo;
};

static inline size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// timing helpers (CUDA events on the ctx stream)
int rid_time_begin(rid_ctx *ctx);
int rid_time_end(rid_ctx *ctx);

void rid_pam_ws_free(PamWs *ws, rid_ctx *ctx);

// gram_i8.cu: exact int8 tensor-core contraction (single GPU, correlation metrics)
int rid_gram_i8(rid_ctx *ctx, const double *Z, int Gp_src, int G, int N, const unsigned char *zero_var, rid_dmat *dm,
                int *used, double *bound_out);
