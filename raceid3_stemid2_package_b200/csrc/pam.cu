// pam.cu — k-medoids (PAM BUILD / SWAP), within-cluster dispersion, medoid refresh, silhouette and the
// clusterboot Jaccard table on a row-block-sharded, fixed-point distance matrix resident in HBM.
//
// Replaces cluster::pam as called at reference R/RaceID_utils.R:113,221 (algorithm: cluster/src/pam.c
// bswap()+cstat()+dark(), restated in oracle/oracle.c), W.k R/RaceID_utils.R:39-53, fpc::clusterboot as called
// at R/RaceID_utils.R:132-133, compmedoids R/RaceID.R:1298-1316 and the refresh R/RaceID.R:1044-1054.
//
// Formulation.  The matrix is symmetric, so the cost of opening candidate h while closing
// medoid slot c,
//        TD'(h,c) = sum_j min( D[j][h], c == nearest(j) ? dsecond(j) : dnearest(j) ),
// is a column sum over point rows j with a per-row cap.  One streaming pass over the owned rows gives, for all
// columns h at once,  SP[h] = sum_j min(D[j][h], dnearest_j)  and, with rows grouped by nearest medoid,
// AQ[c][h] = sum_{j in c} (min(D[j][h], dsecond_j) - min(D[j][h], dnearest_j));  TD'(h,c) = SP[h] + AQ[c][h].
// BUILD's gain is TD - SP[h].  Distances are 32-bit fixed point, all sums are exact 64-bit integers, so the
// result does not depend on summation order, tiling or the number of GPUs, and ties are exact ties resolved
// with the upstream rules (BUILD: largest index; SWAP: first (h, then medoid) in index order).
#include <limits.h>
#include <math.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <future>
#include <vector>

#include "common.cuh"

// ----------------------------------------------------------------------------------------------------------
#define K2MAX 16384  // incremental SWAP keys rows by (old group, new group): k*k <= K2MAX, i.e. k <= 128

struct PamState {
    u64 td;       // total deviation at the last SWAP decision (fixed point)
    u64 td_build; // total deviation after BUILD
    u64 best_val;
    int best_h, best_slot;
    int done;
    int nswaps;
    int nmed;
    unsigned int ticket;
    int pad;
};

struct Cand {
    u64 val;
    int p, mp, slot;
};

struct HostSubset {
    int m = 0, N = 0;
    bool want_compact = false; // a proper subset large enough to be worth a compact copy
    int err = 0;
    char errmsg[96] = {0};
    std::vector<int> cells;    // cell by position
    std::vector<int> pos_of;   // [N] position by cell or -1
    std::vector<int> src;      // owned subset rows (cell order) as rows of the shard
    std::vector<int> rows_pos; // their positions
    std::vector<int> iota;     // 0..nls-1 (rows of the compact copy)
    std::vector<int> subset_c; // compact column by position
    std::vector<int> pos_c;    // position by compact column
    std::vector<int> ccells;   // cell by compact column (the sorted subset)
};

struct PamWs {
    HostSubset hs; // tables of the current subset when nobody prepared them ahead
    int N = 0;
    size_t ld = 0;
    int nloc = 0; // owned rows
    int kcap = 0;
    // per subset position / per cell
    int *subset = nullptr;         // [N] cell index by position
    int *pos_of = nullptr;         // [N] position by cell or -1
    unsigned char *is_med = nullptr; // [N] by cell
    int *lab_pos = nullptr;        // [N] group by position
    // owned subset rows
    int *rows = nullptr, *rows_pos = nullptr;     // [nloc]
    u32 *capP = nullptr, *capQ = nullptr;         // [nloc]
    int *grp = nullptr;                           // [nloc]
    int *s_row = nullptr, *s_grp = nullptr;       // sorted by group
    u32 *s_capP = nullptr, *s_capQ = nullptr;
    // per group
    int *med = nullptr, *medpos = nullptr; // [kcap]
    int *hist = nullptr;                   // [3*kcap]: hist, base, cursor
    u64 *grp_sum = nullptr;                // [kcap] T_c
    int *grp_cnt = nullptr;                // [kcap]
    u64 *grp_min = nullptr;                // [kcap]
    int *grp_arg = nullptr;                // [kcap]
    // accumulators: hdr[4] | SP[ld] | AQ[kcap][ld]
    u64 *acc = nullptr;
    Cand *blockbest = nullptr; // [maxblocks]
    int maxblocks = 0;
    double *sil = nullptr; // [N]
    double *dscratch = nullptr; // [8]
    PamState *state = nullptr;
    // active matrix view: the rank's shard of D, or a compacted copy D[subset, subset] (columns in sorted-cell order)
    const u32 *mat = nullptr;
    bool compact = false;
    bool sp0_ready = false;     // the compaction pass already left BUILD's step-0 sums in acc
    // BUILD's first pick known before the compaction (rid_clusterboot computes the step-0 column sums of ALL replicates in
    // a few batched passes over the resident matrix): the compaction then accumulates the sums of the SECOND step,
    // sum_j min(D[j][h], D[j][h0]), and the full rescan that step used to cost disappears
    bool sp1_ready = false;
    const int2 *first_dev = nullptr; // (cell, position) of the first medoid, on the device
    u32 *firstcap = nullptr;         // [nloc] D[j][h0] by compact row
    u32 *cbuf = nullptr;        // compact copy (owned subset rows x ldc)
    size_t cbuf_bytes = 0;
    int *subset_cells = nullptr; // [N] cell by position (== subset unless compact)
    int *colmap = nullptr;       // [N] compact column of a cell, or -1
    int *src_rows = nullptr;     // [nloc] shard row of each owned subset row (compaction source)
    // incremental SWAP: previous assignment of the owned rows, the list of rows whose assignment changed (old / new
    // caps and groups, sorted by the (old group, new group) pair) and a delta accumulator with acc's layout
    u32 *o_capP = nullptr, *o_capQ = nullptr;
    int *o_grp = nullptr;
    int *d_row = nullptr, *d_key = nullptr;
    u32 *d_po = nullptr, *d_qo = nullptr, *d_pn = nullptr, *d_qn = nullptr;
    int *hist2 = nullptr; // [3 * K2MAX]: counts, bases, cursors of the pair keys
    u64 *acc2 = nullptr;
    int *chg_count = nullptr; // rows whose cap changed in the last BUILD step
    // lazy BUILD (complete views): per candidate column an upper bound of its gain, the step it was last exact at, the
    // "cap dropped in the last step" flag; lz_thr = {stage-A threshold, best exact gain so far, rows evaluated, -}
    u64 *lz_ub = nullptr;
    int *lz_stamp = nullptr;
    unsigned char *lz_chg = nullptr;
    u64 *lz_thr = nullptr;
    // host mirrors for the current subset
    int m = 0, nls = 0;
    // the two pooled slabs every device array above is carved from
    char *slab = nullptr, *gslab = nullptr;
    size_t slab_bytes = 0, gslab_bytes = 0;
};

// All per-matrix work arrays live in two slabs taken from the context's buffer pool (one for the per-cell / per-row
// arrays, one for the per-group arrays and accumulators): creating and dropping a workspace costs no cudaMalloc /
// cudaFree once the pool is warm.  (Dozens of small cudaFree calls per clustexp() showed up as 0.1-0.5 s of host stalls
// on some boxes.)
template <class T> static inline void carve(char *&p, T *&dst, size_t count)
{
    dst = reinterpret_cast<T *>(p);
    p += (count * sizeof(T) + 255) & ~(size_t)255;
}

static size_t ws_carve_base(PamWs *w, char *base)
{
    char *p = base;
    const size_t N = (size_t)w->N, nl = (size_t)std::max(w->nloc, 1);
    carve(p, w->subset, N); carve(p, w->pos_of, N); carve(p, w->is_med, N); carve(p, w->lab_pos, N);
    carve(p, w->rows, nl); carve(p, w->rows_pos, nl); carve(p, w->capP, nl); carve(p, w->capQ, nl); carve(p, w->grp, nl);
    carve(p, w->s_row, nl); carve(p, w->s_grp, nl); carve(p, w->s_capP, nl); carve(p, w->s_capQ, nl);
    carve(p, w->blockbest, (size_t)w->maxblocks); carve(p, w->sil, N); carve(p, w->dscratch, 8);
    carve(p, w->state, 1); carve(p, w->chg_count, 1);
    carve(p, w->o_capP, nl); carve(p, w->o_capQ, nl); carve(p, w->o_grp, nl); carve(p, w->d_row, nl); carve(p, w->d_key, nl);
    carve(p, w->d_po, nl); carve(p, w->d_qo, nl); carve(p, w->d_pn, nl); carve(p, w->d_qn, nl);
    carve(p, w->hist2, (size_t)3 * K2MAX);
    carve(p, w->subset_cells, N); carve(p, w->colmap, N); carve(p, w->src_rows, nl);
    carve(p, w->lz_ub, N + 128); carve(p, w->lz_stamp, N + 128); carve(p, w->lz_chg, N + 128); carve(p, w->lz_thr, 4);
    carve(p, w->firstcap, nl + 128);
    return (size_t)(p - base);
}

static size_t ws_carve_groups(PamWs *w, char *base, int kc, size_t ld)
{
    char *p = base;
    const size_t k = (size_t)kc, nacc = 4 + (k + 1) * ld;
    carve(p, w->med, k); carve(p, w->medpos, k); carve(p, w->hist, 3 * k); carve(p, w->grp_sum, k); carve(p, w->grp_cnt, k);
    carve(p, w->grp_min, k); carve(p, w->grp_arg, k);
    carve(p, w->acc, nacc); carve(p, w->acc2, nacc);
    return (size_t)(p - base);
}

static void ws_free_groups(PamWs *w, rid_ctx *ctx)
{
    rid_pool_free(ctx, w->gslab, w->gslab_bytes);
    w->gslab = nullptr; w->gslab_bytes = 0;
    w->acc2 = nullptr;
    w->med = w->medpos = w->hist = w->grp_cnt = w->grp_arg = nullptr;
    w->grp_sum = w->grp_min = w->acc = nullptr;
}

void rid_pam_ws_free(PamWs *w, rid_ctx *ctx)
{
    if (!w) return;
    if (w->cbuf) rid_pool_free(ctx, w->cbuf, w->cbuf_bytes);
    rid_pool_free(ctx, w->slab, w->slab_bytes);
    ws_free_groups(w, ctx);
    delete w;
}

static int ws_get(rid_dmat *dm, int k, PamWs **out)
{
    PamWs *w = dm->ws;
    rid_ctx *ctx = dm->ctx;
    RID_CUDA(cudaSetDevice(ctx->device));
    if (!w) {
        w = new PamWs();
        dm->ws = w;
        w->N = dm->N;
        w->ld = dm->ld;
        w->nloc = dm->row1 - dm->row0;
        w->maxblocks = (int)(((size_t)dm->N + 255) / 256);
        w->slab_bytes = ws_carve_base(w, nullptr);
        if (rid_pool_alloc(ctx, (void **)&w->slab, w->slab_bytes) != cudaSuccess) {
            cudaGetLastError();
            rid_set_error("out of device memory for the PAM work arrays (%zu bytes)", w->slab_bytes);
            dm->ws = nullptr;
            delete w;
            return RID_ERR_NOMEM;
        }
        ws_carve_base(w, w->slab);
    }
    if (k > w->kcap) {
        ws_free_groups(w, ctx);
        w->kcap = 0;
        int kc = std::max(k, 32);
        w->gslab_bytes = ws_carve_groups(w, nullptr, kc, dm->ld);
        if (rid_pool_alloc(ctx, (void **)&w->gslab, w->gslab_bytes) != cudaSuccess) {
            cudaGetLastError();
            w->gslab = nullptr; w->gslab_bytes = 0;
            ws_free_groups(w, ctx);
            rid_set_error("out of device memory for %d group accumulators", kc);
            return RID_ERR_NOMEM;
        }
        ws_carve_groups(w, w->gslab, kc, dm->ld);
        w->kcap = kc;
    }
    *out = w;
    return RID_OK;
}

// ----------------------------------------------------------------------------------------------------------
// kernels
// ----------------------------------------------------------------------------------------------------------
#define LAUNCH(ctx, kern, grid, block, smem, ...)                         \
    do {                                                                  \
        kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);    \
        (ctx)->launches++;                                                \
    } while (0)

__global__ void k_fill_u32(u32 *p, u32 v, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void k_zero(u64 *acc, size_t n_acc, int *hist, int n_hist, const PamState *st, int check_done,
                       int *hist2 = nullptr, int n_hist2 = 0, int *counter = nullptr)
{
    if (check_done && st->done) return;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = i; t < n_acc; t += stride) acc[t] = 0ull;
    if (i < (size_t)n_hist) hist[i] = 0;
    for (size_t t = i; t < (size_t)n_hist2; t += stride) hist2[t] = 0; // (incremental SWAP: pair-key counts, changed-row counter)
    if (i == 0 && counter) *counter = 0;
}

__global__ void k_state_init(PamState *st)
{
    st->td = 0; st->td_build = 0; st->best_val = 0; st->best_h = -1; st->best_slot = -1;
    st->done = 0; st->nswaps = 0; st->nmed = 0; st->ticket = 0u;
}

// medoids := given cells (slots 0..k-1); used to start SWAP(k) from the shared BUILD prefix
__global__ void k_set_medoids(const int *cells, int k, int *med, int *medpos, const int *pos_of,
                              unsigned char *is_med, PamState *st)
{
    int c = threadIdx.x;
    if (c < k) {
        int h = cells[c];
        med[c] = h;
        medpos[c] = pos_of[h];
        is_med[h] = 1;
    }
    if (c == 0) { st->nmed = k; st->done = 0; st->nswaps = 0; st->ticket = 0u; }
}

// The streaming pass.  Thread = 4 consecutive columns (one 128-bit load per row), CTA = 512 columns x a chunk
// of the (group-sorted) row list.  32-bit partial sums are promoted to 64 bits every FLUSH rows
// (q <= 2^28: 8 rows; q <= 2^31: 2 rows), and pushed to SP / AQ with 64-bit integer reductions.
#define SCAN_THREADS 128
#define SCAN_MAXROWS 512

__device__ __forceinline__ uint4 ld_stream(const u32 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void red_add_u64(u64 *p, u64 v)
{
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// MODE 0: SP[h] += sum_j min(d, capP_j)                                   (BUILD step 0)
// MODE 1: MODE 0 plus AQ[g_j][h] += min(d, capQ_j) - min(d, capP_j)       (SWAP, grouped column sums)
// MODE 2: only AQ[g_j][h] += min(d, capQ_j) - min(d, capP_j)              (incremental BUILD: capQ = old, capP = new cap)
// A CTA owns 512 columns and walks row chunks blockIdx.y, blockIdx.y + gridDim.y, ...; the number of rows may live on
// the device (nrows_dev), so the host never has to read it back.
template <int MODE, int FLUSH>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan(const u32 *__restrict__ D, size_t ld, const int *__restrict__ s_row,
       const u32 *__restrict__ s_capP, const u32 *__restrict__ s_capQ, const int *__restrict__ s_grp, int nrows,
       const int *__restrict__ nrows_dev, int rows_per_chunk, u64 *__restrict__ SP, u64 *__restrict__ AQ,
       const PamState *st, int check_done)
{
    if (check_done > 0 && st->done) return; // (check_done < 0: A/B switch, static row chunks)
    constexpr bool WITH_P = MODE != 2, WITH_Q = MODE != 0;
    __shared__ int sh_row[SCAN_MAXROWS];
    __shared__ u32 sh_cp[SCAN_MAXROWS];
    __shared__ u32 sh_cq[SCAN_MAXROWS];
    __shared__ int sh_g[SCAN_MAXROWS];

    const int n = nrows_dev ? *nrows_dev : nrows;
    // The grid was sized for the host's upper bound of the row count.  When the real count is only known here, spread it
    // over ALL row chunks of the grid (>= 8 rows each) instead of leaving most CTAs without work: a small incremental
    // pass then keeps every SM streaming.
    if (nrows_dev && check_done >= 0) rows_per_chunk = min(rows_per_chunk, max(8, (((n + (int)gridDim.y - 1) / (int)gridDim.y) + 7) & ~7));
    const size_t col = ((size_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * 4;
    const bool active = col < ld;
    const u32 *base = D + col;

    u64 accP[4] = {0, 0, 0, 0}, accQ[4] = {0, 0, 0, 0};
    u32 p32[4] = {0, 0, 0, 0}, q32[4] = {0, 0, 0, 0};
    int cur_g = -1;
    bool any = false;

    for (int chunk = blockIdx.y; (long long)chunk * rows_per_chunk < n; chunk += gridDim.y) {
        const int r_begin = chunk * rows_per_chunk;
        const int nr = min(n - r_begin, rows_per_chunk);
        const int nr8 = (nr + 7) & ~7;
        __syncthreads(); // everyone is done with the previous chunk's shared arrays
        for (int i = threadIdx.x; i < nr8; i += SCAN_THREADS) {
            bool ok = i < nr;
            int src = r_begin + (ok ? i : nr - 1);
            sh_row[i] = ok ? s_row[src] : -1;
            sh_cp[i] = ok ? s_capP[src] : 0u;
            if (WITH_Q) {
                sh_cq[i] = ok ? s_capQ[src] : 0u;
                sh_g[i] = s_grp[src];
            }
        }
        __syncthreads();
        if (!active) continue;
        any = true;
        if (WITH_Q && cur_g < 0) cur_g = sh_g[0];
        for (int r = 0; r < nr8; r += 8) {
            uint4 d[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                int lr = sh_row[r + u];
                d[u] = lr >= 0 ? ld_stream(base + (size_t)lr * ld) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const u32 cp = sh_cp[r + u];
                if (WITH_Q) {
                    const int g = sh_g[r + u];
                    if (g != cur_g) { // uniform across the CTA
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (WITH_P) { accP[c] += p32[c]; p32[c] = 0; }
                            accQ[c] += q32[c]; q32[c] = 0;
                            red_add_u64(AQ + (size_t)cur_g * ld + col + c, accQ[c]);
                            accQ[c] = 0;
                        }
                        cur_g = g;
                    }
                    const u32 cq = sh_cq[r + u];
                    u32 p, q;
                    p = min(d[u].x, cp); q = min(d[u].x, cq); if (WITH_P) p32[0] += p; q32[0] += q - p;
                    p = min(d[u].y, cp); q = min(d[u].y, cq); if (WITH_P) p32[1] += p; q32[1] += q - p;
                    p = min(d[u].z, cp); q = min(d[u].z, cq); if (WITH_P) p32[2] += p; q32[2] += q - p;
                    p = min(d[u].w, cp); q = min(d[u].w, cq); if (WITH_P) p32[3] += p; q32[3] += q - p;
                } else {
                    p32[0] += min(d[u].x, cp);
                    p32[1] += min(d[u].y, cp);
                    p32[2] += min(d[u].z, cp);
                    p32[3] += min(d[u].w, cp);
                }
                if ((u % FLUSH) == FLUSH - 1) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (WITH_P) { accP[c] += p32[c]; p32[c] = 0; }
                        if (WITH_Q) { accQ[c] += q32[c]; q32[c] = 0; }
                    }
                }
            }
        }
    }
    if (!any) return;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (WITH_P) red_add_u64(SP + col + c, accP[c]);
        if (WITH_Q) red_add_u64(AQ + (size_t)cur_g * ld + col + c, accQ[c]);
    }
}

// ---- incremental SWAP ------------------------------------------------------------------------------------------
// After a swap only the rows whose (nearest, second nearest) pair changed contribute differently to SP / AQ.  For those
// rows the pass below subtracts the old contribution and adds the new one:
//     SP[h]      += min(d, dn_new) - min(d, dn_old)
//     AQ[g_old][h] -= min(d, ds_old) - min(d, dn_old)          AQ[g_new][h] += min(d, ds_new) - min(d, dn_new)
// in wrap-around 64-bit integer arithmetic (exact), so SP / AQ stay equal to what a full pass would give.
__global__ void k_prefix2(int *hist2, int k2, int *total, const PamState *st, u64 *prof_dev, int m, size_t ld)
{
    if (st->done) return;
    // one warp: exclusive prefix of the k2 <= K2MAX pair-key counts (bases at [K2MAX + c], cursors at [2 K2MAX + c] zeroed)
    const int lane = threadIdx.x;
    const int per = (k2 + 31) / 32, b = lane * per, e = min(k2, b + per);
    int cnt = 0;
    for (int c = b; c < e; ++c) cnt += hist2[c];
    int incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    int run = incl - cnt;
    for (int c = b; c < e; ++c) {
        hist2[K2MAX + c] = run;
        hist2[2 * K2MAX + c] = 0;
        run += hist2[c];
    }
    const int tot = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 0) {
        *total = tot;
        if (prof_dev) { // rows the delta pass is about to rescan
            atomicAdd((unsigned long long *)prof_dev + 2, (unsigned long long)tot * (unsigned long long)m);
            atomicAdd((unsigned long long *)prof_dev + 3, (unsigned long long)tot * (unsigned long long)ld);
        }
    }
}

__global__ void k_diff_scatter(const int *__restrict__ rows, const u32 *__restrict__ nP, const u32 *__restrict__ nQ,
                               const int *__restrict__ nG, const u32 *__restrict__ oP, const u32 *__restrict__ oQ,
                               const int *__restrict__ oG, int nls, int k, int *hist2, int *d_row, int *d_key, u32 *d_po,
                               u32 *d_qo, u32 *d_pn, u32 *d_qn, const PamState *st)
{
    if (st->done) return;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nls && (nP[t] != oP[t] || nQ[t] != oQ[t] || nG[t] != oG[t])) {
        int key = oG[t] * k + nG[t];
        int dst = hist2[K2MAX + key] + atomicAdd(&hist2[2 * K2MAX + key], 1);
        d_row[dst] = rows[t];
        d_key[dst] = key;
        d_po[dst] = oP[t]; d_qo[dst] = oQ[t];
        d_pn[dst] = nP[t]; d_qn[dst] = nQ[t];
    }
}

template <int FLUSH>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_delta(const u32 *__restrict__ D, size_t ld, const int *__restrict__ d_row, const int *__restrict__ d_key,
             const u32 *__restrict__ d_po, const u32 *__restrict__ d_qo, const u32 *__restrict__ d_pn,
             const u32 *__restrict__ d_qn, const int *__restrict__ nrows_dev, int rows_per_chunk, int k,
             u64 *__restrict__ SP, u64 *__restrict__ AQ, const PamState *st)
{
    if (st->done) return;
    __shared__ int sh_row[SCAN_MAXROWS];
    __shared__ int sh_key[SCAN_MAXROWS];
    __shared__ u32 sh_po[SCAN_MAXROWS], sh_qo[SCAN_MAXROWS], sh_pn[SCAN_MAXROWS], sh_qn[SCAN_MAXROWS];
    const int n = *nrows_dev;
    // (the grid was sized for all owned rows: spread the changed rows over all of its row chunks, see k_scan)
    rows_per_chunk = min(rows_per_chunk, max(8, (((n + (int)gridDim.y - 1) / (int)gridDim.y) + 7) & ~7));
    const size_t col = ((size_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * 4;
    const bool active = col < ld;
    const u32 *base = D + col;
    long long accS[4] = {0, 0, 0, 0};
    u64 accO[4] = {0, 0, 0, 0}, accN[4] = {0, 0, 0, 0};
    int s32[4] = {0, 0, 0, 0};
    u32 o32[4] = {0, 0, 0, 0}, n32[4] = {0, 0, 0, 0};
    int cur = -1;
    bool any = false;
    for (int chunk = blockIdx.y; (long long)chunk * rows_per_chunk < n; chunk += gridDim.y) {
        const int r_begin = chunk * rows_per_chunk;
        const int nr = min(n - r_begin, rows_per_chunk);
        const int nr8 = (nr + 7) & ~7;
        __syncthreads();
        for (int i = threadIdx.x; i < nr8; i += SCAN_THREADS) {
            bool ok = i < nr;
            int src = r_begin + (ok ? i : nr - 1);
            sh_row[i] = ok ? d_row[src] : -1;
            sh_key[i] = d_key[src];
            sh_po[i] = ok ? d_po[src] : 0u; sh_qo[i] = ok ? d_qo[src] : 0u;
            sh_pn[i] = ok ? d_pn[src] : 0u; sh_qn[i] = ok ? d_qn[src] : 0u;
        }
        __syncthreads();
        if (!active) continue;
        any = true;
        if (cur < 0) cur = sh_key[0];
        for (int r = 0; r < nr8; r += 8) {
            uint4 d[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                int lr = sh_row[r + u];
                d[u] = lr >= 0 ? ld_stream(base + (size_t)lr * ld) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int key = sh_key[r + u];
                if (key != cur) { // uniform across the CTA
                    const int go = cur / k, gn = cur - go * k;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        accS[c] += s32[c]; s32[c] = 0;
                        accO[c] += o32[c]; o32[c] = 0;
                        accN[c] += n32[c]; n32[c] = 0;
                        red_add_u64(AQ + (size_t)go * ld + col + c, 0ull - accO[c]);
                        red_add_u64(AQ + (size_t)gn * ld + col + c, accN[c]);
                        accO[c] = 0; accN[c] = 0;
                    }
                    cur = key;
                }
                const u32 po = sh_po[r + u], qo = sh_qo[r + u], pn = sh_pn[r + u], qn = sh_qn[r + u];
                const u32 dv[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const u32 a = min(dv[c], po), b = min(dv[c], pn);
                    s32[c] += (int)(b - a); // |b - a| <= q_max
                    o32[c] += min(dv[c], qo) - a;
                    n32[c] += min(dv[c], qn) - b;
                }
                if ((u % (FLUSH / 2)) == FLUSH / 2 - 1) { // the signed partial holds half as many rows
#pragma unroll
                    for (int c = 0; c < 4; ++c) { accS[c] += s32[c]; s32[c] = 0; }
                }
                if ((u % FLUSH) == FLUSH - 1) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        accO[c] += o32[c]; o32[c] = 0;
                        accN[c] += n32[c]; n32[c] = 0;
                    }
                }
            }
        }
    }
    if (!any) return;
    const int go = cur / k, gn = cur - go * k;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        red_add_u64(SP + col + c, (u64)accS[c]);
        red_add_u64(AQ + (size_t)go * ld + col + c, 0ull - accO[c]);
        red_add_u64(AQ + (size_t)gn * ld + col + c, accN[c]);
    }
}

// acc[0] = delta[0] (the new TD), acc[4..] += delta[4..]   (wrap-around adds: exact)
__global__ void k_acc_add(u64 *acc, const u64 *__restrict__ delta, size_t n, const PamState *st)
{
    if (st->done) return;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    if (i == 0) acc[0] = delta[0];
    for (size_t t = 4 + i; t < n; t += stride) acc[t] += delta[t];
}

// nearest / second-nearest medoid of every owned subset row (cstat + the dysma/dysmb refresh of bswap):
// strict '<' with medoids visited in index (position) order => the lowest-position medoid wins ties.
// DELTA (incremental SWAP): the previous assignment of every row is kept in o_capP / o_capQ / o_grp and the rows whose
// (nearest, second nearest) pair changed are counted by their (old group, new group) key in hist2.
template <bool DELTA>
__global__ void k_assign(const u32 *__restrict__ D, size_t ld, const int *__restrict__ rows, int nls,
                         const int *__restrict__ med, const int *__restrict__ medpos, int k, u32 *capP, u32 *capQ,
                         int *grp, int *hist, u64 *td, const PamState *st, int check_done,
                         u32 *o_capP = nullptr, u32 *o_capQ = nullptr, int *o_grp = nullptr, int *hist2 = nullptr)
{
    if (check_done && st->done) return;
    extern __shared__ int sh[];
    int *sh_med = sh, *sh_mp = sh + k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) { sh_med[i] = med[i]; sh_mp[i] = medpos[i]; }
    __syncthreads();
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    u64 mine = 0;
    if (t < nls) {
        const u32 *row = D + (size_t)rows[t] * ld;
        u32 d1 = 0xFFFFFFFFu, d2 = 0xFFFFFFFFu;
        int g = -1, gp = INT_MAX;
        for (int c = 0; c < k; ++c) {
            u32 d = row[sh_med[c]];
            int mp = sh_mp[c];
            if (d < d1 || (d == d1 && mp < gp)) {
                d2 = d1; d1 = d; g = c; gp = mp;
            } else if (d < d2) {
                d2 = d;
            }
        }
        if (DELTA) {
            const u32 p0 = capP[t], q0 = capQ[t];
            const int g0 = grp[t];
            o_capP[t] = p0; o_capQ[t] = q0; o_grp[t] = g0;
            if (d1 != p0 || d2 != q0 || g != g0) atomicAdd(&hist2[g0 * k + g], 1);
        }
        capP[t] = d1;
        capQ[t] = d2;
        grp[t] = g;
        atomicAdd(&hist[g], 1);
        mine = d1;
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(td, mine);
}

// groups given per position (W.k, compmedoids, silhouette): plain grouped column sums (capP=0, capQ=inf)
__global__ void k_groups_from_labels(const int *__restrict__ rows_pos, int nls, const int *__restrict__ lab_pos,
                                     u32 *capP, u32 *capQ, int *grp, int *hist)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nls) {
        int g = lab_pos[rows_pos[t]];
        capP[t] = 0u;
        capQ[t] = 0xFFFFFFFFu;
        grp[t] = g;
        atomicAdd(&hist[g], 1);
    }
}

__global__ void k_prefix(int *hist, int k, const PamState *st, int check_done)
{
    if (check_done && st->done) return;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int run = 0;
        for (int c = 0; c < k; ++c) {
            hist[k + c] = run;   // base
            hist[2 * k + c] = 0; // cursor
            run += hist[c];
        }
    }
}

__global__ void k_scatter(const int *__restrict__ rows, const u32 *__restrict__ capP, const u32 *__restrict__ capQ,
                          const int *__restrict__ grp, int nls, int *hist, int k, int *s_row, u32 *s_capP,
                          u32 *s_capQ, int *s_grp, const PamState *st, int check_done)
{
    if (check_done && st->done) return;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nls) {
        int g = grp[t];
        int dst = hist[k + g] + atomicAdd(&hist[2 * k + g], 1);
        s_row[dst] = rows[t];
        s_capP[dst] = capP[t];
        s_capQ[dst] = capQ[t];
        s_grp[dst] = g;
    }
}

__device__ __forceinline__ bool cand_better_swap(const Cand &a, const Cand &b)
{
    if (a.val != b.val) return a.val < b.val;
    if (a.p != b.p) return a.p < b.p;
    return a.mp < b.mp;
}
__device__ __forceinline__ bool cand_better_build(const Cand &a, const Cand &b)
{
    if (a.val != b.val) return a.val < b.val;
    return a.p > b.p; // upstream BUILD: `if (ammax <= beter[i])` => the largest index wins ties
}

template <bool BUILD>
__device__ __forceinline__ Cand cand_reduce_block(Cand c, Cand *sh)
{
    for (int o = 16; o > 0; o >>= 1) {
        Cand b;
        b.val = __shfl_down_sync(0xffffffffu, c.val, o);
        b.p = __shfl_down_sync(0xffffffffu, c.p, o);
        b.mp = __shfl_down_sync(0xffffffffu, c.mp, o);
        b.slot = __shfl_down_sync(0xffffffffu, c.slot, o);
        if (BUILD ? cand_better_build(b, c) : cand_better_swap(b, c)) c = b;
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (l == 0) sh[w] = c;
    __syncthreads();
    if (w == 0) {
        Cand e;
        e.val = ~0ull; e.p = BUILD ? -1 : INT_MAX; e.mp = INT_MAX; e.slot = -1;
        c = l < nw ? sh[l] : e;
        for (int o = 16; o > 0; o >>= 1) {
            Cand b;
            b.val = __shfl_down_sync(0xffffffffu, c.val, o);
            b.p = __shfl_down_sync(0xffffffffu, c.p, o);
            b.mp = __shfl_down_sync(0xffffffffu, c.mp, o);
            b.slot = __shfl_down_sync(0xffffffffu, c.slot, o);
            if (BUILD ? cand_better_build(b, c) : cand_better_swap(b, c)) c = b;
        }
    }
    __syncthreads();
    return c; // valid in thread 0
}

// Candidate selection.  SWAP: argmin over (h non-medoid, medoid slot c) of SP[h]+AQ[c][h], first (h, medoid) in
// position order on ties; accepted iff it lowers TD (upstream: dzsky < -16*DBL_EPSILON*|sky|, which on exact
// fixed-point sums is dz < 0).  BUILD: argmin SP[h], largest position on ties.
template <bool BUILD>
__global__ void __launch_bounds__(256)
k_select(const int *__restrict__ subset, int m, const unsigned char *is_med_c,
         u64 *acc, size_t ld, int k, int *med, int *medpos, unsigned char *is_med,
         Cand *blockbest, PamState *st, int check_done, u64 *sp_sub, int *zero_cnt, u64 *prof_dev, int *order_out = nullptr)
{
    // order_out (BUILD, k-sweep): only report the pick -- order_out[k] = its column -- and leave the medoid tables alone
    // sp_sub (incremental BUILD): SP[h] -= sp_sub[h] and sp_sub[h] = 0 on the way (every subset column is visited once);
    // zero_cnt: a counter the next kernel fills again
    if (check_done && st->done) return;
    __shared__ Cand sh[32];
    __shared__ int sh_mp[1024];
    __shared__ bool is_last;
    u64 *SP = acc + 4;
    const u64 *AQ = SP + ld;
    if (!BUILD)
        for (int i = threadIdx.x; i < k && i < 1024; i += blockDim.x) sh_mp[i] = medpos[i];
    __syncthreads();
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    Cand best;
    best.val = ~0ull; best.p = BUILD ? -1 : INT_MAX; best.mp = INT_MAX; best.slot = -1;
    if (p < m) {
        int h = subset[p];
        u64 sp = SP[h];
        if (BUILD && sp_sub) {
            sp -= sp_sub[h];
            SP[h] = sp;
            sp_sub[h] = 0ull;
        }
        if (!is_med_c[h]) {
            if (BUILD) {
                best.val = sp; best.p = p; best.mp = 0; best.slot = 0;
            } else {
                for (int c = 0; c < k; ++c) {
                    Cand x;
                    x.val = sp + AQ[(size_t)c * ld + h];
                    x.p = p;
                    x.mp = c < 1024 ? sh_mp[c] : medpos[c];
                    x.slot = c;
                    if (cand_better_swap(x, best)) best = x;
                }
            }
        }
    }
    best = cand_reduce_block<BUILD>(best, sh);
    if (threadIdx.x == 0) {
        blockbest[blockIdx.x] = best;
        __threadfence();
        unsigned int t = atomicAdd(&st->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    Cand c;
    c.val = ~0ull; c.p = BUILD ? -1 : INT_MAX; c.mp = INT_MAX; c.slot = -1;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
        Cand b = blockbest[i];
        if (BUILD ? cand_better_build(b, c) : cand_better_swap(b, c)) c = b;
    }
    c = cand_reduce_block<BUILD>(c, sh);
    if (threadIdx.x == 0) {
        st->ticket = 0u;
        if (zero_cnt) {
            if (prof_dev && sp_sub) { // rows the incremental pass of this step rescanned
                atomicAdd((unsigned long long *)prof_dev, (unsigned long long)*zero_cnt * (unsigned long long)m);
                atomicAdd((unsigned long long *)prof_dev + 1, (unsigned long long)*zero_cnt * (unsigned long long)ld);
            }
            *zero_cnt = 0;
        }
        u64 td = acc[0];
        if (BUILD && order_out) {
            order_out[k] = subset[c.p];
        } else if (BUILD) {
            int h = subset[c.p];
            int s = st->nmed;
            med[s] = h;
            medpos[s] = c.p;
            is_med[h] = 1;
            st->nmed = s + 1;
            st->best_h = h; st->best_slot = s; st->best_val = c.val;
        } else {
            st->td = td;
            st->best_val = c.val;
            if (c.slot >= 0 && c.val < td) {
                int h = subset[c.p];
                is_med[med[c.slot]] = 0;
                is_med[h] = 1;
                med[c.slot] = h;
                medpos[c.slot] = c.p;
                st->best_h = h; st->best_slot = c.slot;
                st->nswaps += 1;
            } else {
                st->done = 1;
            }
        }
    }
}

// after BUILD picked medoid `best_h`: dnearest_j = min(dnearest_j, D[j][best_h]).  Rows whose cap changed are appended
// (old cap, new cap) to the delta list: the next BUILD step only has to rescan those rows,
//      SP_new[h] = SP_old[h] - sum_{changed j} ( min(D[j][h], old_j) - min(D[j][h], new_j) ).
__global__ void k_build_update(const u32 *__restrict__ D, size_t ld, const int *__restrict__ rows,
                               int nls, u32 *capP, const PamState *st, int *chg_count, int *c_row, u32 *c_new,
                               u32 *c_old, int *c_grp)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nls) {
        u32 d = D[(size_t)rows[t] * ld + st->best_h];
        u32 old = capP[t];
        if (d < old) {
            capP[t] = d;
            int i = atomicAdd(chg_count, 1);
            c_row[i] = rows[t];
            c_new[i] = d;
            c_old[i] = old;
            c_grp[i] = 0;
        }
    }
}

__global__ void k_capture_td_build(PamState *st, const u64 *acc) { st->td_build = acc[0]; }

__global__ void k_labels_to_pos(const int *__restrict__ rows_pos, const int *__restrict__ grp, int nls, int *lab_pos)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nls) lab_pos[rows_pos[t]] = grp[t];
}

// T_c = sum_{h in c} S_h with S_h = AQ[c][h]; n_c
__global__ void k_wdisp(const int *__restrict__ subset, const int *__restrict__ lab_pos, int m,
                        const u64 *__restrict__ AQ, size_t ld, u64 *grp_sum, int *grp_cnt)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < m) {
        int c = lab_pos[p];
        atomicAdd(&grp_sum[c], AQ[(size_t)c * ld + subset[p]]);
        atomicAdd(&grp_cnt[c], 1);
    }
}

// compmedoids: first argmin of S_h within each cluster (two passes of integer atomicMin)
__global__ void k_medoid_min(const int *__restrict__ subset, const int *__restrict__ lab_pos, int m,
                             const u64 *__restrict__ AQ, size_t ld, u64 *grp_min, int *grp_arg, int pass)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < m) {
        int c = lab_pos[p];
        u64 s = AQ[(size_t)c * ld + subset[p]];
        if (pass == 0) atomicMin(&grp_min[c], s);
        else if (s == grp_min[c]) atomicMin(&grp_arg[c], p);
    }
}

// silhouette width of every subset position (pam.c dark())
__global__ void k_silhouette(const int *__restrict__ subset, const int *__restrict__ lab_pos, int m, int k,
                             const u64 *__restrict__ AQ, size_t ld, const int *__restrict__ grp_cnt, double *sil)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < m) {
        int h = subset[p], cj = lab_pos[p];
        double sw = 0.0;
        int nj = grp_cnt[cj];
        if (nj > 1) {
            double a = (double)AQ[(size_t)cj * ld + h] / (double)(nj - 1);
            double b = INFINITY;
            for (int c = 0; c < k; ++c)
                if (c != cj && grp_cnt[c] > 0) {
                    double v = (double)AQ[(size_t)c * ld + h] / (double)grp_cnt[c];
                    if (v < b) b = v;
                }
            double mx = a > b ? a : b;
            sw = mx > 0.0 ? (b - a) / mx : 0.0;
        }
        sil[p] = sw;
    }
}

// cluster::silhouette(part, dist) per cell: a = mean distance to the other members of the own cluster, b = smallest mean
// distance to another cluster (the FIRST such cluster in sorted cluster-id order on ties, like which.min), neighbour = that
// cluster, s = (b - a) / max(a, b); singleton clusters get s = 0.  Sums are exact integers; the ratios are FP64.
__global__ void k_silhouette_full(const int *__restrict__ subset, const int *__restrict__ lab_pos, int m, int k,
                                  const u64 *__restrict__ AQ, size_t ld, const int *__restrict__ grp_cnt, double *sil,
                                  int *neighbor)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < m) {
        const int h = subset[p], cj = lab_pos[p];
        const int nj = grp_cnt[cj];
        double b = INFINITY;
        int nb = -1;
        for (int c = 0; c < k; ++c)
            if (c != cj && grp_cnt[c] > 0) {
                double v = (double)AQ[(size_t)c * ld + h] / (double)grp_cnt[c];
                if (v < b) { b = v; nb = c; }
            }
        double sw = 0.0;
        if (nj > 1) {
            const double a = (double)AQ[(size_t)cj * ld + h] / (double)(nj - 1);
            const double mx = a > b ? a : b;
            sw = mx > 0.0 ? (b - a) / mx : 0.0;
        }
        sil[p] = sw;
        neighbor[p] = nb;
    }
}

// ---- W.k for several partitions in ONE pass (the sweep) ---------------------------------------------------------------
// S_q[h] = sum over rows j in h's cluster (under partition q) of D[j][h], for WB partitions at once: the matrix is read
// once for WB values of k instead of once per k.  Labels are bytes; a thread keeps the labels of its 4 columns for the
// WB partitions packed in WB registers and compares them with the row's labels four at a time (__vcmpeq4).
#define WB 6
__global__ void k_labels_by_column(const int *__restrict__ subset, const int *__restrict__ lab_pos, int m,
                                   unsigned char *__restrict__ lab_col)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < m) lab_col[subset[p]] = (unsigned char)lab_pos[p];
}

template <int FLUSH>
__global__ void __launch_bounds__(SCAN_THREADS)
k_wk_batch(const u32 *__restrict__ D, size_t ld, const int *__restrict__ rows, const int *__restrict__ rows_pos,
           const int *__restrict__ subset, int nrows, int rows_per_chunk,
           const unsigned char *__restrict__ lab_col /* [WB][lab_pitch] by column */, size_t lab_pitch, int nq,
           u64 *__restrict__ S /* [WB][ld] */)
{
    __shared__ int sh_row[SCAN_MAXROWS];
    __shared__ unsigned char sh_lab[WB][SCAN_MAXROWS];
    const size_t col = ((size_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * 4;
    const bool active = col < ld;
    u32 clab[WB];
#pragma unroll
    for (int q = 0; q < WB; ++q)
        clab[q] = (active && q < nq) ? *reinterpret_cast<const u32 *>(lab_col + (size_t)q * lab_pitch + col) : 0xFFFFFFFFu;
    u64 acc[WB][4];
    u32 p32[WB][4];
#pragma unroll
    for (int q = 0; q < WB; ++q)
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[q][c] = 0; p32[q][c] = 0; }
    bool any = false;
    for (int chunk = blockIdx.y; (long long)chunk * rows_per_chunk < nrows; chunk += gridDim.y) {
        const int r_begin = chunk * rows_per_chunk;
        const int nr = min(nrows - r_begin, rows_per_chunk);
        const int nr8 = (nr + 7) & ~7;
        __syncthreads();
        for (int i = threadIdx.x; i < nr8; i += SCAN_THREADS) {
            const bool ok = i < nr;
            const int lr = ok ? rows[r_begin + i] : -1;
            sh_row[i] = lr;
#pragma unroll
            for (int q = 0; q < WB; ++q) sh_lab[q][i] = 0xFE; // matches no column label (0xFF pads columns)
        }
        __syncthreads();
        // the row's own labels: owned row slot -> position -> column id of the same cell
        for (int i = threadIdx.x; i < nr; i += SCAN_THREADS) {
            const size_t cid = (size_t)subset[rows_pos[r_begin + i]];
#pragma unroll
            for (int q = 0; q < WB; ++q)
                if (q < nq) sh_lab[q][i] = lab_col[(size_t)q * lab_pitch + cid];
        }
        __syncthreads();
        if (!active) continue;
        any = true;
        for (int r = 0; r < nr8; r += 8) {
            uint4 d[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int lr = sh_row[r + u];
                d[u] = lr >= 0 ? ld_stream(D + (size_t)lr * ld + col) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int q = 0; q < WB; ++q) {
                    // one bit per equal byte; the masked add is a multiply-add by that bit: IMAD runs on the FMA pipe, so the
                    // integer pipe (half rate, and what bounded this kernel at 2.7 TB/s) only extracts the bits
                    const u32 b4 = __vcmpeq4(clab[q], (u32)sh_lab[q][r + u] * 0x01010101u) & 0x01010101u;
                    p32[q][0] += d[u].x * __byte_perm(b4, 0u, 0x4440u);
                    p32[q][1] += d[u].y * __byte_perm(b4, 0u, 0x4441u);
                    p32[q][2] += d[u].z * __byte_perm(b4, 0u, 0x4442u);
                    p32[q][3] += d[u].w * __byte_perm(b4, 0u, 0x4443u);
                }
                if ((u % FLUSH) == FLUSH - 1) {
#pragma unroll
                    for (int q = 0; q < WB; ++q)
#pragma unroll
                        for (int c = 0; c < 4; ++c) { acc[q][c] += p32[q][c]; p32[q][c] = 0; }
                }
            }
        }
    }
    if (!any) return;
#pragma unroll
    for (int q = 0; q < WB; ++q)
        if (q < nq)
#pragma unroll
            for (int c = 0; c < 4; ++c) red_add_u64(S + (size_t)q * ld + col + c, acc[q][c]);
}

// T_c += S[h] for the cells h of cluster c; n_c
__global__ void k_wdisp_cols(const int *__restrict__ subset, int m, const unsigned char *__restrict__ lab_col,
                             const u64 *__restrict__ S, u64 *grp_sum, int *grp_cnt)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < m) {
        const int h = subset[p], c = lab_col[h];
        atomicAdd(&grp_sum[c], S[h]);
        atomicAdd(&grp_cnt[c], 1);
    }
}

// ---- W.k of ALL partitions of a sweep from ONE pass over the matrix: atoms ------------------------------------------
// The partitions PAM(1..Kmax) of one matrix are strongly related (mostly one cluster split per k), so the common
// refinement of all of them -- the "atoms": cells with the same label under every k -- has few classes (A ~ Kmax on
// clustered data).  With the rows sorted by atom, ONE grouped-column-sum pass (k_scan<2>, caps (0, inf)) gives
// AQ[a][h] = sum_{j in atom a} D[j][h], and every partition's within-cluster sums follow from that small table:
//     T_q[c] = sum_{h in c} sum_{a subset of c} AQ[a][h]
// (exact integers, the same numbers k_wk_batch + k_wdisp_cols produce with one 6-partition pass per 6 values of k).
// The atoms are refined on the device, one partition at a time: key = atom * k + label, dense ranks of the used keys.
struct AtomState { int A, overflow, pad0, pad1; };

__global__ void k_atom_init(int *atom, int m, AtomState *as)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < m) atom[p] = 0;
    if (p == 0) { as->A = 1; as->overflow = 0; }
}

__global__ void k_atom_mark(const int *__restrict__ atom, const int *__restrict__ lab_pos, int m, int k, int *table,
                            const AtomState *as)
{
    if (as->overflow) return;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < m) table[(size_t)atom[p] * k + lab_pos[p]] = 1;
}

// table[0 .. A*k): 1 = key in use  ->  1-based dense rank of the key (0 stays 0); A := number of keys in use
__global__ void __launch_bounds__(1024) k_atom_rank(int *table, int k, AtomState *as, int a_cap)
{
    if (as->overflow) return;
    __shared__ int sh[1024];
    const int n = as->A * k;
    const int per = (n + 1023) / 1024;
    const int b = threadIdx.x * per, e = min(n, b + per);
    int cnt = 0;
    for (int i = b; i < e; ++i) cnt += table[i];
    sh[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) { // inclusive scan of the per-thread counts
        int v = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += v;
        __syncthreads();
    }
    int run = sh[threadIdx.x] - cnt;
    for (int i = b; i < e; ++i)
        if (table[i]) table[i] = ++run;
    __syncthreads();
    if (threadIdx.x == 0) {
        const int tot = sh[1023];
        if (tot > a_cap) as->overflow = 1;
        else as->A = tot;
    }
}

// atom := rank of its key (the host clears the table before the next partition's marks)
__global__ void k_atom_relabel(int *atom, const int *__restrict__ lab_pos, int m, int k, const int *__restrict__ table,
                               const AtomState *as)
{
    if (as->overflow) return;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < m) atom[p] = table[(size_t)atom[p] * k + lab_pos[p]] - 1;
}

// counting sort of the owned rows by atom: histogram, exclusive prefix (one block), scatter with the caps (0, inf) that
// turn k_scan<2> into plain grouped column sums
__global__ void k_atom_hist(const int *__restrict__ atom, const int *__restrict__ rows_pos, int nls, int *hist)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nls) atomicAdd(&hist[atom[rows_pos[t]]], 1);
}
__global__ void __launch_bounds__(1024) k_atom_prefix(int *hist, int A)
{
    __shared__ int sh[1024];
    const int per = (A + 1023) / 1024;
    const int b = threadIdx.x * per, e = min(A, b + per);
    int cnt = 0;
    for (int i = b; i < e; ++i) cnt += hist[i];
    sh[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int v = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += v;
        __syncthreads();
    }
    int run = sh[threadIdx.x] - cnt;
    for (int i = b; i < e; ++i) { int c = hist[i]; hist[i] = run; run += c; }
}
__global__ void k_atom_scatter(const int *__restrict__ atom, const int *__restrict__ rows, const int *__restrict__ rows_pos,
                               int nls, int *cursor, int *s_row, int *s_grp, u32 *s_capP, u32 *s_capQ)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nls) {
        const int a = atom[rows_pos[t]];
        const int i = atomicAdd(&cursor[a], 1);
        s_row[i] = rows[t];
        s_grp[i] = a;
        s_capP[i] = 0u;
        s_capQ[i] = 0xFFFFFFFFu;
    }
}

// atomlab[q][a] = label of atom a under partition q (every cell of an atom carries the same one)
__global__ void k_atom_labels(const int *__restrict__ atom, const int *__restrict__ subset, int m,
                              const unsigned char *__restrict__ lab_col, size_t lab_pitch, int nq, int A,
                              unsigned char *__restrict__ atomlab)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < m) {
        const int a = atom[p];
        const size_t h = (size_t)subset[p];
        for (int q = 0; q < nq; ++q) atomlab[(size_t)q * A + a] = lab_col[(size_t)q * lab_pitch + h];
    }
}

// T[q][c] += sum_{h in c} sum_{a in c} AQ[a][h]; cnt[q][c] = cells of cluster c.  Thread = one column h and WKA_PB
// partitions; blockIdx.y walks a chunk of atoms, blockIdx.z a batch of partitions.
#define WKA_PB 8
#define WKA_ACH 512
__global__ void __launch_bounds__(256)
k_wk_atoms(const int *__restrict__ subset, int m, const unsigned char *__restrict__ lab_col, size_t lab_pitch, int nq,
           const u64 *__restrict__ AQ, size_t ld, int A, const unsigned char *__restrict__ atomlab, int kmax,
           u64 *__restrict__ T /* [nq][kmax] */, int *__restrict__ cnt /* [nq][kmax] */)
{
    __shared__ unsigned char sh_al[WKA_PB][WKA_ACH];
    __shared__ u64 sh_T[WKA_PB * 256];
    __shared__ int sh_C[WKA_PB * 256];
    const int q0 = blockIdx.z * WKA_PB, nqb = min(WKA_PB, nq - q0);
    const int a0 = blockIdx.y * WKA_ACH, na = min(WKA_ACH, A - a0);
    for (int i = threadIdx.x; i < WKA_PB * WKA_ACH; i += blockDim.x) {
        const int qq = i / WKA_ACH, a = i % WKA_ACH;
        sh_al[qq][a] = (qq < nqb && a < na) ? atomlab[(size_t)(q0 + qq) * A + a0 + a] : 0xFE;
    }
    for (int i = threadIdx.x; i < WKA_PB * 256; i += blockDim.x) { sh_T[i] = 0ull; sh_C[i] = 0; }
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < m) {
        const size_t h = (size_t)subset[p];
        u32 l[WKA_PB];
        u64 acc[WKA_PB];
#pragma unroll
        for (int qq = 0; qq < WKA_PB; ++qq) {
            l[qq] = qq < nqb ? (u32)lab_col[(size_t)(q0 + qq) * lab_pitch + h] : 0xFFu;
            acc[qq] = 0ull;
        }
        const u64 *col = AQ + (size_t)a0 * ld + h;
        for (int a = 0; a < na; ++a) {
            const u64 v = col[(size_t)a * ld];
#pragma unroll
            for (int qq = 0; qq < WKA_PB; ++qq) acc[qq] += (u32)sh_al[qq][a] == l[qq] ? v : 0ull;
        }
#pragma unroll
        for (int qq = 0; qq < WKA_PB; ++qq)
            if (qq < nqb) {
                atomicAdd(reinterpret_cast<unsigned long long *>(&sh_T[qq * 256 + l[qq]]), (unsigned long long)acc[qq]);
                if (blockIdx.y == 0) atomicAdd(&sh_C[qq * 256 + l[qq]], 1);
            }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < WKA_PB * 256; i += blockDim.x) {
        const int qq = i >> 8, c = i & 255;
        if (qq < nqb && c < kmax) {
            if (sh_T[i]) atomicAdd(reinterpret_cast<unsigned long long *>(&T[(size_t)(q0 + qq) * kmax + c]), (unsigned long long)sh_T[i]);
            if (sh_C[i]) atomicAdd(&cnt[(size_t)(q0 + qq) * kmax + c], sh_C[i]);
        }
    }
}

// deterministic sum of n doubles (fixed order) -> out[0]
__global__ void k_sum_f64(const double *__restrict__ x, int n, double *out)
{
    __shared__ double sh[1024];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

// the final-partition refresh (RaceID.R:1049-1052): nearest of the given medoid columns, first minimum
__global__ void k_refresh(const u32 *__restrict__ D, size_t ld, int row0, int nloc, const int *__restrict__ med,
                          int k, int *labels)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nloc) {
        const u32 *row = D + (size_t)t * ld;
        u32 best = row[med[0]];
        int g = 0;
        for (int c = 1; c < k; ++c) {
            u32 d = row[med[c]];
            if (d < best) { best = d; g = c; }
        }
        labels[row0 + t] = g + 1;
    }
}

// D[subset, subset] -> compact copy, fused with BUILD's step 0.  A thread owns 4 consecutive COMPACT columns (their
// source columns are looked up once), walks a chunk of owned subset rows gathering 4 x 8 independent 32-bit loads (the
// gathers of neighbouring lanes fall into the same sectors, so DRAM sees whole rows streamed once) and writes one
// aligned 128-bit store per row.  The column sums it keeps are exactly BUILD's first pass:
// SP0[c] += sum_rows D[row][c].
#define COMPACT_ROWS 64
__global__ void __launch_bounds__(128)
k_compact(const u32 *__restrict__ D, size_t ld, const int *__restrict__ src_rows, int nrows,
          const int *__restrict__ ccells, u32 *__restrict__ out, size_t ldc, int m, u64 *__restrict__ SP0)
{
    __shared__ int sh_src[COMPACT_ROWS];
    const int r0 = blockIdx.x * COMPACT_ROWS;
    const int nr = min(COMPACT_ROWS, nrows - r0);
    for (int i = threadIdx.x; i < COMPACT_ROWS; i += blockDim.x) sh_src[i] = i < nr ? src_rows[r0 + i] : 0;
    __syncthreads();
    const size_t c0 = ((size_t)blockIdx.y * blockDim.x + threadIdx.x) * 4;
    if (c0 >= ldc) return;
    int src[4];
    bool keep[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        keep[c] = c0 + c < (size_t)m;
        src[c] = keep[c] ? ccells[c0 + c] : 0;
    }
    u64 sum[4] = {0, 0, 0, 0};
    for (int r = 0; r < nr; r += 8) {
        u32 v[8][4];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const u32 *row = D + (size_t)sh_src[min(r + u, nr - 1)] * ld;
#pragma unroll
            for (int c = 0; c < 4; ++c) v[u][c] = __ldg(row + src[c]);
        }
        u32 p32[4] = {0, 0, 0, 0};
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (r + u < nr) {
                uint4 o = make_uint4(keep[0] ? v[u][0] : 0u, keep[1] ? v[u][1] : 0u, keep[2] ? v[u][2] : 0u,
                                     keep[3] ? v[u][3] : 0u);
                *reinterpret_cast<uint4 *>(out + (size_t)(r0 + r + u) * ldc + c0) = o;
#pragma unroll
                for (int c = 0; c < 4; ++c) p32[c] += v[u][c];
            }
            if (u & 1) { // two rows of q < 2^31 fit 32 bits
#pragma unroll
                for (int c = 0; c < 4; ++c) { sum[c] += p32[c]; p32[c] = 0; }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (keep[c]) red_add_u64(SP0 + c0 + c, sum[c]);
}

// Symmetric variant for a complete matrix view (one GPU, or a full replica): D[subset, subset] is symmetric and its rows
// and columns are the same sorted cell list, so only super-tiles on or above the diagonal are gathered from the
// resident matrix (half the source bytes) and every off-diagonal super-tile is written twice: as it is, straight from
// registers, and transposed through a swizzled shared-memory stage so that both writes are whole 128-byte lines.
// BUILD's step-0 column sums come from the same registers: column sums of the tile for its own columns, row sums of
// the tile for the columns of the mirrored tile.
#define CS_S 512       // super-tile edge (compact rows and columns)
#define CS_R 32        // rows staged per transposition
#define CS_THREADS 256
// CAPS: `cap` (by compact index) holds D[j][h0] for the first medoid h0 and the sums become BUILD's second-step sums
// sum_j min(D[j][h], cap_j): the tile's column sums with the caps of its rows, the mirrored tile's with those of its columns.
#define CS_SMEM (CS_S * CS_R * 4 + CS_S * 8 + CS_S * 4 + 2 * CS_S * 4)
template <bool CAPS>
__global__ void __launch_bounds__(CS_THREADS, 2)
k_compact_sym(const u32 *__restrict__ D, size_t ld, const int *__restrict__ ccells, u32 *__restrict__ out, size_t ldc,
              int m, u64 *__restrict__ SP0, const u32 *__restrict__ cap)
{
    const int bj = blockIdx.x, bi = blockIdx.y;
    if (bi > bj) return;
    extern __shared__ __align__(16) unsigned char cs_smem[];
    uint4 *tile = reinterpret_cast<uint4 *>(cs_smem);                   // [CS_S columns][8 chunks of 4 rows], swizzled
    u64 *rowacc = reinterpret_cast<u64 *>(cs_smem + CS_S * CS_R * 4);   // [CS_S] row sums of the super-tile
    int *sh_src = reinterpret_cast<int *>(rowacc + CS_S);               // [CS_S] source row of each tile row
    u32 *rowcap = reinterpret_cast<u32 *>(sh_src + CS_S);               // [CS_S] caps of the tile's rows / columns (CAPS)
    u32 *colcap = rowcap + CS_S;
    const bool mirror = bi < bj;
    const int tid = threadIdx.x;
    const int I0 = bi * CS_S, J0 = bj * CS_S;
    const int nr = min(CS_S, m - I0);
    for (int i = tid; i < CS_S; i += CS_THREADS) {
        sh_src[i] = ccells[I0 + min(i, nr - 1)];
        rowacc[i] = 0ull;
        if (CAPS) {
            rowcap[i] = I0 + i < m ? cap[I0 + i] : 0u;
            colcap[i] = J0 + i < m ? cap[J0 + i] : 0u;
        }
    }
    __syncthreads();
    const int t = tid & 127, h = tid >> 7;
    const size_t c0 = (size_t)J0 + 4 * (size_t)t;
    const bool col_in = c0 < ldc;
    int src[4];
    bool keep[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        keep[c] = c0 + c < (size_t)m;
        src[c] = keep[c] ? ccells[c0 + c] : 0;
    }
    u64 sum[4] = {0, 0, 0, 0};
    const int warp = tid >> 5, lane = tid & 31;
    for (int rs = 0; rs < nr; rs += CS_R) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int r = rs + 16 * h + 8 * g; // first of this thread's 8 rows, tile-relative
            u32 v[8][4];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const u32 *row = D + (size_t)sh_src[min(r + u, nr - 1)] * ld;
#pragma unroll
                for (int c = 0; c < 4; ++c) v[u][c] = __ldg(row + src[c]);
            }
            u32 p32[4] = {0, 0, 0, 0};
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool row_in = r + u < nr;
#pragma unroll
                for (int c = 0; c < 4; ++c) v[u][c] = (row_in && keep[c]) ? v[u][c] : 0u;
                if (row_in && col_in)
                    *reinterpret_cast<uint4 *>(out + (size_t)(I0 + r + u) * ldc + c0) = make_uint4(v[u][0], v[u][1], v[u][2], v[u][3]);
                if (CAPS) {
                    const u32 rc = rowcap[min(r + u, CS_S - 1)];
#pragma unroll
                    for (int c = 0; c < 4; ++c) p32[c] += min(v[u][c], rc);
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) p32[c] += v[u][c];
                }
                if (u & 1) { // two rows of q < 2^31 fit 32 bits
#pragma unroll
                    for (int c = 0; c < 4; ++c) { sum[c] += p32[c]; p32[c] = 0; }
                }
            }
            if (mirror) {
                const int q0 = 4 * h + 2 * g; // chunk (4 rows) of v[0..3]; v[4..7] is the next one
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int cc = 4 * t + c;
                    tile[cc * 8 + (q0 ^ (t & 7))] = make_uint4(v[0][c], v[1][c], v[2][c], v[3][c]);
                    tile[cc * 8 + ((q0 + 1) ^ (t & 7))] = make_uint4(v[4][c], v[5][c], v[6][c], v[7][c]);
                }
            }
        }
        if (mirror) { // (uniform over the CTA; a mirrored super-tile always has all CS_S rows)
            __syncthreads();
            u64 racc[4] = {0, 0, 0, 0};
            const int q = lane & 7;
#pragma unroll 4
            for (int p = 0; p < CS_S / 32; ++p) {
                const int cc = p * 32 + warp * 4 + (lane >> 3);
                const uint4 x = tile[cc * 8 + (q ^ ((cc >> 2) & 7))];
                if (J0 + cc < m) *reinterpret_cast<uint4 *>(out + (size_t)(J0 + cc) * ldc + I0 + rs + 4 * q) = x;
                if (CAPS) {
                    const u32 cc_cap = colcap[cc];
                    racc[0] += min(x.x, cc_cap); racc[1] += min(x.y, cc_cap); racc[2] += min(x.z, cc_cap); racc[3] += min(x.w, cc_cap);
                } else {
                    racc[0] += x.x; racc[1] += x.y; racc[2] += x.z; racc[3] += x.w;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                racc[j] += __shfl_xor_sync(0xffffffffu, racc[j], 8);
                racc[j] += __shfl_xor_sync(0xffffffffu, racc[j], 16);
            }
            if (lane < 8) {
#pragma unroll
                for (int j = 0; j < 4; ++j) atomicAdd(reinterpret_cast<unsigned long long *>(&rowacc[rs + 4 * q + j]), (unsigned long long)racc[j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (keep[c]) red_add_u64(SP0 + c0 + c, sum[c]);
    if (mirror)
        for (int i = tid; i < CS_S; i += CS_THREADS) red_add_u64(SP0 + I0 + i, rowacc[i]);
}

// clusterboot: k x k contingency table of (c1 label of the resampled cell, bootstrap label)
__global__ void k_contingency(const int *__restrict__ subset, const int *__restrict__ lab_pos, int m,
                              const int *__restrict__ part_slot /* [N] c1 group by cell */, int k, int *tab)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < m) atomicAdd(&tab[part_slot[subset[p]] * k + lab_pos[p]], 1);
}


// ---- lazy BUILD ---------------------------------------------------------------------------------------------------
// BUILD's gain of candidate h, g(h) = sum_j max(cap_j - D[j][h], 0), can only shrink when medoids are added (caps only
// drop), so a gain that was exact at an earlier step is an upper bound now.  From the third pick on, a complete view
// (one GPU holds every row, rows and columns are the same list) therefore evaluates exactly only the candidates that can
// still win -- by symmetry g(h) is a reduction over ROW h, one contiguous read -- instead of rescanning every row whose
// cap dropped:
//   stage A: the candidates whose cap did not drop in the last step and whose bound lies within 2^-6 of the largest
//            such bound are made exact; L = the best exact gain among them;
//   stage B: every candidate whose bound is still >= L is made exact.  Now every candidate that is not exact has
//            gain <= bound < L <= best, so the arg max over the exact ones (ties: largest position, like
//            `ammax <= beter[i]`) is the arg max over all candidates: the pick is identical to a full evaluation.
#define LZ_THREADS 256
#define LZ_LOCAL 512

// after the second pick's selection (SP = sum_j min(D[j][h], cap_j) is exact for the caps BEFORE that pick):
// bound[h] = sum_j cap_j - SP[h]
__global__ void __launch_bounds__(1024)
k_lazy_init(const u64 *__restrict__ SP, const u32 *__restrict__ capP, int m, u64 *ub, int *stamp, unsigned char *chg, u64 *thr)
{
    __shared__ u64 sh[32];
    u64 s = 0;
    for (int t = threadIdx.x; t < m; t += blockDim.x) s += capP[t];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0ull;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) sh[0] = s;
    }
    __syncthreads();
    const u64 td = sh[0];
    for (int h = threadIdx.x; h < m; h += blockDim.x) {
        const u64 sp = SP[h];
        ub[h] = td >= sp ? td - sp : 0ull;
        stamp[h] = -1;
        chg[h] = 0;
    }
    if (threadIdx.x == 0) { thr[0] = 0ull; thr[1] = 0ull; }
}

struct LzBest { u64 g; int pos, h; };
__device__ __forceinline__ bool lz_better(const LzBest &a, const LzBest &b) // a beats b
{
    if (a.h < 0) return false;
    if (b.h < 0) return true;
    if (a.g != b.g) return a.g > b.g;
    return a.pos > b.pos; // upstream BUILD: the largest index wins ties
}
__device__ __forceinline__ LzBest lz_block_best(LzBest v, LzBest *sh)
{
    for (int o = 16; o > 0; o >>= 1) {
        LzBest b;
        b.g = __shfl_down_sync(0xffffffffu, v.g, o);
        b.pos = __shfl_down_sync(0xffffffffu, v.pos, o);
        b.h = __shfl_down_sync(0xffffffffu, v.h, o);
        if (lz_better(b, v)) v = b;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        LzBest e;
        e.g = 0; e.pos = -1; e.h = -1;
        v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : e;
        for (int o = 16; o > 0; o >>= 1) {
            LzBest b;
            b.g = __shfl_down_sync(0xffffffffu, v.g, o);
            b.pos = __shfl_down_sync(0xffffffffu, v.pos, o);
            b.h = __shfl_down_sync(0xffffffffu, v.h, o);
            if (lz_better(b, v)) v = b;
        }
        if (threadIdx.x == 0) sh[0] = v;
    }
    __syncthreads();
    return sh[0];
}

// One CTA.  finish_step >= 0: the pick of that step = arg max of the gains that are exact at it; the medoid is
// installed and the caps drop along its row (D is symmetric: column h* of the point rows is row h*).
// prepare != 0: thresholds of the next step's stage A.
__global__ void __launch_bounds__(1024)
k_lazy_serial(const u32 *__restrict__ M, size_t ld, int m, const int *__restrict__ pos_of, u32 *capP, unsigned char *is_med,
              int *med, int *medpos, PamState *st, u64 *ub, const int *__restrict__ stamp, unsigned char *chg, u64 *thr,
              int finish_step, int prepare)
{
    __shared__ LzBest shb[32];
    if (finish_step >= 0) {
        LzBest v;
        v.g = 0; v.pos = -1; v.h = -1;
        for (int h = threadIdx.x; h < m; h += blockDim.x) {
            if (stamp[h] == finish_step && !is_med[h]) {
                LzBest c;
                c.g = ub[h]; c.pos = pos_of[h]; c.h = h;
                if (lz_better(c, v)) v = c;
            }
        }
        v = lz_block_best(v, shb);
        const int hs = v.h; // (>= 0: stage A always evaluates at least the candidate with the largest bound)
        if (threadIdx.x == 0 && hs >= 0) {
            const int s = st->nmed;
            med[s] = hs;
            medpos[s] = v.pos;
            is_med[hs] = 1;
            st->nmed = s + 1;
            st->best_h = hs; st->best_slot = s; st->best_val = v.g;
        }
        __syncthreads();
        if (hs >= 0) {
            const u32 *row = M + (size_t)hs * ld;
            for (int t = threadIdx.x; t < m; t += blockDim.x) {
                const u32 d = row[t], old = capP[t];
                chg[t] = d < old ? 1 : 0;
                if (d < old) capP[t] = d;
            }
        }
        __syncthreads();
    }
    if (prepare) {
        // the largest bound among the candidates whose own cap did not just drop (those are the stale ones: the cluster
        // the last pick covers); without any such candidate, among all of them
        u64 a = 0, b = 0;
        for (int h = threadIdx.x; h < m; h += blockDim.x) {
            if (is_med[h]) continue;
            const u64 u = ub[h];
            b = u > b ? u : b;
            if (!chg[h]) a = u > a ? u : a;
        }
        LzBest v;
        v.g = a; v.pos = 0; v.h = 0;
        v = lz_block_best(v, shb);
        a = v.g;
        v.g = b; v.pos = 0; v.h = 0;
        v = lz_block_best(v, shb);
        b = v.g;
        if (threadIdx.x == 0) {
            const u64 top = a ? a : b;
            thr[0] = top - (top >> 6);
            thr[1] = 0ull;  // L: best exact gain of the step (atomicMax by the evaluations)
            thr[3] = a ? 0ull : 1ull; // 1: no unchanged candidate is left, stage A looks at all of them
        }
    }
}

// stage 0 (A): candidates with bound >= thr[0] whose cap did not just drop; stage 1 (B): bound >= thr[1] = L.
// A CTA checks a strided share of the candidates, then makes its eligible ones exact, one row each.
__global__ void __launch_bounds__(LZ_THREADS)
k_lazy_eval(const u32 *__restrict__ M, size_t ld, int m, const u32 *__restrict__ capP, const unsigned char *__restrict__ is_med,
            const unsigned char *__restrict__ chg, u64 *ub, int *stamp, int step, int stage, u64 *thr, u64 *prof_dev)
{
    __shared__ int sh_list[LZ_LOCAL];
    __shared__ int sh_n;
    __shared__ u64 sh_red[LZ_THREADS / 32];
    if (threadIdx.x == 0) sh_n = 0;
    __syncthreads();
    const u64 t = thr[stage];
    const bool all = stage == 1 || thr[3] != 0ull;
    for (int i = threadIdx.x; (long long)i * gridDim.x + blockIdx.x < m; i += LZ_THREADS) {
        const int h = i * gridDim.x + blockIdx.x;
        if (!is_med[h] && stamp[h] != step && ub[h] >= t && (all || !chg[h])) {
            const int q = atomicAdd(&sh_n, 1);
            if (q < LZ_LOCAL) sh_list[q] = h;
        }
    }
    __syncthreads();
    const int n = min(sh_n, LZ_LOCAL); // (the grid is sized so that a CTA's share never exceeds LZ_LOCAL)
    const int m4 = m & ~3;
    for (int q = 0; q < n; ++q) {
        const int h = sh_list[q];
        const u32 *row = M + (size_t)h * ld;
        u64 acc = 0;
        for (int c = threadIdx.x * 4; c < m4; c += LZ_THREADS * 4) {
            const uint4 d = ld_stream(row + c);
            const uint4 cp = *reinterpret_cast<const uint4 *>(capP + c);
            acc += (u64)(cp.x - min(d.x, cp.x)) + (u64)(cp.y - min(d.y, cp.y)) + (u64)(cp.z - min(d.z, cp.z)) +
                   (u64)(cp.w - min(d.w, cp.w));
        }
        if (threadIdx.x < m - m4) {
            const u32 d = row[m4 + threadIdx.x], cp = capP[m4 + threadIdx.x];
            acc += (u64)(cp - min(d, cp));
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sh_red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            u64 g = 0;
#pragma unroll
            for (int wv = 0; wv < LZ_THREADS / 32; ++wv) g += sh_red[wv];
            ub[h] = g;
            stamp[h] = step;
            atomicMax((unsigned long long *)&thr[1], (unsigned long long)g);
            if (prof_dev) {
                atomicAdd((unsigned long long *)prof_dev, (unsigned long long)m);
                atomicAdd((unsigned long long *)prof_dev + 1, (unsigned long long)ld);
            }
        }
    }
}


// ---- BUILD's first pick of MANY subsets in a few passes over the resident matrix ------------------------------------
// Step 0 of BUILD needs the column sums over the subset's rows, SP0_b[h] = sum_{j in S_b} D[j][h].  For the bootstrap
// replicates of one clusterboot call the subsets are known up front, so the sums of BP replicates are taken in ONE pass
// over the rows this rank owns (a row contributes to the replicates whose bit is set in its mask): 50 replicates cost 5
// passes instead of a share of 50 compaction passes, and each replicate's first medoid is known before its compaction.
#define CSB 10
template <int FLUSH>
__global__ void __launch_bounds__(SCAN_THREADS)
k_colsum_batch(const u32 *__restrict__ D, size_t ld, int nrows, int rows_per_chunk,
               const unsigned short *__restrict__ mask /* [nrows]: bit b = the row belongs to replicate b of the batch */,
               u64 *__restrict__ out /* [CSB][ld] */)
{
    __shared__ unsigned short sh_m[SCAN_MAXROWS];
    const size_t col = ((size_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * 4;
    const bool active = col < ld;
    u64 acc[CSB][4];
    u32 p32[CSB][4];
#pragma unroll
    for (int b = 0; b < CSB; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc[b][c] = 0; p32[b][c] = 0; }
    bool any = false;
    for (int chunk = blockIdx.y; (long long)chunk * rows_per_chunk < nrows; chunk += gridDim.y) {
        const int r_begin = chunk * rows_per_chunk;
        const int nr = min(nrows - r_begin, rows_per_chunk);
        const int nr8 = (nr + 7) & ~7;
        __syncthreads();
        for (int i = threadIdx.x; i < nr8; i += SCAN_THREADS) sh_m[i] = i < nr ? mask[r_begin + i] : (unsigned short)0;
        __syncthreads();
        if (!active) continue;
        any = true;
        for (int r = 0; r < nr8; r += 8) {
            uint4 d[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                d[u] = sh_m[r + u] ? ld_stream(D + (size_t)(r_begin + r + u) * ld + col) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned mk = sh_m[r + u]; // uniform across the CTA
#pragma unroll
                for (int b = 0; b < CSB; ++b)
                    if ((mk >> b) & 1u) { p32[b][0] += d[u].x; p32[b][1] += d[u].y; p32[b][2] += d[u].z; p32[b][3] += d[u].w; }
                if ((u % FLUSH) == FLUSH - 1) {
#pragma unroll
                    for (int b = 0; b < CSB; ++b)
#pragma unroll
                        for (int c = 0; c < 4; ++c) { acc[b][c] += p32[b][c]; p32[b][c] = 0; }
                }
            }
        }
    }
    if (!any) return;
#pragma unroll
    for (int b = 0; b < CSB; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (acc[b][c]) red_add_u64(out + (size_t)b * ld + col + c, acc[b][c]);
}

// The same sums on the tensor cores, 16 replicates per pass: SP0 = M x D is a matrix product of the 0/1 membership matrix
// M [16 replicates x rows] with the matrix itself, and D splits exactly into its four byte planes,
//     sum_j M[b][j] D[j][h] = sum_d 256^d  sum_j M[b][j] byte_d(D[j][h]),
// each an u8 x u8 -> s32 product that mma.sync.m16n8k32 evaluates without rounding (255 x rows < 2^31).  A warp owns 32
// columns: lane (g, t) streams rows 4t..4t+3 and 16+4t..16+4t+3 of every 32-row block at columns 4g..4g+3 (128-bit loads, a
// warp instruction covers 4 rows x 128 contiguous bytes), transposes 4 rows x 4 bytes per column into the B fragments of
// the four planes (8 PRMT) and issues 16 MMAs per block; the A fragments (membership bits of the block's rows) are
// built once per row chunk in shared memory.  The IMAD kernel above needs one multiply-add per element and replicate
// (FMA-pipe bound at 10 replicates per pass, 4.3 TB/s); this one is HBM bound.
#define CSM 16
#define CM_ROWS 512
__device__ __forceinline__ void mma_u8_16832(int (&c)[4], const uint4 &a, u32 b0, u32 b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint2 ld_stream2(const u32 *p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// MB = 1: 16 replicates per pass, a lane streams 4 columns (128-bit loads), a CTA of 4 warps owns 128 columns.
// MB = 2: 32 replicates per pass (two A fragments share every B fragment), a lane streams 2 columns (64-bit loads: the
//         same 64 accumulator registers), a CTA owns 64 columns -- half the passes over the matrix for 50 replicates.
template <int MB>
__global__ void __launch_bounds__(128)
k_colsum_mma(const u32 *__restrict__ D, size_t ld, int nrows, int rows_per_chunk /* multiple of 32, <= CM_ROWS */,
             const void *__restrict__ mask_v /* [nrows] u16 (MB = 1) / u32 (MB = 2): bit b = the row belongs to replicate b */,
             u64 *__restrict__ out /* [16 MB][ld] */)
{
    constexpr int NC = 4 / MB; // columns per lane
    __shared__ u32 sh_m[CM_ROWS];
    __shared__ uint4 sh_a[CM_ROWS / 32][MB][32]; // A fragments of every 32-row block, by lane
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const size_t col0 = (size_t)blockIdx.x * (32 * NC) + (size_t)warp * (8 * NC); // (ld is a multiple of 128)
    const u32 *base = D + col0 + NC * g;
    int acc[MB][NC][4][4]; // [replicate block][column of the lane][byte plane][C fragment]
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int pl = 0; pl < 4; ++pl)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[mb][c][pl][r] = 0;
    for (int chunk = blockIdx.y; (long long)chunk * rows_per_chunk < nrows; chunk += gridDim.y) {
        const int r_begin = chunk * rows_per_chunk;
        const int nr = min(nrows - r_begin, rows_per_chunk);
        const int nb32 = (nr + 31) >> 5;
        __syncthreads();
        for (int i = threadIdx.x; i < nb32 * 32; i += 128)
            sh_m[i] = i >= nr ? 0u : (MB == 1 ? (u32)reinterpret_cast<const unsigned short *>(mask_v)[r_begin + i]
                                              : reinterpret_cast<const u32 *>(mask_v)[r_begin + i]);
        __syncthreads();
        for (int i = threadIdx.x; i < nb32 * 32 * MB; i += 128) {
            const int mb = i / (nb32 * 32), j = i - mb * (nb32 * 32);
            const int gg = (j & 31) >> 2, tt = j & 3;
            const u32 *m = sh_m + (j & ~31);
            u32 a[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) { // a0: (replicate gg, rows 4tt..), a1: (gg+8, 4tt..), a2: (gg, 16+4tt..), a3: (gg+8, 16+4tt..)
                const int rp = 16 * mb + gg + ((r & 1) ? 8 : 0), r0 = ((r & 2) ? 16 : 0) + 4 * tt;
                u32 v = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) v |= ((m[r0 + e] >> rp) & 1u) << (8 * e);
                a[r] = v;
            }
            sh_a[j >> 5][mb][j & 31] = make_uint4(a[0], a[1], a[2], a[3]);
        }
        __syncthreads();
        // MB = 2 streams 64-bit words: two 32-row blocks are loaded before the first MMA so that as many bytes are in flight
        // per lane as with the 128-bit loads of MB = 1
        constexpr int NBLK = MB;
        for (int blk0 = 0; blk0 < nb32; blk0 += NBLK) {
            u32 d[NBLK][8][NC];
#pragma unroll
            for (int bb = 0; bb < NBLK; ++bb) {
                const int rb = r_begin + (blk0 + bb) * 32 + 4 * t;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int row = min(rb + (u < 4 ? u : 12 + u), nrows - 1); // (rows past the end carry mask 0)
                    if (MB == 1) {
                        const uint4 v = ld_stream(base + (size_t)row * ld);
                        d[bb][u][0] = v.x; d[bb][u][1 % NC] = v.y; d[bb][u][2 % NC] = v.z; d[bb][u][3 % NC] = v.w;
                    } else {
                        const uint2 v = ld_stream2(base + (size_t)row * ld);
                        d[bb][u][0] = v.x; d[bb][u][1 % NC] = v.y;
                    }
                }
            }
#pragma unroll
            for (int bb = 0; bb < NBLK; ++bb) {
                if (blk0 + bb >= nb32) break; // (its loads were clamped to valid rows)
                uint4 a[MB];
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) a[mb] = sh_a[blk0 + bb][mb][lane];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    u32 b0[4], b1[4];
#pragma unroll
                    for (int h = 0; h < 2; ++h) { // rows 4t.. (b0) and 16+4t.. (b1): 4 x 4 byte transpose
                        const u32 w0 = d[bb][4 * h][c], w1 = d[bb][4 * h + 1][c], w2 = d[bb][4 * h + 2][c], w3 = d[bb][4 * h + 3][c];
                        const u32 t0 = __byte_perm(w0, w1, 0x5140), t1 = __byte_perm(w2, w3, 0x5140);
                        const u32 t2 = __byte_perm(w0, w1, 0x7362), t3 = __byte_perm(w2, w3, 0x7362);
                        u32 *b = h ? b1 : b0;
                        b[0] = __byte_perm(t0, t1, 0x5410); b[1] = __byte_perm(t0, t1, 0x7632);
                        b[2] = __byte_perm(t2, t3, 0x5410); b[3] = __byte_perm(t2, t3, 0x7632);
                    }
#pragma unroll
                    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                        for (int pl = 0; pl < 4; ++pl) mma_u8_16832(acc[mb][c][pl], a[mb], b0[pl], b1[pl]);
                }
            }
        }
    }
    // C fragment r of lane (g, t): replicate g (+8 if r & 2), MMA column 2t + (r & 1) = the column group whose lanes fed it
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const u64 v = (u64)(u32)acc[mb][c][0][r] + ((u64)(u32)acc[mb][c][1][r] << 8) + ((u64)(u32)acc[mb][c][2][r] << 16) +
                              ((u64)(u32)acc[mb][c][3][r] << 24);
                const int rp = 16 * mb + g + ((r & 2) ? 8 : 0), n = 2 * t + (r & 1);
                if (v) red_add_u64(out + (size_t)rp * ld + col0 + NC * n + c, v);
            }
}

// one CTA per replicate: arg min of SP0 over the replicate's cells, ties to the LARGEST position (cand_better_build)
__global__ void __launch_bounds__(256)
k_first_medoid(const u64 *__restrict__ sp0 /* [.][ld] */, size_t ld, const int *__restrict__ cat, const long long *__restrict__ off,
               const int *__restrict__ rep_of_slot, int2 *__restrict__ first /* by replicate */, int N)
{
    __shared__ Cand sh[32];
    const int b = rep_of_slot[blockIdx.x];
    const int *cells = cat + off[b];
    const int m = (int)(off[b + 1] - off[b]);
    const u64 *sp = sp0 + (size_t)blockIdx.x * ld;
    Cand best;
    best.val = ~0ull; best.p = -1; best.mp = 0; best.slot = 0;
    for (int p = threadIdx.x; p < m; p += blockDim.x) {
        const int h = cells[p];
        if (h < 0 || h >= N) continue; // (the replicate's own host-side validation reports it)
        Cand x;
        x.val = sp[h]; x.p = p; x.mp = 0; x.slot = 0;
        if (cand_better_build(x, best)) best = x;
    }
    best = cand_reduce_block<true>(best, sh);
    if (threadIdx.x == 0) first[b] = best.p >= 0 ? make_int2(cells[best.p], best.p) : make_int2(0, 0);
}

// row masks of the batched first-pick passes, built from the samples on the device: bit (b mod CSB) of
// mask[b / CSB][row] is set iff owned row `row` belongs to replicate b (16-bit masks, OR-ed in as halves of 32-bit words)
__global__ void k_first_masks(const int *__restrict__ cat, const long long *__restrict__ off, int row0, int nloc, int csb,
                              unsigned int *__restrict__ mask32)
{
    // (csb > 16: one 32-bit mask per row)
    const int b = blockIdx.y;
    const int *cells = cat + off[b];
    const int len = (int)(off[b + 1] - off[b]);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < len; p += gridDim.x * blockDim.x) {
        const int r = cells[p] - row0;
        if (r >= 0 && r < nloc) {
            const size_t e = (size_t)(b / csb) * nloc + r;
            if (csb > 16) atomicOr(&mask32[e], 1u << (b % csb));
            else atomicOr(&mask32[e >> 1], (1u << (b % csb)) << ((e & 1) * 16));
        }
    }
}

// cap[c] = D[cell_c][h0] for the compact columns c of a replicate; read as D[min][max] so that a replica that holds only
// the upper part of the peers' row blocks serves it
__global__ void k_first_caps(const u32 *__restrict__ D, size_t ld, const int2 *__restrict__ first, const int *__restrict__ ccells,
                             int m, u32 *__restrict__ cap)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const int h0 = first->x, cell = ccells[c];
    const int lo = min(h0, cell), hi = max(h0, cell);
    cap[c] = D[(size_t)lo * ld + hi];
}

// the first medoid of a run whose compaction already produced the second step's sums: caps, medoid tables, state
__global__ void k_install_first(const int2 *__restrict__ first, const int *__restrict__ subset /* column by position */,
                                const u32 *__restrict__ cap, int nls, u32 *capP, int *med, int *medpos, unsigned char *is_med,
                                PamState *st)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nls) capP[t] = cap[t];
    if (t == 0) {
        const int p0 = first->y, c0 = subset[p0];
        med[0] = c0;
        medpos[0] = p0;
        is_med[c0] = 1;
        st->nmed = 1;
        st->best_h = c0; st->best_slot = 0;
    }
}

// ----------------------------------------------------------------------------------------------------------
// host orchestration
// ----------------------------------------------------------------------------------------------------------
static inline int nblk(int n, int b) { return (n + b - 1) / b; }

static int pam_check_na(rid_dmat *dm, const int *subset, int m)
{
    if (!dm->has_na) return RID_OK;
    if (dm->na_cell.empty()) {
        rid_set_error("No clustering performed, NAs in the computed dissimilarity matrix.");
        return RID_ERR_NA;
    }
    for (int p = 0; p < m; ++p) {
        int h = subset ? subset[p] : p;
        if (dm->na_cell[h]) {
            rid_set_error("No clustering performed, NAs in the computed dissimilarity matrix.");
            return RID_ERR_NA;
        }
    }
    return RID_OK;
}

// ---- subset installation -----------------------------------------------------------------------------------------
// The host tables of a subset depend only on (N, owned row range, subset): they can be prepared on another thread while
// the GPU is still busy with the previous bootstrap replicate.
#define RID_COMPACT_MIN_BYTES ((size_t)1 << 20)
static void host_subset_prepare(HostSubset *H, int N, int row0, int row1, const int *subset, int m, bool want_compact)
{
    H->err = 0;
    H->m = m;
    H->N = N;
    if (m < 1 || m > N) {
        H->err = RID_ERR_ARG;
        snprintf(H->errmsg, sizeof(H->errmsg), "subset size %d out of range", m);
        return;
    }
    H->cells.resize(m);
    H->pos_of.assign(N, -1);
    for (int p = 0; p < m; ++p) {
        int h = subset ? subset[p] : p;
        if (h < 0 || h >= N || H->pos_of[h] != -1) {
            H->err = RID_ERR_ARG;
            snprintf(H->errmsg, sizeof(H->errmsg), h < 0 || h >= N ? "subset index %d out of range" : "subset holds cell %d twice", h);
            return;
        }
        H->cells[p] = h;
        H->pos_of[h] = p;
    }
    H->want_compact = want_compact && m < N && (size_t)m * (size_t)m * sizeof(u32) >= RID_COMPACT_MIN_BYTES;
    // owned subset rows in cell order (any order works: all sums are exact integers)
    H->src.clear(); H->rows_pos.clear();
    for (int h = row0; h < row1; ++h)
        if (H->pos_of[h] >= 0) {
            H->src.push_back(h - row0);
            H->rows_pos.push_back(H->pos_of[h]);
        }
    if (H->want_compact) {
        // compact column of a cell = its rank among the subset's cells; positions index compact columns
        H->ccells.resize(m);
        H->subset_c.resize(m);
        H->pos_c.resize(m);
        int c = 0;
        for (int h = 0; h < N; ++h)
            if (H->pos_of[h] >= 0) {
                H->ccells[c] = h;
                H->subset_c[H->pos_of[h]] = c;
                H->pos_c[c] = H->pos_of[h];
                ++c;
            }
        H->iota.resize(H->src.size());
        for (size_t t = 0; t < H->iota.size(); ++t) H->iota[t] = (int)t;
    }
}

// want_compact: the caller will make many passes (PAM): a proper subset is copied once into a compact matrix
// D[subset, subset] (columns in sorted-cell order; tie-breaking still follows the POSITIONS) so that every later pass
// reads only the distances it needs instead of whole rows of the resident matrix.
// `pin`: optional page-locked arena of >= 8 N ints.  With it the table uploads are true asynchronous copies (the host
// does not wait for the stream to drain, as it does when the driver stages pageable memory) and H may be rewritten
// as soon as the call returns; the arena itself must stay untouched until the copies have run.
static inline const int *stage_table(int *&pin, const std::vector<int> &v)
{
    if (!pin) return v.data();
    int *dst = pin;
    memcpy(dst, v.data(), v.size() * sizeof(int));
    pin += v.size();
    return dst;
}

static bool replica_upper_suffices(const rid_dmat *parent, const std::vector<int> &ccells);
// a replica view whose parent has pulled only part of the peers' row blocks: fetch the rest and wait for it
static int replica_need_whole(rid_dmat *dm)
{
    if (!dm->parent || dm->parent->replica_state == 2) return RID_OK;
    RID_TRY(rid_replica_complete(dm->parent));
    RID_CUDA(cudaStreamWaitEvent(dm->ctx->stream, dm->parent->replica_ev, 0));
    return RID_OK;
}

static int pam_install_subset(rid_dmat *dm, PamWs *w, HostSubset &H, int *pin = nullptr, const int2 *first = nullptr)
{
    rid_ctx *ctx = dm->ctx;
    const int N = dm->N, m = H.m;
    if (H.err) { rid_set_error("%s", H.errmsg); return H.err; }
    w->m = m;
    w->nls = (int)H.src.size();
    bool compact = H.want_compact;
    if (compact) {
        const size_t ldc = round_up((size_t)m, 128);
        const size_t need = std::max<size_t>((size_t)w->nls, 1) * ldc * sizeof(u32);
        int ok = 1;
        if (need > w->cbuf_bytes) {
            if (w->cbuf) { rid_pool_free(ctx, w->cbuf, w->cbuf_bytes); w->cbuf = nullptr; w->cbuf_bytes = 0; }
            // head-room (bootstrap resamples differ by a fraction of a percent), rounded so that the context's
            // exact-size reuse pool hits again on the next clustexp() call
            size_t want = round_up(need + need / 32, (size_t)256 << 20);
            if (rid_pool_alloc(ctx, (void **)&w->cbuf, want) == cudaSuccess) w->cbuf_bytes = want;
            else { cudaGetLastError(); ok = 0; }
        }
        if (ctx->world > 1) { // every rank must take the same path
            int *flag = (int *)w->dscratch;
            int miss = ok ? 0 : 1;
            RID_CUDA(cudaMemcpyAsync(flag, &miss, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            RID_TRY(rid_allreduce_i32(ctx, flag, 1));
            RID_CUDA(cudaMemcpyAsync(&miss, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            RID_CUDA(cudaStreamSynchronize(ctx->stream));
            ok = miss == 0;
        }
        if (!ok) compact = false; // not enough memory somewhere: scan the resident matrix through the row list
        else {
            RID_CUDA(cudaMemcpyAsync(w->subset, stage_table(pin, H.subset_c), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            RID_CUDA(cudaMemcpyAsync(w->pos_of, stage_table(pin, H.pos_c), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            RID_CUDA(cudaMemcpyAsync(w->colmap, stage_table(pin, H.ccells), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            RID_CUDA(cudaMemcpyAsync(w->subset_cells, stage_table(pin, H.cells), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            // SP (BUILD step 0) is accumulated by the compaction pass
            RID_CUDA(cudaMemsetAsync(w->acc, 0, (4 + ldc) * sizeof(u64), ctx->stream));
            w->sp1_ready = false;
            if (w->nls) {
                RID_CUDA(cudaMemcpyAsync(w->rows, stage_table(pin, H.iota), (size_t)w->nls * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
                RID_CUDA(cudaMemcpyAsync(w->rows_pos, stage_table(pin, H.rows_pos), (size_t)w->nls * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
                RID_CUDA(cudaMemcpyAsync(w->src_rows, stage_table(pin, H.src), (size_t)w->nls * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
                // complete view (rows == columns == the sorted subset): gather the upper triangle only and mirror it
                const bool sym = w->nls == m && dm->row0 == 0 && dm->row1 == N && !getenv("RID_COMPACT_FULL");
                // (a replica that holds only the upper part of the peers' blocks serves the symmetric gather, nothing else)
                if (dm->parent && !(sym && replica_upper_suffices(dm->parent, H.ccells))) RID_TRY(replica_need_whole(dm));
                cudaEvent_t pe = rid_prof_begin(ctx, RID_PROF_COMPACT);
                if (sym) {
                    // (per device and cheap: set on every call rather than once per process)
                    // (tried: __launch_bounds__(256, 3) = 80 registers with spills, three CTAs per SM instead of two: 6.70 ms
                    // instead of 5.23 ms per launch -- the spills cost more than the extra resident warps give)
                    auto kc0 = k_compact_sym<false>;
                    auto kc1 = k_compact_sym<true>;
                    RID_CUDA(cudaFuncSetAttribute(kc0, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM));
                    RID_CUDA(cudaFuncSetAttribute(kc1, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM));
                    const unsigned nb = (unsigned)((m + CS_S - 1) / CS_S);
                    if (first) {
                        // the first medoid is known: its distances cap every term, the sums are those of BUILD's second step
                        LAUNCH(ctx, k_first_caps, nblk(m, 256), 256, 0, dm->D, dm->ld, first, w->colmap, m, w->firstcap);
                        LAUNCH(ctx, kc1, dim3(nb, nb), CS_THREADS, CS_SMEM, dm->D, dm->ld, w->colmap, w->cbuf, ldc, m,
                               w->acc + 4, w->firstcap);
                        w->sp1_ready = true;
                        w->first_dev = first;
                    } else {
                        LAUNCH(ctx, kc0, dim3(nb, nb), CS_THREADS, CS_SMEM, dm->D, dm->ld, w->colmap, w->cbuf, ldc, m,
                               w->acc + 4, (const u32 *)nullptr);
                    }
                    // useful bytes: the upper triangle gathered once + the square written; requested: the source spans
                    // of the gathered tiles (whole sectors: ~N/m source columns per compact column) + the square
                    const double tri = 0.5 * (double)m * ((double)m + 1.0);
                    rid_prof_end(ctx, pe, RID_PROF_COMPACT, (tri + (double)m * (double)m) * 4.0,
                                 (tri * (double)N / (double)m + (double)m * (double)ldc) * 4.0);
                } else {
                    dim3 grid((unsigned)((w->nls + COMPACT_ROWS - 1) / COMPACT_ROWS), (unsigned)((ldc + 511) / 512));
                    LAUNCH(ctx, k_compact, grid, 128, 0, dm->D, dm->ld, w->src_rows, w->nls, w->colmap, w->cbuf, ldc, m,
                           w->acc + 4);
                    rid_prof_end(ctx, pe, RID_PROF_COMPACT, (double)w->nls * ((double)N + (double)m) * 4.0,
                                 (double)w->nls * ((double)dm->ld + (double)ldc) * 4.0);
                }
                RID_CUDA(cudaGetLastError());
            }
            // pageable host buffers: the async copies above have been staged by the time the calls return
            w->mat = w->cbuf;
            w->ld = ldc;
            w->compact = true;
            w->sp0_ready = true;
            return RID_OK;
        }
    }
    // view mode: rows of the resident shard through a row list, columns masked by the candidate list
    RID_TRY(replica_need_whole(dm)); // (whole rows are read)
    w->compact = false;
    w->sp0_ready = false;
    w->sp1_ready = false;
    w->mat = dm->D;
    w->ld = dm->ld;
    const int *h_cells = stage_table(pin, H.cells);
    RID_CUDA(cudaMemcpyAsync(w->subset, h_cells, (size_t)m * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    RID_CUDA(cudaMemcpyAsync(w->subset_cells, h_cells, (size_t)m * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    RID_CUDA(cudaMemcpyAsync(w->pos_of, stage_table(pin, H.pos_of), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if (w->nls) {
        RID_CUDA(cudaMemcpyAsync(w->rows, stage_table(pin, H.src), (size_t)w->nls * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        RID_CUDA(cudaMemcpyAsync(w->rows_pos, stage_table(pin, H.rows_pos), (size_t)w->nls * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
    return RID_OK;
}

static int pam_set_subset(rid_dmat *dm, PamWs *w, const int *subset, int m, bool want_compact, HostSubset *prepared = nullptr)
{
    if (prepared) return pam_install_subset(dm, w, *prepared);
    host_subset_prepare(&w->hs, dm->N, dm->row0, dm->row1, subset, m, want_compact);
    int rc = pam_install_subset(dm, w, w->hs);
    // w->hs is reused by the next call: make sure the (pageable, hence already staged) copies are really done
    if (rc == RID_OK) RID_CUDA(cudaStreamSynchronize(dm->ctx->stream));
    return rc;
}

// nrows_ub: upper bound of the row count known on the host (grid sizing); nrows_dev: optional exact count on the
// device.  work_rows: rows to credit in the roofline accounting (<0: nrows_ub).
template <int MODE>
static int launch_scan(rid_dmat *dm, PamWs *w, const int *rows, const u32 *capP, const u32 *capQ, const int *grp,
                       int nrows_ub, const int *nrows_dev, int check_done, int prof_kind)
{
    rid_ctx *ctx = dm->ctx;
    if (nrows_ub <= 0) return RID_OK;
    int strips = (int)((w->ld + SCAN_THREADS * 4 - 1) / (SCAN_THREADS * 4));
    int want_y = std::max(1, (ctx->sm_count * 16 + strips - 1) / strips);
    int rpc = (nrows_ub + want_y - 1) / want_y;
    rpc = (rpc + 7) & ~7;
    rpc = std::min(std::max(rpc, 8), SCAN_MAXROWS);
    int chunks = (nrows_ub + rpc - 1) / rpc;
    dim3 grid(strips, std::min(chunks, want_y * 2));
    u64 *SP = w->acc + 4, *AQ = SP + w->ld;
    // algorithmic bytes: every distance of the (rows x subset) block once; traffic: whole rows are read
    cudaEvent_t pe = rid_prof_begin(ctx, prof_kind);
    auto scan8 = k_scan<MODE, 8>;
    auto scan2 = k_scan<MODE, 2>;
    if (dm->bits <= 28)
        LAUNCH(ctx, scan8, grid, SCAN_THREADS, 0, w->mat, w->ld, rows, capP, capQ, grp, nrows_ub, nrows_dev,
               rpc, SP, AQ, w->state, check_done);
    else
        LAUNCH(ctx, scan2, grid, SCAN_THREADS, 0, w->mat, w->ld, rows, capP, capQ, grp, nrows_ub, nrows_dev,
               rpc, SP, AQ, w->state, check_done);
    // a pass sized on the device (nrows_dev) is credited through ctx->prof_dev (k_select / k_prefix2)
    rid_prof_end(ctx, pe, prof_kind, nrows_dev ? 0.0 : (double)nrows_ub * (double)w->m * 4.0,
                 nrows_dev ? 0.0 : (double)nrows_ub * (double)w->ld * 4.0);
    RID_CUDA(cudaGetLastError());
    return RID_OK;
}

static int zero_acc(rid_dmat *dm, PamWs *w, int ngroups, int check_done)
{
    size_t n = 4 + (size_t)(1 + ngroups) * w->ld;
    int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)dm->ctx->sm_count * 8);
    blocks = std::max(blocks, nblk(3 * w->kcap, 256));
    LAUNCH(dm->ctx, k_zero, blocks, 256, 0, w->acc, n, w->hist, 3 * w->kcap, w->state, check_done);
    RID_CUDA(cudaGetLastError());
    return RID_OK;
}

// BUILD: k greedy picks.  Step 0 is one full pass (SP[h] = sum_j D[j][h]); every later step rescans only the rows whose
// nearest-medoid distance just dropped (about N/(s+1) of them) and corrects SP in place: ~N*H_k row reads instead of
// N*k, with bit-identical picks because all sums are exact integers.
static int pam_build(rid_dmat *dm, PamWs *w, int k_from, int k_to)
{
    rid_ctx *ctx = dm->ctx;
    u64 *SP = w->acc + 4, *AQ0 = SP + w->ld;
    auto sel_build = k_select<true>;
    // lazy evaluation from the third pick on (see k_lazy_eval): complete views only -- every row on this GPU, rows and
    // columns the same list in the same order (the compact copy, or the whole matrix)
    // OPT-IN (RID_BUILD_LAZY=1): exact, but measured 2.6x SLOWER than the incremental rescans on the benchmark data
    // (19 ms instead of 6.5 ms of BUILD per bootstrap replicate at cfg4): in a few thousand dimensions every new medoid
    // lowers the caps of 1/(s+1) of ALL points, so every candidate's gain keeps shrinking and bounds that are a few
    // steps old certify nothing -- stage B ends up re-evaluating most rows.  Kept for low-dimensional / well-separated
    // inputs and as a cross-check of the incremental path (tests/test_gpu_round2.py).
    const bool lazy = k_from == 0 && k_to > 2 && ctx->world == 1 && w->nls == w->m && dm->row0 == 0 &&
                      (w->compact || w->m == dm->N) && getenv("RID_BUILD_LAZY") != nullptr;
    const int k_scan_to = lazy ? 2 : k_to;
    // the compaction already produced the sums of the second pick (first medoid known up front): install the first medoid
    // and go straight to that pick's selection
    const bool have_first = w->sp1_ready && w->sp0_ready && k_from == 0 && k_to >= 2;
    w->sp1_ready = false;
    if (have_first) {
        w->sp0_ready = false;
        LAUNCH(ctx, k_install_first, nblk(std::max(w->nls, 1), 256), 256, 0, w->first_dev, w->subset, w->firstcap, w->nls, w->capP,
               w->med, w->medpos, w->is_med, w->state);
        RID_CUDA(cudaMemsetAsync(AQ0, 0, w->ld * sizeof(u64), ctx->stream)); // the delta buffer of the next pick's rescan
    }
    for (int s = have_first ? 1 : k_from; s < k_scan_to; ++s) {
        if (have_first && s == 1) {
            // SP is exact already: nothing to rescan, nothing to subtract
        } else if (s == 0) {
            if (!w->sp0_ready) {
                RID_TRY(zero_acc(dm, w, 0, 0));
                RID_TRY((launch_scan<0>(dm, w, w->rows, w->capP, nullptr, nullptr, w->nls, nullptr, 0, RID_PROF_SCAN_BUILD)));
            }
            w->sp0_ready = false;
            RID_TRY(rid_allreduce_u64(ctx, w->acc, 4 + w->ld));
        } else {
            if (s == std::max(k_from, 1)) // the selection kernel leaves AQ0 zeroed for the next step
                RID_CUDA(cudaMemsetAsync(AQ0, 0, w->ld * sizeof(u64), ctx->stream));
            RID_TRY((launch_scan<2>(dm, w, w->s_row, w->s_capP, w->s_capQ, w->s_grp, w->nls, w->chg_count,
                                    getenv("RID_SCAN_STATIC") ? -1 : 0, RID_PROF_SCAN_BUILD)));
            RID_TRY(rid_allreduce_u64(ctx, AQ0, w->ld));
        }
        // picks the next medoid; for s > 0 it first applies SP -= AQ0 (and clears AQ0 and the changed-row counter)
        LAUNCH(ctx, sel_build, nblk(w->m, 256), 256, 0, w->subset, w->m, w->is_med, w->acc, w->ld, s, w->med, w->medpos,
               w->is_med, w->blockbest, w->state, 0, (s > 0 && !(have_first && s == 1)) ? AQ0 : (u64 *)nullptr, w->chg_count,
               ((ctx->prof_on >> RID_PROF_SCAN_BUILD) & 1) ? ctx->prof_dev : (u64 *)nullptr, (int *)nullptr);
        if (lazy && s == 1) // SP is exact for the caps before this pick: the gains it implies bound every later gain
            LAUNCH(ctx, k_lazy_init, 1, 1024, 0, SP, w->capP, w->m, w->lz_ub, w->lz_stamp, w->lz_chg, w->lz_thr);
        if (w->nls)
            LAUNCH(ctx, k_build_update, nblk(w->nls, 256), 256, 0, w->mat, w->ld, w->rows, w->nls, w->capP,
                   w->state, w->chg_count, w->s_row, w->s_capP, w->s_capQ, w->s_grp);
        RID_CUDA(cudaGetLastError());
    }
    if (lazy) {
        // a CTA of the evaluation kernels checks ceil(m / grid) candidates: at most LZ_LOCAL
        const int grid = std::max(std::min(w->m, ctx->sm_count * 8), (w->m + LZ_LOCAL - 1) / LZ_LOCAL);
        u64 *pd = ((ctx->prof_on >> RID_PROF_SCAN_BUILD) & 1) ? ctx->prof_dev : (u64 *)nullptr;
        for (int s = 2; s <= k_to; ++s) {
            // finishes step s - 1 (pick + caps) when that was a lazy step, prepares step s
            LAUNCH(ctx, k_lazy_serial, 1, 1024, 0, w->mat, w->ld, w->m, w->pos_of, w->capP, w->is_med, w->med, w->medpos,
                   w->state, w->lz_ub, w->lz_stamp, w->lz_chg, w->lz_thr, s > 2 ? s - 1 : -1, s < k_to ? 1 : 0);
            if (s == k_to) break;
            cudaEvent_t pe = rid_prof_begin(ctx, RID_PROF_SCAN_BUILD);
            for (int stage = 0; stage < 2; ++stage)
                LAUNCH(ctx, k_lazy_eval, grid, LZ_THREADS, 0, w->mat, w->ld, w->m, w->capP, w->is_med, w->lz_chg, w->lz_ub,
                       w->lz_stamp, s, stage, w->lz_thr, pd);
            rid_prof_end(ctx, pe, RID_PROF_SCAN_BUILD, 0.0, 0.0);
        }
        RID_CUDA(cudaGetLastError());
    }
    return RID_OK;
}

static int pam_begin(rid_dmat *dm, PamWs *w)
{
    rid_ctx *ctx = dm->ctx;
    LAUNCH(ctx, k_state_init, 1, 1, 0, w->state);
    RID_CUDA(cudaMemsetAsync(w->is_med, 0, (size_t)dm->N, ctx->stream));
    if (w->nls) LAUNCH(ctx, k_fill_u32, nblk(w->nls, 256), 256, 0, w->capP, 0xFFFFFFFFu, w->nls);
    RID_CUDA(cudaGetLastError());
    return RID_OK;
}

// one assign pass: caps, groups, TD -> acc[0] (all-reduced).  ngroups = k accumulators are zeroed too.
static int pam_assign(rid_dmat *dm, PamWs *w, int k, int check_done, u64 *acc = nullptr)
{
    rid_ctx *ctx = dm->ctx;
    if (!acc) acc = w->acc;
    {
        size_t n = 4 + (size_t)(1 + k) * w->ld;
        int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 8);
        blocks = std::max(blocks, nblk(3 * w->kcap, 256));
        LAUNCH(ctx, k_zero, blocks, 256, 0, acc, n, w->hist, 3 * w->kcap, w->state, check_done);
    }
    if (w->nls)
        LAUNCH(ctx, k_assign<false>, nblk(w->nls, 128), 128, 2 * k * sizeof(int), w->mat, w->ld, w->rows, w->nls,
               w->med, w->medpos, k, w->capP, w->capQ, w->grp, w->hist, acc, w->state, check_done, (u32 *)nullptr,
               (u32 *)nullptr, (int *)nullptr, (int *)nullptr);
    RID_CUDA(cudaGetLastError());
    return RID_OK;
}

static int sort_rows_by_group(rid_dmat *dm, PamWs *w, int k, int check_done)
{
    rid_ctx *ctx = dm->ctx;
    LAUNCH(ctx, k_prefix, 1, 32, 0, w->hist, w->kcap, w->state, check_done);
    if (w->nls)
        LAUNCH(ctx, k_scatter, nblk(w->nls, 256), 256, 0, w->rows, w->capP, w->capQ, w->grp, w->nls, w->hist, w->kcap,
               w->s_row, w->s_capP, w->s_capQ, w->s_grp, w->state, check_done);
    RID_CUDA(cudaGetLastError());
    return RID_OK;
}

// One incremental SWAP iteration (all kernels are no-ops once state.done is set).
// With do_select = false it only moves the accumulators and the assignment to the current medoid set (used by the
// k-sweep to step from the BUILD medoids of one k to those of the next).
static int pam_swap_delta_iteration(rid_dmat *dm, PamWs *w, int k, int *n_delta, bool do_select = true)
{
    rid_ctx *ctx = dm->ctx;
    const size_t nacc = 4 + (size_t)(1 + k) * w->ld;
    // one iteration = 7 launches (all early-exit once the local optimum is reached): zero acc2 / key counts; new
    // assignment (keeps the old one, counts the changed rows by key, TD -> acc2[0]); prefix; scatter; the delta pass; add; select
    {
        int blocks = (int)std::min<size_t>((nacc + 255) / 256, (size_t)ctx->sm_count * 8);
        blocks = std::max(blocks, nblk(3 * w->kcap, 256));
        LAUNCH(ctx, k_zero, blocks, 256, 0, w->acc2, nacc, w->hist, 3 * w->kcap, w->state, 1, w->hist2, k * k, w->chg_count);
    }
    if (w->nls)
        LAUNCH(ctx, k_assign<true>, nblk(w->nls, 128), 128, 2 * k * sizeof(int), w->mat, w->ld, w->rows, w->nls,
               w->med, w->medpos, k, w->capP, w->capQ, w->grp, w->hist, w->acc2, w->state, 1, w->o_capP, w->o_capQ, w->o_grp,
               w->hist2);
    if (w->nls) {
        LAUNCH(ctx, k_prefix2, 1, 32, 0, w->hist2, k * k, w->chg_count, w->state,
               ((ctx->prof_on >> RID_PROF_SCAN_DELTA) & 1) ? ctx->prof_dev : (u64 *)nullptr, w->m, w->ld);
        LAUNCH(ctx, k_diff_scatter, nblk(w->nls, 256), 256, 0, w->rows, w->capP, w->capQ, w->grp, w->o_capP, w->o_capQ,
               w->o_grp, w->nls, k, w->hist2, w->d_row, w->d_key, w->d_po, w->d_qo, w->d_pn, w->d_qn, w->state);
        int strips = (int)((w->ld + SCAN_THREADS * 4 - 1) / (SCAN_THREADS * 4));
        int want_y = std::max(1, (ctx->sm_count * 16 + strips - 1) / strips);
        int rpc = (w->nls + want_y - 1) / want_y;
        rpc = std::min(std::max((rpc + 7) & ~7, 8), 256); // six per-row arrays live in shared memory
        int chunks = (w->nls + rpc - 1) / rpc;
        dim3 grid(strips, std::min(chunks, want_y * 2));
        u64 *SP2 = w->acc2 + 4, *AQ2 = SP2 + w->ld;
        cudaEvent_t pe = rid_prof_begin(ctx, RID_PROF_SCAN_DELTA);
        // (tried: __launch_bounds__(128, 5) = 96 registers, five CTAs per SM instead of four: 0.323 vs 0.322 ms per launch,
        // no difference -- occupancy is not what holds this pass at 0.79 of the copy peak)
        auto d8 = k_scan_delta<8>;
        auto d2 = k_scan_delta<2>;
        if (dm->bits <= 28)
            LAUNCH(ctx, d8, grid, SCAN_THREADS, 0, w->mat, w->ld, w->d_row, w->d_key, w->d_po, w->d_qo, w->d_pn, w->d_qn,
                   w->chg_count, rpc, k, SP2, AQ2, w->state);
        else
            LAUNCH(ctx, d2, grid, SCAN_THREADS, 0, w->mat, w->ld, w->d_row, w->d_key, w->d_po, w->d_qo, w->d_pn, w->d_qn,
                   w->chg_count, rpc, k, SP2, AQ2, w->state);
        rid_prof_end(ctx, pe, RID_PROF_SCAN_DELTA, 0.0, 0.0);
        if ((ctx->prof_on >> RID_PROF_SCAN_DELTA) & 1) ++*n_delta;
    }
    RID_TRY(rid_allreduce_u64(ctx, w->acc2, nacc));
    {
        int blocks = (int)std::min<size_t>((nacc + 255) / 256, (size_t)ctx->sm_count * 8);
        LAUNCH(ctx, k_acc_add, blocks, 256, 0, w->acc, w->acc2, nacc, w->state);
    }
    if (do_select) {
        auto sel_swap = k_select<false>;
        LAUNCH(ctx, sel_swap, nblk(w->m, 256), 256, 0, w->subset, w->m, w->is_med, w->acc, w->ld, k, w->med, w->medpos,
               w->is_med, w->blockbest, w->state, 1, (u64 *)nullptr, (int *)nullptr, (u64 *)nullptr, (int *)nullptr);
    }
    RID_CUDA(cudaGetLastError());
    return RID_OK;
}

// First SWAP iteration up to (not including) the selection: assignment, TD of the BUILD medoids, full scan.  k >= 2.
static int pam_swap_first_scan(rid_dmat *dm, PamWs *w, int k)
{
    rid_ctx *ctx = dm->ctx;
    RID_TRY(pam_assign(dm, w, k, 0));
    // objective["build"]: TD of the BUILD medoids
    if (ctx->world > 1) RID_TRY(rid_allreduce_u64(ctx, w->acc, 4));
    LAUNCH(ctx, k_capture_td_build, 1, 1, 0, w->state, w->acc);
    // acc[0] now holds the global TD on every rank; the big all-reduce below would add it world times, so keep only
    // this rank's share: rank 0 keeps it, the others restart from zero.
    if (ctx->world > 1 && ctx->rank != 0) RID_CUDA(cudaMemsetAsync(w->acc, 0, sizeof(u64), ctx->stream));
    RID_TRY(sort_rows_by_group(dm, w, k, 0));
    RID_TRY((launch_scan<1>(dm, w, w->s_row, w->s_capP, w->s_capQ, w->s_grp, w->nls, nullptr, 0, RID_PROF_SCAN_SWAP)));
    RID_TRY(rid_allreduce_u64(ctx, w->acc, 4 + (size_t)(1 + k) * w->ld));
    return RID_OK;
}

// SWAP until no improving swap is left.  Medoids (k of them) must be installed in w->med / is_med.
// On return: state.td_build = TD before the first swap, state.td = final TD, w->grp / capP / capQ describe the
// final medoids for the owned rows.
// The first SWAP iteration (full scan) and `ahead` incremental iterations, enqueued without any host round trip: every
// kernel after the local optimum is a no-op (state.done).  One rank, 2 <= k, k*k <= K2MAX only.  The caller reads
// state.done later and, in the rare case that more swaps were needed, resumes with pam_swap(..., resume = true).
static int pam_swap_enqueue(rid_dmat *dm, PamWs *w, int k, int ahead, int *n_delta)
{
    rid_ctx *ctx = dm->ctx;
    RID_TRY(pam_assign(dm, w, k, 0));
    LAUNCH(ctx, k_capture_td_build, 1, 1, 0, w->state, w->acc);
    RID_TRY(sort_rows_by_group(dm, w, k, 0));
    RID_TRY((launch_scan<1>(dm, w, w->s_row, w->s_capP, w->s_capQ, w->s_grp, w->nls, nullptr, 0, RID_PROF_SCAN_SWAP)));
    auto sel_swap = k_select<false>;
    LAUNCH(ctx, sel_swap, nblk(w->m, 256), 256, 0, w->subset, w->m, w->is_med, w->acc, w->ld, k, w->med, w->medpos,
           w->is_med, w->blockbest, w->state, 0, (u64 *)nullptr, (int *)nullptr, (u64 *)nullptr, (int *)nullptr);
    RID_CUDA(cudaGetLastError());
    for (int i = 0; i < ahead; ++i) RID_TRY(pam_swap_delta_iteration(dm, w, k, n_delta));
    return RID_OK;
}

static int pam_swap(rid_dmat *dm, PamWs *w, int k, PamState *h_state, bool resume = false)
{
    rid_ctx *ctx = dm->ctx;
    // how many iterations to enqueue between host checks: per-iteration work is ~ nls*ld*4 bytes
    double bytes = (double)w->nls * (double)w->ld * 4.0;
    int batch = bytes > 2.0e9 ? 1 : (bytes > 2.0e8 ? 2 : 8);
    bool first = !resume;
    const bool incremental = k >= 2 && k * k <= K2MAX && !getenv("RID_SWAP_FULL");
    if (incremental) batch = std::max(batch, 2);
    int n_delta = 0;
    for (;;) {
        for (int b = 0; b < batch; ++b) {
            if (incremental && !first) { RID_TRY(pam_swap_delta_iteration(dm, w, k, &n_delta)); continue; }
            int chk = first ? 0 : 1;
            RID_TRY(pam_assign(dm, w, k, chk));
            if (first) {
                // objective["build"]: TD of the BUILD medoids
                if (ctx->world > 1) RID_TRY(rid_allreduce_u64(ctx, w->acc, 4));
                LAUNCH(ctx, k_capture_td_build, 1, 1, 0, w->state, w->acc);
                if (ctx->world > 1) {
                    // acc[0] now holds the global TD on every rank; the big all-reduce below would add it world
                    // times, so keep only this rank's share: rank 0 keeps it, the others restart from zero.
                    if (ctx->rank != 0) RID_CUDA(cudaMemsetAsync(w->acc, 0, sizeof(u64), ctx->stream));
                }
                first = false;
            }
            if (k < 2) break;
            {
                RID_TRY(sort_rows_by_group(dm, w, k, chk));
                RID_TRY((launch_scan<1>(dm, w, w->s_row, w->s_capP, w->s_capQ, w->s_grp, w->nls, nullptr, chk, RID_PROF_SCAN_SWAP)));
                RID_TRY(rid_allreduce_u64(ctx, w->acc, 4 + (size_t)(1 + k) * w->ld));
                auto sel_swap = k_select<false>;
                LAUNCH(ctx, sel_swap, nblk(w->m, 256), 256, 0, w->subset, w->m, w->is_med, w->acc, w->ld, k,
                       w->med, w->medpos, w->is_med, w->blockbest, w->state, chk, (u64 *)nullptr, (int *)nullptr, (u64 *)nullptr, (int *)nullptr);
                RID_CUDA(cudaGetLastError());
            }
            // every later iteration only rescans the rows whose (nearest, second nearest) pair changed
            while (incremental && ++b < batch) RID_TRY(pam_swap_delta_iteration(dm, w, k, &n_delta));
        }
        RID_CUDA(cudaMemcpyAsync(h_state, w->state, sizeof(PamState), cudaMemcpyDeviceToHost, ctx->stream));
        RID_CUDA(cudaStreamSynchronize(ctx->stream));
        if (k < 2) {
            h_state->td = h_state->td_build;
            break;
        }
        if (h_state->done) { if (ctx->rank == 0) ctx->swaps += h_state->nswaps; break; } // (cooperating ranks count once)
        RID_TRY(rid_check_abort(ctx, false));
    }
    return RID_OK;
}

// labels by position for the current medoids -> w->lab_pos (device, all ranks) and host copy
static int pam_collect_labels(rid_dmat *dm, PamWs *w, std::vector<int> &h_lab)
{
    rid_ctx *ctx = dm->ctx;
    RID_CUDA(cudaMemsetAsync(w->lab_pos, 0, (size_t)w->m * sizeof(int), ctx->stream));
    if (w->nls) LAUNCH(ctx, k_labels_to_pos, nblk(w->nls, 256), 256, 0, w->rows_pos, w->grp, w->nls, w->lab_pos);
    RID_CUDA(cudaGetLastError());
    RID_TRY(rid_allreduce_i32(ctx, w->lab_pos, (size_t)w->m));
    h_lab.resize(w->m);
    RID_CUDA(cudaMemcpyAsync(h_lab.data(), w->lab_pos, (size_t)w->m * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RID_CUDA(cudaStreamSynchronize(ctx->stream));
    return RID_OK;
}

// grouped column sums for the groups in w->lab_pos (0..k-1 by position): AQ[c][h] = sum_{j in c} D[j][h]
static int grouped_colsums(rid_dmat *dm, PamWs *w, int k)
{
    rid_ctx *ctx = dm->ctx;
    RID_TRY(zero_acc(dm, w, k, 0));
    if (w->nls)
        LAUNCH(ctx, k_groups_from_labels, nblk(w->nls, 256), 256, 0, w->rows_pos, w->nls, w->lab_pos, w->capP, w->capQ,
               w->grp, w->hist);
    RID_TRY(sort_rows_by_group(dm, w, k, 0));
    RID_TRY((launch_scan<1>(dm, w, w->s_row, w->s_capP, w->s_capQ, w->s_grp, w->nls, nullptr, 0, RID_PROF_SCAN_GROUP)));
    RID_TRY(rid_allreduce_u64(ctx, w->acc, 4 + (size_t)(1 + k) * w->ld));
    return RID_OK;
}

// W.k from w->lab_pos
static int within_disp(rid_dmat *dm, PamWs *w, int k, double *W, std::vector<int> *counts)
{
    rid_ctx *ctx = dm->ctx;
    RID_TRY(grouped_colsums(dm, w, k));
    RID_CUDA(cudaMemsetAsync(w->grp_sum, 0, (size_t)k * sizeof(u64), ctx->stream));
    RID_CUDA(cudaMemsetAsync(w->grp_cnt, 0, (size_t)k * sizeof(int), ctx->stream));
    LAUNCH(ctx, k_wdisp, nblk(w->m, 256), 256, 0, w->subset, w->lab_pos, w->m, w->acc + 4 + w->ld, w->ld, w->grp_sum,
           w->grp_cnt);
    RID_CUDA(cudaGetLastError());
    std::vector<u64> hs(k);
    std::vector<int> hc(k);
    RID_CUDA(cudaMemcpyAsync(hs.data(), w->grp_sum, (size_t)k * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    RID_CUDA(cudaMemcpyAsync(hc.data(), w->grp_cnt, (size_t)k * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RID_CUDA(cudaStreamSynchronize(ctx->stream));
    long double tot = 0.0L;
    for (int c = 0; c < k; ++c)
        if (hc[c] > 0) tot += (long double)((double)((long double)hs[c] * (long double)dm->step / (long double)hc[c]));
    *W = 0.5 * (double)tot;
    if (counts) *counts = hc;
    return RID_OK;
}

// cstat numbering: clusters numbered by first appearance scanning positions 0..m-1
static void number_clusters(const std::vector<int> &slot_by_pos, int k, const std::vector<int> &medpos_by_slot,
                            int *labels, int *medoids)
{
    std::vector<int> num(k, 0);
    int next = 0;
    for (size_t p = 0; p < slot_by_pos.size(); ++p) {
        int s = slot_by_pos[p];
        if (!num[s]) {
            num[s] = ++next;
            if (medoids) medoids[next - 1] = medpos_by_slot[s];
        }
        if (labels) labels[p] = num[s];
    }
}

static int pam_run(rid_dmat *dm, const int *subset, int m, int k, int *medoids, int *labels, double *objective,
                   double *avg_sil, int *n_swaps, HostSubset *prepared = nullptr)
{
    if (!dm) { rid_set_error("null distance matrix"); return RID_ERR_ARG; }
    if (subset == nullptr) m = dm->N;
    if (k < 1 || k >= m) {
        rid_set_error("Number of clusters 'k' must be in {1,2, .., n-1}");
        return RID_ERR_ARG;
    }
    RID_TRY(pam_check_na(dm, subset, m));
    PamWs *w;
    RID_TRY(ws_get(dm, k, &w));
    RID_TRY(pam_set_subset(dm, w, subset, m, true, prepared));
    RID_TRY(pam_begin(dm, w));
    RID_TRY(pam_build(dm, w, 0, k));
    PamState hs;
    RID_TRY(pam_swap(dm, w, k, &hs));
    std::vector<int> slot;
    RID_TRY(pam_collect_labels(dm, w, slot));
    std::vector<int> h_medpos(k);
    RID_CUDA(cudaMemcpyAsync(h_medpos.data(), w->medpos, (size_t)k * sizeof(int), cudaMemcpyDeviceToHost, dm->ctx->stream));
    RID_CUDA(cudaStreamSynchronize(dm->ctx->stream));
    number_clusters(slot, k, h_medpos, labels, medoids);
    if (objective) {
        objective[0] = (double)hs.td_build * dm->step / (double)m;
        objective[1] = (double)hs.td * dm->step / (double)m;
    }
    if (n_swaps) *n_swaps = hs.nswaps;
    if (avg_sil) {
        if (k == 1) {
            *avg_sil = 0.0;
        } else {
            rid_ctx *ctx = dm->ctx;
            RID_TRY(grouped_colsums(dm, w, k));
            RID_CUDA(cudaMemsetAsync(w->grp_sum, 0, (size_t)k * sizeof(u64), ctx->stream));
            RID_CUDA(cudaMemsetAsync(w->grp_cnt, 0, (size_t)k * sizeof(int), ctx->stream));
            LAUNCH(ctx, k_wdisp, nblk(w->m, 256), 256, 0, w->subset, w->lab_pos, w->m, w->acc + 4 + w->ld, w->ld,
                   w->grp_sum, w->grp_cnt);
            LAUNCH(ctx, k_silhouette, nblk(w->m, 256), 256, 0, w->subset, w->lab_pos, w->m, k, w->acc + 4 + w->ld, w->ld,
                   w->grp_cnt, w->sil);
            LAUNCH(ctx, k_sum_f64, 1, 1024, 0, w->sil, w->m, w->dscratch);
            RID_CUDA(cudaGetLastError());
            double s = 0.0;
            RID_CUDA(cudaMemcpyAsync(&s, w->dscratch, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            RID_CUDA(cudaStreamSynchronize(ctx->stream));
            *avg_sil = s / (double)m;
        }
    }
    return RID_OK;
}

extern "C" int rid_pam(rid_dmat *dm, const int *subset, int m, int k, int *medoids, int *labels, double *objective,
                       double *avg_sil, int *n_swaps)
{
    if (!dm) { rid_set_error("null distance matrix"); return RID_ERR_ARG; }
    RID_TRY(rid_time_begin(dm->ctx));
    int rc = pam_run(dm, subset, m, k, medoids, labels, objective, avg_sil, n_swaps);
    if (rc == RID_OK) rc = rid_time_end(dm->ctx);
    return rc;
}

extern "C" int rid_within_disp(rid_dmat *dm, const int *subset, int m, const int *labels, int k, double *W)
{
    if (!dm || !labels || !W || k < 1) { rid_set_error("rid_within_disp: bad arguments"); return RID_ERR_ARG; }
    if (subset == nullptr) m = dm->N;
    RID_TRY(pam_check_na(dm, subset, m));
    PamWs *w;
    RID_TRY(ws_get(dm, k, &w));
    RID_TRY(pam_set_subset(dm, w, subset, m, false));
    std::vector<int> g(m);
    for (int p = 0; p < m; ++p) {
        if (labels[p] < 1 || labels[p] > k) { rid_set_error("label %d out of 1..%d", labels[p], k); return RID_ERR_ARG; }
        g[p] = labels[p] - 1;
    }
    // (on the library's stream, which is not ordered against the legacy default stream; pageable source: staged on return)
    RID_CUDA(cudaMemcpyAsync(w->lab_pos, g.data(), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, dm->ctx->stream));
    RID_TRY(rid_time_begin(dm->ctx));
    RID_TRY(within_disp(dm, w, k, W, nullptr));
    return rid_time_end(dm->ctx);
}

// ---- multi-GPU scheduling of independent PAM runs ------------------------------------------------------------------
// Bootstrap replicates (and the k values of the sweep) are independent problems.  When the whole matrix fits on every
// GPU next to its shard, the shards are all-gathered once over NVLink and each rank runs its share of the problems on
// the replica at single-GPU efficiency (no per-step all-reduce, no latency-bound 1/world-sized passes); otherwise all
// ranks cooperate on every problem through the row shards.  Results are identical either way (exact integer sums).
struct FullReplica {
    rid_dmat *full = nullptr; // view of dm->full as a matrix one GPU owns completely (lives as long as dm does)
    int saved_world = 1, saved_rank = 0;
    bool active = false;
};

// The replica buffer was allocated with the matrix (dist.cu dmat_alloc) and the copy engines have been pulling the other
// ranks' row blocks since the distance kernels ended; here the library's stream starts to depend on those copies.
static int replica_begin(rid_dmat *dm, FullReplica *R)
{
    rid_ctx *ctx = dm->ctx;
    R->active = false;
    if (dm->world == 1 || ctx->world == 1 || !dm->full || !dm->peers_ok || getenv("RID_NO_REPLICA")) return RID_OK;
    RID_TRY(rid_replica_start(dm)); // (no-op when the pulls are already under way)
    if (dm->replica_state == 0) return RID_OK;
    RID_CUDA(cudaStreamWaitEvent(ctx->stream, dm->replica_ev, 0));
    if (!dm->replica_view) {
        rid_dmat *v = new rid_dmat();
        v->ctx = ctx;
        v->N = dm->N;
        v->row0 = 0;
        v->row1 = dm->N;
        v->ld = dm->ld;
        v->D = dm->full;
        v->alloc_rows = (size_t)dm->world * dm->alloc_rows;
        v->bits = dm->bits;
        v->step = dm->step;
        v->parent = dm;
        dm->replica_view = v;
    }
    dm->replica_view->has_na = dm->has_na;
    dm->replica_view->na_cell = dm->na_cell;
    R->full = dm->replica_view;
    R->saved_world = ctx->world;
    R->saved_rank = ctx->rank;
    R->active = true;
    return RID_OK;
}
// local phases run with the context switched to "one GPU"; collectives need the real world back
static void replica_local(FullReplica *R, bool on)
{
    if (!R->active) return;
    rid_ctx *ctx = R->full->ctx;
    ctx->world = on ? 1 : R->saved_world;
    ctx->rank = on ? 0 : R->saved_rank;
}
static void replica_end(FullReplica *R)
{
    if (!R->active) return;
    replica_local(R, false);
    R->active = false;
}

// A view whose parent has only pulled the columns >= replica_lo[s] of every peer block s: true if the symmetric
// compaction of this (sorted) subset reads nothing else.  A compact super-tile (bi, bj), bi <= bj, reads source rows
// ccells[bi*512 ..] x source columns ccells[bj*512 ..]; the smallest column a row of block s ever meets is therefore the
// first cell of the super-tile row that holds the block's first subset cell.
static bool replica_upper_suffices(const rid_dmat *parent, const std::vector<int> &ccells)
{
    const int world = parent->world;
    const size_t per = parent->alloc_rows;
    for (int s = 0; s < world && s < 16; ++s) {
        const int lo = parent->replica_lo[s];
        if (lo <= 0) continue;
        const int r0 = (int)std::min<size_t>((size_t)s * per, (size_t)parent->N);
        const size_t p = (size_t)(std::lower_bound(ccells.begin(), ccells.end(), r0) - ccells.begin());
        if (p >= ccells.size()) continue; // no subset row in this block
        const size_t I0 = p / 512 * 512;  // (CS_S)
        if (ccells[I0] < lo) return false;
    }
    return true;
}

// State of the BUILD medoids of the last k a rank handled in the sweep: accumulators (hdr | SP | AQ[k]) and assignment.
struct SweepChain {
    int chain_k = 0;
    u64 *acc = nullptr;
    u32 *capP = nullptr, *capQ = nullptr;
    int *grp = nullptr;
};

// SWAP(k) of the sweep starts from the first k BUILD picks.  Their accumulators and assignment are kept from the previous
// k of this rank and moved to the new medoid set by ONE delta pass over the rows whose (nearest, second nearest) pair
// changes when the extra medoids are added, instead of a full scan per k.  Medoids must already be installed.
static int sweep_chain_step(rid_dmat *dm, PamWs *w, int k, SweepChain &B, PamState *hs, int *next_pick = nullptr)
{
    rid_ctx *ctx = dm->ctx;
    const size_t nl = (size_t)std::max(w->nls, 1);
    const size_t nacc = 4 + (size_t)(1 + k) * w->ld;
    int n_delta = 0;
    if (B.chain_k == 0) {
        RID_TRY(pam_swap_first_scan(dm, w, k));
    } else {
        const size_t nold = 4 + (size_t)(1 + B.chain_k) * w->ld;
        RID_CUDA(cudaMemcpyAsync(w->capP, B.capP, nl * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
        RID_CUDA(cudaMemcpyAsync(w->capQ, B.capQ, nl * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
        RID_CUDA(cudaMemcpyAsync(w->grp, B.grp, nl * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
        RID_CUDA(cudaMemcpyAsync(w->acc, B.acc, nold * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
        // the accumulators of the groups added since chain_k start from zero
        RID_CUDA(cudaMemsetAsync(w->acc + nold, 0, (nacc - nold) * sizeof(u64), ctx->stream));
        RID_TRY(pam_swap_delta_iteration(dm, w, k, &n_delta, false));
        LAUNCH(ctx, k_capture_td_build, 1, 1, 0, w->state, w->acc);
    }
    RID_CUDA(cudaMemcpyAsync(B.capP, w->capP, nl * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
    RID_CUDA(cudaMemcpyAsync(B.capQ, w->capQ, nl * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
    RID_CUDA(cudaMemcpyAsync(B.grp, w->grp, nl * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    RID_CUDA(cudaMemcpyAsync(B.acc, w->acc, nacc * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
    B.chain_k = k;
    if (next_pick) {
        // SP of the first k BUILD picks is exactly what BUILD's next step selects from (gain(h) = TD - SP[h]): the chain
        // makes BUILD's own incremental passes redundant.  next_pick[k] = column of pick k + 1 (medoid tables untouched).
        auto sel_build = k_select<true>;
        LAUNCH(ctx, sel_build, nblk(w->m, 256), 256, 0, w->subset, w->m, w->is_med, w->acc, w->ld, k, w->med, w->medpos,
               w->is_med, w->blockbest, w->state, 0, (u64 *)nullptr, (int *)nullptr, (u64 *)nullptr, next_pick);
    }
    auto sel_swap = k_select<false>;
    LAUNCH(ctx, sel_swap, nblk(w->m, 256), 256, 0, w->subset, w->m, w->is_med, w->acc, w->ld, k, w->med, w->medpos,
           w->is_med, w->blockbest, w->state, 0, (u64 *)nullptr, (int *)nullptr, (u64 *)nullptr, (int *)nullptr);
    RID_CUDA(cudaGetLastError());
    return pam_swap(dm, w, k, hs, true);
}

// after a phase in which every rank polled for cancellation on its own: one answer for all
static int agree_interrupt(rid_ctx *ctx, int rc)
{
    if (ctx->world == 1) return rc;
    int none = 0;
    int r2 = rid_agree_all(ctx, rc != RID_ERR_INTERRUPTED, &none);
    if (r2 != RID_OK) return r2;
    if (!none && rc == RID_OK) {
        rid_set_error("interrupted: the caller asked for cancellation");
        return RID_ERR_INTERRUPTED;
    }
    return rc;
}

// clusGapExt's k loop.  BUILD is greedy and does not depend on k, so BUILD(k) is the first k picks of
// BUILD(Kmax): run it once and start SWAP(k) from the k-prefix (identical results, Kmax instead of
// Kmax(Kmax+1)/2 BUILD passes).
static int gap_sweep_on(rid_dmat *dm, const int *subset, int m, int Kmax, double *logW, int nshare, int myshare)
{
    rid_ctx *ctx = dm->ctx;
    PamWs *w;
    RID_TRY(ws_get(dm, Kmax, &w));
    RID_TRY(pam_set_subset(dm, w, subset, m, true));
    RID_TRY(pam_begin(dm, w));
    // state of the BUILD medoids of the last k this rank handled (see sweep_chain_step)
    const bool chain = !getenv("RID_SWAP_FULL") && !getenv("RID_SWEEP_FULL");
    SweepChain B;
    char *b_buf = nullptr;
    const size_t b_nl = (size_t)std::max(w->nls, 1);
    const size_t b_nacc = 4 + (size_t)(1 + Kmax) * w->ld;
    const size_t b_bytes = std::max<size_t>(((b_nacc * sizeof(u64) + 255) & ~(size_t)255) + 3 * ((b_nl * 4 + 255) & ~(size_t)255),
                                            (size_t)256 << 10);
    {
        int mine = chain && rid_pool_alloc(ctx, (void **)&b_buf, b_bytes) == cudaSuccess, all = 0;
        if (!mine) { cudaGetLastError(); b_buf = nullptr; }
        RID_TRY(rid_agree_all(ctx, mine, &all)); // (chained and plain SWAP(k) use different collectives)
        if (!all && b_buf) { rid_pool_free(ctx, b_buf, b_bytes); b_buf = nullptr; }
        if (b_buf) {
            char *p = b_buf;
            carve(p, B.acc, b_nacc); carve(p, B.capP, b_nl); carve(p, B.capQ, b_nl); carve(p, B.grp, b_nl);
        }
    }
    // With the chain, BUILD only makes its first two picks here: pick k + 1 is selected from the chain's sums of the
    // first k picks (sweep_chain_step), which are the sums BUILD's own incremental pass would produce.
    const bool interleave = b_buf && nshare == 1 && Kmax >= 3 && Kmax * Kmax <= K2MAX && !getenv("RID_SWEEP_BUILD_FIRST");
    RID_TRY(pam_build(dm, w, 0, interleave ? 2 : Kmax));
    int *d_order = nullptr;
    RID_CUDA(rid_tmp_alloc(ctx, &d_order, (size_t)Kmax * sizeof(int)));
    RID_CUDA(cudaMemcpyAsync(d_order, w->med, (size_t)(interleave ? 2 : Kmax) * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    int rc = RID_OK;
    // Labels of every k this rank owns are kept by column (bytes); W.k is then evaluated for WB values of k per pass
    // over the matrix instead of one pass per k.
    const bool batched = Kmax <= 255 && !getenv("RID_WK_SINGLE");
    const size_t lab_pitch = round_up(w->ld, 16);
    unsigned char *d_lab = nullptr;
    u64 *d_S = nullptr;
    std::vector<int> ks;
    if (batched) {
        if (rid_tmp_alloc(ctx, &d_lab, (size_t)Kmax * lab_pitch) != cudaSuccess ||
            rid_tmp_alloc(ctx, &d_S, (size_t)WB * w->ld * sizeof(u64)) != cudaSuccess) {
            cudaGetLastError();
            rid_tmp_free(ctx, d_lab, (size_t)Kmax * lab_pitch); rid_tmp_free(ctx, d_S, (size_t)WB * w->ld * sizeof(u64));
            rid_tmp_free(ctx, d_order, (size_t)Kmax * sizeof(int));
            rid_set_error("rid_gap_sweep: out of device memory");
            return RID_ERR_NOMEM;
        }
        cudaMemsetAsync(d_lab, 0xFF, (size_t)Kmax * lab_pitch, ctx->stream); // 0xFF = column outside the subset
    }
    // W.k of all partitions from one pass (atoms, see k_wk_atoms): the common refinement of the partitions is built on the
    // device while the sweep runs; at most a_cap classes (4 GB of grouped column sums), else the batched passes below
    const int a_cap = getenv("RID_WK_ATOM_CAP") ? std::max(1, atoi(getenv("RID_WK_ATOM_CAP")))
                                                : (int)std::min<size_t>(8192, ((size_t)4 << 30) / (w->ld * sizeof(u64)));
    bool atoms = batched && nshare == 1 && Kmax >= 2 && (a_cap >= 2 * Kmax || getenv("RID_WK_ATOM_CAP")) && !getenv("RID_WK_BATCH");
    int *d_atom = nullptr, *d_table = nullptr;
    AtomState *d_as = nullptr;
    const size_t atom_bytes = (size_t)m * sizeof(int), table_bytes = (size_t)a_cap * Kmax * sizeof(int);
    if (atoms) {
        int mine = 1, all = 0;
        if (rid_tmp_alloc(ctx, &d_atom, atom_bytes) != cudaSuccess || rid_tmp_alloc(ctx, &d_table, table_bytes) != cudaSuccess ||
            rid_tmp_alloc(ctx, &d_as, sizeof(AtomState)) != cudaSuccess) {
            cudaGetLastError();
            mine = 0;
        }
        RID_TRY(rid_agree_all(ctx, mine, &all)); // (the two W.k routes use different collectives: every rank takes the same one)
        atoms = all != 0;
        if (atoms) LAUNCH(ctx, k_atom_init, nblk(m, 256), 256, 0, d_atom, m, d_as);
    }
    for (int k = 1; k <= Kmax && rc == RID_OK; ++k) {
        logW[k - 1] = 0.0;
        if ((k - 1) % nshare != myshare) continue; // another rank's k
        if ((rc = rid_check_abort(ctx, nshare == 1)) != RID_OK) break;
        double W = 0.0;
        if (k == 1) {
            rc = cudaMemsetAsync(w->lab_pos, 0, (size_t)m * sizeof(int), ctx->stream) == cudaSuccess ? RID_OK : RID_ERR_CUDA;
        } else {
            cudaMemsetAsync(w->is_med, 0, (size_t)dm->N, ctx->stream);
            LAUNCH(ctx, k_set_medoids, 1, std::max(32, (k + 31) / 32 * 32), 0, d_order, k, w->med, w->medpos, w->pos_of,
                   w->is_med, w->state);
            PamState hs;
            const bool chain_ok = chain && b_buf && k * k <= K2MAX;
            if (!chain_ok) {
                B.chain_k = 0;
                rc = pam_swap(dm, w, k, &hs);
            } else {
                rc = sweep_chain_step(dm, w, k, B, &hs, (interleave && k < Kmax) ? d_order : (int *)nullptr);
            }
            std::vector<int> slot, h_medpos(k);
            const bool memo = rc == RID_OK && subset == nullptr && nshare == 1 && !dm->parent && !getenv("RID_NO_MEMO");
            if (memo && cudaMemcpyAsync(h_medpos.data(), w->medpos, (size_t)k * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
                rc = RID_ERR_CUDA;
            if (rc == RID_OK) rc = pam_collect_labels(dm, w, slot); // (synchronises the stream)
            if (memo && rc == RID_OK) {
                // PAM(k) on all cells is what clusterboot starts with (c1): keep it with the matrix
                rid_dmat::PamMemo mm;
                mm.k = k; mm.nswaps = hs.nswaps;
                mm.labels.resize(m); mm.medoids.resize(k);
                number_clusters(slot, k, h_medpos, mm.labels.data(), mm.medoids.data());
                bool found = false;
                for (auto &o : dm->pam_memo)
                    if (o.k == k) { o = std::move(mm); found = true; break; }
                if (!found) dm->pam_memo.push_back(std::move(mm));
            }
        }
        if (rc != RID_OK) break;
        if (batched) {
            LAUNCH(ctx, k_labels_by_column, nblk(m, 256), 256, 0, w->subset, w->lab_pos, m, d_lab + (size_t)ks.size() * lab_pitch);
            ks.push_back(k);
            if (atoms && k >= 2) { // refine the atoms by this partition
                cudaMemsetAsync(d_table, 0, (size_t)a_cap * k * sizeof(int), ctx->stream);
                LAUNCH(ctx, k_atom_mark, nblk(m, 256), 256, 0, d_atom, w->lab_pos, m, k, d_table, d_as);
                LAUNCH(ctx, k_atom_rank, 1, 1024, 0, d_table, k, d_as, a_cap);
                LAUNCH(ctx, k_atom_relabel, nblk(m, 256), 256, 0, d_atom, w->lab_pos, m, k, d_table, d_as);
            }
        } else {
            rc = within_disp(dm, w, k, &W, nullptr);
            logW[k - 1] = log(W);
        }
    }
    bool wk_done = false;
    if (atoms && rc == RID_OK && !ks.empty()) {
        AtomState as;
        if (cudaMemcpyAsync(&as, d_as, sizeof(as), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = RID_ERR_CUDA;
        const int A = as.A, nq = (int)ks.size();
        u64 *d_AQ = nullptr, *d_T = nullptr;
        int *d_cnt = nullptr, *d_hist = nullptr;
        unsigned char *d_al = nullptr;
        const size_t aq_bytes = (size_t)A * w->ld * sizeof(u64), t_bytes = (size_t)nq * Kmax * sizeof(u64),
                     c_bytes = (size_t)nq * Kmax * sizeof(int), h_bytes = (size_t)A * sizeof(int), al_bytes = (size_t)nq * A;
        int mine = rc == RID_OK && !as.overflow && A >= 1 &&
                   rid_tmp_alloc(ctx, &d_AQ, aq_bytes) == cudaSuccess && rid_tmp_alloc(ctx, &d_T, t_bytes) == cudaSuccess &&
                   rid_tmp_alloc(ctx, &d_cnt, c_bytes) == cudaSuccess && rid_tmp_alloc(ctx, &d_hist, h_bytes) == cudaSuccess &&
                   rid_tmp_alloc(ctx, &d_al, al_bytes) == cudaSuccess;
        int all = 0;
        if (rc == RID_OK) rc = rid_agree_all(ctx, mine, &all);
        if (rc == RID_OK && all) {
            cudaMemsetAsync(d_AQ, 0, aq_bytes, ctx->stream);
            cudaMemsetAsync(d_T, 0, t_bytes, ctx->stream);
            cudaMemsetAsync(d_cnt, 0, c_bytes, ctx->stream);
            cudaMemsetAsync(d_hist, 0, h_bytes, ctx->stream);
            if (w->nls) {
                LAUNCH(ctx, k_atom_hist, nblk(w->nls, 256), 256, 0, d_atom, w->rows_pos, w->nls, d_hist);
                LAUNCH(ctx, k_atom_prefix, 1, 1024, 0, d_hist, A);
                LAUNCH(ctx, k_atom_scatter, nblk(w->nls, 256), 256, 0, d_atom, w->rows, w->rows_pos, w->nls, d_hist, w->s_row,
                       w->s_grp, w->s_capP, w->s_capQ);
                int strips = (int)((w->ld + SCAN_THREADS * 4 - 1) / (SCAN_THREADS * 4));
                int want_y = std::max(1, (ctx->sm_count * 16 + strips - 1) / strips);
                int rpc = std::min(std::max((((w->nls + want_y - 1) / want_y) + 7) & ~7, 8), SCAN_MAXROWS);
                dim3 grid(strips, std::min((w->nls + rpc - 1) / rpc, want_y * 2));
                cudaEvent_t pe = rid_prof_begin(ctx, RID_PROF_SCAN_GROUP);
                auto s8 = k_scan<2, 8>;
                auto s2 = k_scan<2, 2>;
                if (dm->bits <= 28)
                    LAUNCH(ctx, s8, grid, SCAN_THREADS, 0, w->mat, w->ld, w->s_row, w->s_capP, w->s_capQ, w->s_grp, w->nls,
                           (const int *)nullptr, rpc, (u64 *)nullptr, d_AQ, w->state, 0);
                else
                    LAUNCH(ctx, s2, grid, SCAN_THREADS, 0, w->mat, w->ld, w->s_row, w->s_capP, w->s_capQ, w->s_grp, w->nls,
                           (const int *)nullptr, rpc, (u64 *)nullptr, d_AQ, w->state, 0);
                rid_prof_end(ctx, pe, RID_PROF_SCAN_GROUP, (double)w->nls * (double)w->m * 4.0, (double)w->nls * (double)w->ld * 4.0);
            }
            LAUNCH(ctx, k_atom_labels, nblk(m, 256), 256, 0, d_atom, w->subset, m, d_lab, lab_pitch, nq, A, d_al);
            dim3 g2(nblk(m, 256), (A + WKA_ACH - 1) / WKA_ACH, (nq + WKA_PB - 1) / WKA_PB);
            LAUNCH(ctx, k_wk_atoms, g2, 256, 0, w->subset, m, d_lab, lab_pitch, nq, d_AQ, w->ld, A, d_al, Kmax, d_T, d_cnt);
            if (cudaGetLastError() != cudaSuccess) rc = RID_ERR_CUDA;
            // every rank summed its own rows: the partial within-cluster sums add up (the counts are complete on every rank)
            if (rc == RID_OK) rc = rid_allreduce_u64(ctx, d_T, (size_t)nq * Kmax);
            std::vector<u64> hT((size_t)nq * Kmax);
            std::vector<int> hC((size_t)nq * Kmax);
            if (rc == RID_OK && (cudaMemcpyAsync(hT.data(), d_T, t_bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                                 cudaMemcpyAsync(hC.data(), d_cnt, c_bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                                 cudaStreamSynchronize(ctx->stream) != cudaSuccess)) rc = RID_ERR_CUDA;
            for (int q = 0; q < nq && rc == RID_OK; ++q) {
                const int k = ks[q];
                long double tot = 0.0L;
                for (int c = 0; c < k; ++c)
                    if (hC[(size_t)q * Kmax + c] > 0)
                        tot += (long double)((double)((long double)hT[(size_t)q * Kmax + c] * (long double)dm->step /
                                                      (long double)hC[(size_t)q * Kmax + c]));
                logW[k - 1] = log(0.5 * (double)tot);
            }
            wk_done = rc == RID_OK;
            ctx->stat[10] = (double)A;
        } else {
            cudaGetLastError();
            ctx->stat[10] = as.overflow ? -1.0 : 0.0; // more classes than a_cap (or no memory): the batched passes do it
        }
        rid_tmp_free(ctx, d_AQ, aq_bytes); rid_tmp_free(ctx, d_T, t_bytes); rid_tmp_free(ctx, d_cnt, c_bytes);
        rid_tmp_free(ctx, d_hist, h_bytes); rid_tmp_free(ctx, d_al, al_bytes);
    }
    rid_tmp_free(ctx, d_atom, atom_bytes); rid_tmp_free(ctx, d_table, table_bytes); rid_tmp_free(ctx, d_as, sizeof(AtomState));
    for (size_t q0 = 0; batched && !wk_done && rc == RID_OK && q0 < ks.size(); q0 += WB) {
        const int nq = (int)std::min<size_t>(WB, ks.size() - q0);
        cudaMemsetAsync(d_S, 0, (size_t)WB * w->ld * sizeof(u64), ctx->stream);
        if (w->nls) {
            int strips = (int)((w->ld + SCAN_THREADS * 4 - 1) / (SCAN_THREADS * 4));
            int want_y = std::max(1, (ctx->sm_count * 16 + strips - 1) / strips);
            int rpc = std::min(std::max((((w->nls + want_y - 1) / want_y) + 7) & ~7, 8), SCAN_MAXROWS);
            dim3 grid(strips, std::min((w->nls + rpc - 1) / rpc, want_y * 2));
            cudaEvent_t pe = rid_prof_begin(ctx, RID_PROF_SCAN_GROUP);
            auto k8 = k_wk_batch<8>;
            auto k2 = k_wk_batch<2>;
            if (dm->bits <= 28)
                LAUNCH(ctx, k8, grid, SCAN_THREADS, 0, w->mat, w->ld, w->rows, w->rows_pos, w->subset, w->nls, rpc, d_lab + q0 * lab_pitch, lab_pitch, nq, d_S);
            else
                LAUNCH(ctx, k2, grid, SCAN_THREADS, 0, w->mat, w->ld, w->rows, w->rows_pos, w->subset, w->nls, rpc, d_lab + q0 * lab_pitch, lab_pitch, nq, d_S);
            rid_prof_end(ctx, pe, RID_PROF_SCAN_GROUP, (double)w->nls * (double)w->m * 4.0, (double)w->nls * (double)w->ld * 4.0);
        }
        if (cudaGetLastError() != cudaSuccess) { rc = RID_ERR_CUDA; break; }
        rc = rid_allreduce_u64(ctx, d_S, (size_t)WB * w->ld);
        for (int q = 0; q < nq && rc == RID_OK; ++q) {
            const int k = ks[q0 + q];
            cudaMemsetAsync(w->grp_sum, 0, (size_t)k * sizeof(u64), ctx->stream);
            cudaMemsetAsync(w->grp_cnt, 0, (size_t)k * sizeof(int), ctx->stream);
            LAUNCH(ctx, k_wdisp_cols, nblk(m, 256), 256, 0, w->subset, m, d_lab + (q0 + q) * lab_pitch, d_S + (size_t)q * w->ld,
                   w->grp_sum, w->grp_cnt);
            std::vector<u64> hs(k);
            std::vector<int> hc(k);
            if (cudaMemcpyAsync(hs.data(), w->grp_sum, (size_t)k * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                cudaMemcpyAsync(hc.data(), w->grp_cnt, (size_t)k * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = RID_ERR_CUDA; break; }
            long double tot = 0.0L;
            for (int c = 0; c < k; ++c)
                if (hc[c] > 0) tot += (long double)((double)((long double)hs[c] * (long double)dm->step / (long double)hc[c]));
            logW[k - 1] = log(0.5 * (double)tot);
        }
    }
    // (pooled buffers: every later user is ordered behind this work on ctx->stream)
    rid_tmp_free(ctx, d_lab, (size_t)Kmax * lab_pitch);
    rid_tmp_free(ctx, d_S, (size_t)WB * w->ld * sizeof(u64));
    rid_tmp_free(ctx, d_order, (size_t)Kmax * sizeof(int));
    if (b_buf) { cudaStreamSynchronize(ctx->stream); rid_pool_free(ctx, b_buf, b_bytes); }
    return rc;
}

extern "C" int rid_gap_sweep(rid_dmat *dm, const int *subset, int m, int Kmax, double *logW)
{
    if (!dm || !logW || Kmax < 1) { rid_set_error("rid_gap_sweep: bad arguments"); return RID_ERR_ARG; }
    if (subset == nullptr) m = dm->N;
    if (Kmax >= m) { rid_set_error("Number of clusters 'k' must be in {1,2, .., n-1}"); return RID_ERR_ARG; }
    RID_TRY(pam_check_na(dm, subset, m));
    rid_ctx *ctx = dm->ctx;
    RID_TRY(rid_time_begin(ctx));
    // Several ranks cooperate on the row shards: every pass reads 1/world of the bytes per rank and needs no replica
    // (measured equal to splitting the k values over full replicas, which repeats the shared BUILD on every rank and
    // needs the complete matrix everywhere first).  RID_SWEEP_REPLICA=1 selects the replica schedule: every rank repeats
    // the shared BUILD and takes every world-th k.
    FullReplica R;
    if (getenv("RID_SWEEP_REPLICA")) {
        RID_TRY(replica_begin(dm, &R));
        if (R.active) RID_TRY(replica_need_whole(R.full));
    }
    // RID_REPLICA_EARLY=1: the copy engines start pulling the peers' row blocks now, under the (cooperative) sweep
    if (!R.active && getenv("RID_REPLICA_EARLY") && atoi(getenv("RID_REPLICA_EARLY")) && ctx->world > 1 && dm->full && dm->peers_ok &&
        !getenv("RID_NO_REPLICA"))
        RID_TRY(rid_replica_start(dm));
    replica_local(&R, true);
    int rc = gap_sweep_on(R.active ? R.full : dm, subset, m, Kmax, logW, R.active ? R.saved_world : 1,
                          R.active ? R.saved_rank : 0);
    replica_local(&R, false);
    if (R.active) rc = agree_interrupt(ctx, rc);
    // the sweep is usually followed by clusterboot: its replica starts to fill now, behind the host work in between
    if (rc == RID_OK && ctx->world > 1 && dm->full && dm->peers_ok && !getenv("RID_NO_REPLICA")) rc = rid_replica_start(dm);
    if (R.active && rc == RID_OK) {
        double *d_lw = nullptr;
        if (rid_tmp_alloc(ctx, &d_lw, (size_t)Kmax * sizeof(double)) != cudaSuccess) rc = RID_ERR_NOMEM;
        if (rc == RID_OK && cudaMemcpyAsync(d_lw, logW, (size_t)Kmax * sizeof(double), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = RID_ERR_CUDA;
        if (rc == RID_OK) rc = rid_allreduce_f64(ctx, d_lw, (size_t)Kmax);
        if (rc == RID_OK && (cudaMemcpyAsync(logW, d_lw, (size_t)Kmax * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                             cudaStreamSynchronize(ctx->stream) != cudaSuccess)) rc = RID_ERR_CUDA;
        rid_tmp_free(ctx, d_lw, (size_t)Kmax * sizeof(double));
    }
    replica_end(&R);
    if (rc != RID_OK) return rc;
    return rid_time_end(ctx);
}

// BUILD's first pick of every bootstrap replicate, before any replicate starts: the step-0 column sums of CSB replicates
// per pass over this rank's rows (k_colsum_batch), summed over the ranks, then the arg min per replicate
// (k_first_medoid).  Runs on the sharded matrix with the real world size (before any switch to a replica).
// The row masks of all batches are built from the samples on the device (k_first_masks).
struct FirstMasks {
    int csb = 2 * CSM;          // replicates per pass: 32 (or 16: RID_COLSUM_MB=1) on the tensor cores, CSB with the IMAD kernel (RID_COLSUM_IMAD=1)
    int nbatch = 0;
    std::vector<int> rep;       // [nbatch][csb] replicate of every slot (padding repeats the batch's last one)
    std::vector<long long> off; // [B + 1]
};
static void boot_first_tables(FirstMasks *M, const int *sample_len, int B)
{
    const int csb = M->csb = getenv("RID_COLSUM_IMAD") ? CSB : ((getenv("RID_COLSUM_MB") && atoi(getenv("RID_COLSUM_MB")) == 1) ? CSM : 2 * CSM);
    M->nbatch = (B + csb - 1) / csb;
    M->off.assign(B + 1, 0);
    for (int b = 0; b < B; ++b) M->off[b + 1] = M->off[b] + sample_len[b];
    M->rep.assign((size_t)M->nbatch * csb, 0);
    for (int t = 0; t < M->nbatch; ++t) {
        const int b0 = t * csb, nb = std::min(csb, B - b0);
        for (int r = 0; r < csb; ++r) M->rep[(size_t)t * csb + r] = b0 + std::min(r, nb - 1);
    }
}

// d_first[b] = (cell, position) for every replicate b; d_cat / d_off are the samples on the device.
static int boot_first_medoids(rid_dmat *dm, const FirstMasks &M, const int *samples, int B, int2 *d_first, int *d_cat,
                              long long *d_off, u64 *d_sp0 /* [csb][ld] */, int *d_rep /* [nbatch][csb] */,
                              unsigned short *d_mask /* [nbatch][nloc] */)
{
    rid_ctx *ctx = dm->ctx;
    const int nloc = dm->row1 - dm->row0, csb = M.csb;
    // (pageable sources: staged by the time the calls return)
    RID_CUDA(cudaMemcpyAsync(d_cat, samples, (size_t)M.off[B] * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    RID_CUDA(cudaMemcpyAsync(d_off, M.off.data(), (size_t)(B + 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    RID_CUDA(cudaMemcpyAsync(d_rep, M.rep.data(), M.rep.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if (nloc) {
        const size_t mask_bytes = csb > 16 ? (size_t)M.nbatch * nloc * sizeof(u32) : ((((size_t)M.nbatch * nloc * sizeof(unsigned short)) + 3) & ~(size_t)3);
        RID_CUDA(cudaMemsetAsync(d_mask, 0, mask_bytes, ctx->stream));
        dim3 gm(std::max(1, std::min(64, nblk(dm->N, 256))), B);
        LAUNCH(ctx, k_first_masks, gm, 256, 0, d_cat, d_off, dm->row0, nloc, csb, reinterpret_cast<unsigned int *>(d_mask));
    }
    int strips = (int)((dm->ld + SCAN_THREADS * 4 - 1) / (SCAN_THREADS * 4));
    int want_y = std::max(1, (ctx->sm_count * 16 + strips - 1) / strips);
    int rpc = std::min(std::max((((nloc + want_y - 1) / want_y) + 7) & ~7, 8), SCAN_MAXROWS);
    dim3 grid(strips, std::max(1, std::min((nloc + rpc - 1) / rpc, want_y * 2)));
    // tensor-core kernel: a CTA owns 128 (16 replicates per pass) or 64 (32 per pass) columns and walks row chunks of
    // <= CM_ROWS rows (multiples of 32)
    const int strips_m = (int)(dm->ld / (csb > 16 ? 64 : 128));
    const int want_ym = std::max(1, (ctx->sm_count * 16 + strips_m - 1) / strips_m);
    const int rpc_m = std::min(std::max((((nloc + want_ym - 1) / want_ym) + 31) & ~31, 32), CM_ROWS);
    dim3 grid_m(strips_m, std::max(1, std::min((nloc + rpc_m - 1) / rpc_m, want_ym * 2)));
    for (int t = 0; t < M.nbatch; ++t) {
        const int nb = std::min(csb, B - t * csb);
        RID_CUDA(cudaMemsetAsync(d_sp0, 0, (size_t)csb * dm->ld * sizeof(u64), ctx->stream));
        if (nloc) {
            cudaEvent_t pe = rid_prof_begin(ctx, RID_PROF_SCAN_BUILD);
            auto k8 = k_colsum_batch<8>;
            auto k2 = k_colsum_batch<2>;
            if (csb == 2 * CSM)
                LAUNCH(ctx, k_colsum_mma<2>, grid_m, 128, 0, dm->D, dm->ld, nloc, rpc_m,
                       (const void *)(reinterpret_cast<const u32 *>(d_mask) + (size_t)t * nloc), d_sp0);
            else if (csb == CSM)
                LAUNCH(ctx, k_colsum_mma<1>, grid_m, 128, 0, dm->D, dm->ld, nloc, rpc_m, (const void *)(d_mask + (size_t)t * nloc), d_sp0);
            else if (dm->bits <= 28) LAUNCH(ctx, k8, grid, SCAN_THREADS, 0, dm->D, dm->ld, nloc, rpc, d_mask + (size_t)t * nloc, d_sp0);
            else LAUNCH(ctx, k2, grid, SCAN_THREADS, 0, dm->D, dm->ld, nloc, rpc, d_mask + (size_t)t * nloc, d_sp0);
            rid_prof_end(ctx, pe, RID_PROF_SCAN_BUILD, (double)nloc * (double)dm->N * 4.0, (double)nloc * (double)dm->ld * 4.0);
        }
        RID_CUDA(cudaGetLastError());
        RID_TRY(rid_allreduce_u64(ctx, d_sp0, (size_t)csb * dm->ld));
        LAUNCH(ctx, k_first_medoid, nb, 256, 0, d_sp0, dm->ld, d_cat, d_off, d_rep + (size_t)t * csb, d_first, dm->N);
        RID_CUDA(cudaGetLastError());
    }
    return RID_OK;
}

extern "C" int rid_clusterboot(rid_dmat *dm, int k, const int *samples, const int *sample_len, int B, int *partition,
                               int *medoids, double *bootresult)
{
    if (!dm || !partition || (B > 0 && (!samples || !sample_len || !bootresult))) {
        rid_set_error("rid_clusterboot: bad arguments");
        return RID_ERR_ARG;
    }
    rid_ctx *ctx = dm->ctx;
    int N = dm->N;
    RID_TRY(rid_time_begin(ctx));
    // c1 = PAM on all cells: every rank scans its own row shard (1/world of the bytes), before any replica exists
    // the copy engines start (or go on) pulling the peers' row blocks while c1 and the first picks run on the shard
    if (ctx->world > 1 && dm->full && dm->peers_ok && !getenv("RID_NO_REPLICA")) RID_TRY(rid_replica_start(dm));
    // BUILD's first pick of every replicate comes from a few batched passes over this rank's rows (boot_first_medoids);
    // only the enqueue-only schedule on a complete view consumes it.
    const bool want_first = B > 0 && k >= 2 && k * k <= K2MAX && !getenv("RID_SWAP_FULL") && !getenv("RID_BOOT_SYNC") &&
                            !getenv("RID_NO_FIRST") &&
                            (ctx->world == 1 || (dm->full && dm->peers_ok && !getenv("RID_NO_REPLICA")));
    FirstMasks fmasks;
    if (want_first) boot_first_tables(&fmasks, sample_len, B);
    std::vector<int> med_pos(k);
    int c1_swaps = 0;
    int rc = RID_OK;
    const rid_dmat::PamMemo *memo = nullptr;
    if (!getenv("RID_NO_MEMO"))
        for (const auto &o : dm->pam_memo)
            if (o.k == k && (int)o.labels.size() == N) memo = &o;
    if (memo) {
        // the k sweep of the same clustexp call already ran pam(as.dist(D), k) on all cells (rid_gap_sweep): identical call,
        // identical result -- c1 is not repeated
        std::copy(memo->labels.begin(), memo->labels.end(), partition);
        std::copy(memo->medoids.begin(), memo->medoids.end(), med_pos.begin());
        c1_swaps = memo->nswaps;
        ctx->stat[9] += 1.0;
    } else {
        rc = pam_run(dm, nullptr, N, k, med_pos.data(), partition, nullptr, nullptr, &c1_swaps);
    }
    if (rc != RID_OK) return rc;
    for (int i = 0; i < 4; ++i)
        if (!ctx->phase_ev[i]) RID_CUDA(cudaEventCreate(&ctx->phase_ev[i]));
    RID_CUDA(cudaEventRecord(ctx->phase_ev[0], ctx->stream)); // c1 done
    char *first_buf = nullptr;
    size_t first_bytes = 0;
    int2 *d_first = nullptr;
    if (want_first) {
        const int nloc = dm->row1 - dm->row0;
        const long long tot = fmasks.off.empty() ? 0 : fmasks.off[B];
        auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t o_first = 0, o_off = o_first + al((size_t)B * sizeof(int2)), o_rep = o_off + al((size_t)(B + 1) * sizeof(long long)),
                     o_sp0 = o_rep + al((size_t)std::max(fmasks.nbatch, 1) * fmasks.csb * sizeof(int)),
                     o_mask = o_sp0 + al((size_t)fmasks.csb * dm->ld * sizeof(u64)),
                     o_cat = o_mask + al((size_t)std::max(fmasks.nbatch, 1) * std::max(nloc, 1) * sizeof(u32));
        first_bytes = std::max<size_t>(o_cat + al((size_t)std::max<long long>(tot, 1) * sizeof(int)), (size_t)256 << 10);
        if (rid_pool_alloc(ctx, (void **)&first_buf, first_bytes) != cudaSuccess) { cudaGetLastError(); first_buf = nullptr; }
        int all_ok = 0;
        RID_TRY(rid_agree_all(ctx, first_buf != nullptr, &all_ok)); // (every rank takes the same path)
        if (!all_ok && first_buf) { rid_pool_free(ctx, first_buf, first_bytes); first_buf = nullptr; }
        if (first_buf) {
            d_first = (int2 *)(first_buf + o_first);
            rc = boot_first_medoids(dm, fmasks, samples, B, d_first, (int *)(first_buf + o_cat), (long long *)(first_buf + o_off),
                                    (u64 *)(first_buf + o_sp0), (int *)(first_buf + o_rep), (unsigned short *)(first_buf + o_mask));
            if (rc != RID_OK) { cudaStreamSynchronize(ctx->stream); rid_pool_free(ctx, first_buf, first_bytes); return rc; }
        }
    }
    FullReplica R;
    RID_TRY(replica_begin(dm, &R));
    RID_CUDA(cudaEventRecord(ctx->phase_ev[1], ctx->stream)); // the replica (if any) has arrived
    rid_dmat *work = R.active ? R.full : dm; // the matrix the bootstrap PAM runs read
    const int nshare = R.active ? R.saved_world : 1, myshare = R.active ? R.saved_rank : 0;
    replica_local(&R, true);
    int *d_part = nullptr, *d_tab = nullptr;
    if (rc == RID_OK && (rid_tmp_alloc(ctx, &d_part, (size_t)N * sizeof(int)) != cudaSuccess ||
                         rid_tmp_alloc(ctx, &d_tab, (size_t)k * k * sizeof(int)) != cudaSuccess)) {
        rid_set_error("rid_clusterboot: out of device memory");
        rc = RID_ERR_NOMEM;
    }
    if (rc == RID_OK) {
        if (medoids) for (int c = 0; c < k; ++c) medoids[c] = med_pos[c]; // positions == cells for the full set
        std::vector<int> p0(N);
        for (int i = 0; i < N; ++i) p0[i] = partition[i] - 1; // c1 labels by cell (0-based) on the device
        if (cudaMemcpyAsync(d_part, p0.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = RID_ERR_CUDA;
    }
    std::vector<int> tab((size_t)k * k), blab, bmed(k);
    // this rank's replicates; the host tables of the next one are prepared on a helper thread while the GPU works
    std::vector<int> mine;
    std::vector<const int *> ptr(B);
    {
        const int *sp = samples;
        for (int b = 0; b < B; ++b) {
            ptr[b] = sp;
            sp += sample_len[b];
            for (int j = 0; j < k; ++j) bootresult[(size_t)b * k + j] = 0.0;
            if (b % nshare == myshare) mine.push_back(b);
        }
    }
    HostSubset prep[2];
    std::future<void> fut;
    auto launch_prep = [&](size_t i) {
        const int b = mine[i];
        fut = std::async(std::launch::async, host_subset_prepare, &prep[i & 1], work->N, work->row0, work->row1, ptr[b],
                         sample_len[b], true);
    };
    auto jaccard_row = [&](const int *t, int b) {
        std::vector<long long> ca(k, 0), cb(k, 0);
        for (int a = 0; a < k; ++a)
            for (int c = 0; c < k; ++c) { ca[a] += t[(size_t)a * k + c]; cb[c] += t[(size_t)a * k + c]; }
        for (int j = 0; j < k; ++j) {
            double best = 0.0;
            for (int c = 0; c < k; ++c) {
                long long inter = t[(size_t)j * k + c], uni = ca[j] + cb[c] - inter;
                double cj = uni == 0 ? 0.0 : (double)inter / (double)uni; // clujaccard(zerobyzero=0)
                if (cj > best) best = cj;
            }
            bootresult[(size_t)b * k + j] = best;
        }
    };
    // Replicates are independent and all of a replicate's decisions are taken on the device, so with the whole matrix
    // on this GPU the host only ENQUEUES: tables from page-locked memory, compaction, BUILD, the first SWAP scan plus a
    // few incremental iterations that turn into no-ops at the optimum, labels, contingency table, result copies into the
    // replicate's own page-locked slot -- no host round trip until the end of a wave of replicates.  The launch queue
    // keeps the GPU several replicates ahead of the host, so host hiccups (a descheduled thread, a slow driver call) no
    // longer idle it; stream order alone makes one device workspace enough.
    // iterations enqueued ahead per replicate: c1 (the same data, all cells) shows how hard SWAP has to work; every
    // unused iteration costs a handful of early-exit launches, a replicate that needs more is redone from scratch
    const int ahead = getenv("RID_BOOT_AHEAD") ? std::max(0, atoi(getenv("RID_BOOT_AHEAD")))
                                               : std::min(48, 6 + c1_swaps + c1_swaps / 2);
    bool pipelined = rc == RID_OK && work->ctx->world == 1 && k >= 2 && k * k <= K2MAX && !getenv("RID_SWAP_FULL") &&
                     !getenv("RID_BOOT_SYNC") && !mine.empty();
    // slot layout (ints): subset tables [8 N] | PamState [64, 8-byte aligned] | contingency table [k k]
    const size_t tab_off = (((size_t)8 * N + 1) & ~(size_t)1) + 64;
    const size_t slot_ints = (tab_off + (size_t)k * k + 1) & ~(size_t)1;
    size_t wave = 1;
    int *pin = nullptr, *d_tabs = nullptr;
    if (pipelined) {
        const size_t budget = (size_t)768 << 20; // page-locked bytes per wave
        // (at most 16 replicates per wave: the cancellation poll runs between waves)
        wave = std::max<size_t>(2, std::min<size_t>(std::min<size_t>(mine.size(), 16), budget / (slot_ints * sizeof(int))));
        pin = (int *)rid_pinned(ctx, wave * slot_ints * sizeof(int));
        if (!pin || rid_tmp_alloc(ctx, &d_tabs, wave * (size_t)k * k * sizeof(int)) != cudaSuccess) { cudaGetLastError(); pipelined = false; }
    }
    auto run_sync = [&](int b, HostSubset *H) -> int { // one replicate, start to finish, results on the host
        const int m = sample_len[b];
        blab.resize(m);
        RID_TRY(pam_run(work, ptr[b], m, k, bmed.data(), blab.data(), nullptr, nullptr, nullptr, H));
        PamWs *w = work->ws;
        // Jaccard needs cluster NUMBERS only up to a permutation (max over k'), so the slot ids in lab_pos do.
        RID_CUDA(cudaMemsetAsync(d_tab, 0, (size_t)k * k * sizeof(int), ctx->stream));
        LAUNCH(ctx, k_contingency, nblk(m, 256), 256, 0, w->subset_cells, w->lab_pos, m, d_part, k, d_tab);
        RID_CUDA(cudaMemcpyAsync(tab.data(), d_tab, (size_t)k * k * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        RID_CUDA(cudaStreamSynchronize(ctx->stream));
        jaccard_row(tab.data(), b);
        return RID_OK;
    };
    auto enqueue = [&](int b, HostSubset &H, int *slot, int *dt) -> int { // nothing here waits for the GPU
        const int m = sample_len[b];
        if (k >= m) { rid_set_error("Number of clusters 'k' must be in {1,2, .., n-1}"); return RID_ERR_ARG; }
        RID_TRY(pam_check_na(work, ptr[b], m));
        PamWs *w = nullptr;
        RID_TRY(ws_get(work, k, &w));
        int n_delta = 0;
        RID_TRY(pam_install_subset(work, w, H, slot, d_first ? d_first + b : (const int2 *)nullptr));
        RID_TRY(pam_begin(work, w));
        RID_TRY(pam_build(work, w, 0, k));
        RID_TRY(pam_swap_enqueue(work, w, k, ahead, &n_delta));
        RID_CUDA(cudaMemsetAsync(w->lab_pos, 0, (size_t)w->m * sizeof(int), ctx->stream));
        if (w->nls) LAUNCH(ctx, k_labels_to_pos, nblk(w->nls, 256), 256, 0, w->rows_pos, w->grp, w->nls, w->lab_pos);
        RID_CUDA(cudaMemsetAsync(dt, 0, (size_t)k * k * sizeof(int), ctx->stream));
        LAUNCH(ctx, k_contingency, nblk(m, 256), 256, 0, w->subset_cells, w->lab_pos, m, d_part, k, dt);
        RID_CUDA(cudaMemcpyAsync(slot + tab_off - 64, w->state, sizeof(PamState), cudaMemcpyDeviceToHost, ctx->stream));
        RID_CUDA(cudaMemcpyAsync(slot + tab_off, dt, (size_t)k * k * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        RID_CUDA(cudaGetLastError());
        return RID_OK;
    };
    double host_enqueue_ms = 0.0; // host time spent inside the enqueue calls (a stalled host shows up here)
    if (!mine.empty() && rc == RID_OK) launch_prep(0);
    for (size_t i0 = 0; i0 < mine.size() && rc == RID_OK; i0 += wave) {
        const size_t i1 = std::min(mine.size(), i0 + wave);
        for (size_t i = i0; i < i1 && rc == RID_OK; ++i) {
            fut.get();
            if (i == i0 || !pipelined) rc = rid_check_abort(ctx, !R.active);
            if (rc != RID_OK) break;
            if (i + 1 < mine.size()) launch_prep(i + 1);
            if (pipelined) {
                struct timespec ta, tb;
                clock_gettime(CLOCK_MONOTONIC, &ta);
                rc = enqueue(mine[i], prep[i & 1], pin + (i - i0) * slot_ints, d_tabs + (i - i0) * (size_t)k * k);
                clock_gettime(CLOCK_MONOTONIC, &tb);
                host_enqueue_ms += (tb.tv_sec - ta.tv_sec) * 1e3 + (tb.tv_nsec - ta.tv_nsec) * 1e-6;
            } else rc = run_sync(mine[i], &prep[i & 1]);
        }
        if (!pipelined || rc != RID_OK) continue;
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            rid_set_error("rid_clusterboot: %s", cudaGetErrorString(cudaGetLastError()));
            rc = RID_ERR_CUDA;
            break;
        }
        if (fut.valid()) fut.wait(); // run_sync below prepares its own tables in work->ws->hs; keep the helper thread out
        for (size_t i = i0; i < i1 && rc == RID_OK; ++i) {
            const int *slot = pin + (i - i0) * slot_ints;
            const PamState *st = (const PamState *)(slot + tab_off - 64);
            if (st->done) { ctx->swaps += st->nswaps; jaccard_row(slot + tab_off, mine[i]); }
            else { ctx->stat[1] += 1.0; rc = run_sync(mine[i], nullptr); } // more swaps than were enqueued ahead (rare): redo this one the plain way
        }
    }
    rid_tmp_free(ctx, d_tabs, wave * (size_t)k * k * sizeof(int));
    if (fut.valid()) fut.wait();
    rid_tmp_free(ctx, d_part, (size_t)N * sizeof(int));
    rid_tmp_free(ctx, d_tab, (size_t)k * k * sizeof(int));
    replica_local(&R, false);
    if (R.active) rc = agree_interrupt(ctx, rc);
    if (rc != RID_OK) cudaStreamSynchronize(ctx->stream);
    if (R.active && rc == RID_OK && B > 0) {
        // every rank holds the columns of its own replicates and zeros elsewhere: one sum makes them whole
        double *d_br = nullptr;
        const size_t nb = (size_t)k * B;
        if (rid_tmp_alloc(ctx, &d_br, nb * sizeof(double)) != cudaSuccess) rc = RID_ERR_NOMEM;
        if (rc == RID_OK && cudaMemcpyAsync(d_br, bootresult, nb * sizeof(double), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = RID_ERR_CUDA;
        if (rc == RID_OK) rc = rid_allreduce_f64(ctx, d_br, nb);
        if (rc == RID_OK && (cudaMemcpyAsync(bootresult, d_br, nb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                             cudaStreamSynchronize(ctx->stream) != cudaSuccess)) rc = RID_ERR_CUDA;
        rid_tmp_free(ctx, d_br, nb * sizeof(double));
    }
    replica_end(&R);
    if (first_buf) { cudaStreamSynchronize(ctx->stream); rid_pool_free(ctx, first_buf, first_bytes); }
    if (rc != RID_OK) return rc;
    RID_CUDA(cudaEventRecord(ctx->phase_ev[2], ctx->stream));
    RID_TRY(rid_time_end(ctx));
    {
        float a = 0.f, b = 0.f, c = 0.f;
        cudaEventElapsedTime(&a, ctx->ev0, ctx->phase_ev[0]);
        cudaEventElapsedTime(&b, ctx->phase_ev[0], ctx->phase_ev[1]);
        cudaEventElapsedTime(&c, ctx->phase_ev[1], ctx->phase_ev[2]);
        ctx->stat[2] = a; ctx->stat[3] = b; ctx->stat[4] = c; ctx->stat[6] = (double)ahead; ctx->stat[7] = host_enqueue_ms;
        if (dm->pull_t0 && cudaEventQuery(dm->pull_t1) == cudaSuccess) {
            float pm = 0.f;
            if (cudaEventElapsedTime(&pm, dm->pull_t0, dm->pull_t1) == cudaSuccess) ctx->stat[8] = pm;
        }
        cudaGetLastError();
    }
    return RID_OK;
}

extern "C" int rid_medoids(rid_dmat *dm, const int *labels, int N, int *medoid_idx, int *cluster_ids, int *n_clusters)
{
    if (!dm || !labels || !medoid_idx || N != dm->N) { rid_set_error("rid_medoids: bad arguments"); return RID_ERR_ARG; }
    rid_ctx *ctx = dm->ctx;
    // sort(unique(part)) and the member list in cell order
    std::vector<int> ids;
    for (int i = 0; i < N; ++i) if (labels[i] > 0) ids.push_back(labels[i]);
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    int k = (int)ids.size();
    if (n_clusters) *n_clusters = k;
    if (k == 0) return RID_OK;
    std::vector<int> sub, g;
    for (int i = 0; i < N; ++i)
        if (labels[i] > 0) {
            sub.push_back(i);
            g.push_back((int)(std::lower_bound(ids.begin(), ids.end(), labels[i]) - ids.begin()));
        }
    int m = (int)sub.size();
    RID_TRY(pam_check_na(dm, sub.data(), m));
    PamWs *w;
    RID_TRY(ws_get(dm, k, &w));
    RID_TRY(pam_set_subset(dm, w, sub.data(), m, false));
    // (on the library's stream, which is not ordered against the legacy default stream; pageable source: staged on return)
    RID_CUDA(cudaMemcpyAsync(w->lab_pos, g.data(), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, dm->ctx->stream));
    RID_TRY(rid_time_begin(ctx));
    RID_TRY(grouped_colsums(dm, w, k));
    RID_CUDA(cudaMemsetAsync(w->grp_min, 0xFF, (size_t)k * sizeof(u64), ctx->stream));
    RID_CUDA(cudaMemsetAsync(w->grp_arg, 0x7F, (size_t)k * sizeof(int), ctx->stream));
    for (int pass = 0; pass < 2; ++pass)
        LAUNCH(ctx, k_medoid_min, nblk(m, 256), 256, 0, w->subset, w->lab_pos, m, w->acc + 4 + w->ld, w->ld, w->grp_min,
               w->grp_arg, pass);
    RID_CUDA(cudaGetLastError());
    std::vector<int> arg(k);
    RID_CUDA(cudaMemcpyAsync(arg.data(), w->grp_arg, (size_t)k * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RID_TRY(rid_time_end(ctx));
    for (int c = 0; c < k; ++c) {
        medoid_idx[c] = sub[arg[c]];
        if (cluster_ids) cluster_ids[c] = ids[c];
    }
    return RID_OK;
}

// cluster::silhouette(x = part, dist) as plotsilhouette calls it (reference R/RaceID.R:1283-1284): per cell the neighbour
// cluster and the silhouette width.  labels[N]: positive cluster numbers, 0 = cell not in the partition (gets NaN / 0).
extern "C" int rid_silhouette(rid_dmat *dm, const int *labels, int N, int *neighbor, double *width)
{
    if (!dm || !labels || !neighbor || !width || N != dm->N) { rid_set_error("rid_silhouette: bad arguments"); return RID_ERR_ARG; }
    rid_ctx *ctx = dm->ctx;
    std::vector<int> ids;
    for (int i = 0; i < N; ++i) if (labels[i] > 0) ids.push_back(labels[i]);
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    const int k = (int)ids.size();
    std::vector<int> sub, g;
    for (int i = 0; i < N; ++i)
        if (labels[i] > 0) {
            sub.push_back(i);
            g.push_back((int)(std::lower_bound(ids.begin(), ids.end(), labels[i]) - ids.begin()));
        }
    const int m = (int)sub.size();
    if (k < 2 || k > m - 1) { // silhouette.default: "need 2 <= k <= n-1"
        rid_set_error("silhouette needs 2 <= number of clusters <= n - 1 (got %d clusters for %d cells)", k, m);
        return RID_ERR_ARG;
    }
    RID_TRY(pam_check_na(dm, sub.data(), m));
    PamWs *w;
    RID_TRY(ws_get(dm, k, &w));
    RID_TRY(pam_set_subset(dm, w, sub.data(), m, false));
    // (on the library's stream, which is not ordered against the legacy default stream; pageable source: staged on return)
    RID_CUDA(cudaMemcpyAsync(w->lab_pos, g.data(), (size_t)m * sizeof(int), cudaMemcpyHostToDevice, dm->ctx->stream));
    int *d_nb = nullptr;
    RID_CUDA(cudaMalloc(&d_nb, (size_t)m * sizeof(int)));
    int rc = rid_time_begin(ctx);
    if (rc == RID_OK) rc = grouped_colsums(dm, w, k);
    std::vector<int> h_nb(m);
    std::vector<double> h_sw(m);
    if (rc == RID_OK) {
        cudaMemsetAsync(w->grp_sum, 0, (size_t)k * sizeof(u64), ctx->stream);
        cudaMemsetAsync(w->grp_cnt, 0, (size_t)k * sizeof(int), ctx->stream);
        LAUNCH(ctx, k_wdisp, nblk(m, 256), 256, 0, w->subset, w->lab_pos, m, w->acc + 4 + w->ld, w->ld, w->grp_sum, w->grp_cnt);
        LAUNCH(ctx, k_silhouette_full, nblk(m, 256), 256, 0, w->subset, w->lab_pos, m, k, w->acc + 4 + w->ld, w->ld, w->grp_cnt,
               w->sil, d_nb);
        if (cudaMemcpyAsync(h_nb.data(), d_nb, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync(h_sw.data(), w->sil, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) {
            rid_set_error("rid_silhouette: %s", cudaGetErrorString(cudaGetLastError()));
            rc = RID_ERR_CUDA;
        }
    }
    if (rc == RID_OK) rc = rid_time_end(ctx);
    cudaFree(d_nb);
    if (rc != RID_OK) return rc;
    for (int i = 0; i < N; ++i) { neighbor[i] = 0; width[i] = NAN; }
    for (int p = 0; p < m; ++p) {
        neighbor[sub[p]] = h_nb[p] >= 0 ? ids[h_nb[p]] : 0;
        width[sub[p]] = h_sw[p];
    }
    return RID_OK;
}

extern "C" int rid_refresh(rid_dmat *dm, const int *medoid_idx, int k, int *labels)
{
    if (!dm || !medoid_idx || !labels || k < 1) { rid_set_error("rid_refresh: bad arguments"); return RID_ERR_ARG; }
    rid_ctx *ctx = dm->ctx;
    int N = dm->N;
    for (int c = 0; c < k; ++c)
        if (medoid_idx[c] < 0 || medoid_idx[c] >= N) { rid_set_error("medoid index out of range"); return RID_ERR_ARG; }
    RID_TRY(pam_check_na(dm, nullptr, N));
    PamWs *w;
    RID_TRY(ws_get(dm, k, &w));
    RID_TRY(rid_time_begin(ctx));
    RID_CUDA(cudaMemcpyAsync(w->med, medoid_idx, (size_t)k * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    RID_CUDA(cudaMemsetAsync(w->lab_pos, 0, (size_t)N * sizeof(int), ctx->stream));
    int nloc = dm->row1 - dm->row0;
    if (nloc) LAUNCH(ctx, k_refresh, nblk(nloc, 128), 128, 0, dm->D, dm->ld, dm->row0, nloc, w->med, k, w->lab_pos);
    RID_CUDA(cudaGetLastError());
    RID_TRY(rid_allreduce_i32(ctx, w->lab_pos, (size_t)N));
    RID_CUDA(cudaMemcpyAsync(labels, w->lab_pos, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RID_TRY(rid_time_end(ctx));
    // for (i in max(cpart):1) if (sum(cpart==i)==0) cpart[cpart>i] <- cpart[cpart>i]-1
    std::vector<int> cnt(k + 1, 0), remap(k + 1, 0);
    for (int i = 0; i < N; ++i) cnt[labels[i]]++;
    int next = 0;
    for (int c = 1; c <= k; ++c) if (cnt[c]) remap[c] = ++next;
    for (int i = 0; i < N; ++i) labels[i] = remap[labels[i]];
    return RID_OK;
}

// ----------------------------------------------------------------------------------------------------------
// Row-wise nearest neighbours over the resident matrix: apply(D, 1, function(x) head(order(x), K)) as used at
// reference R/RaceID.R:774 (compdist's knn branch), :1123 (compfr), R/VarID_functions.R:549 (pruneKnn), R/StemID.R:192.
// order() is stable, so ties go to the lower cell index.  One CTA per owned row: exact radix select of the K-th
// smallest 32-bit distance (12 + 12 + 8 bits, three histogram passes over the row), then an index-ordered pick of the
// ties, then a rank sort of the K winners.
// ----------------------------------------------------------------------------------------------------------
#define KNN_THREADS 256
#define KNN_MAXK 256

__device__ __forceinline__ int knn_block_excl_scan(int v, int *sh, int *total)
{
    // exclusive prefix sum over the CTA's threads (KNN_THREADS), result per thread; *total = sum
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();
    if (lane == 31) sh[w] = x;
    __syncthreads();
    if (w == 0) {
        int s = lane < KNN_THREADS / 32 ? sh[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < KNN_THREADS / 32) sh[lane] = s; // inclusive over warps
    }
    __syncthreads();
    int base = w ? sh[w - 1] : 0;
    *total = sh[KNN_THREADS / 32 - 1];
    return base + x - v;
}

__global__ void __launch_bounds__(KNN_THREADS)
k_knn(const u32 *__restrict__ D, size_t ld, int row0, int nloc, int N, int K, double step, int *__restrict__ idx_out,
      double *__restrict__ dist_out)
{
    __shared__ int hist[4096];
    __shared__ int sh_scan[KNN_THREADS / 32];
    __shared__ u32 sel_v[KNN_MAXK];
    __shared__ int sel_i[KNN_MAXK];
    __shared__ u32 s_prefix;
    __shared__ int s_need, s_cnt;
    const int lrow = blockIdx.x;
    if (lrow >= nloc) return;
    const u32 *row = D + (size_t)lrow * ld;

    // ---- radix select: the K-th smallest value v* and how many entries are strictly below it ----
    u32 prefix = 0;   // bits fixed so far
    int need = K;     // rank still to find inside the current prefix bucket
    const int shifts[3] = {20, 8, 0}, widths[3] = {12, 12, 8};
    u32 mask = 0;
    for (int pass = 0; pass < 3; ++pass) {
        const int nb = 1 << widths[pass];
        for (int i = threadIdx.x; i < nb; i += KNN_THREADS) hist[i] = 0;
        __syncthreads();
        for (int j = threadIdx.x; j < N; j += KNN_THREADS) {
            u32 v = row[j];
            if ((v & mask) == prefix) atomicAdd(&hist[(v >> shifts[pass]) & (nb - 1)], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = 0, b = 0;
            for (; b < nb; ++b) {
                if (run + hist[b] >= need) break;
                run += hist[b];
            }
            s_prefix = prefix | ((u32)b << shifts[pass]);
            s_need = need - run;
        }
        __syncthreads();
        prefix = s_prefix;
        need = s_need;
        mask |= (u32)(nb - 1) << shifts[pass];
        __syncthreads();
    }
    const u32 vstar = prefix; // K-th smallest value; `need` of the entries equal to it are wanted (lowest indices)
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    // ---- everything below v*, in any order (sorted afterwards) ----
    for (int j = threadIdx.x; j < N; j += KNN_THREADS) {
        u32 v = row[j];
        if (v < vstar) {
            int p = atomicAdd(&s_cnt, 1);
            sel_v[p] = v;
            sel_i[p] = j;
        }
    }
    __syncthreads();
    const int nless = s_cnt;
    // ---- the first `need` entries equal to v*, by index ----
    int taken = 0;
    for (int j0 = 0; j0 < N && taken < need; j0 += KNN_THREADS) {
        const int j = j0 + threadIdx.x;
        const int hit = (j < N && row[j] == vstar) ? 1 : 0;
        int total;
        const int off = knn_block_excl_scan(hit, sh_scan, &total);
        if (hit && taken + off < need) {
            sel_v[nless + taken + off] = vstar;
            sel_i[nless + taken + off] = j;
        }
        taken += total;
        __syncthreads();
    }
    __syncthreads();
    // ---- rank sort of the K winners by (value, index) ----
    const int g = row0 + lrow;
    for (int t = threadIdx.x; t < K; t += KNN_THREADS) {
        const u32 v = sel_v[t];
        const int i = sel_i[t];
        int rank = 0;
        for (int u = 0; u < K; ++u) rank += (sel_v[u] < v) || (sel_v[u] == v && sel_i[u] < i);
        idx_out[(size_t)g * K + rank] = i;
        if (dist_out) dist_out[(size_t)g * K + rank] = v == RID_Q_NA ? __longlong_as_double(0x7ff8000000000000LL) : (double)v * step;
    }
}

extern "C" int rid_knn(rid_dmat *dm, int K, int *idx, double *dist)
{
    if (!dm || !idx || K < 1 || K > KNN_MAXK || K > dm->N) {
        rid_set_error("rid_knn: K must be in 1..min(N, %d)", KNN_MAXK);
        return RID_ERR_ARG;
    }
    rid_ctx *ctx = dm->ctx;
    RID_CUDA(cudaSetDevice(ctx->device));
    const int N = dm->N, nloc = dm->row1 - dm->row0;
    const size_t n = (size_t)N * K;
    int *d_idx = nullptr;
    double *d_dist = nullptr;
    RID_CUDA(cudaMalloc(&d_idx, n * sizeof(int)));
    if (dist) RID_CUDA(cudaMalloc(&d_dist, n * sizeof(double)));
    RID_TRY(rid_time_begin(ctx));
    RID_CUDA(cudaMemsetAsync(d_idx, 0, n * sizeof(int), ctx->stream));
    if (dist) RID_CUDA(cudaMemsetAsync(d_dist, 0, n * sizeof(double), ctx->stream));
    if (nloc) LAUNCH(ctx, k_knn, nloc, KNN_THREADS, 0, dm->D, dm->ld, dm->row0, nloc, N, K, dm->step, d_idx, d_dist);
    RID_CUDA(cudaGetLastError());
    RID_TRY(rid_allreduce_i32(ctx, d_idx, n)); // disjoint rows per rank
    if (dist) RID_TRY(rid_allreduce_f64(ctx, d_dist, n));
    RID_CUDA(cudaMemcpyAsync(idx, d_idx, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (dist) RID_CUDA(cudaMemcpyAsync(dist, d_dist, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    int rc = rid_time_end(ctx);
    cudaFree(d_idx);
    cudaFree(d_dist);
    return rc;
}
