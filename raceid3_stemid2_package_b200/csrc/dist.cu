// dist.cu — the cell-by-cell distance matrix: replaces dist.gen (reference R/RaceID_utils.R:9-15), i.e.
// 1 - stats::cor(x, method = pearson|spearman) / 1 - cor(log2(x)) / as.matrix(stats::dist(t(x))), as called from
// compdist at R/RaceID.R:769 (algorithms: R's cov.c / distance.c, restated in oracle/oracle.c).
//
// Pipeline:  x (G x N, one column per cell)  --k_prep-->  Z (N x Gp, one unit-norm centred row per cell, FP64)
//            --k_gram_dmma-->  r = Z Z^T on the FP64 tensor path (mma.sync m8n8k4 DMMA; tcgen05 has no FP64 kind)
//            --fused epilogue-->  D = 1 - r (or sqrt(|xi|^2+|xj|^2-2 xi.xj)), clamped, exact-zero diagonal,
//            quantised to 32-bit fixed point and written once into the row-block shard of the HBM-resident matrix.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

#define LAUNCH(ctx, kern, grid, block, smem, ...)                         \
    do {                                                                  \
        kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);    \
        (ctx)->launches++;                                                \
    } while (0)

// ----------------------------------------------------------------------------------------------------------
// k_prep: one CTA per cell.  transform (identity | log2 | average-tie rank), two-pass mean, L2 normalisation.
// ----------------------------------------------------------------------------------------------------------
#define PREP_THREADS 256

__device__ __forceinline__ double block_sum(double v, double *sh)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        v = l < (PREP_THREADS / 32) ? sh[l] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (l == 0) sh[0] = v;
    }
    __syncthreads();
    return sh[0];
}

// metric 0 pearson, 1 spearman, 2 logpearson, 3 euclidean.
// Z row c: unit-norm centred values (correlation metrics) or the raw values (euclidean, with nrm2[c] = |x_c|^2).
// Spearman: rank2 = 2*rank = 2*#less + #equal + 1 is an exact integer; the centred value 2*rank-(G+1) too.
__global__ void __launch_bounds__(PREP_THREADS)
k_prep(const double *__restrict__ x, int G, int N, int Gp, int metric, double *__restrict__ Z,
       unsigned char *__restrict__ zero_var, double *__restrict__ nrm2)
{
    extern __shared__ double sh_keys[]; // spearman only: G keys
    __shared__ double sh_red[PREP_THREADS / 32];
    const int c = blockIdx.x;
    const double *col = x + (size_t)c * G;
    double *zrow = Z + (size_t)c * Gp;

    if (metric == 3) {
        double s = 0.0;
        for (int k = threadIdx.x; k < G; k += PREP_THREADS) {
            double v = col[k];
            zrow[k] = v;
            s += v * v;
        }
        s = block_sum(s, sh_red);
        if (threadIdx.x == 0) { nrm2[c] = s; zero_var[c] = 0; }
        return;
    }
    if (metric == 1) {
        for (int k = threadIdx.x; k < G; k += PREP_THREADS) sh_keys[k] = col[k];
        __syncthreads();
        double ss = 0.0;
        for (int k = threadIdx.x; k < G; k += PREP_THREADS) {
            const double v = sh_keys[k];
            int less = 0, eq = 0;
            for (int l = 0; l < G; ++l) {
                const double u = sh_keys[l];
                less += (u < v);
                eq += (u == v);
            }
            double cv = (double)(2 * less + eq + 1 - (G + 1)); // 2*(rank - mean rank), exact
            zrow[k] = cv;
            ss += cv * cv; // exact: integers below 2^53
        }
        ss = block_sum(ss, sh_red);
        const double inv = ss > 0.0 ? 1.0 / sqrt(ss) : 0.0;
        for (int k = threadIdx.x; k < G; k += PREP_THREADS) zrow[k] *= inv; // same thread wrote it
        if (threadIdx.x == 0) zero_var[c] = ss > 0.0 ? 0 : 1;
        return;
    }
    // pearson / logpearson: two-pass mean like cov.c, then centred sum of squares
    double s = 0.0;
    for (int k = threadIdx.x; k < G; k += PREP_THREADS) {
        double v = col[k];
        if (metric == 2) v = log2(v);
        s += v;
    }
    double mean = block_sum(s, sh_red) / (double)G;
    s = 0.0;
    for (int k = threadIdx.x; k < G; k += PREP_THREADS) {
        double v = col[k];
        if (metric == 2) v = log2(v);
        s += v - mean;
    }
    mean += block_sum(s, sh_red) / (double)G;
    double ss = 0.0;
    for (int k = threadIdx.x; k < G; k += PREP_THREADS) {
        double v = col[k];
        if (metric == 2) v = log2(v);
        v -= mean;
        ss += v * v;
    }
    ss = block_sum(ss, sh_red);
    const double inv = ss > 0.0 ? 1.0 / sqrt(ss) : 0.0;
    for (int k = threadIdx.x; k < G; k += PREP_THREADS) {
        double v = col[k];
        if (metric == 2) v = log2(v);
        zrow[k] = (v - mean) * inv;
    }
    if (threadIdx.x == 0) zero_var[c] = ss > 0.0 ? 0 : 1;
}

// Spearman through a sort: one CTA per cell, bitonic sort of (value, gene) in shared memory (G padded to a power of two
// with +inf), then for every sorted position the tie group [lo, hi) by two binary searches: 2*rank = lo + hi + 1 (the
// average of the ranks lo+1 .. hi), an exact integer like the counting kernel's 2*#less + #equal + 1 -- O(G log^2 G)
// per cell instead of O(G^2): 100k cells x 3k genes take ~3 ms instead of ~250 ms.  G <= 16384 (196 KB of shared memory).
#define RANK_THREADS 512
__global__ void __launch_bounds__(RANK_THREADS)
k_prep_rank_sort(const double *__restrict__ x, int G, int P /* power of two >= G */, int Gp, double *__restrict__ Z,
                 unsigned char *__restrict__ zero_var)
{
    extern __shared__ double sh_keys[];              // [P]
    int *sh_idx = reinterpret_cast<int *>(sh_keys + P); // [P]
    __shared__ double sh_red[RANK_THREADS / 32];
    const int c = blockIdx.x;
    const double *col = x + (size_t)c * G;
    double *zrow = Z + (size_t)c * Gp;
    for (int i = threadIdx.x; i < P; i += RANK_THREADS) {
        sh_keys[i] = i < G ? col[i] : __longlong_as_double(0x7ff0000000000000LL);
        sh_idx[i] = i;
    }
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += RANK_THREADS) {
                // the t-th pair of this stage: i has bit j clear, partner i | j
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool asc = (i & k) == 0;
                const double a = sh_keys[i], b = sh_keys[p];
                if ((a > b) == asc && a != b) {
                    sh_keys[i] = b; sh_keys[p] = a;
                    const int ia = sh_idx[i];
                    sh_idx[i] = sh_idx[p]; sh_idx[p] = ia;
                }
            }
            __syncthreads();
        }
    }
    double ss = 0.0;
    for (int i = threadIdx.x; i < G; i += RANK_THREADS) {
        const double v = sh_keys[i];
        int lo = 0, hi = G; // lower_bound
        int a = 0, b = i;   // first position with key == v lies in [0, i]
        while (a < b) { const int mid = (a + b) >> 1; if (sh_keys[mid] < v) a = mid + 1; else b = mid; }
        lo = a;
        a = i + 1; b = G;   // first position with key > v lies in [i + 1, G]
        while (a < b) { const int mid = (a + b) >> 1; if (sh_keys[mid] <= v) a = mid + 1; else b = mid; }
        hi = a;
        const double cv = (double)(lo + hi + 1 - (G + 1)); // 2 * (rank - mean rank), exact
        zrow[sh_idx[i]] = cv;
        ss += cv * cv; // exact: integers below 2^53
    }
    // block sum (RANK_THREADS threads)
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_down_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) sh_red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        ss = threadIdx.x < RANK_THREADS / 32 ? sh_red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_down_sync(0xffffffffu, ss, o);
        if (threadIdx.x == 0) sh_red[0] = ss;
    }
    __syncthreads();
    ss = sh_red[0];
    const double inv = ss > 0.0 ? 1.0 / sqrt(ss) : 0.0;
    __syncthreads(); // every zrow write of this CTA is done (different threads wrote them)
    for (int k = threadIdx.x; k < G; k += RANK_THREADS) zrow[k] *= inv;
    if (threadIdx.x == 0) zero_var[c] = ss > 0.0 ? 0 : 1;
}

__global__ void k_max_f64(const double *__restrict__ x, int n, double *out)
{
    __shared__ double sh[256];
    double m = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, x[i]);
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

// ----------------------------------------------------------------------------------------------------------
// k_gram_dmma: 128 x 128 tile of Z Z^T per CTA, K in slabs of 16, 3-stage cp.async pipeline, 8 warps (2 x 4),
// each warp a 64 x 32 sub-tile = 8 x 4 DMMA m8n8k4 accumulators.  Shared-memory rows are padded to 20 doubles so
// that the half-warp fragment loads (4 rows x 4 k) hit 16 distinct 8-byte bank pairs.
// ----------------------------------------------------------------------------------------------------------
#define GM_BM 128
#define GM_BN 128
#define GM_BK 16
#define GM_PITCH 20
#define GM_STAGES 3
#define GM_THREADS 256
#define GM_SMEM (GM_STAGES * 2 * GM_BM * GM_PITCH * 8)

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct GramParams {
    const double *Z;   // Np x Gp (rows beyond N are zero)
    int Gp, N, row0, nloc;
    u32 *D;
    size_t ld;
    const unsigned char *zero_var;
    int euclid;
    const double *nrm2; // euclid
    const double *xraw; // euclid fallback: == Z
    int G;
    double inv_step;    // 1/step
    int sym;            // 1: grid covers only tiles with col-tile >= row-tile; the mirror tile is written too
    int tiles;          // tiles per dimension (sym)
    // multi-rank symmetric mode: every off-diagonal tile pair is computed by exactly one of its two owners, which
    // stores the tile into its own shard and the transposed tile straight into the peer's shard over NVLink
    int psym, world, rank, tiles_per_rank;
    u32 *peer[16];
};

__device__ __forceinline__ u32 quantise_corr(double r, double inv_step)
{
    // cor(): clamp to [-1,1]; dist.gen: 1 - r
    r = fmin(1.0, fmax(-1.0, r));
    return (u32)__double2ull_rn((1.0 - r) * inv_step);
}

__global__ void __launch_bounds__(GM_THREADS, 1) k_gram_dmma(GramParams P)
{
    extern __shared__ __align__(16) double smem[];
    double *As = smem;                                   // [stage][128][20]
    double *Bs = smem + GM_STAGES * GM_BM * GM_PITCH;    // [stage][128][20]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3; // 2 x 4 warps
    int bm = blockIdx.y, bn = blockIdx.x;
    if (P.sym) {
        // linear index over the upper triangle (row-major): t = bm*T - bm(bm-1)/2 + (bn - bm)
        const long long T = P.tiles, t = blockIdx.x;
        long long b = (long long)((2.0 * T + 1.0 - sqrt((2.0 * T + 1.0) * (2.0 * T + 1.0) - 8.0 * (double)t)) * 0.5);
        while (b > 0 && b * T - b * (b - 1) / 2 > t) --b;
        while ((b + 1) * T - (b + 1) * b / 2 <= t) ++b;
        bm = (int)b;
        bn = (int)(t - (b * T - b * (b - 1) / 2) + b);
    }
    int tile_i = bm;      // global tile row
    int owner_j = P.rank; // rank owning tile row `bn` (the mirror's destination)
    if (P.psym) {
        tile_i = P.row0 / GM_BM + bm;
        owner_j = min(bn / P.tiles_per_rank, P.world - 1);
        bool mine;
        if (owner_j == P.rank) {
            mine = bn >= tile_i;
        } else {
            const int d = (owner_j - P.rank + P.world) % P.world;
            mine = 2 * d < P.world || (2 * d == P.world && ((tile_i + bn) & 1) == (P.rank < owner_j ? 0 : 1));
        }
        if (!mine) return;
    }
    const size_t a_row0 = (size_t)P.row0 + (size_t)bm * GM_BM; // global cell index of the tile's first row
    const size_t b_row0 = (size_t)bn * GM_BN;
    const int nslab = P.Gp / GM_BK;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // each stage: A 128 rows x 16 doubles = 1024 16-byte chunks; same for B; 256 threads -> 4 + 4 chunks
    auto load_stage = [&](int stage, int slab) {
        double *a = As + (size_t)stage * GM_BM * GM_PITCH;
        double *b = Bs + (size_t)stage * GM_BN * GM_PITCH;
        const int k0 = slab * GM_BK;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int ch = tid + i * GM_THREADS; // 0..1023
            int r = ch >> 3, c2 = (ch & 7) * 2;
            cp_async16(a + r * GM_PITCH + c2, P.Z + (a_row0 + r) * P.Gp + k0 + c2);
            cp_async16(b + r * GM_PITCH + c2, P.Z + (b_row0 + r) * P.Gp + k0 + c2);
        }
    };

#pragma unroll
    for (int s = 0; s < GM_STAGES - 1; ++s) {
        if (s < nslab) load_stage(s, s);
        cp_async_commit();
    }
    for (int slab = 0; slab < nslab; ++slab) {
        cp_async_wait<GM_STAGES - 2>();
        __syncthreads();
        {
            int nxt = slab + GM_STAGES - 1;
            if (nxt < nslab) load_stage(nxt % GM_STAGES, nxt);
            cp_async_commit();
        }
        const double *a = As + (size_t)(slab % GM_STAGES) * GM_BM * GM_PITCH + (wm * 64 + (lane >> 2)) * GM_PITCH + (lane & 3);
        const double *b = Bs + (size_t)(slab % GM_STAGES) * GM_BN * GM_PITCH + (wn * 32 + (lane >> 2)) * GM_PITCH + (lane & 3);
#pragma unroll
        for (int kk = 0; kk < GM_BK; kk += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = a[i * 8 * GM_PITCH + kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = b[j * 8 * GM_PITCH + kk];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: c0 -> (row = lane/4, col = 2*(lane%4)), c1 -> col+1
    const bool mirror = (P.sym || P.psym) && bn != tile_i;
    u32 *stage = reinterpret_cast<u32 *>(smem); // [128][129] u32 = 66 KB, reuses the (drained) pipeline buffers
    if (mirror) __syncthreads();                // every warp is done reading the last slab
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int trow = wm * 64 + i * 8 + (lane >> 2);
        const int lrow = bm * GM_BM + trow;
        const bool row_ok = lrow < P.nloc;
        const int grow = P.row0 + lrow;
        const bool zr = row_ok ? P.zero_var[grow] : false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int tcol = wn * 32 + j * 8 + 2 * (lane & 3);
            const int gcol = bn * GM_BN + tcol;
            u32 q[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int cc = gcol + e;
                u32 v;
                if (cc >= P.N || !row_ok) {
                    v = 0u;
                } else if (cc == grow) {
                    v = zr ? RID_Q_NA : 0u; // diagonal: exactly 0 (1 - 1), NA for a zero-variance cell
                } else if (!P.euclid) {
                    v = (zr || P.zero_var[cc]) ? RID_Q_NA : quantise_corr(acc[i][j][e], P.inv_step);
                } else {
                    const double s = P.nrm2[grow] + P.nrm2[cc];
                    double d2 = s - 2.0 * acc[i][j][e];
                    if (d2 < 1e-9 * s) {
                        // near-duplicate cells: the Gram form cancels; redo this pair in difference form (dist())
                        const double *xa = P.xraw + (size_t)grow * P.Gp, *xb = P.xraw + (size_t)cc * P.Gp;
                        d2 = 0.0;
                        for (int k = 0; k < P.G; ++k) {
                            double dv = xa[k] - xb[k];
                            d2 += dv * dv;
                        }
                    }
                    v = (u32)min(__double2ull_rn(sqrt(fmax(d2, 0.0)) * P.inv_step), 0x7FFFFFFFull); // q < 2^31: two rows add up in 32 bits
                }
                q[e] = v;
            }
            if (row_ok) *reinterpret_cast<uint2 *>(P.D + (size_t)lrow * P.ld + gcol) = make_uint2(q[0], q[1]);
            if (mirror) {
                stage[trow * 129 + tcol] = q[0];
                stage[trow * 129 + tcol + 1] = q[1];
            }
        }
    }
    if (mirror) {
        // D is symmetric: the same values, transposed, are the tile (bn, bm).  Column reads of the padded staging
        // tile are bank-conflict free; every warp store is 128 contiguous bytes.
        __syncthreads();
        u32 *mbase = P.psym ? P.peer[owner_j] : P.D;
        const int mrow0 = P.psym ? owner_j * P.tiles_per_rank * GM_BM : 0; // first row of the destination shard
        for (int c = warp; c < GM_BN; c += GM_THREADS / 32) {
            const int mrow = bn * GM_BN + c; // global row of the mirror tile (a column of this tile)
            if (mrow >= P.N) break;
            u32 *dst = mbase + (size_t)(mrow - mrow0) * P.ld + (size_t)tile_i * GM_BM;
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[lane + 32 * i] = stage[(lane + 32 * i) * 129 + c];
        }
    }
}

// ----------------------------------------------------------------------------------------------------------
// import / export
// ----------------------------------------------------------------------------------------------------------
// src: N x N column-major doubles (device).  Lower triangle is authoritative (as.dist): D(a,b) = src[max,min].
__global__ void k_import(const double *__restrict__ src, int N, int row0, int nloc, u32 *__restrict__ D, size_t ld,
                         double inv_step, u32 qmax, int *has_na)
{
    size_t col = (size_t)blockIdx.y * blockDim.x + threadIdx.x;
    int lrow = blockIdx.x;
    if (col >= ld || lrow >= nloc) return;
    int r = row0 + lrow;
    u32 v = 0u;
    if (col < (size_t)N && (int)col != r) {
        int hi = max(r, (int)col), lo = min(r, (int)col);
        double d = src[(size_t)lo * N + hi];
        if (isnan(d)) { v = RID_Q_NA; *has_na = 1; }
        else v = (u32)min(__double2ull_rn(d * inv_step), (unsigned long long)qmax);
    }
    D[(size_t)lrow * ld + col] = v;
}

// out: m x n column-major doubles; rows ridx (global cells) owned by this rank are filled, others left 0
__global__ void k_export(const u32 *__restrict__ D, size_t ld, int row0, int row1, const int *__restrict__ ridx, int m,
                         const int *__restrict__ cidx, int n, double step, double *__restrict__ out)
{
    int a = blockIdx.x * blockDim.x + threadIdx.x; // row position (fast index of the column-major output)
    int b = blockIdx.y;
    if (a >= m || b >= n) return;
    int r = ridx ? ridx[a] : a, c = cidx ? cidx[b] : b;
    if (r < row0 || r >= row1) return;
    u32 q = D[(size_t)(r - row0) * ld + c];
    out[(size_t)b * m + a] = q == RID_Q_NA ? __longlong_as_double(0x7ff8000000000000LL) : (double)q * step;
}

// euclidean matrices whose range makes the 31-bit grid coarser than the 1e-6 budget: exported distances are recomputed
// from the retained feature rows in FP64 difference form, like stats::dist (sum of squared differences, then sqrt)
__global__ void k_export_exact(const double *__restrict__ x, int ldx, int G, int row0, int row1,
                               const int *__restrict__ ridx, int m, const int *__restrict__ cidx, int c_first, int n,
                               double *__restrict__ out)
{
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    int b = blockIdx.y;
    if (a >= m || b >= n) return;
    int r = ridx ? ridx[a] : a, c = cidx ? cidx[b] : c_first + b;
    if (r < row0 || r >= row1) return;
    const double *xa = x + (size_t)r * ldx, *xb = x + (size_t)c * ldx;
    double d2 = 0.0;
    for (int k = 0; k < G; ++k) {
        const double dv = xa[k] - xb[k];
        d2 += dv * dv;
    }
    out[(size_t)b * m + a] = sqrt(d2);
}

// ----------------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------------
static void shard_rows(int N, int rank, int world, int *r0, int *r1)
{
    // contiguous row blocks; block size rounded to 128 rows so GEMM tiles do not straddle shards
    long long per = ((long long)N + world - 1) / world;
    per = (per + 127) / 128 * 128;
    long long a = std::min<long long>((long long)rank * per, N), b = std::min<long long>(a + per, N);
    *r0 = (int)a;
    *r1 = (int)b;
}

// host-only: the row block rank `rank` of `world` owns (no GPU needed; the N>1 CPU tests check it)
extern "C" int rid_shard_rows(int N, int rank, int world, int *row0, int *row1)
{
    if (N < 1 || world < 1 || rank < 0 || rank >= world || !row0 || !row1) {
        rid_set_error("rid_shard_rows: bad arguments");
        return RID_ERR_ARG;
    }
    shard_rows(N, rank, world, row0, row1);
    return RID_OK;
}

// Multi-rank: when the whole matrix fits next to the work buffers on EVERY rank, each rank allocates all N rows and
// its shard is the row block [row0, row1) inside that buffer; the other blocks are pulled over NVLink later
// (rid_replica_start).  Otherwise only the shard is allocated and every clustering phase runs cooperatively.
static int dmat_alloc(rid_ctx *ctx, int N, rid_dmat **out)
{
    rid_dmat *dm = new rid_dmat();
    dm->ctx = ctx;
    dm->N = N;
    dm->world = ctx->world;
    dm->rank = ctx->rank;
    shard_rows(N, ctx->rank, ctx->world, &dm->row0, &dm->row1);
    dm->ld = round_up((size_t)N, 128);
    size_t nloc = (size_t)(dm->row1 - dm->row0);
    if (ctx->world > 1) { // every shard is allocated at the size of rank 0's
        int r0, r1;
        shard_rows(N, 0, ctx->world, &r0, &r1);
        dm->alloc_rows = (size_t)std::max(r1 - r0, 1);
    } else {
        dm->alloc_rows = std::max<size_t>(nloc, 1);
    }
    const size_t shard_bytes = dm->alloc_rows * dm->ld * sizeof(u32);
    cudaError_t e = cudaSuccess;
    if (ctx->world > 1) {
        // mappings of buffers nobody uses any more pile up when shapes change: drop them while no sharded matrix lives
        if (ctx->live_sharded == 0 && ctx->exported.size() > 4) {
            int rc = rid_ipc_release_all(ctx);
            if (rc != RID_OK) { delete dm; return rc; }
        }
        const size_t full_bytes = (size_t)ctx->world * shard_bytes;
        int want = getenv("RID_NO_REPLICA") ? 0 : 1;
        void *buf = nullptr;
        if (want) {
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            for (auto &b : ctx->pool) free_b += b.bytes; // cached buffers are given back on demand
            // replica + a compact bootstrap copy (~0.45 N^2 words) + operands / accumulators + slack
            const size_t need = full_bytes + (size_t)(0.45 * (double)N * (double)dm->ld * 4.0) + ((size_t)6 << 30);
            if (free_b < need || rid_pool_alloc(ctx, &buf, full_bytes) != cudaSuccess) { cudaGetLastError(); buf = nullptr; want = 0; }
        }
        int all = 0;
        int rc = rid_agree_all(ctx, want, &all);
        if (rc != RID_OK) { if (buf) rid_pool_free(ctx, buf, full_bytes); delete dm; return rc; }
        if (all) {
            dm->full = (u32 *)buf;
            dm->full_bytes = full_bytes;
            dm->D = dm->full + (size_t)ctx->rank * dm->alloc_rows * dm->ld;
        } else {
            if (buf) rid_pool_free(ctx, buf, full_bytes); // some rank cannot hold the replica: everybody shards
            e = rid_pool_alloc(ctx, (void **)&dm->D, shard_bytes);
        }
    } else {
        e = rid_pool_alloc(ctx, (void **)&dm->D, shard_bytes);
    }
    if (e == cudaSuccess && dm->alloc_rows > nloc) // rows nobody owns (last rank): keep them defined
        cudaMemsetAsync(dm->D + nloc * dm->ld, 0, (dm->alloc_rows - nloc) * dm->ld * sizeof(u32), ctx->stream);
    if (e != cudaSuccess) {
        rid_set_error("cannot allocate %.2f GB for the distance-matrix shard (%zu x %zu): %s",
                      (double)nloc * dm->ld * 4 / 1e9, nloc, dm->ld, cudaGetErrorString(e));
        delete dm;
        return RID_ERR_NOMEM;
    }
    if (ctx->world > 1) {
        // every rank's row block, mapped once per matrix (CUDA IPC; cached per allocation): the distance kernels store
        // mirror tiles through these pointers, the replica pulls read through them
        void *peers[16];
        int ok = 0;
        int rc = rid_peer_pointers(ctx, dm->D, peers, &ok);
        if (rc != RID_OK) { rid_pool_free(ctx, dm->full ? (void *)dm->full : (void *)dm->D, dm->full ? dm->full_bytes : shard_bytes); delete dm; return rc; }
        dm->peers_ok = ok;
        for (int r = 0; r < ctx->world; ++r) {
            dm->peer_shard[r] = ok ? (u32 *)peers[r] : nullptr;
            dm->peer_full[r] = (ok && dm->full) ? (u32 *)peers[r] - (size_t)r * dm->alloc_rows * dm->ld : nullptr;
        }
        ctx->live_sharded++;
    }
    *out = dm;
    return RID_OK;
}

// Fill dm->full with what the bootstrap replicates read of the other ranks' row blocks.  Call after the barrier that
// ends the distance kernels (every shard complete everywhere).  One-sided: no rank waits for another here.
//
// The symmetric compaction of a replicate reads, of row block s, only the columns >= (first row of s) - margin (upper
// triangle plus a margin for its diagonal super-tiles); rid_replica_complete fetches the rest on demand.  Fetching that
// upper part of every block from its owner would make rank 0 send seven times its whole block and rank W-1 almost
// nothing (measured at 8 GPUs: 27 ms of pulls on rank 0, 68 ms on the slowest rank, and the early ranks' replicates
// slowed down by the peers still reading their memory).  D is symmetric, so block (s, t), s < t, also exists transposed
// as block (t, s) in the shard of t's owner: each off-diagonal block is fetched from whichever of its two holders keeps
// the senders balanced (the same circulant rule for every reader) -- as it is, by the copy engines
// (cudaMemcpy2DAsync from the owner of s), or transposed on the way by a small kernel that reads the peer's rows over
// NVLink (k_pull_transpose from the owner of t).  Blocks (s, me) are transposes of this rank's own rows: no NVLink.
#define RID_REPLICA_MARGIN 4096
__global__ void __launch_bounds__(256)
k_pull_transpose(const u32 *__restrict__ src, u32 *__restrict__ dst, size_t ld, int nr /* rows of src = columns of dst */,
                 int nc /* columns of src = rows of dst */)
{
    __shared__ u32 tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int r = by + j, c = bx + threadIdx.x;
        tile[j][threadIdx.x] = (r < nr && c < nc) ? src[(size_t)r * ld + c] : 0u;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int r = bx + j, c = by + threadIdx.x; // dst row = src column, dst column = src row
        if (r < nc && c < nr) dst[(size_t)r * ld + c] = tile[threadIdx.x][j];
    }
}

int rid_replica_start(rid_dmat *dm)
{
    rid_ctx *ctx = dm->ctx;
    if (dm->world == 1 || !dm->full || !dm->peers_ok || dm->replica_state != 0) return RID_OK;
    if (!dm->replica_ev) RID_CUDA(cudaEventCreateWithFlags(&dm->replica_ev, cudaEventDisableTiming));
    RID_CUDA(cudaEventRecord(dm->replica_ev, ctx->stream));
    RID_CUDA(cudaStreamWaitEvent(ctx->copy_stream, dm->replica_ev, 0));
    if (!dm->pull_t0) { RID_CUDA(cudaEventCreate(&dm->pull_t0)); RID_CUDA(cudaEventCreate(&dm->pull_t1)); }
    RID_CUDA(cudaEventRecord(dm->pull_t0, ctx->copy_stream));
    const bool whole = getenv("RID_REPLICA_FULL") != nullptr;
    const bool direct_only = whole || getenv("RID_REPLICA_DIRECT") != nullptr; // every block from the owner of its rows
    double pulled = 0.0;
    const int margin = getenv("RID_REPLICA_MARGIN") ? std::max(0, atoi(getenv("RID_REPLICA_MARGIN"))) : RID_REPLICA_MARGIN;
    const int W = dm->world, me = dm->rank;
    const size_t per = dm->alloc_rows, pitch = dm->ld * sizeof(u32);
    // block (s, t), s < t: from the owner of s as it is (true) or from the owner of t, transposed (false)
    auto fwd = [&](int s, int t) {
        if (direct_only) return true;
        if (t == me) return false; // a transpose of this rank's own rows
        const int d = t - s;
        return 2 * d < W || (2 * d == W && (((s + t) & 1) == 0));
    };
    auto copy_cols = [&](int s, size_t c0, size_t c1, int rows) -> int { // columns [c0, c1) of row block s from its owner
        if (c1 <= c0 || rows <= 0) return RID_OK;
        const size_t off = (size_t)s * per * dm->ld + c0;
        RID_CUDA(cudaMemcpy2DAsync(dm->full + off, pitch, dm->peer_full[s] + off, pitch, (c1 - c0) * sizeof(u32), (size_t)rows,
                                   cudaMemcpyDeviceToDevice, ctx->copy_stream));
        pulled += (double)(c1 - c0) * sizeof(u32) * (double)rows;
        return RID_OK;
    };
    bool all_whole = true;
    for (int i = 1; i < W; ++i) {
        const int s = (me + i) % W; // staggered: in round i every source serves one destination
        int r0, r1;
        shard_rows(dm->N, s, W, &r0, &r1);
        if (r1 <= r0) continue;
        const size_t lo = whole ? 0 : (size_t)std::max(0, r0 - margin) & ~(size_t)31;
        dm->replica_lo[s] = (int)lo;
        if (lo) all_whole = false;
        // runs of column blocks taken as they are from the owner of s: the diagonal block (with its margin) and every
        // block t > s with fwd(s, t)
        size_t run0 = lo, run1 = std::min(dm->ld, (size_t)(s + 1) * per);
        for (int t = s + 1; t < W; ++t) {
            int q0, q1;
            shard_rows(dm->N, t, W, &q0, &q1);
            if (q1 <= q0) break;
            const size_t c0 = (size_t)t * per, c1 = t == W - 1 ? dm->ld : std::min(dm->ld, (size_t)(t + 1) * per);
            if (fwd(s, t)) {
                if (run1 == c0) run1 = c1;
                else { RID_TRY(copy_cols(s, run0, run1, r1 - r0)); run0 = c0; run1 = c1; }
            } else {
                // transposed from the owner of t: its rows [q0, q1) x columns [r0, r1) -> my rows [r0, r1) x columns [q0, q1)
                const u32 *src = (t == me ? dm->full : dm->peer_full[t]) + (size_t)t * per * dm->ld + (size_t)s * per;
                u32 *dst = dm->full + (size_t)s * per * dm->ld + (size_t)t * per;
                dim3 grid((unsigned)((r1 - r0 + 31) / 32), (unsigned)((q1 - q0 + 31) / 32));
                k_pull_transpose<<<grid, dim3(32, 8), 0, ctx->copy_stream>>>(src, dst, dm->ld, q1 - q0, r1 - r0);
                ctx->launches++;
                if (t != me) pulled += (double)(q1 - q0) * (double)(r1 - r0) * sizeof(u32);
            }
        }
        RID_TRY(copy_cols(s, run0, run1, r1 - r0));
    }
    RID_CUDA(cudaGetLastError());
    ctx->stat[5] = pulled;
    dm->replica_lo[me] = 0;
    RID_CUDA(cudaEventRecord(dm->pull_t1, ctx->copy_stream));
    RID_CUDA(cudaEventRecord(dm->replica_ev, ctx->copy_stream));
    dm->replica_state = all_whole ? 2 : 1;
    return RID_OK;
}

int rid_replica_complete(rid_dmat *dm)
{
    rid_ctx *ctx = dm->ctx;
    if (dm->world == 1 || !dm->full) return RID_OK;
    if (dm->replica_state == 0) RID_TRY(rid_replica_start(dm));
    if (dm->replica_state == 2) return RID_OK;
    for (int i = 1; i < dm->world; ++i) {
        const int s = (dm->rank + i) % dm->world;
        int r0, r1;
        shard_rows(dm->N, s, dm->world, &r0, &r1);
        const size_t lo = (size_t)dm->replica_lo[s];
        if (r1 <= r0 || lo == 0) continue;
        const size_t off = (size_t)s * dm->alloc_rows * dm->ld;
        RID_CUDA(cudaMemcpy2DAsync(dm->full + off, dm->ld * sizeof(u32), dm->peer_full[s] + off, dm->ld * sizeof(u32),
                                   lo * sizeof(u32), (size_t)(r1 - r0), cudaMemcpyDeviceToDevice, ctx->copy_stream));
        dm->replica_lo[s] = 0;
    }
    RID_CUDA(cudaEventRecord(dm->replica_ev, ctx->copy_stream));
    dm->replica_state = 2;
    return RID_OK;
}

// Multi-rank: every rank must free a sharded matrix before the next rid_dist / rid_dmat_from_host call reuses its
// buffer (the peers' pulls read it); the local synchronisation below is all that is needed then.
extern "C" int rid_dmat_free(rid_dmat *dm)
{
    if (!dm) return RID_OK;
    rid_ctx *ctx = dm->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (dm->replica_view) {
        rid_pam_ws_free(dm->replica_view->ws, ctx);
        delete dm->replica_view;
    }
    rid_pam_ws_free(dm->ws, ctx);
    if (dm->replica_ev) cudaEventDestroy(dm->replica_ev);
    if (dm->pull_t0) { cudaEventDestroy(dm->pull_t0); cudaEventDestroy(dm->pull_t1); }
    if (dm->xraw) rid_pool_free(ctx, dm->xraw, dm->xraw_bytes);
    if (dm->full) rid_pool_free(ctx, dm->full, dm->full_bytes);
    else rid_pool_free(ctx, dm->D, dm->alloc_rows * dm->ld * sizeof(u32));
    if (dm->world > 1 && ctx->live_sharded > 0) ctx->live_sharded--;
    delete dm;
    return RID_OK;
}

extern "C" int rid_dmat_info(const rid_dmat *dm, int *N, int *row0, int *row1, double *step, int *has_na)
{
    if (!dm) { rid_set_error("null distance matrix"); return RID_ERR_ARG; }
    if (N) *N = dm->N;
    if (row0) *row0 = dm->row0;
    if (row1) *row1 = dm->row1;
    if (step) *step = dm->step;
    if (has_na) *has_na = dm->has_na;
    return RID_OK;
}

extern "C" int rid_dist(rid_ctx *ctx, const double *x, int G, int N, int metric, int x_on_device, rid_dmat **out,
                        int *n_zero_var)
{
    if (!ctx || !x || !out || G < 2 || N < 2) { rid_set_error("rid_dist: bad arguments"); return RID_ERR_ARG; }
    if (metric < 0 || metric > 3) {
        rid_set_error("metric has to be one of the following: spearman, pearson, logpearson, euclidean (kendall is not "
                      "implemented on the GPU path)");
        return RID_ERR_ARG;
    }
    RID_CUDA(cudaSetDevice(ctx->device));
    if (metric == RID_SPEARMAN && (size_t)G * 8 > 200 * 1024) {
        rid_set_error("spearman: G=%d features exceed the shared-memory rank kernel's limit (25600)", G);
        return RID_ERR_UNSUPPORTED;
    }
    rid_dmat *dm = nullptr;
    RID_TRY(dmat_alloc(ctx, N, &dm));
    const int Gp = (int)round_up((size_t)G, GM_BK);
    const size_t Np = round_up((size_t)N, 128);
    double *xd = nullptr, *Z = nullptr, *nrm2 = nullptr, *dmax = nullptr;
    unsigned char *zv = nullptr;
    char *aux = nullptr; // nrm2[Np] | dmax[32] | zv[Np], one pooled buffer
    const size_t aux_bytes = std::max<size_t>((Np + 32) * sizeof(double) + Np, (size_t)256 << 10);
    int rc = RID_OK;
    // host input on several ranks: every rank uploads only its 1/world slice of the cells over PCIe and the slices are
    // all-gathered over NVLink (the buffer is padded to world equal slices)
    const size_t cells_per = ((size_t)N + ctx->world - 1) / ctx->world;
    const bool sliced_upload = !x_on_device && ctx->world > 1 && !getenv("RID_FULL_UPLOAD");
    const size_t z_bytes = Np * (size_t)Gp * sizeof(double);
    const size_t x_bytes = (sliced_upload ? cells_per * ctx->world : (size_t)N) * (size_t)G * sizeof(double);
    auto fail = [&](int code) {
        rid_pool_free(ctx, Z, z_bytes); rid_pool_free(ctx, aux, aux_bytes);
        if (!x_on_device) rid_pool_free(ctx, xd, x_bytes);
        rid_dmat_free(dm);
        return code;
    };
#define DCHK(call)                                                                                         \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) {                                                                          \
            rid_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));     \
            return fail(RID_ERR_CUDA);                                                                     \
        }                                                                                                  \
    } while (0)
    DCHK(rid_pool_alloc(ctx, (void **)&Z, z_bytes));
    DCHK(rid_pool_alloc(ctx, (void **)&aux, aux_bytes));
    nrm2 = (double *)aux; dmax = nrm2 + Np; zv = (unsigned char *)(dmax + 32);
    if ((rc = rid_time_begin(ctx)) != RID_OK) return fail(rc);
    if (x_on_device) {
        xd = const_cast<double *>(x);
    } else {
        DCHK(rid_pool_alloc(ctx, (void **)&xd, x_bytes));
        if (sliced_upload) {
            const size_t c0 = std::min(cells_per * ctx->rank, (size_t)N), c1 = std::min(c0 + cells_per, (size_t)N);
            const size_t slice = cells_per * (size_t)G * sizeof(double);
            if (c1 > c0)
                DCHK(cudaMemcpyAsync(xd + c0 * G, x + c0 * G, (c1 - c0) * (size_t)G * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            if ((rc = rid_allgather(ctx, (const char *)xd + slice * ctx->rank, xd, slice)) != RID_OK) return fail(rc);
        } else {
            DCHK(cudaMemcpyAsync(xd, x, (size_t)G * N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    DCHK(cudaMemsetAsync(Z, 0, Np * (size_t)Gp * sizeof(double), ctx->stream));
    DCHK(cudaMemsetAsync(nrm2, 0, Np * sizeof(double), ctx->stream));
    DCHK(cudaMemsetAsync(zv, 0, Np, ctx->stream));
    size_t prep_smem = metric == RID_SPEARMAN ? (size_t)G * sizeof(double) : 0;
    if (prep_smem > 48 * 1024)
        DCHK(cudaFuncSetAttribute(k_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prep_smem));
    {
        cudaEvent_t pe = rid_prof_begin(ctx, RID_PROF_PREP);
        if (metric == RID_SPEARMAN && G <= 16384 && !getenv("RID_RANK_COUNT")) {
            // ranks by sorting (RID_RANK_COUNT=1 keeps the O(G^2) counting kernel, which also serves 16384 < G <= 25600)
            int P = 2;
            while (P < G) P <<= 1;
            const size_t smem = (size_t)P * (sizeof(double) + sizeof(int));
            DCHK(cudaFuncSetAttribute(k_prep_rank_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            LAUNCH(ctx, k_prep_rank_sort, N, RANK_THREADS, smem, xd, G, P, Gp, Z, zv);
        } else {
            LAUNCH(ctx, k_prep, N, PREP_THREADS, prep_smem, xd, G, N, Gp, metric, Z, zv, nrm2);
        }
        rid_prof_end(ctx, pe, RID_PROF_PREP, (double)G * N * 16.0, (double)G * N * 32.0);
    }
    DCHK(cudaGetLastError());

    if (metric == RID_EUCLIDEAN) {
        // d(i,j) <= |x_i| + |x_j| <= 2 max|x|: pick the power-of-two step so that every q fits 31 bits
        LAUNCH(ctx, k_max_f64, 1, 256, 0, nrm2, N, dmax);
        double h = 0.0;
        DCHK(cudaMemcpyAsync(&h, dmax, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        DCHK(cudaStreamSynchronize(ctx->stream));
        double bound = 2.0 * sqrt(h);
        int e = 0;
        frexp(bound > 0 ? bound : 1.0, &e); // bound < 2^e
        dm->bits = 31;
        dm->step = ldexp(1.0, e - 31);
    } else {
        dm->bits = 28;
        dm->step = ldexp(1.0, -27); // D in [0,2] -> q in [0, 2^28]
    }
    // exact int8 tensor-core path (tcgen05) for correlation metrics on one GPU; RID_GRAM=f64 forces the FP64 DMMA kernel,
    // RID_GRAM=i8 makes a fallback an error (tests)
    bool gram_done = false;
    {
        const char *mode = getenv("RID_GRAM");
        const bool want_i8 = metric != RID_EUCLIDEAN && ctx->world <= 16 && !(mode && strcmp(mode, "f64") == 0);
        if (want_i8) {
            int used = 0;
            double bound = 0.0;
            if ((rc = rid_gram_i8(ctx, Z, Gp, G, N, zv, dm, &used, &bound)) != RID_OK) return fail(rc);
            gram_done = used != 0;
            if (!gram_done && mode && strcmp(mode, "i8") == 0) {
                rid_set_error("RID_GRAM=i8 but the int8 path was not taken (a-posteriori bound %.3g)", bound);
                return fail(RID_ERR_UNSUPPORTED);
            }
        }
    }
    if (!gram_done) {
        GramParams P;
        P.Z = Z; P.Gp = Gp; P.N = N; P.row0 = dm->row0; P.nloc = dm->row1 - dm->row0; P.D = dm->D; P.ld = dm->ld;
        P.zero_var = zv; P.euclid = metric == RID_EUCLIDEAN; P.nrm2 = nrm2; P.xraw = Z; P.G = G;
        P.inv_step = 1.0 / dm->step;
        // one GPU owns every row: compute only the tiles on or above the diagonal and mirror them (half the flops)
        P.sym = ctx->world == 1 ? 1 : 0;
        P.tiles = (int)(Np / GM_BN);
        P.psym = 0; P.world = ctx->world; P.rank = ctx->rank;
        {
            int r0, r1;
            shard_rows(N, 0, ctx->world, &r0, &r1);
            P.tiles_per_rank = std::max(1, (int)(round_up((size_t)std::max(r1 - r0, 1), GM_BM) / GM_BM));
        }
        for (int r = 0; r < 16; ++r) P.peer[r] = nullptr;
        if (ctx->world > 1 && dm->peers_ok) {
            P.psym = 1;
            for (int r = 0; r < ctx->world; ++r) P.peer[r] = dm->peer_shard[r];
        }
        double tiles_done = 0.0;
        {
            const int tloc = (P.nloc + GM_BM - 1) / GM_BM;
            if (P.sym) tiles_done = (double)P.tiles * (P.tiles + 1) / 2.0;
            else if (!P.psym) tiles_done = (double)P.tiles * tloc;
            else
                for (int bm = 0; bm < tloc; ++bm)
                    for (int bn = 0; bn < P.tiles; ++bn) {
                        int ti = P.row0 / GM_BM + bm, oj = std::min(bn / P.tiles_per_rank, P.world - 1);
                        if (oj == P.rank) tiles_done += bn >= ti;
                        else {
                            int d = (oj - P.rank + P.world) % P.world;
                            tiles_done += (2 * d < P.world || (2 * d == P.world && ((ti + bn) & 1) == (P.rank < oj ? 0 : 1)));
                        }
                    }
        }
        if (P.nloc > 0) {
            DCHK(cudaFuncSetAttribute(k_gram_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, GM_SMEM));
            dim3 grid((unsigned)(Np / GM_BN), (unsigned)((P.nloc + GM_BM - 1) / GM_BM));
            if (P.sym) grid = dim3((unsigned)((long long)P.tiles * (P.tiles + 1) / 2), 1u);
            cudaEvent_t pe = rid_prof_begin(ctx, RID_PROF_GRAM);
            LAUNCH(ctx, k_gram_dmma, grid, GM_THREADS, GM_SMEM, P);
            // work: flops actually issued (tiles computed x 2*128*128*G); the symmetric path issues about half of
            // the algorithmic 2*N*N*G and must not be credited with the tiles it did not compute
            rid_prof_end(ctx, pe, RID_PROF_GRAM, tiles_done * 2.0 * GM_BM * GM_BN * (double)G, (double)P.nloc * dm->ld * 4.0);
            DCHK(cudaGetLastError());
        }
        // peers have written mirror tiles into this rank's shard: nobody may go on before every kernel is done
        if (P.psym && (rc = rid_barrier(ctx)) != RID_OK) return fail(rc);
    }
    // every shard is complete on every rank (both tensor paths end with a barrier): the copy engines start pulling the
    // other ranks' row blocks into the replica while the first clustering phase runs on the shard
    // When to start: measured at 8 GPUs, the 19 GB per rank of pulls and the latency-bound cooperative phases slow each
    // other down (k sweep 56 ms alone, 91 ms next to the pulls; c1 7.9 vs 15.6 ms), so the pulls start as late as they
    // can still hide: at the end of the sweep / the beginning of clusterboot (pam.cu).  RID_REPLICA_EAGER=1: right here.
    if (ctx->world > 1 && dm->full && getenv("RID_REPLICA_EAGER")) {
        // (the pulls need the peers' pointers, and with those both paths have just passed their closing barrier)
        if ((rc = rid_replica_start(dm)) != RID_OK) return fail(rc);
    }
    // zero-variance cells (cor(): NA + warning "the standard deviation is zero")
    dm->na_cell.assign(N, 0);
    DCHK(cudaMemcpyAsync(dm->na_cell.data(), zv, (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = rid_time_end(ctx)) != RID_OK) return fail(rc);
    int nz = 0;
    for (int i = 0; i < N; ++i) nz += dm->na_cell[i];
    dm->has_na = nz > 0;
    if (n_zero_var) *n_zero_var = nz;
    if (metric == RID_EUCLIDEAN && dm->step > 2e-6) {
        // the grid alone would miss the 1e-6 budget (error up to step / 2): keep the raw rows for exact exports
        dm->xraw = Z; dm->xraw_bytes = z_bytes; dm->xraw_G = G; dm->xraw_ld = Gp;
    } else {
        rid_pool_free(ctx, Z, z_bytes);
    }
    rid_pool_free(ctx, aux, aux_bytes);
    if (!x_on_device) rid_pool_free(ctx, xd, x_bytes);
    *out = dm;
    return RID_OK;
#undef DCHK
}

extern "C" int rid_dmat_from_host(rid_ctx *ctx, const double *D, int N, rid_dmat **out)
{
    if (!ctx || !D || !out || N < 2) { rid_set_error("rid_dmat_from_host: bad arguments"); return RID_ERR_ARG; }
    RID_CUDA(cudaSetDevice(ctx->device));
    // range of the lower triangle decides the fixed-point step
    double mx = 0.0;
    bool neg = false;
    for (int c = 0; c < N; ++c)
        for (int r = c + 1; r < N; ++r) {
            double d = D[(size_t)c * N + r];
            if (d < 0) neg = true;
            if (d > mx) mx = d;
        }
    if (neg) { rid_set_error("negative dissimilarities are not supported"); return RID_ERR_ARG; }
    rid_dmat *dm = nullptr;
    RID_TRY(dmat_alloc(ctx, N, &dm));
    if (mx > 1.0 && mx <= 2.0) {
        dm->bits = 28;
        dm->step = ldexp(1.0, -27);
    } else {
        // (also for small ranges: a fixed 2^-27 step would turn a matrix with maximum 1e-4 into a few thousand levels
        // and create ties the real values do not have)
        int e = 0;
        frexp(mx > 0.0 ? mx : 1.0, &e); // mx < 2^e (or == 2^(e-1))
        dm->bits = 31;
        dm->step = ldexp(1.0, e - 31);
    }
    double *src = nullptr;
    int *flag = nullptr;
    if (cudaMalloc(&src, (size_t)N * N * sizeof(double)) != cudaSuccess || cudaMalloc(&flag, sizeof(int)) != cudaSuccess) {
        rid_set_error("cannot stage the %d x %d host matrix on the device", N, N);
        cudaFree(src); rid_dmat_free(dm);
        return RID_ERR_NOMEM;
    }
    int rc = rid_time_begin(ctx);
    cudaMemcpyAsync(src, D, (size_t)N * N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream);
    int nloc = dm->row1 - dm->row0;
    if (nloc > 0) {
        dim3 grid((unsigned)nloc, (unsigned)((dm->ld + 255) / 256));
        LAUNCH(ctx, k_import, grid, 256, 0, src, N, dm->row0, nloc, dm->D, dm->ld, 1.0 / dm->step,
               dm->bits <= 28 ? (u32)(1u << 28) : 0x7FFFFFFFu, flag);
    }
    int h = 0;
    cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (rc == RID_OK) rc = rid_time_end(ctx);
    cudaError_t e = cudaGetLastError();
    cudaFree(src); cudaFree(flag);
    if (rc != RID_OK || e != cudaSuccess) {
        if (e != cudaSuccess) rid_set_error("rid_dmat_from_host: %s", cudaGetErrorString(e));
        rid_dmat_free(dm);
        return RID_ERR_CUDA;
    }
    // NA anywhere: every rank must agree, so look at the host copy
    for (int c = 0; c < N && !h; ++c)
        for (int r = c + 1; r < N; ++r)
            if (D[(size_t)c * N + r] != D[(size_t)c * N + r]) { h = 1; break; }
    dm->has_na = h;
    if (ctx->world > 1 && dm->full && dm->peers_ok) {
        // every rank has imported its own rows: let the copy engines fetch the others
        if ((rc = rid_barrier(ctx)) != RID_OK) { rid_dmat_free(dm); return rc; }
        if (getenv("RID_REPLICA_EAGER") && (rc = rid_replica_start(dm)) != RID_OK) { rid_dmat_free(dm); return rc; }
    }
    *out = dm;
    return RID_OK;
}

extern "C" int rid_dmat_sub(rid_dmat *dm, const int *ridx, int m, const int *cidx, int n, double *out)
{
    if (!dm || !out || m < 1 || n < 1) { rid_set_error("rid_dmat_sub: bad arguments"); return RID_ERR_ARG; }
    rid_ctx *ctx = dm->ctx;
    RID_CUDA(cudaSetDevice(ctx->device));
    if ((!ridx && m > dm->N) || (!cidx && n > dm->N)) { rid_set_error("rid_dmat_sub: block exceeds the matrix"); return RID_ERR_ARG; }
    for (int i = 0; ridx && i < m; ++i) if (ridx[i] < 0 || ridx[i] >= dm->N) { rid_set_error("row index out of range"); return RID_ERR_ARG; }
    for (int i = 0; cidx && i < n; ++i) if (cidx[i] < 0 || cidx[i] >= dm->N) { rid_set_error("column index out of range"); return RID_ERR_ARG; }
    double *d_out = nullptr;
    int *d_r = nullptr, *d_c = nullptr;
    // export in column panels to bound the staging buffer
    const int panel = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>((size_t)n, 65535), ((size_t)256 << 20) / ((size_t)m * 8)));
    RID_CUDA(cudaMalloc(&d_out, (size_t)m * panel * sizeof(double)));
    // (uploads on the library's own stream: it is non-blocking, i.e. not ordered against the legacy default stream)
    if (ridx) { RID_CUDA(cudaMalloc(&d_r, (size_t)m * sizeof(int))); RID_CUDA(cudaMemcpyAsync(d_r, ridx, (size_t)m * sizeof(int), cudaMemcpyHostToDevice, ctx->stream)); }
    if (cidx) { RID_CUDA(cudaMalloc(&d_c, (size_t)n * sizeof(int))); RID_CUDA(cudaMemcpyAsync(d_c, cidx, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream)); }
    int rc = RID_OK;
    for (int c0 = 0; c0 < n && rc == RID_OK; c0 += panel) {
        int nc = std::min(panel, n - c0);
        cudaMemsetAsync(d_out, 0, (size_t)m * nc * sizeof(double), ctx->stream);
        dim3 grid((unsigned)((m + 127) / 128), (unsigned)nc);
        // without a column list the panel's columns are c0.. ; shift through a pointer offset on the matrix
        if (dm->xraw)
            LAUNCH(ctx, k_export_exact, grid, 128, 0, dm->xraw, dm->xraw_ld, dm->xraw_G, dm->row0, dm->row1, d_r, m,
                   d_c ? d_c + c0 : (const int *)nullptr, c0, nc, d_out);
        else if (d_c)
            LAUNCH(ctx, k_export, grid, 128, 0, dm->D, dm->ld, dm->row0, dm->row1, d_r, m, d_c + c0, nc, dm->step, d_out);
        else
            LAUNCH(ctx, k_export, grid, 128, 0, dm->D + c0, dm->ld, dm->row0, dm->row1, d_r, m, (const int *)nullptr, nc, dm->step, d_out);
        rc = rid_allreduce_f64(ctx, d_out, (size_t)m * nc);
        if (rc == RID_OK && cudaMemcpyAsync(out + (size_t)c0 * m, d_out, (size_t)m * nc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = RID_ERR_CUDA;
        if (rc == RID_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = RID_ERR_CUDA;
    }
    if (rc != RID_OK && rc != RID_ERR_CUDA) {}
    else if (rc == RID_ERR_CUDA) rid_set_error("rid_dmat_sub: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_out); cudaFree(d_r); cudaFree(d_c);
    return rc;
}

// ----------------------------------------------------------------------------------------------------------
// measurement hooks: the ceilings the two hot kernels are held against, measured on the spot
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dmma_peak(double *out, int iters, double seed)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = seed * i; c[i][1] = seed; }
    double a = seed + threadIdx.x, b = seed * 0.5 + blockIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s; // never true: keeps the loop alive
}

// issue-rate ceiling of the FP64 tensor pipe (mma.sync.m8n8k4.f64), TFLOP/s
extern "C" int rid_measure_dmma_peak(rid_ctx *ctx, double *tflops)
{
    if (!ctx || !tflops) return RID_ERR_ARG;
    RID_CUDA(cudaSetDevice(ctx->device));
    double *out = nullptr;
    RID_CUDA(cudaMalloc(&out, sizeof(double)));
    const int iters = 20000, blocks = ctx->sm_count * 2;
    k_dmma_peak<<<blocks, 256, 0, ctx->stream>>>(out, 2000, 1e-3); // warm-up
    cudaEvent_t a, b;
    RID_CUDA(cudaEventCreate(&a));
    RID_CUDA(cudaEventCreate(&b));
    RID_CUDA(cudaEventRecord(a, ctx->stream));
    k_dmma_peak<<<blocks, 256, 0, ctx->stream>>>(out, iters, 1e-3);
    RID_CUDA(cudaEventRecord(b, ctx->stream));
    RID_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    RID_CUDA(cudaEventElapsedTime(&ms, a, b));
    *tflops = (double)blocks * 8.0 /*warps*/ * iters * 8.0 * 512.0 / (ms * 1e-3) / 1e12;
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
    return RID_OK;
}

__global__ void __launch_bounds__(256) k_read_peak(const uint4 *__restrict__ p, size_t n16, unsigned *out)
{
    unsigned acc = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 v0 = p[i], v1 = p[i + stride], v2 = p[i + 2 * stride], v3 = p[i + 3 * stride];
        acc += v0.x ^ v0.y ^ v0.z ^ v0.w ^ v1.x ^ v1.y ^ v1.z ^ v1.w ^ v2.x ^ v2.y ^ v2.z ^ v2.w ^ v3.x ^ v3.y ^ v3.z ^ v3.w;
    }
    for (; i < n16; i += stride) { uint4 v = p[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) out[0] = acc;
}

// read-only streaming ceiling over `gib` GiB of device memory (a plain grid-stride 128-bit load loop), GB/s
extern "C" int rid_measure_read_peak(rid_ctx *ctx, double gib, double *gbs)
{
    if (!ctx || !gbs || gib <= 0) return RID_ERR_ARG;
    RID_CUDA(cudaSetDevice(ctx->device));
    size_t bytes = (size_t)(gib * 1024.0 * 1024.0 * 1024.0) / 16 * 16;
    void *buf = nullptr;
    unsigned *out = nullptr;
    if (rid_pool_alloc(ctx, &buf, bytes) != cudaSuccess) { rid_set_error("rid_measure_read_peak: out of memory"); return RID_ERR_NOMEM; }
    RID_CUDA(cudaMalloc(&out, sizeof(unsigned)));
    RID_CUDA(cudaMemsetAsync(buf, 1, bytes, ctx->stream));
    cudaEvent_t a, b;
    RID_CUDA(cudaEventCreate(&a));
    RID_CUDA(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        RID_CUDA(cudaEventRecord(a, ctx->stream));
        k_read_peak<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>((const uint4 *)buf, bytes / 16, out);
        RID_CUDA(cudaEventRecord(b, ctx->stream));
        RID_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        RID_CUDA(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    *gbs = (double)bytes / (best * 1e-3) / 1e9;
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
    rid_pool_free(ctx, buf, bytes);
    return RID_OK;
}

// ----------------------------------------------------------------------------------------------------------
// Fused getfdata -> prep from the package's native storage ("next" row of SURVEY §8f-2): the filtered count matrix is
// a dgCMatrix (CSC: i, p, x); getfdata (R/RaceID.R:723-728) makes (ndata * min(counts))[genes, cells] + 0.1 dense on the
// host and compdist transposes it twice (R/RaceID.R:764,769).  Here the CSC triplet goes to the device as is (about a
// tenth of the dense bytes for UMI data) and one kernel writes the dense feature matrix the prep kernel reads.
// ----------------------------------------------------------------------------------------------------------
__global__ void k_csc_to_features(const int *__restrict__ row_idx, const int *__restrict__ col_ptr,
                                  const double *__restrict__ val, const int *__restrict__ feat_of_row, int G,
                                  const double *__restrict__ counts, double mincount, double *__restrict__ x)
{
    const int c = blockIdx.x;
    double *col = x + (size_t)c * G;
    for (int k = threadIdx.x; k < G; k += blockDim.x) col[k] = 0.1; // 0 * min(counts) + .1
    __syncthreads();
    const double cnt = counts[c];
    for (int t = col_ptr[c] + threadIdx.x; t < col_ptr[c + 1]; t += blockDim.x) {
        const int f = feat_of_row[row_idx[t]];
        // ((expdata / counts) * min(counts)) + .1 in R's order, each step rounded (no FMA contraction)
        if (f >= 0) col[f] = __dadd_rn(__dmul_rn(__ddiv_rn(val[t], cnt), mincount), 0.1);
    }
}

extern "C" int rid_dist_csc(rid_ctx *ctx, const int *row_idx, const int *col_ptr, const double *val, int n_rows_all, int N,
                            const int *feature_rows, int G, const double *counts, int metric, rid_dmat **out,
                            int *n_zero_var)
{
    if (!ctx || !row_idx || !col_ptr || !val || !feature_rows || !counts || !out || N < 2 || G < 2 || n_rows_all < G) {
        rid_set_error("rid_dist_csc: bad arguments");
        return RID_ERR_ARG;
    }
    RID_CUDA(cudaSetDevice(ctx->device));
    const size_t nnz = (size_t)col_ptr[N];
    std::vector<int> feat(n_rows_all, -1);
    for (int g = 0; g < G; ++g) {
        if (feature_rows[g] < 0 || feature_rows[g] >= n_rows_all) { rid_set_error("feature row out of range"); return RID_ERR_ARG; }
        feat[feature_rows[g]] = g;
    }
    double mincount = counts[0];
    for (int c = 1; c < N; ++c) mincount = std::min(mincount, counts[c]);
    // one pooled staging buffer (exact-size reuse: a repeated compdist() on the same object allocates nothing)
    int *d_i = nullptr, *d_p = nullptr, *d_f = nullptr;
    double *d_v = nullptr, *d_c = nullptr, *d_x = nullptr;
    const size_t x_bytes = (size_t)G * N * sizeof(double);
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_v = 0, o_c = o_v + al(std::max<size_t>(nnz, 1) * sizeof(double)), o_i = o_c + al((size_t)N * sizeof(double)),
                 o_p = o_i + al(std::max<size_t>(nnz, 1) * sizeof(int)), o_f = o_p + al(((size_t)N + 1) * sizeof(int)),
                 st_bytes = std::max<size_t>(o_f + al((size_t)n_rows_all * sizeof(int)), (size_t)256 << 10);
    char *stage = nullptr;
    int rc = RID_OK;
    auto done = [&](int code) {
        rid_pool_free(ctx, stage, st_bytes);
        rid_pool_free(ctx, d_x, x_bytes);
        return code;
    };
    if (rid_pool_alloc(ctx, (void **)&stage, st_bytes) != cudaSuccess || rid_pool_alloc(ctx, (void **)&d_x, x_bytes) != cudaSuccess) {
        rid_set_error("rid_dist_csc: out of device memory");
        cudaGetLastError();
        return done(RID_ERR_NOMEM);
    }
    d_v = (double *)(stage + o_v); d_c = (double *)(stage + o_c); d_i = (int *)(stage + o_i); d_p = (int *)(stage + o_p);
    d_f = (int *)(stage + o_f);
    // Several ranks: every rank uploads only the non-zeros of its 1/world slice of the cells over PCIe and the slices
    // are exchanged over NVLink (one in-place broadcast per rank) -- eight full 2 GB uploads through one host's memory
    // were a third of the 8-GPU end-to-end step.  RID_FULL_UPLOAD=1: every rank uploads everything.
    bool sliced = ctx->world > 1 && !getenv("RID_FULL_UPLOAD");
    if (sliced) {
        const size_t cells_per = ((size_t)N + ctx->world - 1) / ctx->world;
        std::vector<size_t> off_i(ctx->world + 1), off_v(ctx->world + 1);
        for (int r = 0; r <= ctx->world; ++r) {
            const size_t c = std::min(cells_per * (size_t)r, (size_t)N);
            off_i[r] = (size_t)col_ptr[c] * sizeof(int);
            off_v[r] = (size_t)col_ptr[c] * sizeof(double);
        }
        const size_t lo = off_i[ctx->rank] / sizeof(int), hi = off_i[ctx->rank + 1] / sizeof(int);
        if (hi > lo) {
            cudaMemcpyAsync(d_i + lo, row_idx + lo, (hi - lo) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
            cudaMemcpyAsync(d_v + lo, val + lo, (hi - lo) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        }
        rc = rid_bcast_ranges(ctx, d_i, off_i.data());
        if (rc == RID_OK) rc = rid_bcast_ranges(ctx, d_v, off_v.data());
        if (rc == RID_ERR_UNSUPPORTED) { sliced = false; rc = RID_OK; } // (the same answer on every rank: one library)
        else if (rc != RID_OK) return done(rc);
    }
    if (!sliced) {
        cudaMemcpyAsync(d_i, row_idx, nnz * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(d_v, val, nnz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    }
    cudaMemcpyAsync(d_p, col_ptr, ((size_t)N + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_f, feat.data(), (size_t)n_rows_all * sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_c, counts, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    LAUNCH(ctx, k_csc_to_features, N, 256, 0, d_i, d_p, d_v, d_f, G, d_c, mincount, d_x);
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        rid_set_error("rid_dist_csc: building the feature matrix failed");
        return done(RID_ERR_CUDA);
    }
    rc = rid_dist(ctx, d_x, G, N, metric, 1, out, n_zero_var);
    return done(rc);
}
