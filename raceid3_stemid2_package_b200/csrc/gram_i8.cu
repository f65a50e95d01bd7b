// gram_i8.cu — the correlation contraction r = Z Z^T on the 5th-generation tensor cores, EXACTLY.
//
// tcgen05 has no FP64 kind, and TF32/BF16 kinds accumulate in FP32, which cannot hold |r| <= 1 to 1e-7 over G ~ 3000
// terms.  Integers can: every row of Z has unit norm, so it is a natural fixed-point operand.  Each row i is scaled by
// a power of two (max|z| * 2^e_i in [0.496, 0.992]), rounded to a 27-bit integer m_ik = rint(z_ik * 2^(e_i+27)), and split into
// four balanced radix-128 digits d_s in [-64, 63]  (an "Ozaki split"):  m = sum_s d_s 128^s.  Then
//      sum_k m_ik m_jk = sum_{s,t} 128^(s+t) * P_st,     P_st = sum_k d_s(i,k) d_t(j,k)
// and every P_st is an int8 x int8 -> int32 GEMM that `tcgen05.mma.kind::i8` computes without any rounding
// (G * 64^2 < 2^31).  The five diagonals s+t = 2..6 (13 digit products) are accumulated in five TMEM accumulators; the
// two lowest diagonals (s+t <= 1) are below 2^-54 * 257 * G * 64^2 and are covered by an a-posteriori bound computed
// from the actual digit norms (the host falls back to the FP64 DMMA kernel if the bound exceeds 8e-7).
//
// Kernel shape: 128 x 96 output tiles (5 accumulators x 96 columns = 480 of the 512 TMEM columns), K blocks of 128
// bytes, two 112 KB shared-memory stages filled by TMA (SWIZZLE_128B, K-major), one elected thread issuing the
// tcgen05.mma stream (7 instructions per K = 32 step: products (s,t),(s,t+1) share one N = 192 instruction),
// tcgen05.commit -> mbarrier to recycle stages and to hand the accumulators to the epilogue warps, which read TMEM with
// tcgen05.ld, rebuild the 64-bit integer dot product, turn it into the fixed-point distance and store the tile and its
// mirror image (D is symmetric; only tiles touching the upper triangle are computed).
// Kernels in this file: k_gram_i8_persist (the default: persistent CTAs over a tile list, prefetch across tiles, integer
// epilogue), k_gram_i8<1> (one tile per CTA, FP64 epilogue), k_gram_i8<2> (2 x 2 cluster, TMA multicast),
// k_gram_i8_2sm (CTA pair, tcgen05.mma.cta_group::2) -- all bit-identical (tests/test_gpu_parity.py).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

#define I8_BM 128
#define I8_BN 96
#define I8_BK 128
#define I8_STAGES 2
#define I8_A_PLANE (I8_BM * I8_BK)                 // 16384
#define I8_B_PLANE (I8_BN * I8_BK)                 // 12288
#define I8_STAGE_BYTES (4 * I8_A_PLANE + 4 * I8_B_PLANE) // 114688
#define I8_SMEM (I8_STAGES * I8_STAGE_BYTES + 1024)
#define I8_TMEM_COLS 512
#define I8_THREADS 256
#define I8_GROUP 16

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must end in a trap, not in a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    for (unsigned long long spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1ull << 26)) __trap();
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tma_load_2d_mc(void *smem_dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar,
                                               uint16_t cta_mask)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
                 " [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO=1 | SBO=1024>>4 |
// version=1 | layout=SWIZZLE_128B(2)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c=S32 (2) @4, a=INT8 (1) @7, b=INT8 (1) @10, K-major both,
// N>>3 @17, M>>4 @24
#define I8_IDESC_N(n) ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((n) >> 3) << 17) | ((uint32_t)(I8_BM >> 4) << 24))
#define I8_IDESC I8_IDESC_N(I8_BN)

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate,
                                        uint32_t idesc = I8_IDESC)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- CTA-pair (cta_group::2) variants -------------------------------------------------------------------
#define I8_IDESC_MN(m, n) ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((n) >> 3) << 17) | ((uint32_t)((m) >> 4) << 24))
#define I8_PEER_MASK 0xFEFFFFFFu // clears the CTA-pair peer bit of a shared::cluster address: "the even CTA's copy"

// TMA load issued by either CTA of a pair into its OWN shared memory, completing on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(void *smem_dst, const CUtensorMap *tm, int c0, int c1, uint32_t leader_bar)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(leader_bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void umma_i8_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t idesc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}

struct I8Params {
    int N, nkb;
    u32 *D;
    size_t ld;
    int plane_rows; // rows per digit plane in the operand matrix
    const unsigned char *zero_var;
    const double *rowscale; // 2^-(e_i+27)
    double inv_step;
    // row-block sharding: this rank owns rows [row0, row0+nloc); rank b owns [b*per, ...).  Every off-diagonal pair of
    // shards is computed by exactly one of its two owners, which stores the mirror image into the peer's shard (NVLink).
    int row0, nloc, world, rank, per;
    int tiles_loc, tiles_n;
    u32 *peer[16];
};

// Does rank `me` compute the elements (rows of me) x (columns of rank b)?  Antipodal pairs (world even) are split by the
// parity of the 384 x 384 super-block (384 = lcm of the 128-row and 96-column tiles), which both owners evaluate alike.
__device__ __host__ __forceinline__ bool i8_pair_mine(int me, int b, int world, int i0, int j0)
{
    if (b == me) return j0 + I8_BN > i0;
    const int d = (b - me + world) % world;
    if (2 * d != world) return 2 * d < world;
    return (((i0 / 384) + (j0 / 384)) & 1) == (me < b ? 0 : 1);
}

// ===== epilogue of one 128 x 96 tile, all 8 warps of the CTA =====
__device__ __forceinline__ void i8_epilogue(const I8Params &P, uint32_t tmem_base, uint64_t *tmem_full_bar, int warp, int lane,
                                            int i0, int j0, bool tile_valid)
{
    // ===== epilogue, all 8 warps: TMEM -> registers -> fixed-point distances -> global (tile + mirror) =====
    // warp w may touch TMEM lanes 32*(w%4)..+31; warps w and w+4 share a lane quarter and split the 96 columns
    mbar_wait(tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const int r = q * 32 + lane;     // tile row = TMEM lane
    const int gi = i0 + r;
    const bool row_ok = tile_valid && gi < P.row0 + P.nloc;
    const double rs_i = row_ok ? P.rowscale[gi] * 16384.0 : 0.0; // 128^2: the lowest kept diagonal is s+t = 2
    const bool zr = row_ok ? P.zero_var[gi] : false;
    const int cbeg = (warp >> 2) * (I8_BN / 2);
    for (int c0 = cbeg; c0 < cbeg + I8_BN / 2; c0 += 16) {
        int a2[16], a3[16], a4[16], a5[16], a6[16];
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        tmem_ld16(ta + 0 * I8_BN, a2);
        tmem_ld16(ta + 1 * I8_BN, a3);
        tmem_ld16(ta + 2 * I8_BN, a4);
        tmem_ld16(ta + 3 * I8_BN, a5);
        tmem_ld16(ta + 4 * I8_BN, a6);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        u32 qv[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const int gj = j0 + c0 + c;
            u32 v = 0u;
            if (row_ok && gj < P.N) {
                if (gj == gi) {
                    v = zr ? RID_Q_NA : 0u;
                } else if (zr || P.zero_var[gj]) {
                    v = RID_Q_NA;
                } else {
                    long long acc = (long long)a6[c];
                    acc = acc * 128 + a5[c];
                    acc = acc * 128 + a4[c];
                    acc = acc * 128 + a3[c];
                    acc = acc * 128 + a2[c];
                    double rr = (double)acc * rs_i * P.rowscale[gj];
                    rr = fmin(1.0, fmax(-1.0, rr));
                    v = (u32)__double2ull_rn((1.0 - rr) * P.inv_step);
                }
            }
            qv[c] = v;
        }
        // direct store: 16 consecutive columns of row gi (64-byte aligned: j0 and c0 are multiples of 16)
        if (row_ok) {
            u32 *dst = P.D + (size_t)(gi - P.row0) * P.ld + j0 + c0;
            if ((size_t)(j0 + c0 + 16) <= P.ld) {
#pragma unroll
                for (int c = 0; c < 16; c += 4)
                    *reinterpret_cast<uint4 *>(dst + c) = make_uint4(qv[c], qv[c + 1], qv[c + 2], qv[c + 3]);
            } else {
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    if ((size_t)(j0 + c0 + c) < P.ld) dst[c] = qv[c];
            }
        }
        // mirror store: column gi of rows j0+c0..+15; a warp's 32 lanes write 128 contiguous bytes
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const int gj = j0 + c0 + c;
            if (tile_valid && gj < P.N && (size_t)gi < P.ld) {
                const int ow = min(gj / P.per, P.world - 1);
                P.peer[ow][(size_t)(gj - ow * P.per) * P.ld + gi] = row_ok ? qv[c] : 0u;
            }
        }
    }
}

// CL = 1: one CTA per tile.  CL = 2: 2 x 2 thread-block clusters; the two CTAs of a cluster row share their A panel and
// the two CTAs of a cluster column their B panel: each loads HALF of either panel and TMA-multicasts it to its partner,
// which halves the operand traffic out of L2 (the kernel's bottleneck).  A stage may be refilled only after this CTA and
// both partners have consumed it, so every MMA commit arrives on three CTAs' "empty" barriers.
template <int CL>
__global__ void __launch_bounds__(I8_THREADS, 1)
k_gram_i8(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, I8Params P)
{
    // Tile order: consecutive clusters walk I8_GROUP tile-rows for one tile-column before moving to the next column, so
    // the ~148 CTAs in flight share 16 A panels and ~9 B panels out of L2 (the plain row-major order re-reads every B
    // panel from HBM for every tile-row: 479 GB of DRAM reads per launch, L2 hit rate 53 %).
    int tile_m, tile_n;
    const int cx = CL == 2 ? (int)(blockIdx.x & 1) : 0, cy = CL == 2 ? (int)(blockIdx.y & 1) : 0;
    {
        const int gx = (int)gridDim.x / CL, gy = (int)gridDim.y / CL, grp_rows = I8_GROUP / CL;
        const long long L = (long long)(blockIdx.y / CL) * gx + (blockIdx.x / CL);
        const long long per_group = (long long)grp_rows * gx;
        const int grp = (int)(L / per_group);
        const int rows_in_grp = min(grp_rows, gy - grp * grp_rows);
        const long long r = L - (long long)grp * per_group;
        tile_m = (grp * grp_rows + (int)(r % rows_in_grp)) * CL + cy;
        tile_n = (int)(r / rows_in_grp) * CL + cx;
    }
    const int i0 = P.row0 + tile_m * I8_BM, j0 = tile_n * I8_BN; // global row / column of the tile origin
    const bool tile_valid = tile_m < P.tiles_loc && tile_n < P.tiles_n;
    {
        // a column tile may straddle two shards: compute it if either owner's pair is ours; a cluster runs if any of its
        // tiles is needed (the decision must be the same in all of its CTAs)
        bool need = false;
#pragma unroll
        for (int dy = 0; dy < CL; ++dy)
#pragma unroll
            for (int dx = 0; dx < CL; ++dx) {
                const int tm = tile_m - cy + dy, tn = tile_n - cx + dx;
                if (tm >= P.tiles_loc || tn >= P.tiles_n) continue;
                const int ii = P.row0 + tm * I8_BM, jj = tn * I8_BN;
                const int b1 = min(jj / P.per, P.world - 1), b2 = min(min(jj + I8_BN - 1, P.N - 1) / P.per, P.world - 1);
                need = need || i8_pair_mine(P.rank, b1, P.world, ii, jj) || i8_pair_mine(P.rank, b2, P.world, ii, jj);
            }
        if (!need) return;
    }
    const uint16_t mask_a = CL == 2 ? (uint16_t)(0x3u << (2 * cy)) : (uint16_t)1;          // the CTAs sharing my A panel
    const uint16_t mask_b = CL == 2 ? (uint16_t)((1u << cx) | (1u << (cx + 2))) : (uint16_t)1; // ... my B panel

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[I8_STAGES], empty_bar[I8_STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < I8_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CL == 2 ? 3 : 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(I8_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL == 2) cluster_sync_all(); // every CTA's barriers exist before anybody multicasts into them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < P.nkb; ++kb) {
                const int st = kb % I8_STAGES;
                const uint32_t ph = (kb / I8_STAGES) & 1;
                mbar_wait(&empty_bar[st], ph ^ 1); // passes immediately the first time round
                uint8_t *sa = smem + (size_t)st * I8_STAGE_BYTES, *sb = sa + 4 * I8_A_PLANE;
                mbar_expect_tx(&full_bar[st], I8_STAGE_BYTES); // bytes landing in THIS CTA (own halves + partners')
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    if (CL == 2) {
                        // tmA / tmB are the half-panel boxes here: rows [64 cx, +64) of A, rows [48 cy, +48) of B
                        tma_load_2d_mc(sa + s * I8_A_PLANE + cx * (I8_A_PLANE / 2), &tmA, kb * I8_BK,
                                       s * P.plane_rows + i0 + cx * (I8_BM / 2), &full_bar[st], mask_a);
                        tma_load_2d_mc(sb + s * I8_B_PLANE + cy * (I8_B_PLANE / 2), &tmB, kb * I8_BK,
                                       s * P.plane_rows + j0 + cy * (I8_BN / 2), &full_bar[st], mask_b);
                    } else {
                        tma_load_2d(sa + s * I8_A_PLANE, &tmA, kb * I8_BK, s * P.plane_rows + i0, &full_bar[st]);
                        tma_load_2d(sb + s * I8_B_PLANE, &tmB, kb * I8_BK, s * P.plane_rows + j0, &full_bar[st]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            for (int kb = 0; kb < P.nkb; ++kb) {
                const int st = kb % I8_STAGES;
                const uint32_t ph = (kb / I8_STAGES) & 1;
                mbar_wait(&full_bar[st], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + (size_t)st * I8_STAGE_BYTES), sb = sa + 4 * I8_A_PLANE;
                // Digit products (s,t) and (s,t+1) share their A panel, their B panels are adjacent in shared memory and
                // their accumulators (diagonals s+t, s+t+1) adjacent in TMEM: ONE N = 192 instruction does both and reads
                // the A panel once (7 instead of 13 instructions per K step: 26 % fewer operand bytes out of shared
                // memory, which is what bounds this kernel).  Only the very first K step of a tile uses single products,
                // because there the "overwrite" flag differs between the two diagonals of a pair.
#pragma unroll
                for (int kk = 0; kk < I8_BK / 32; ++kk) {
                    // advancing K inside the 128-byte swizzle atom: +32 bytes = +2 in the (addr >> 4) field
                    const uint64_t koff = (uint64_t)(2 * kk);
                    if (kb == 0 && kk == 0) {
                        uint32_t started = 0;
#pragma unroll
                        for (int s = 0; s < 4; ++s)
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                if (s + t < 2) continue;
                                const int d = s + t - 2;
                                umma_i8(tmem_base + (uint32_t)(d * I8_BN), umma_desc_sw128(sa + s * I8_A_PLANE) + koff,
                                        umma_desc_sw128(sb + t * I8_B_PLANE) + koff, (started >> d) & 1u);
                                started |= 1u << d;
                            }
                    } else {
#pragma unroll
                        for (int s = 0; s < 4; ++s) {
                            const uint64_t ad = umma_desc_sw128(sa + s * I8_A_PLANE) + koff;
                            const int t_lo = s >= 2 ? 0 : 2 - s; // first t with s + t >= 2
                            int t = t_lo;
                            for (; t + 1 < 4; t += 2) // pairs (t, t+1): N = 192
                                umma_i8(tmem_base + (uint32_t)((s + t - 2) * I8_BN), ad,
                                        umma_desc_sw128(sb + t * I8_B_PLANE) + koff, 1u, I8_IDESC_N(2 * I8_BN));
                            if (t < 4)                // one product left over: N = 96
                                umma_i8(tmem_base + (uint32_t)((s + t - 2) * I8_BN), ad,
                                        umma_desc_sw128(sb + t * I8_B_PLANE) + koff, 1u);
                        }
                    }
                }
                // frees this stage once the MMAs above have read it (in all CTAs that fill it, for a cluster)
                if (CL == 2) umma_commit_mc(&empty_bar[st], (uint16_t)(mask_a | mask_b));
                else umma_commit(&empty_bar[st]);
            }
            umma_commit(&tmem_full_bar);     // all accumulators final
        }
    }
    __syncwarp();
    {
        i8_epilogue(P, tmem_base, &tmem_full_bar, warp, lane, i0, j0, tile_valid);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (CL == 2) cluster_sync_all(); // partners may still arrive on this CTA's barriers / multicast into its stages
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(I8_TMEM_COLS) : "memory");
    }
}

// CTA-pair kernel: two CTAs on the two SMs of a TPC own a 256 x 96 output tile (128 rows each) and run ONE stream of
// tcgen05.mma.cta_group::2 (M = 256) issued by the even CTA.  In pair mode the tensor cores of both SMs read the B
// operand half from either SM's shared memory, so an N = 192 instruction (two digit products) costs each SM 4 KB of A
// + 3 KB of B per K = 32 step instead of 4 + 6 KB: 118 B/clk of shared-memory traffic (operand reads + TMA writes)
// instead of 152 B/clk against the 128 B/clk the SM sustains -- the single-CTA kernel's bound.
// What each CTA stages per K block: its own 128 rows of the four A digit planes, and B "slots" whose two halves live
// in the two CTAs:  P01 = plane 0 | plane 1,  P12 = plane 1 | plane 2,  P23 = plane 2 | plane 3 (even CTA | odd CTA),
// H3 = rows 0..47 | rows 48..95 of plane 3.  The 13 digit products are the 7 instructions
//   (A0,P23) (A1,P12) (A1,H3) (A3,P23) (A2,P01) (A2,P23) (A3,P01)
// in an order in which both diagonals an N = 192 instruction touches are either both fresh or both started, so the very
// first K step needs no special case.
#define I8P_B_BYTES (3 * I8_B_PLANE + I8_B_PLANE / 2)   // 43008
#define I8P_STAGE_BYTES (4 * I8_A_PLANE + I8P_B_BYTES)  // 108544
#define I8P_SMEM (I8_STAGES * I8P_STAGE_BYTES + 1024)
__global__ void __launch_bounds__(I8_THREADS, 1)
k_gram_i8_2sm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB48, I8Params P)
{
    const int cta = (int)(blockIdx.x & 1); // rank in the pair (cluster dims (2,1,1))
    int tile_m, tile_n;
    {
        const int gx = P.tiles_n, gy = (P.tiles_loc + 1) / 2, grp_rows = I8_GROUP / 2;
        const long long L = (long long)(blockIdx.x >> 1);
        const long long per_group = (long long)grp_rows * gx;
        const int grp = (int)(L / per_group);
        const int rows_in_grp = min(grp_rows, gy - grp * grp_rows);
        const long long r = L - (long long)grp * per_group;
        tile_m = (grp * grp_rows + (int)(r % rows_in_grp)) * 2 + cta;
        tile_n = (int)(r / rows_in_grp);
    }
    const int i0 = P.row0 + tile_m * I8_BM, j0 = tile_n * I8_BN;
    const bool tile_valid = tile_m < P.tiles_loc && tile_n < P.tiles_n;
    {
        bool need = false; // the pair runs if either of its tiles is needed (same decision in both CTAs)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const int tm = tile_m - cta + dy;
            if (tm >= P.tiles_loc) continue;
            const int ii = P.row0 + tm * I8_BM;
            const int b1 = min(j0 / P.per, P.world - 1), b2 = min(min(j0 + I8_BN - 1, P.N - 1) / P.per, P.world - 1);
            need = need || i8_pair_mine(P.rank, b1, P.world, ii, j0) || i8_pair_mine(P.rank, b2, P.world, ii, j0);
        }
        if (!need) return;
    }
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[I8_STAGES], empty_bar[I8_STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB48) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < I8_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) { // the same warp of BOTH CTAs allocates for the pair
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(I8_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all(); // both CTAs' barriers exist before the partner's TMA / commits arrive on them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own A rows, own halves of the B slots; bytes are counted on the leader's barrier
        if (lane == 0) {
            for (int kb = 0; kb < P.nkb; ++kb) {
                const int st = kb % I8_STAGES;
                const uint32_t ph = (kb / I8_STAGES) & 1;
                mbar_wait(&empty_bar[st], ph ^ 1);
                uint8_t *sa = smem + (size_t)st * I8P_STAGE_BYTES, *sb = sa + 4 * I8_A_PLANE;
                if (cta == 0) mbar_expect_tx(&full_bar[st], 2 * I8P_STAGE_BYTES);
                const uint32_t lbar = smem_u32(&full_bar[st]) & I8_PEER_MASK;
#pragma unroll
                for (int s = 0; s < 4; ++s)
                    tma_load_2d_2sm(sa + s * I8_A_PLANE, &tmA, kb * I8_BK, s * P.plane_rows + i0, lbar);
#pragma unroll
                for (int slot = 0; slot < 3; ++slot) { // P01, P12, P23: plane (slot + cta), all 96 rows as two 48-row boxes
                    const int row = (slot + cta) * P.plane_rows + j0;
                    tma_load_2d_2sm(sb + slot * I8_B_PLANE, &tmB48, kb * I8_BK, row, lbar);
                    tma_load_2d_2sm(sb + slot * I8_B_PLANE + I8_B_PLANE / 2, &tmB48, kb * I8_BK, row + I8_BN / 2, lbar);
                }
                tma_load_2d_2sm(sb + 3 * I8_B_PLANE, &tmB48, kb * I8_BK, 3 * P.plane_rows + j0 + cta * (I8_BN / 2), lbar); // H3
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the even CTA drives the tensor cores of both SMs =====
        if (lane == 0 && cta == 0) {
            constexpr uint32_t ID192 = I8_IDESC_MN(256, 2 * I8_BN), ID96 = I8_IDESC_MN(256, I8_BN);
            for (int kb = 0; kb < P.nkb; ++kb) {
                const int st = kb % I8_STAGES;
                const uint32_t ph = (kb / I8_STAGES) & 1;
                mbar_wait(&full_bar[st], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + (size_t)st * I8P_STAGE_BYTES), sb = sa + 4 * I8_A_PLANE;
                const uint64_t A0 = umma_desc_sw128(sa), A1 = umma_desc_sw128(sa + I8_A_PLANE),
                               A2 = umma_desc_sw128(sa + 2 * I8_A_PLANE), A3 = umma_desc_sw128(sa + 3 * I8_A_PLANE);
                const uint64_t P01 = umma_desc_sw128(sb), P12 = umma_desc_sw128(sb + I8_B_PLANE),
                               P23 = umma_desc_sw128(sb + 2 * I8_B_PLANE), H3 = umma_desc_sw128(sb + 3 * I8_B_PLANE);
                const uint32_t d0 = tmem_base, d1 = tmem_base + I8_BN, d2 = tmem_base + 2 * I8_BN, d3 = tmem_base + 3 * I8_BN;
#pragma unroll
                for (int kk = 0; kk < I8_BK / 32; ++kk) {
                    const uint64_t koff = (uint64_t)(2 * kk); // +32 bytes inside the 128-byte swizzle atom
                    const uint32_t acc = (kb == 0 && kk == 0) ? 0u : 1u;
                    umma_i8_2sm(d0, A0 + koff, P23 + koff, acc, ID192); // (0,2) (0,3) -> diagonals 0, 1
                    umma_i8_2sm(d0, A1 + koff, P12 + koff, 1u, ID192);  // (1,1) (1,2) -> 0, 1
                    umma_i8_2sm(d2, A1 + koff, H3 + koff, acc, ID96);   // (1,3)       -> 2
                    umma_i8_2sm(d3, A3 + koff, P23 + koff, acc, ID192); // (3,2) (3,3) -> 3, 4
                    umma_i8_2sm(d0, A2 + koff, P01 + koff, 1u, ID192);  // (2,0) (2,1) -> 0, 1
                    umma_i8_2sm(d2, A2 + koff, P23 + koff, 1u, ID192);  // (2,2) (2,3) -> 2, 3
                    umma_i8_2sm(d1, A3 + koff, P01 + koff, 1u, ID192);  // (3,0) (3,1) -> 1, 2
                }
                umma_commit_2sm(&empty_bar[st]); // both CTAs may refill this stage once the MMAs above have read it
            }
            umma_commit_2sm(&tmem_full_bar);     // accumulators final, in both CTAs
        }
    }
    __syncwarp();
    i8_epilogue(P, tmem_base, &tmem_full_bar, warp, lane, i0, j0, tile_valid);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all(); // the partner may still read this CTA's operand halves / arrive on its barriers
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(I8_TMEM_COLS) : "memory");
    }
}

// ----------------------------------------------------------------------------------------------------------
// Persistent kernel (the default): one CTA per SM walks the tile list with stride gridDim.x.  TMEM holds exactly one
// tile's five accumulators, so the epilogue cannot overlap the next tile's MMAs; what the single-shot kernels above
// lose on top of that is removed here:
//  * the TMA producer prefetches the next tile's first two K blocks BEFORE it joins the epilogue (no pipeline refill,
//    no CTA launch / TMEM allocation / barrier setup per tile),
//  * the six otherwise idle warps prepare the tile's column tables (scale exponent, NA flag, mirror pointer) in shared
//    memory while the MMAs run, so the epilogue does no global loads and no divisions per element,
//  * the epilogue is integer only.  Row scales are exact powers of two (2^-(e_i+27)) and the step of D is 2^-27, so
//        q = rint((1 - r) 2^27),  r = acc 2^-(e_i+e_j+54)   ==>   q = 2^27 - RNE(acc >> (e_i+e_j+27)),
//    an exact round-half-even shift of the 64-bit dot product: no int64 -> double conversion, no FP64 chain.
// ----------------------------------------------------------------------------------------------------------
// (tile_m, tile_n) of the L-th tile in launch order: groups of I8_GROUP tile rows, column-major inside a group, so that
// the CTAs in flight share a handful of A and B panels out of L2.  Host and device use the same function.
__device__ __host__ __forceinline__ bool i8_tile_at(const I8Params &P, long long L, int &tile_m, int &tile_n)
{
    const int gx = P.tiles_n, gy = P.tiles_loc;
    const long long per_group = (long long)I8_GROUP * gx;
    const int grp = (int)(L / per_group);
    const int rows_in_grp = (I8_GROUP < gy - grp * I8_GROUP) ? I8_GROUP : gy - grp * I8_GROUP;
    const long long r = L - (long long)grp * per_group;
    tile_m = grp * I8_GROUP + (int)(r % rows_in_grp);
    tile_n = (int)(r / rows_in_grp);
    const int ii = P.row0 + tile_m * I8_BM, jj = tile_n * I8_BN;
    const int je = (jj + I8_BN - 1 < P.N - 1) ? jj + I8_BN - 1 : P.N - 1;
    const int b1 = (jj / P.per < P.world - 1) ? jj / P.per : P.world - 1, b2 = (je / P.per < P.world - 1) ? je / P.per : P.world - 1;
    return i8_pair_mine(P.rank, b1, P.world, ii, jj) || i8_pair_mine(P.rank, b2, P.world, ii, jj);
}

__global__ void __launch_bounds__(I8_THREADS, 1)
k_gram_i8_persist(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, I8Params P,
                  const int2 *__restrict__ tiles, long long ntiles)
{
    // `tiles` lists only the tiles this rank computes, in launch order: CTA c takes entries c, c + gridDim.x, ...  All
    // tiles cost the same, so the CTAs stay within one "wave" of gridDim.x consecutive entries of each other and keep
    // the L2 locality of the order (striding over ALL tiles and skipping the unneeded ones lets CTAs drift apart:
    // 321 GB of DRAM reads per launch instead of ~35 GB).
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[I8_STAGES], empty_bar[I8_STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ u32 *col_ptr[I8_BN];   // mirror destination of column c: row (j0+c) of its owner's shard, or null
    __shared__ int col_e[I8_BN];      // e_j + 27 of column j0+c (0: no such column)
    __shared__ unsigned char col_na[I8_BN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long total = ntiles;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < I8_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(I8_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    int tile_m = 0, tile_n = 0;
    long long L = blockIdx.x;
    if (L < total) { const int2 t0 = tiles[L]; tile_m = t0.x; tile_n = t0.y; }
    uint32_t it = 0;      // K blocks produced (warp 0) / consumed (warp 1) so far, over all tiles
    uint32_t tcount = 0;  // tiles done by this CTA
    int prefetched = 0;   // K blocks of the current tile already requested (producer)

    auto produce = [&](int kb, int i0, int j0) {
        const int st = it % I8_STAGES;
        const uint32_t ph = (it / I8_STAGES) & 1;
        mbar_wait(&empty_bar[st], ph ^ 1);
        uint8_t *sa = smem + (size_t)st * I8_STAGE_BYTES, *sb = sa + 4 * I8_A_PLANE;
        mbar_expect_tx(&full_bar[st], I8_STAGE_BYTES);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            tma_load_2d(sa + s * I8_A_PLANE, &tmA, kb * I8_BK, s * P.plane_rows + i0, &full_bar[st]);
            tma_load_2d(sb + s * I8_B_PLANE, &tmB, kb * I8_BK, s * P.plane_rows + j0, &full_bar[st]);
        }
        ++it;
    };

    while (L < total) {
        int next_m = 0, next_n = 0;
        const long long Ln = L + gridDim.x;
        if (Ln < total) { const int2 t1 = tiles[Ln]; next_m = t1.x; next_n = t1.y; }
        const int i0 = P.row0 + tile_m * I8_BM, j0 = tile_n * I8_BN;

        if (warp == 0) {
            // ===== TMA producer: the rest of this tile, then the head of the next one =====
            if (lane == 0) {
                for (int kb = prefetched; kb < P.nkb; ++kb) produce(kb, i0, j0);
                prefetched = 0;
                if (Ln < total) {
                    const int ni0 = P.row0 + next_m * I8_BM, nj0 = next_n * I8_BN;
                    for (; prefetched < min(I8_STAGES, P.nkb); ++prefetched) produce(prefetched, ni0, nj0);
                }
            }
        } else if (warp == 1) {
            // ===== MMA issuer =====
            if (lane == 0) {
                for (int kb = 0; kb < P.nkb; ++kb) {
                    const int st = it % I8_STAGES;
                    const uint32_t ph = (it / I8_STAGES) & 1;
                    mbar_wait(&full_bar[st], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(smem + (size_t)st * I8_STAGE_BYTES), sb = sa + 4 * I8_A_PLANE;
                    const uint64_t A0 = umma_desc_sw128(sa), A1 = umma_desc_sw128(sa + I8_A_PLANE),
                                   A2 = umma_desc_sw128(sa + 2 * I8_A_PLANE), A3 = umma_desc_sw128(sa + 3 * I8_A_PLANE);
                    // B planes t and t+1 are adjacent: an N = 192 instruction multiplies one A plane with both
                    const uint64_t B0 = umma_desc_sw128(sb), B1 = umma_desc_sw128(sb + I8_B_PLANE),
                                   B2 = umma_desc_sw128(sb + 2 * I8_B_PLANE), B3 = umma_desc_sw128(sb + 3 * I8_B_PLANE);
                    const uint32_t d0 = tmem_base, d1 = tmem_base + I8_BN, d2 = tmem_base + 2 * I8_BN, d3 = tmem_base + 3 * I8_BN;
                    constexpr uint32_t ID192 = I8_IDESC_N(2 * I8_BN), ID96 = I8_IDESC_N(I8_BN);
#pragma unroll
                    for (int kk = 0; kk < I8_BK / 32; ++kk) {
                        const uint64_t koff = (uint64_t)(2 * kk); // +32 bytes inside the 128-byte swizzle atom
                        // order: both diagonals an N = 192 instruction touches are either both fresh or both started
                        const uint32_t acc = (kb == 0 && kk == 0) ? 0u : 1u;
                        umma_i8(d0, A0 + koff, B2 + koff, acc, ID192); // (0,2) (0,3) -> diagonals 0, 1
                        umma_i8(d0, A1 + koff, B1 + koff, 1u, ID192);  // (1,1) (1,2) -> 0, 1
                        umma_i8(d2, A1 + koff, B3 + koff, acc, ID96);  // (1,3)       -> 2
                        umma_i8(d3, A3 + koff, B2 + koff, acc, ID192); // (3,2) (3,3) -> 3, 4
                        umma_i8(d0, A2 + koff, B0 + koff, 1u, ID192);  // (2,0) (2,1) -> 0, 1
                        umma_i8(d2, A2 + koff, B2 + koff, 1u, ID192);  // (2,2) (2,3) -> 2, 3
                        umma_i8(d1, A3 + koff, B0 + koff, 1u, ID192);  // (3,0) (3,1) -> 1, 2
                    }
                    umma_commit(&empty_bar[st]);
                    ++it;
                }
                umma_commit(&tmem_full_bar);
            }
        } else {
            // ===== column tables of this tile (warps 2..7, while the MMAs run) =====
            for (int c = (int)threadIdx.x - 64; c < I8_BN; c += I8_THREADS - 64) {
                const int gj = j0 + c;
                u32 *ptr = nullptr;
                int e = 0;
                unsigned char na = 0;
                if (gj < P.N) {
                    const int ow = min(gj / P.per, P.world - 1);
                    ptr = P.peer[ow] + (size_t)(gj - ow * P.per) * P.ld;
                    e = 1023 - (int)((__double2hiint(P.rowscale[gj]) >> 20) & 0x7ff);
                    na = P.zero_var[gj];
                }
                col_ptr[c] = ptr; col_e[c] = e; col_na[c] = na;
            }
        }
        __syncthreads(); // tables written; every role is through with this tile's mainloop issue

        // ===== epilogue, all 8 warps: TMEM -> registers -> fixed-point distances -> global (tile + mirror) =====
        mbar_wait(&tmem_full_bar, tcount & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            const int q = warp & 3;
            const int gi = i0 + q * 32 + lane; // tile row = TMEM lane
            const bool row_ok = gi < P.row0 + P.nloc;
            const int e_i = row_ok ? 1023 - (int)((__double2hiint(P.rowscale[gi]) >> 20) & 0x7ff) : 0;
            const bool zr = row_ok ? P.zero_var[gi] : false;
            u32 *drow = P.D + (size_t)(gi - P.row0) * P.ld + j0;
            const bool mirror_ok = (size_t)gi < P.ld;
            const int cbeg = (warp >> 2) * (I8_BN / 2);
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + I8_BN / 2; c0 += 16) {
                int a2[16], a3[16], a4[16], a5[16], a6[16];
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
                tmem_ld16(ta + 0 * I8_BN, a2);
                tmem_ld16(ta + 1 * I8_BN, a3);
                tmem_ld16(ta + 2 * I8_BN, a4);
                tmem_ld16(ta + 3 * I8_BN, a5);
                tmem_ld16(ta + 4 * I8_BN, a6);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                u32 qv[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int gj = j0 + c0 + c;
                    const int e_j = col_e[c0 + c];
                    // sum_d a_d 128^(d-2):  (a6 128 + a5) 2^21 + (a4 128 + a3) 2^7 + a2
                    const long long hi = (long long)a6[c] * 128 + a5[c], mid = (long long)a4[c] * 128 + a3[c];
                    const long long acc = (hi << 21) + (mid << 7) + a2[c];
                    // acc carries 14 more fraction bits than sum m_i m_j 2^-0 (the lowest kept diagonal is 128^2)
                    const int sh = e_i + e_j - 27 - 14;
                    long long t;
                    if (sh <= 0) t = acc << min(-sh, 20);             // (only for rows of huge dynamic range; exact)
                    else if (sh > 62) t = 0;
                    else t = (acc + ((1ll << (sh - 1)) - 1) + ((acc >> sh) & 1)) >> sh; // round half to even
                    const int ti = (int)max(min(t, (long long)(1 << 27)), -(long long)(1 << 27)); // clamp r to [-1, 1]
                    u32 v = (u32)((1 << 27) - ti);
                    if (gj == gi) v = 0u;
                    if (zr || col_na[c0 + c]) v = RID_Q_NA;
                    if (!row_ok || e_j == 0) v = 0u;
                    qv[c] = v;
                }
                // direct store: 16 consecutive columns of row gi (64-byte aligned: j0 and c0 are multiples of 16)
                if (row_ok) {
                    u32 *dst = drow + c0;
                    if ((size_t)(j0 + c0 + 16) <= P.ld) {
#pragma unroll
                        for (int c = 0; c < 16; c += 4)
                            *reinterpret_cast<uint4 *>(dst + c) = make_uint4(qv[c], qv[c + 1], qv[c + 2], qv[c + 3]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            if ((size_t)(j0 + c0 + c) < P.ld) dst[c] = qv[c];
                    }
                }
                // mirror store: column gi of rows j0+c0..+15; a warp's 32 lanes write 128 contiguous bytes
                if (mirror_ok) {
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        u32 *mp = col_ptr[c0 + c];
                        if (mp) mp[gi] = qv[c];
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads(); // TMEM drained, tables free
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        L = Ln; tile_m = next_m; tile_n = next_n;
        ++tcount;
    }
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(I8_TMEM_COLS) : "memory");
    }
}

// ----------------------------------------------------------------------------------------------------------
// Z (FP64, unit rows) -> four int8 digit planes + per-row scale; statistics for the a-posteriori error bound
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_quant_i8(const double *__restrict__ Z, int Gp_src, int G, int N, int8_t *__restrict__ planes, size_t plane_rows,
           int Kp, double *__restrict__ rowscale, double *__restrict__ stats /* [4]: max 2^-e*|d0|, 2^-e*|d1|, 2^-e, |z|_1 */)
{
    __shared__ double sh[256];
    const int i = blockIdx.x;
    const double *z = Z + (size_t)i * Gp_src;
    double amax = 0.0, l1 = 0.0;
    for (int k = threadIdx.x; k < G; k += blockDim.x) { double a = fabs(z[k]); amax = fmax(amax, a); l1 += a; }
    sh[threadIdx.x] = amax;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]); __syncthreads(); }
    amax = sh[0];
    __syncthreads();
    sh[threadIdx.x] = l1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    l1 = sh[0];
    __syncthreads();
    int x = 0;
    double f = 0.0;
    if (amax > 0.0) f = frexp(amax, &x); // amax = f * 2^x, f in [0.5, 1)
    // amax * 2^e in [0.496, 0.992]: |m| <= 0.992 * 2^27 fits four balanced radix-128 digits (max 63*(128^4-1)/127)
    const int e = f > 0.992 ? -x - 1 : -x;
    const double up = ldexp(1.0, e + 27);
    double n0 = 0.0, n1 = 0.0;
    int8_t *p0 = planes + (size_t)i * Kp;
    const size_t ps = plane_rows * (size_t)Kp;
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
        long long m = k < G ? llrint(z[k] * up) : 0;
        int d[4];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            int dd = (int)(((m + 64) % 128 + 128) % 128) - 64; // balanced digit in [-64, 63]
            d[s] = dd;
            m = (m - dd) / 128;
        }
        d[3] = (int)m;
        p0[k] = (int8_t)d[0];
        p0[ps + k] = (int8_t)d[1];
        p0[2 * ps + k] = (int8_t)d[2];
        p0[3 * ps + k] = (int8_t)d[3];
        n0 += (double)d[0] * d[0];
        n1 += (double)d[1] * d[1];
    }
    sh[threadIdx.x] = n0;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    n0 = sh[0];
    __syncthreads();
    sh[threadIdx.x] = n1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    n1 = sh[0];
    if (threadIdx.x == 0) {
        const double inv2e = ldexp(1.0, -e);
        rowscale[i] = ldexp(1.0, -(e + 27));
        // positive doubles order like their bit patterns
        atomicMax((unsigned long long *)&stats[0], (unsigned long long)__double_as_longlong(sqrt(n0) * inv2e));
        atomicMax((unsigned long long *)&stats[1], (unsigned long long)__double_as_longlong(sqrt(n1) * inv2e));
        atomicMax((unsigned long long *)&stats[2], (unsigned long long)__double_as_longlong(inv2e));
        atomicMax((unsigned long long *)&stats[3], (unsigned long long)__double_as_longlong(l1));
    }
}

// ----------------------------------------------------------------------------------------------------------
// host
// ----------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// Runs the exact tensor-core contraction for a single-GPU correlation matrix.  Returns RID_OK and *used = 1 when the
// matrix was produced, *used = 0 when the caller must use the FP64 DMMA kernel instead (bound not met / unsupported).
int rid_gram_i8(rid_ctx *ctx, const double *Z, int Gp_src, int G, int N, const unsigned char *zero_var, rid_dmat *dm,
                int *used, double *bound_out)
{
    *used = 0;
    // multi-rank: the mirror tiles go straight into the peers' shards, so every rank needs every shard mapped
    void *peers[16];
    for (int r = 0; r < 16; ++r) peers[r] = nullptr;
    peers[0] = dm->D;
    if (ctx->world > 1) {
        if (!dm->peers_ok) return RID_OK; // (mapped once per matrix by dmat_alloc; the same answer on every rank)
        for (int r = 0; r < ctx->world; ++r) peers[r] = dm->peer_shard[r];
    }
    PFN_encodeTiled encode = get_encode();
    if (!encode) return RID_OK;
    const int Kp = (int)round_up((size_t)G, I8_BK);
    const int nloc = dm->row1 - dm->row0;
    const int tiles_m = (N + I8_BM - 1) / I8_BM, tiles_n = (N + I8_BN - 1) / I8_BN, tiles_loc = (nloc + I8_BM - 1) / I8_BM;
    const size_t plane_rows = std::max((size_t)tiles_m * I8_BM, (size_t)tiles_n * I8_BN);
    const size_t planes_bytes = 4 * plane_rows * (size_t)Kp;
    int8_t *planes = nullptr;
    double *rowscale = nullptr, *stats = nullptr;
    if (rid_pool_alloc(ctx, (void **)&planes, planes_bytes) != cudaSuccess) { cudaGetLastError(); return RID_OK; }
    char *aux = nullptr; // stats[32] | rowscale[plane_rows], one pooled buffer
    const size_t aux_bytes = std::max<size_t>((plane_rows + 32) * sizeof(double), (size_t)256 << 10);
    auto cleanup = [&]() { rid_pool_free(ctx, planes, planes_bytes); rid_pool_free(ctx, aux, aux_bytes); };
    if (rid_pool_alloc(ctx, (void **)&aux, aux_bytes) != cudaSuccess) { cudaGetLastError(); aux = nullptr; cleanup(); return RID_OK; }
    stats = (double *)aux; rowscale = stats + 32;
    cudaMemsetAsync(planes, 0, planes_bytes, ctx->stream);
    cudaMemsetAsync(rowscale, 0, plane_rows * sizeof(double), ctx->stream);
    cudaMemsetAsync(stats, 0, 4 * sizeof(double), ctx->stream);
    k_quant_i8<<<N, 256, 0, ctx->stream>>>(Z, Gp_src, G, N, planes, plane_rows, Kp, rowscale, stats);
    ctx->launches++;
    double h[4];
    if (cudaMemcpyAsync(h, stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        rid_set_error("rid_gram_i8: quantisation failed: %s", cudaGetErrorString(cudaGetLastError()));
        cleanup();
        return RID_ERR_CUDA;
    }
    // a-posteriori bound on |r - r_exact|:  dropped diagonals (s+t <= 1) by Cauchy-Schwarz on the digit norms, plus the
    // rounding of z to 27 bits:  2^-54 (256 U0 U1 + U0^2)  +  2^-27 V W1     (U_s = max 2^-e |d_s|, V = max 2^-e, W1 = max |z|_1)
    const double bound = ldexp(256.0 * h[0] * h[1] + h[0] * h[0], -54) + ldexp(h[2] * h[3], -27);
    if (bound_out) *bound_out = bound;
    // (every rank quantises all rows, so every rank sees the same statistics and takes the same branch)
    if (!(bound <= 8e-7)) { cleanup(); return RID_OK; } // + 2^-28 quantisation of D stays under the 1e-6 bar

    CUtensorMap tmA, tmB;
    cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)(4 * plane_rows)};
    cuuint64_t gstr[1] = {(cuuint64_t)Kp};
    // RID_I8_CLUSTER=2 selects the 2 x 2 cluster / TMA-multicast variant (every CTA then loads half-panel boxes).  It is
    // correct but measured SLOWER at cfg4 (199.6 ms vs 176 ms): halving the L2 operand traffic does not help because the
    // kernel is bound by shared-memory bandwidth (13 digit products re-read their A and B panels: 146 B/clk of operand
    // reads + 45 B/clk of TMA writes against 128 B/clk), and 4-CTA clusters strand SMs (148 is not a multiple of 4 per GPC).
    const char *cl_env = getenv("RID_I8_CLUSTER");
    const int CL = (cl_env && atoi(cl_env) == 2) ? 2 : 1;
    // Default: the persistent single-CTA kernel.  RID_I8_PAIR=1 selects the CTA-pair (cta_group::2) kernel -- correct, but
    // measured no faster than the single-shot single-CTA kernel at cfg4 (161.6 vs 158.8 ms): operand bandwidth out of
    // shared memory is not what bounds these kernels, the serialised epilogue is.  RID_I8_PERSIST=0 selects the
    // single-shot single-CTA kernel.
    const char *pair_env = getenv("RID_I8_PAIR"), *pers_env = getenv("RID_I8_PERSIST");
    const bool PAIR = CL == 1 && pair_env && atoi(pair_env) == 1;
    const bool PERSIST = CL == 1 && !PAIR && !(pers_env && atoi(pers_env) == 0);
    cuuint32_t boxA[2] = {I8_BK, (cuuint32_t)(I8_BM / CL)}, boxB[2] = {I8_BK, (cuuint32_t)(PAIR ? I8_BN / 2 : I8_BN / CL)}, estr[2] = {1, 1};
    if (encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, planes, gdim, gstr, boxA, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, planes, gdim, gstr, boxB, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        cleanup();
        return RID_OK;
    }
    I8Params P;
    P.N = N; P.nkb = Kp / I8_BK; P.D = dm->D; P.ld = dm->ld; P.plane_rows = (int)plane_rows; P.zero_var = zero_var;
    P.rowscale = rowscale; P.inv_step = 1.0 / dm->step;
    P.row0 = dm->row0; P.nloc = nloc; P.world = ctx->world; P.rank = ctx->rank;
    {
        int r0, r1;
        rid_shard_rows(N, 0, ctx->world, &r0, &r1);
        P.per = ctx->world > 1 ? std::max(r1 - r0, 1) : std::max(N, 1);
    }
    for (int r = 0; r < 16; ++r) P.peer[r] = (u32 *)peers[r];
    if (ctx->world == 1) P.peer[0] = dm->D;
    P.tiles_loc = tiles_loc; P.tiles_n = tiles_n;
    if (cudaFuncSetAttribute(k_gram_i8_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_gram_i8_2sm, cudaFuncAttributeMaxDynamicSharedMemorySize, I8P_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_gram_i8<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_gram_i8<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM) != cudaSuccess) {
        cudaGetLastError(); cleanup(); return RID_OK;
    }
    // pad columns [N, ld) of every row are never touched by the tiles: clear them
    if (nloc > 0 && dm->ld > (size_t)N)
        cudaMemset2DAsync(dm->D + N, dm->ld * sizeof(u32), 0, (dm->ld - (size_t)N) * sizeof(u32), (size_t)nloc, ctx->stream);
    // nobody may start storing mirror tiles into a peer before that peer has cleared its pad columns
    if (ctx->world > 1) RID_TRY(rid_barrier(ctx));
    dim3 grid((unsigned)round_up((size_t)tiles_n, CL), (unsigned)round_up((size_t)std::max(tiles_loc, 1), CL));
    cudaEvent_t pe = rid_prof_begin(ctx, RID_PROF_GRAM_I8);
    if (nloc > 0) {
        if (PERSIST) {
            // the list of tiles this rank computes, in launch order (cached: it depends on the shape only)
            const long long key[6] = {N, dm->row0, nloc, ctx->world, ctx->rank, P.per};
            if (memcmp(key, ctx->i8_tiles_key, sizeof(key)) != 0) {
                std::vector<int2> list;
                const long long all = (long long)tiles_n * tiles_loc;
                list.reserve((size_t)(all / 2 + tiles_n + 16));
                for (long long L = 0; L < all; ++L) {
                    int tm, tn;
                    if (i8_tile_at(P, L, tm, tn)) list.push_back(make_int2(tm, tn));
                }
                const size_t bytes = std::max<size_t>(list.size(), 1) * sizeof(int2);
                if (bytes > ctx->i8_tiles_bytes) {
                    if (ctx->i8_tiles) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->i8_tiles); ctx->i8_tiles = nullptr; ctx->i8_tiles_bytes = 0; }
                    if (cudaMalloc(&ctx->i8_tiles, bytes) != cudaSuccess) { cudaGetLastError(); cleanup(); return RID_OK; }
                    ctx->i8_tiles_bytes = bytes;
                }
                // (`list` is a local: copy on the library's stream and wait for it)
                if (cudaMemcpyAsync(ctx->i8_tiles, list.data(), list.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
                    cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
                    cudaGetLastError(); cleanup(); return RID_OK;
                }
                memcpy(ctx->i8_tiles_key, key, sizeof(key));
                ctx->i8_ntiles = (long long)list.size();
            }
            const unsigned nct = (unsigned)std::max<long long>(1, std::min<long long>(ctx->i8_ntiles, (long long)ctx->sm_count));
            k_gram_i8_persist<<<nct, I8_THREADS, I8_SMEM, ctx->stream>>>(tmA, tmB, P, (const int2 *)ctx->i8_tiles, ctx->i8_ntiles);
        } else if (PAIR) {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3((unsigned)(2 * (size_t)((tiles_loc + 1) / 2) * (size_t)tiles_n));
            cfg.blockDim = dim3(I8_THREADS);
            cfg.dynamicSmemBytes = I8P_SMEM;
            cfg.stream = ctx->stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            if (cudaLaunchKernelEx(&cfg, k_gram_i8_2sm, tmA, tmB, P) != cudaSuccess) {
                rid_set_error("k_gram_i8_2sm launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                cleanup();
                return RID_ERR_CUDA;
            }
        } else if (CL == 2) {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = grid;
            cfg.blockDim = dim3(I8_THREADS);
            cfg.dynamicSmemBytes = I8_SMEM;
            cfg.stream = ctx->stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 2;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            if (cudaLaunchKernelEx(&cfg, k_gram_i8<2>, tmA, tmB, P) != cudaSuccess) {
                rid_set_error("k_gram_i8 cluster launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                cleanup();
                return RID_ERR_CUDA;
            }
        } else {
            k_gram_i8<1><<<grid, I8_THREADS, I8_SMEM, ctx->stream>>>(tmA, tmB, P);
        }
        ctx->launches++;
    }
    // work: int8 multiply-adds actually issued x 2 (13 digit products per computed tile)
    double tiles_done = 0;
    for (int bm = 0; bm < tiles_loc; ++bm)
        for (int bn = 0; bn < tiles_n; ++bn) {
            const int i0 = P.row0 + bm * I8_BM, j0 = bn * I8_BN;
            const int b1 = std::min(j0 / P.per, P.world - 1), b2 = std::min(std::min(j0 + I8_BN - 1, N - 1) / P.per, P.world - 1);
            tiles_done += (i8_pair_mine(P.rank, b1, P.world, i0, j0) || i8_pair_mine(P.rank, b2, P.world, i0, j0));
        }
    (void)tiles_m;
    rid_prof_end(ctx, pe, RID_PROF_GRAM_I8, tiles_done * 13.0 * 2.0 * I8_BM * I8_BN * (double)Kp, (double)N * dm->ld * 4.0);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    // peers have written mirror tiles into this rank's shard: nobody may go on before every kernel is done
    if (e == cudaSuccess && ctx->world > 1 && rid_barrier(ctx) != RID_OK) e = cudaErrorUnknown;
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (e != cudaSuccess) {
        rid_set_error("k_gram_i8 failed: %s", cudaGetErrorString(e));
        return RID_ERR_CUDA;
    }
    *used = 1;
    return RID_OK;
}

// ----------------------------------------------------------------------------------------------------------
// Ceiling of the int8 tensor pipe, measured on the spot (MEASURED_PEAKS.json holds a bf16 figure only): one CTA per SM
// issues a long stream of the same tcgen05.mma.kind::i8 instructions the distance kernel uses (M = 128, N = 192, K = 32,
// operands in SWIZZLE_128B shared memory, two TMEM accumulators) with no loads and no epilogue.
// ----------------------------------------------------------------------------------------------------------
#define I8PK_SMEM (I8_A_PLANE + 2 * I8_B_PLANE + 1024)
__global__ void __launch_bounds__(128, 1) k_i8_peak(int iters)
{
    extern __shared__ uint8_t pk_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)pk_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t done_bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < (I8_A_PLANE + 2 * I8_B_PLANE) / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x01010101u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 1 && lane == 0) {
        mbar_init(&done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(I8_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    if (warp == 1 && lane == 0) {
        const uint32_t sa = smem_u32(smem), sb = sa + I8_A_PLANE;
        const uint64_t A = umma_desc_sw128(sa), B = umma_desc_sw128(sb);
        constexpr uint32_t ID192 = I8_IDESC_N(2 * I8_BN);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                umma_i8(tmem_base, A + 2 * kk, B + 2 * kk, (it | kk) ? 1u : 0u, ID192);
                umma_i8(tmem_base + 2 * I8_BN, A + 2 * kk, B + 2 * kk, (it | kk) ? 1u : 0u, ID192);
            }
        }
        umma_commit(&done_bar);
    }
    // (bounded wait, 2^26 polls: far beyond the ~ms this loop takes)
    mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(I8_TMEM_COLS) : "memory");
}

extern "C" int rid_measure_i8_peak(rid_ctx *ctx, double *tops)
{
    if (!ctx || !tops) return RID_ERR_ARG;
    RID_CUDA(cudaSetDevice(ctx->device));
    RID_CUDA(cudaFuncSetAttribute(k_i8_peak, cudaFuncAttributeMaxDynamicSharedMemorySize, I8PK_SMEM));
    const int iters = 4000, blocks = ctx->sm_count;
    k_i8_peak<<<blocks, 128, I8PK_SMEM, ctx->stream>>>(200); // warm-up
    cudaEvent_t a, b;
    RID_CUDA(cudaEventCreate(&a));
    RID_CUDA(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        RID_CUDA(cudaEventRecord(a, ctx->stream));
        k_i8_peak<<<blocks, 128, I8PK_SMEM, ctx->stream>>>(iters);
        RID_CUDA(cudaEventRecord(b, ctx->stream));
        RID_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        RID_CUDA(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    RID_CUDA(cudaGetLastError());
    // 8 instructions per iteration, each 2 * 128 * 192 * 32 int8 operations
    *tops = (double)blocks * iters * 8.0 * 2.0 * I8_BM * (2.0 * I8_BN) * 32.0 / (best * 1e-3) / 1e12;
    cudaEventDestroy(a); cudaEventDestroy(b);
    return RID_OK;
}
