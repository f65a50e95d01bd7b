// mds.cu — classical multidimensional scaling of the resident distance matrix: stats::cmdscale(d, k) as the reference
// calls it for the t-SNE start configuration (R/RaceID.R:1093, `Rtsne(di, initial_config = cmdscale(di, k = 2), ...)`;
// also R/VarID_functions.R:2456).  "Next" row (f-4) of SURVEY §8.
//
// cmdscale: B = -1/2 J D^2 J (J = I - 11'/n, D^2 elementwise), points = E_k diag(sqrt(lambda_k)) for the k algebraically
// largest eigenpairs.  R materialises B (n x n doubles) and runs a dense symmetric eigensolver; here B is never formed.
// Block subspace iteration with Rayleigh-Ritz on P = 8 vectors needs only W = D^2 V per iteration, i.e. ONE streaming
// pass over the fixed-point matrix (the same access pattern as the PAM scans: a thread owns 4 columns, rows are
// streamed with 128-bit loads, D is symmetric so column sums serve as row sums), FP64 FMAs in registers, and the
// n x 8 algebra (centring, Gram-Schmidt, the 8 x 8 Ritz problem) on the host in FP64.  Because V is kept centred,
// B V = -1/2 J (D^2 V).  Multi-rank: every rank scans its row shard, partial W's are all-reduced.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

#define MDS_P 8
#define MDS_THREADS 128
#define MDS_ROWS 128

__device__ __forceinline__ uint4 mds_ld_stream(const u32 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// W[c][h] += sum_{rows j of this CTA's chunks} (D[j][h] step)^2 V[j][c]
__global__ void __launch_bounds__(MDS_THREADS)
k_d2_matvec(const u32 *__restrict__ D, size_t ld, int nloc, int row0, const double *__restrict__ V /* [N][P] */,
            double step, int rows_per_chunk, double *__restrict__ W /* [P][ld] */)
{
    __shared__ double sh_v[MDS_ROWS][MDS_P];
    const size_t col = ((size_t)blockIdx.x * MDS_THREADS + threadIdx.x) * 4;
    const bool active = col < ld;
    const u32 *base = D + col;
    double acc[4][MDS_P];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int p = 0; p < MDS_P; ++p) acc[c][p] = 0.0;
    const double step2 = step * step;
    for (int chunk = blockIdx.y; (long long)chunk * rows_per_chunk < nloc; chunk += gridDim.y) {
        const int r_begin = chunk * rows_per_chunk;
        const int r_end = min(nloc, r_begin + rows_per_chunk);
        for (int r0 = r_begin; r0 < r_end; r0 += MDS_ROWS) {
            const int nr = min(MDS_ROWS, r_end - r0);
            __syncthreads();
            for (int i = threadIdx.x; i < nr * MDS_P; i += MDS_THREADS)
                sh_v[i / MDS_P][i % MDS_P] = V[(size_t)(row0 + r0) * MDS_P + i];
            __syncthreads();
            if (!active) continue;
            for (int r = 0; r < nr; r += 4) {
                uint4 d[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    d[u] = r + u < nr ? mds_ld_stream(base + (size_t)(r0 + r + u) * ld) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (r + u >= nr) break;
                    const u32 q[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        // exact u32 -> f64 without the conversion pipe: 2^52 + q, minus 2^52
                        const double x = __hiloint2double(0x43300000, (int)q[c]) - 4503599627370496.0;
                        const double x2 = x * x;
#pragma unroll
                        for (int p = 0; p < MDS_P; ++p) acc[c][p] = fma(x2, sh_v[r + u][p], acc[c][p]);
                    }
                }
            }
        }
    }
    if (!active) return;
#pragma unroll
    for (int p = 0; p < MDS_P; ++p)
#pragma unroll
        for (int c = 0; c < 4; ++c) atomicAdd(W + (size_t)p * ld + col + c, acc[c][p] * step2);
}

// ---- small dense helpers (host, FP64) ------------------------------------------------------------------------------
// cyclic Jacobi for a symmetric P x P matrix: A -> eigenvalues on the diagonal, S = eigenvectors (columns)
static void jacobi_eig(double A[MDS_P][MDS_P], double S[MDS_P][MDS_P])
{
    for (int i = 0; i < MDS_P; ++i)
        for (int j = 0; j < MDS_P; ++j) S[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < MDS_P; ++i)
            for (int j = 0; j < MDS_P; ++j) (i == j ? diag : off) += A[i][j] * A[i][j];
        if (off <= 1e-30 * (diag + 1e-300)) break;
        for (int p = 0; p < MDS_P - 1; ++p)
            for (int q = p + 1; q < MDS_P; ++q) {
                if (fabs(A[p][q]) < 1e-300) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < MDS_P; ++k) { // columns p, q
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < MDS_P; ++k) { // rows p, q
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < MDS_P; ++k) {
                    const double skp = S[k][p], skq = S[k][q];
                    S[k][p] = c * skp - s * skq;
                    S[k][q] = s * skp + c * skq;
                }
            }
    }
}

// G = A' B for two n x P row-major blocks (one pass over the rows, P x P accumulators)
static void gram(const std::vector<double> &A, const std::vector<double> &B, int n, double G[MDS_P][MDS_P])
{
    double acc[MDS_P][MDS_P];
    for (int a = 0; a < MDS_P; ++a)
        for (int b = 0; b < MDS_P; ++b) acc[a][b] = 0.0;
    for (int i = 0; i < n; ++i) {
        const double *ra = &A[(size_t)i * MDS_P], *rb = &B[(size_t)i * MDS_P];
        for (int a = 0; a < MDS_P; ++a)
            for (int b = 0; b < MDS_P; ++b) acc[a][b] += ra[a] * rb[b];
    }
    for (int a = 0; a < MDS_P; ++a)
        for (int b = 0; b < MDS_P; ++b) G[a][b] = acc[a][b];
}

// columns of X (n x P, row-major) -> orthonormal by Cholesky-QR, twice (X = Q R with R from chol(X'X); the blocks here
// have condition numbers of at most a few thousand, far inside what two rounds handle).  false: rank-deficient block
static bool orthonormalise(std::vector<double> &X, int n)
{
    for (int pass = 0; pass < 2; ++pass) {
        double G[MDS_P][MDS_P], R[MDS_P][MDS_P]; // G = R' R, R upper triangular
        gram(X, X, n, G);
        for (int a = 0; a < MDS_P; ++a)
            for (int b = 0; b < MDS_P; ++b) R[a][b] = 0.0;
        for (int j = 0; j < MDS_P; ++j) {
            double d = G[j][j];
            for (int t = 0; t < j; ++t) d -= R[t][j] * R[t][j];
            if (!(d > 1e-280)) return false;
            R[j][j] = sqrt(d);
            for (int c = j + 1; c < MDS_P; ++c) {
                double v = G[j][c];
                for (int t = 0; t < j; ++t) v -= R[t][j] * R[t][c];
                R[j][c] = v / R[j][j];
            }
        }
        for (int i = 0; i < n; ++i) { // row x R^-1 by forward substitution
            double *r = &X[(size_t)i * MDS_P];
            for (int j = 0; j < MDS_P; ++j) {
                double v = r[j];
                for (int t = 0; t < j; ++t) v -= r[t] * R[t][j];
                r[j] = v / R[j][j];
            }
        }
    }
    return true;
}

static void centre_columns(std::vector<double> &X, int n)
{
    double s[MDS_P] = {0};
    for (int i = 0; i < n; ++i)
        for (int p = 0; p < MDS_P; ++p) s[p] += X[(size_t)i * MDS_P + p];
    for (int p = 0; p < MDS_P; ++p) s[p] /= n;
    for (int i = 0; i < n; ++i)
        for (int p = 0; p < MDS_P; ++p) X[(size_t)i * MDS_P + p] -= s[p];
}

// points: N x k column-major (like the matrix cmdscale returns); eig[k]: the eigenvalues; iters (nullable): iterations
// used.  Signs: every axis is flipped so that its largest-magnitude coordinate is positive (LAPACK's sign is arbitrary).
extern "C" int rid_cmdscale(rid_dmat *dm, int k, double tol, int max_iter, double *points, double *eig, int *iters)
{
    if (!dm || !points || !eig || k < 1) { rid_set_error("rid_cmdscale: bad arguments"); return RID_ERR_ARG; }
    const int N = dm->N;
    if (k > MDS_P - 4 || k > N - 1) {
        rid_set_error("'k' must be in {1, 2, ..  n - 1} (and <= %d on this path)", MDS_P - 4);
        return RID_ERR_ARG;
    }
    if (dm->has_na) { rid_set_error("NA values not allowed in 'd'"); return RID_ERR_NA; }
    if (N < 2 * MDS_P) {
        rid_set_error("rid_cmdscale: needs at least %d cells (block of %d vectors); use stats::cmdscale for tiny inputs", 2 * MDS_P, MDS_P);
        return RID_ERR_UNSUPPORTED;
    }
    if (!(tol > 0.0)) tol = 1e-9;
    if (max_iter < 1) max_iter = 1000;
    rid_ctx *ctx = dm->ctx;
    RID_CUDA(cudaSetDevice(ctx->device));
    const int nloc = dm->row1 - dm->row0;
    const size_t ld = dm->ld;
    double *d_V = nullptr, *d_W = nullptr;
    const size_t v_bytes = std::max<size_t>((size_t)N * MDS_P * sizeof(double), (size_t)256 << 10);
    const size_t w_bytes = std::max<size_t>(ld * MDS_P * sizeof(double), (size_t)256 << 10);
    if (rid_pool_alloc(ctx, (void **)&d_V, v_bytes) != cudaSuccess || rid_pool_alloc(ctx, (void **)&d_W, w_bytes) != cudaSuccess) {
        cudaGetLastError();
        rid_pool_free(ctx, d_V, v_bytes);
        rid_set_error("rid_cmdscale: out of device memory");
        return RID_ERR_NOMEM;
    }
    auto cleanup = [&]() { cudaStreamSynchronize(ctx->stream); rid_pool_free(ctx, d_V, v_bytes); rid_pool_free(ctx, d_W, w_bytes); };
    // start block: a fixed pseudo-random matrix (the library draws no random numbers from the caller's stream)
    std::vector<double> V((size_t)N * MDS_P), W((size_t)N * MDS_P), Wd(ld * MDS_P), X((size_t)N * MDS_P);
    {
        unsigned long long s = 0x9E3779B97F4A7C15ull;
        for (size_t i = 0; i < V.size(); ++i) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            V[i] = (double)(s >> 11) * (1.0 / 9007199254740992.0) - 0.5;
        }
    }
    centre_columns(V, N);
    if (!orthonormalise(V, N)) { cleanup(); rid_set_error("rid_cmdscale: degenerate start block"); return RID_ERR_ARG; }

    int rc = rid_time_begin(ctx);
    const int strips = (int)((ld + MDS_THREADS * 4 - 1) / (MDS_THREADS * 4));
    const int want_y = std::max(1, (ctx->sm_count * 8 + strips - 1) / strips);
    int rpc = (std::max(nloc, 1) + want_y - 1) / want_y;
    rpc = std::max((rpc + MDS_ROWS - 1) / MDS_ROWS * MDS_ROWS, MDS_ROWS);
    const dim3 grid(strips, std::max(1, std::min((std::max(nloc, 1) + rpc - 1) / rpc, want_y)));
    double theta[MDS_P], res[MDS_P];
    int order[MDS_P];
    double S[MDS_P][MDS_P];
    int it = 0;
    bool converged = false;
    for (it = 1; it <= max_iter && rc == RID_OK; ++it) {
        if (cudaMemcpyAsync(d_V, V.data(), (size_t)N * MDS_P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaMemsetAsync(d_W, 0, ld * MDS_P * sizeof(double), ctx->stream) != cudaSuccess) { rc = RID_ERR_CUDA; break; }
        if (nloc > 0) {
            k_d2_matvec<<<grid, MDS_THREADS, 0, ctx->stream>>>(dm->D, ld, nloc, dm->row0, d_V, dm->step, rpc, d_W);
            ctx->launches++;
        }
        if ((rc = rid_allreduce_f64(ctx, d_W, ld * MDS_P)) != RID_OK) break;
        if (cudaMemcpyAsync(Wd.data(), d_W, ld * MDS_P * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = RID_ERR_CUDA; break; }
        // W = B V = -1/2 J (D^2 V)
        for (int p = 0; p < MDS_P; ++p)
            for (int i = 0; i < N; ++i) W[(size_t)i * MDS_P + p] = -0.5 * Wd[(size_t)p * ld + i];
        centre_columns(W, N);
        // Rayleigh-Ritz on span(V): H = V' W
        double H[MDS_P][MDS_P];
        gram(V, W, N, H);
        for (int a = 0; a < MDS_P; ++a)
            for (int b = a + 1; b < MDS_P; ++b) H[a][b] = H[b][a] = 0.5 * (H[a][b] + H[b][a]);
        jacobi_eig(H, S);
        for (int p = 0; p < MDS_P; ++p) { theta[p] = H[p][p]; order[p] = p; }
        std::sort(order, order + MDS_P, [&](int a, int b) { return theta[a] > theta[b]; }); // algebraically largest first
        // residuals of the k leading Ritz pairs: || W s - theta V s || / |theta|
        bool ok = true;
        for (int t = 0; t < k; ++t) {
            const int p = order[t];
            long double rr = 0.0L;
            for (int i = 0; i < N; ++i) {
                double ws = 0.0, vs = 0.0;
                for (int a = 0; a < MDS_P; ++a) { ws += W[(size_t)i * MDS_P + a] * S[a][p]; vs += V[(size_t)i * MDS_P + a] * S[a][p]; }
                const double r = ws - theta[p] * vs;
                rr += (long double)r * r;
                X[(size_t)i * MDS_P + t] = vs;
            }
            res[t] = sqrt((double)rr) / std::max(fabs(theta[p]), 1e-300);
            if (!(res[t] < tol)) ok = false;
        }
        if (ok) { converged = true; break; }
        V.swap(W); // next block: orth(B V)
        if (!orthonormalise(V, N)) { rc = RID_ERR_ARG; rid_set_error("rid_cmdscale: the iteration block lost rank"); break; }
        centre_columns(V, N); // (stays centred up to rounding; keep it exact)
    }
    if (rc == RID_OK) rc = rid_time_end(ctx);
    cleanup();
    if (rc != RID_OK) {
        if (rc == RID_ERR_CUDA) rid_set_error("rid_cmdscale: %s", cudaGetErrorString(cudaGetLastError()));
        return rc;
    }
    if (iters) *iters = std::min(it, max_iter);
    if (!converged) {
        rid_set_error("rid_cmdscale: no convergence to %.1e in %d iterations (residuals %.2e ...)", tol, max_iter, res[0]);
        return RID_ERR_UNSUPPORTED;
    }
    for (int t = 0; t < k; ++t) {
        const double ev = theta[order[t]];
        eig[t] = ev;
        // cmdscale: only positive eigenvalues give coordinates ("only %d of the first %d eigenvalues are > 0")
        const double sc = ev > 0.0 ? sqrt(ev) : NAN;
        double big = 0.0;
        for (int i = 0; i < N; ++i) if (fabs(X[(size_t)i * MDS_P + t]) > fabs(big)) big = X[(size_t)i * MDS_P + t];
        const double sign = big < 0.0 ? -1.0 : 1.0;
        for (int i = 0; i < N; ++i) points[(size_t)t * N + i] = sign * sc * X[(size_t)i * MDS_P + t];
    }
    return RID_OK;
}
