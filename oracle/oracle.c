/*
 * oracle.c — CPU restatement of RaceID's clustering hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's library.  The product path (raceid3_stemid2_package_b200) never does.
 *
 * PARITY UNPINNED: the reference's arithmetic for this path lives in R core (stats::cor, stats::dist,
 * rank, sum) and in the un-vendored, un-pinned CRAN packages `cluster` (pam) and `fpc` (clusterboot)
 * (reference DESCRIPTION:14, NAMESPACE:129-130,174,178).  R is not installed in this image and the
 * reference ships no tests or golden vectors, so every "upstream:" rule below is a
 * restatement of the published algorithm (R 4.x src/library/stats/src/cov.c, distance.c; cluster 2.1.x
 * src/pam.c bswap()/cstat()/dark(); fpc 2.2-x R/clusterboot.R).  What pins it here: numpy/scipy
 * cross-checks and brute-force checks in tests/test_oracle_*.py.
 *
 * Call sites restated (reference file:line):
 *   dist.gen                R/RaceID_utils.R:9-15      -> orc_dist()
 *   clusGapExt / W.k        R/RaceID_utils.R:29-61     -> orc_wk(), orc_gap_sweep()
 *   clustfun saturation     R/RaceID_utils.R:116-124   -> orc_saturation()
 *   cluster::pam call       R/RaceID_utils.R:113,221   -> orc_pam()
 *   fpc::clusterboot call   R/RaceID_utils.R:132-133   -> orc_clusterboot()
 *   compmedoids             R/RaceID.R:1298-1316       -> orc_medoids()
 *   final partition refresh R/RaceID.R:1034-1060       -> orc_refresh()
 *   plotsilhouette          R/RaceID.R:1283-1284       -> orc_silhouette()   (cluster::silhouette; pinned by scikit-learn)
 *   comptsne start config   R/RaceID.R:1093            -> oracle.py cmdscale() (stats::cmdscale, numpy eigh)
 *
 * All matrices are column-major like R.  Indices crossing the API are 0-based; labels are 1-based
 * like R's cluster numbers.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef long double LDOUBLE; /* upstream: R accumulates in LDOUBLE in cov.c and in sum()/mean() */

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py --impl reference: launchers such as torchrun export OMP_NUM_THREADS=1; the arm asks for all cores itself */
void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---------------------------------------------------------------------------------------------
 * rank(x, ties.method="average")  — upstream: R rank(); used by cor(method="spearman") through
 * apply(X,2,rank) (RaceID_utils.R:10).  Ranks are over the G genes within one cell.
 * ------------------------------------------------------------------------------------------- */
typedef struct { double v; int i; } orc_vi;
static int cmp_vi(const void *a, const void *b)
{
    double x = ((const orc_vi *)a)->v, y = ((const orc_vi *)b)->v;
    if (x < y) return -1;
    if (x > y) return 1;
    int i = ((const orc_vi *)a)->i, j = ((const orc_vi *)b)->i;
    return (i > j) - (i < j);
}
static void rank_average(const double *x, int n, double *r, orc_vi *tmp)
{
    for (int i = 0; i < n; ++i) { tmp[i].v = x[i]; tmp[i].i = i; }
    qsort(tmp, (size_t)n, sizeof(orc_vi), cmp_vi);
    int i = 0;
    while (i < n) {
        int j = i;
        while (j + 1 < n && tmp[j + 1].v == tmp[i].v) ++j;
        double avg = 0.5 * ((double)(i + 1) + (double)(j + 1));
        for (int t = i; t <= j; ++t) r[tmp[t].i] = avg;
        i = j + 1;
    }
}

/* ---------------------------------------------------------------------------------------------
 * dist.gen (RaceID_utils.R:9-15).  x is G x N column-major: column c holds the G feature values of
 * cell c (this is `t(x)` inside dist.gen, i.e. what cor() sees).  D is N x N column-major.
 * metric: 0 pearson, 1 spearman, 2 logpearson, 3 euclidean.
 * Returns the number of zero-variance cells (their rows/cols are NaN, as cor() gives NA + warning).
 *
 * upstream cov.c (cor, no NAs, pearson): two-pass column means in LDOUBLE; for j<=i the cross
 * product sum((x_i-m_i)(x_j-m_j)) in LDOUBLE divided by n-1; sd=sqrt(diag); r=cov/(sd_i*sd_j),
 * clamped to [-1,1]; diagonal set to exactly 1.  Lower triangle computed, then mirrored.
 * upstream distance.c (euclidean): per pair sum of squared differences in double, then sqrt.
 * ------------------------------------------------------------------------------------------- */
int orc_dist(const double *x, int G, int N, int metric, double *D)
{
    size_t g = (size_t)G, n = (size_t)N;
    if (metric == 3) {
#pragma omp parallel for schedule(dynamic, 8)
        for (int i = 0; i < N; ++i) {
            D[(size_t)i * n + i] = 0.0;
            for (int j = 0; j < i; ++j) {
                double dist = 0.0;
                const double *a = x + (size_t)i * g, *b = x + (size_t)j * g;
                for (int k = 0; k < G; ++k) {
                    double dev = a[k] - b[k];
                    dist += dev * dev;
                }
                dist = sqrt(dist);
                D[(size_t)i * n + j] = dist;
                D[(size_t)j * n + i] = dist;
            }
        }
        return 0;
    }
    double *y = (double *)malloc(g * n * sizeof(double));
    double *xm = (double *)malloc(n * sizeof(double));
    double *sd = (double *)malloc(n * sizeof(double));
    if (metric == 1) {
#pragma omp parallel
        {
            orc_vi *tmp = (orc_vi *)malloc(g * sizeof(orc_vi));
#pragma omp for
            for (int c = 0; c < N; ++c) rank_average(x + (size_t)c * g, G, y + (size_t)c * g, tmp);
            free(tmp);
        }
    } else if (metric == 2) {
        for (size_t t = 0; t < g * n; ++t) y[t] = log2(x[t]);
    } else {
        memcpy(y, x, g * n * sizeof(double));
    }
    /* means: two passes, LDOUBLE */
    for (int c = 0; c < N; ++c) {
        const double *col = y + (size_t)c * g;
        LDOUBLE sum = 0.;
        for (int k = 0; k < G; ++k) sum += col[k];
        LDOUBLE tmp = sum / G;
        if (isfinite((double)tmp)) {
            sum = 0.;
            for (int k = 0; k < G; ++k) sum += (col[k] - tmp);
            tmp = tmp + sum / G;
        }
        xm[c] = (double)tmp;
    }
    int n1 = G - 1;
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < N; ++i) {
        const double *xx = y + (size_t)i * g;
        LDOUBLE xxm = xm[i];
        for (int j = 0; j <= i; ++j) {
            const double *yy = y + (size_t)j * g;
            LDOUBLE yym = xm[j];
            LDOUBLE sum = 0.;
            for (int k = 0; k < G; ++k) sum += (xx[k] - xxm) * (yy[k] - yym);
            double v = (double)(sum / n1);
            D[(size_t)i * n + j] = v;
            D[(size_t)j * n + i] = v;
        }
    }
    int nzero = 0;
    for (int i = 0; i < N; ++i) {
        sd[i] = sqrt(D[(size_t)i * n + i]);
        if (sd[i] == 0.0) ++nzero;
    }
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < N; ++i) {
        for (int j = 0; j < i; ++j) {
            double r;
            if (sd[i] == 0. || sd[j] == 0.) {
                r = NAN;
            } else {
                r = D[(size_t)i * n + j] / (sd[i] * sd[j]);
                if (r >= 1.) r = 1.; else if (r <= -1.) r = -1.;
            }
            /* dist.gen: 1 - cor(...) */
            D[(size_t)i * n + j] = 1. - r;
            D[(size_t)j * n + i] = 1. - r;
        }
    }
    for (int i = 0; i < N; ++i) D[(size_t)i * n + i] = (sd[i] == 0.) ? NAN : 0.0; /* 1 - 1 */
    free(y); free(xm); free(sd);
    return nzero;
}

/* ---------------------------------------------------------------------------------------------
 * cluster::pam(as.dist(x), k) with defaults (diss=TRUE, do.swap=TRUE, variant="original").
 * upstream: cluster/src/pam.c bswap() + cstat() + dark().
 *
 * D: n_full x n_full column-major full matrix.  as.dist() keeps the LOWER triangle, so the value
 * used for the pair (a,b) is D[max(a,b), min(a,b)].
 * subset: m indices into D (order significant: it is the row order of D[subset,subset]); NULL = all.
 * med_out[k]: medoid POSITIONS (0-based, into subset) in cluster-number order (id.med).
 * labels[m]: 1-based cluster numbers by first appearance (cstat).
 * obj[2]: build, swap objective (TD/m).
 * swaps: optional log, 3 doubles per accepted swap (h position, removed medoid position, dz); up to
 * max_swaps entries; *n_swaps receives the count.
 * build_order[k]: optional, medoid positions in the order BUILD selected them.
 * ------------------------------------------------------------------------------------------- */
/* as.dist(D[s,s]) keeps the lower triangle of the SUBSETTED matrix: the value for subset positions
 * (a,b) is the element at row position max(a,b), column position min(a,b). */
static inline double dys_at(const double *dmat, size_t ld, const int *idx, int a, int b)
{
    int r = a > b ? a : b, c = a > b ? b : a;
    return dmat[(size_t)idx[c] * ld + (size_t)idx[r]];
}

int orc_pam(const double *dmat, int n_full, const int *subset, int m, int k, int *med_out, int *labels,
            double *obj, double *avg_sil, double *swaps, int max_swaps, int *n_swaps, int *build_order,
            const int *init_med /* optional: k given medoid positions => skip BUILD (pam(medoids=)) */)
{
    size_t ld = (size_t)n_full;
    int n = m;
    if (k < 1 || k >= n) return -1; /* upstream: "Number of clusters 'k' must be in {1,2, .., n-1}" */
    int *idx = (int *)malloc((size_t)n * sizeof(int));
    for (int i = 0; i < n; ++i) idx[i] = subset ? subset[i] : i;
    int *nrepr = (int *)calloc((size_t)n, sizeof(int));
    double *dysma = (double *)malloc((size_t)n * sizeof(double));
    double *dysmb = (double *)malloc((size_t)n * sizeof(double));
    double *beter = (double *)malloc((size_t)n * sizeof(double));
    int rc = 0;

    /* s := max dys; NA check (upstream pam(): "No clustering performed, NAs in the computed
     * dissimilarity matrix.") */
    double s = 0.;
    for (int i = 0; i < n && rc == 0; ++i)
        for (int j = 0; j < i; ++j) {
            double d = dys_at(dmat, ld, idx, i, j);
            if (isnan(d)) { rc = -2; break; }
            if (s < d) s = d;
        }
    if (rc) goto done;
    s = s * 1.1 + 1.;
    for (int i = 0; i < n; ++i) dysma[i] = s;

    if (init_med) {
        for (int t = 0; t < k; ++t) nrepr[init_med[t]] = 1;
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i)
                if (nrepr[i]) {
                    double d = (i == j) ? 0. : dys_at(dmat, ld, idx, i, j);
                    if (dysma[j] > d) dysma[j] = d;
                }
    } else {
        /* BUILD */
        for (int kk = 0; kk < k; ++kk) {
            int nmax = -1;
            double ammax = 0.;
            /* beter[i] for every non-representative (independent per i; OpenMP only in the _omp build) */
#pragma omp parallel for schedule(static)
            for (int i = 0; i < n; ++i) {
                if (nrepr[i] == 0) {
                    double b = 0.;
                    for (int j = 0; j < n; ++j) {
                        double dij = (i == j) ? 0. : dys_at(dmat, ld, idx, i, j);
                        double cmd = dysma[j] - dij;
                        if (cmd > 0.) b += cmd;
                    }
                    beter[i] = b;
                }
            }
            for (int i = 0; i < n; ++i) {
                if (nrepr[i] == 0) {
                    if (ammax <= beter[i]) { /* upstream: ties go to the LARGEST index */
                        ammax = beter[i];
                        nmax = i;
                    }
                }
            }
            nrepr[nmax] = 1;
            if (build_order) build_order[kk] = nmax;
            for (int j = 0; j < n; ++j) {
                double d = (nmax == j) ? 0. : dys_at(dmat, ld, idx, nmax, j);
                if (dysma[j] > d) dysma[j] = d;
            }
        }
    }
    double sky = 0.;
    for (int j = 0; j < n; ++j) sky += dysma[j];
    obj[0] = sky / n;
    int nsw = 0;

    if (k > 1 || init_med) {
        /* SWAP, original variant */
        for (;;) {
            for (int j = 0; j < n; ++j) {
                dysma[j] = s;
                dysmb[j] = s;
                for (int i = 0; i < n; ++i) {
                    if (nrepr[i]) {
                        double d = (i == j) ? 0. : dys_at(dmat, ld, idx, i, j);
                        if (dysma[j] > d) {
                            dysmb[j] = dysma[j];
                            dysma[j] = d;
                        } else if (dysmb[j] > d) {
                            dysmb[j] = d;
                        }
                    }
                }
            }
            double dzsky = 1.;
            int hbest = -1, nbest = -1;
            /* h-loop split into contiguous chunks (one per thread in the _omp build, one chunk otherwise); chunks
             * are merged in increasing-h order with the same strict '>' so the first (h,i) still wins ties. */
            enum { ORC_MAXT = 256 };
            double cdz[ORC_MAXT];
            int ch[ORC_MAXT], ci[ORC_MAXT], nchunks = 1;
#ifdef _OPENMP
            nchunks = omp_get_max_threads();
            if (nchunks > ORC_MAXT) nchunks = ORC_MAXT;
#endif
#pragma omp parallel for schedule(static, 1)
            for (int t = 0; t < nchunks; ++t) {
                int h0 = (int)((long long)n * t / nchunks), h1 = (int)((long long)n * (t + 1) / nchunks);
                double ldz = 1.;
                int lh = -1, li = -1;
                for (int h = h0; h < h1; ++h) {
                    if (nrepr[h]) continue;
                    for (int i = 0; i < n; ++i) {
                        if (!nrepr[i]) continue;
                        double dz = 0.;
                        for (int j = 0; j < n; ++j) {
                            double dhj = (h == j) ? 0. : dys_at(dmat, ld, idx, h, j);
                            double dij = (i == j) ? 0. : dys_at(dmat, ld, idx, i, j);
                            if (dij == dysma[j]) {
                                double small = dysmb[j] > dhj ? dhj : dysmb[j];
                                dz += (-dysma[j] + small);
                            } else if (dhj < dysma[j]) {
                                dz += (-dysma[j] + dhj);
                            }
                        }
                        if (ldz > dz) { /* upstream: first (h,i) wins ties */
                            ldz = dz;
                            lh = h;
                            li = i;
                        }
                    }
                }
                cdz[t] = ldz; ch[t] = lh; ci[t] = li;
            }
            for (int t = 0; t < nchunks; ++t)
                if (dzsky > cdz[t]) { dzsky = cdz[t]; hbest = ch[t]; nbest = ci[t]; }
            if (dzsky < -16 * DBL_EPSILON * fabs(sky)) {
                nrepr[hbest] = 1;
                nrepr[nbest] = 0;
                if (swaps && nsw < max_swaps) {
                    swaps[3 * nsw + 0] = hbest;
                    swaps[3 * nsw + 1] = nbest;
                    swaps[3 * nsw + 2] = dzsky;
                }
                ++nsw;
                sky += dzsky;
                continue;
            }
            break;
        }
    }
    obj[1] = sky / n;
    if (n_swaps) *n_swaps = nsw;

    /* cstat: nearest medoid (strict <, medoids in index order => lowest index on ties) */
    {
        int *nsend = (int *)malloc((size_t)n * sizeof(int));
        for (int j = 0; j < n; ++j) {
            if (nrepr[j] == 0) {
                double dsmal = s; /* upstream recomputes s*1.1+1 from the same max */
                int ksmal = -1;
                for (int kk = 0; kk < n; ++kk) {
                    if (nrepr[kk] == 1) {
                        double d = dys_at(dmat, ld, idx, kk, j);
                        if (d < dsmal) { dsmal = d; ksmal = kk; }
                    }
                }
                nsend[j] = ksmal;
            } else {
                nsend[j] = j;
            }
        }
        /* cluster numbers by first appearance scanning objects 1..n */
        int jk = 0;
        for (int j = 0; j < n; ++j) labels[j] = 0;
        for (int ja = 0; ja < n; ++ja) {
            int nplac = nsend[ja];
            if (labels[nplac] == 0) {
                ++jk;
                for (int j = 0; j < n; ++j)
                    if (nsend[j] == nplac) labels[j] = jk;
                med_out[jk - 1] = nplac;
                if (jk == k) break;
            }
        }
        free(nsend);
    }

    /* dark(): silhouette widths, only needed for avg.width (RaceID_utils.R:224) */
    if (avg_sil) {
        if (k == 1) {
            *avg_sil = 0.;
        } else {
            int *cnt = (int *)calloc((size_t)k, sizeof(int));
            double *sums = (double *)malloc((size_t)k * sizeof(double));
            for (int j = 0; j < n; ++j) cnt[labels[j] - 1]++;
            LDOUBLE tot = 0.;
            for (int j = 0; j < n; ++j) {
                for (int c = 0; c < k; ++c) sums[c] = 0.;
                for (int l = 0; l < n; ++l)
                    if (l != j) sums[labels[l] - 1] += dys_at(dmat, ld, idx, j, l);
                int cj = labels[j] - 1;
                double sw;
                if (cnt[cj] == 1) {
                    sw = 0.;
                } else {
                    double a = sums[cj] / (cnt[cj] - 1);
                    double b = INFINITY;
                    for (int c = 0; c < k; ++c)
                        if (c != cj) {
                            double v = sums[c] / cnt[c];
                            if (v < b) b = v;
                        }
                    double mx = a > b ? a : b;
                    sw = (mx > 0.) ? (b - a) / mx : 0.;
                }
                tot += sw;
            }
            *avg_sil = (double)(tot / n);
            free(cnt); free(sums);
        }
    }
done:
    free(idx); free(nrepr); free(dysma); free(dysmb); free(beter);
    return rc;
}

/* ---------------------------------------------------------------------------------------------
 * W.k with diss=TRUE (RaceID_utils.R:39-53): 0.5 * sum_c sum(X[I,I]/|I|).  Uses the FULL
 * (both-triangle) matrix like the R code does; sums in LDOUBLE like R's sum().
 * labels are 1-based over the m subset positions.
 * ------------------------------------------------------------------------------------------- */
double orc_wk(const double *dmat, int n_full, const int *subset, int m, const int *labels, int k)
{
    size_t ld = (size_t)n_full;
    LDOUBLE total = 0.;
    int *cnt = (int *)calloc((size_t)k + 1, sizeof(int));
    for (int i = 0; i < m; ++i) cnt[labels[i]]++;
    for (int c = 1; c <= k; ++c) {
        if (!cnt[c]) continue;
        LDOUBLE sc = 0.;
        double nc = (double)cnt[c];
        /* column-major traversal of X[I,I] */
        for (int b = 0; b < m; ++b) {
            if (labels[b] != c) continue;
            size_t cb = (size_t)(subset ? subset[b] : b) * ld;
            for (int a = 0; a < m; ++a) {
                if (labels[a] != c) continue;
                sc += dmat[cb + (size_t)(subset ? subset[a] : a)] / nc;
            }
        }
        total += (double)sc; /* vapply(..., 0) returns doubles, then sum() */
    }
    free(cnt);
    return 0.5 * (double)total;
}

/* clusGapExt's k loop (RaceID_utils.R:58-61) with FUNcluster = pam, random=FALSE, diss=TRUE */
int orc_gap_sweep(const double *dmat, int n_full, const int *subset, int m, int kmax, double *logW)
{
    int *labels = (int *)malloc((size_t)m * sizeof(int));
    int *med = (int *)malloc((size_t)(kmax > 0 ? kmax : 1) * sizeof(int));
    double obj[2];
    int rc = 0;
    for (int k = 1; k <= kmax; ++k) {
        if (k > 1) {
            rc = orc_pam(dmat, n_full, subset, m, k, med, labels, obj, NULL, NULL, 0, NULL, NULL, NULL);
            if (rc) break;
        } else {
            for (int i = 0; i < m; ++i) labels[i] = 1;
        }
        logW[k - 1] = log(orc_wk(dmat, n_full, subset, m, labels, k));
    }
    free(labels); free(med);
    return rc;
}

/* clustfun saturation rule (RaceID_utils.R:116-124).  g = logW[0..K-1].  Returns cln. */
int orc_saturation(const double *g, int K)
{
    int L = K - 1;
    if (L < 1) return 1;
    double *y = (double *)malloc((size_t)L * sizeof(double));
    for (int i = 0; i < L; ++i) y[i] = g[i] - g[i + 1];
    int first = -1;
    for (int i = 0; i < L; ++i) {
        int cnt = L - i;
        LDOUBLE s = 0.;
        for (int t = i; t < L; ++t) s += y[t];
        s /= cnt; /* R mean(): LDOUBLE sum / n, then a second refinement pass */
        LDOUBLE t2 = 0.;
        for (int t = i; t < L; ++t) t2 += (y[t] - s);
        double mm = (double)(s + t2 / cnt);
        double nn;
        if (cnt < 2) {
            nn = NAN; /* var of one value is NA => comparison NA => which() drops it */
        } else {
            LDOUBLE ss = 0.;
            for (int t = i; t < L; ++t) ss += ((LDOUBLE)y[t] - mm) * ((LDOUBLE)y[t] - mm);
            nn = sqrt((double)(ss / (cnt - 1)));
        }
        double v = y[i] - (mm + nn);
        if (!isnan(v) && v < 0 && first < 0) first = i + 1;
    }
    free(y);
    /* max(min(which(...)),1); min(integer(0)) = Inf (with a warning) => cln = Inf upstream.  We
     * return K (clamped) in that degenerate case and flag it by a negative sign. */
    if (first < 0) return -K;
    return first > 1 ? first : 1;
}

/* ---------------------------------------------------------------------------------------------
 * fpc::clusterboot(diM, B, bootmethod="boot"|"subset", clustermethod=pamkdCBI, k, multipleboot=FALSE,
 * bscompare=TRUE) given the index vectors (RNG stays with the caller: R's sample()).
 * samples: concatenated 0-based index vectors (already unique()'d, first-occurrence order);
 * sample_len[B].  partition[n]: labels of c1 (PAM on all cells).  bootresult: k x B column-major.
 * upstream clujaccard(c1,c2,zerobyzero=0) = sum(c1&c2)/sum(c1|c2).
 * ------------------------------------------------------------------------------------------- */
int orc_clusterboot(const double *dmat, int n, int k, const int *samples, const int *sample_len, int B,
                    int *partition, int *medoids, double *bootresult, double *bootmean)
{
    double obj[2];
    int rc = orc_pam(dmat, n, NULL, n, k, medoids, partition, obj, NULL, NULL, 0, NULL, NULL, NULL);
    if (rc) return rc;
    const int *sp = samples;
    int *bl = (int *)malloc((size_t)n * sizeof(int));
    int *bm = (int *)malloc((size_t)k * sizeof(int));
    long *tab = (long *)malloc((size_t)k * k * sizeof(long));
    long *ca = (long *)malloc((size_t)k * sizeof(long));
    long *cb = (long *)malloc((size_t)k * sizeof(long));
    for (int b = 0; b < B; ++b) {
        int m = sample_len[b];
        rc = orc_pam(dmat, n, sp, m, k, bm, bl, obj, NULL, NULL, 0, NULL, NULL, NULL);
        if (rc) break;
        memset(tab, 0, (size_t)k * k * sizeof(long));
        memset(ca, 0, (size_t)k * sizeof(long));
        memset(cb, 0, (size_t)k * sizeof(long));
        for (int t = 0; t < m; ++t) {
            int a = partition[sp[t]] - 1, c = bl[t] - 1;
            tab[a * k + c]++; ca[a]++; cb[c]++;
        }
        for (int j = 0; j < k; ++j) {
            double best = 0.;
            for (int c = 0; c < k; ++c) {
                long inter = tab[j * k + c], uni = ca[j] + cb[c] - inter;
                double cj = (uni == 0) ? 0. : (double)inter / (double)uni;
                if (cj > best) best = cj;
            }
            bootresult[(size_t)b * k + j] = best;
        }
        sp += m;
    }
    if (!rc && bootmean)
        for (int j = 0; j < k; ++j) {
            LDOUBLE s = 0.;
            for (int b = 0; b < B; ++b) s += bootresult[(size_t)b * k + j];
            bootmean[j] = (double)(s / B); /* rowMeans */
        }
    free(bl); free(bm); free(tab); free(ca); free(cb);
    return rc;
}

/* ---------------------------------------------------------------------------------------------
 * compmedoids (RaceID.R:1298-1316): per cluster (increasing cluster number) y = colMeans(D[f,f]);
 * medoid = f[which(y==min(y))[1]].  labels: arbitrary positive ints over n cells (0 = cell not in
 * `part`).  Returns number of clusters; medoid_idx[] 0-based cell indices, cluster_ids[] the
 * sorted unique labels.
 * ------------------------------------------------------------------------------------------- */
int orc_medoids(const double *dmat, int n, const int *labels, int *medoid_idx, int *cluster_ids)
{
    size_t ld = (size_t)n;
    int maxl = 0;
    for (int i = 0; i < n; ++i) if (labels[i] > maxl) maxl = labels[i];
    int nc = 0;
    for (int c = 1; c <= maxl; ++c) {
        int cnt = 0;
        for (int i = 0; i < n; ++i) cnt += (labels[i] == c);
        if (!cnt) continue;
        int best = -1;
        double bestv = INFINITY;
        for (int b = 0; b < n; ++b) { /* column b of D[f,f] */
            if (labels[b] != c) continue;
            if (cnt == 1) { best = b; break; }
            LDOUBLE s = 0.;
            for (int a = 0; a < n; ++a)
                if (labels[a] == c) s += dmat[(size_t)b * ld + a];
            double y = (double)(s / cnt);
            if (y < bestv) { bestv = y; best = b; } /* first minimum */
        }
        medoid_idx[nc] = best;
        if (cluster_ids) cluster_ids[nc] = c;
        ++nc;
    }
    return nc;
}

/* ---------------------------------------------------------------------------------------------
 * cluster::silhouette(x = part, dist), as plotsilhouette calls it (RaceID.R:1283-1284).
 * upstream: cluster/R/silhouette.R silhouette.default() -> src/sildist.c: for cell i in cluster A,
 * a(i) = mean distance to the other members of A, b(i) = min over clusters C != A of the mean
 * distance to C (first minimum in increasing cluster number), neighbor = that C,
 * s(i) = (b - a) / max(a, b); a cell alone in its cluster gets s = 0.
 * labels: positive ints (0 = cell not in the partition -> neighbor 0, width NaN).
 * Returns the number of clusters, or -1 if the 2 <= k <= n-1 precondition fails.
 * ------------------------------------------------------------------------------------------- */
int orc_silhouette(const double *dmat, int n, const int *labels, int *neighbor, double *width)
{
    size_t ld = (size_t)n;
    int maxl = 0, m = 0, k = 0;
    for (int i = 0; i < n; ++i) { if (labels[i] > maxl) maxl = labels[i]; m += labels[i] > 0; }
    int *cnt = (int *)calloc((size_t)maxl + 1, sizeof(int));
    double *sum = (double *)malloc(((size_t)maxl + 1) * sizeof(double));
    for (int i = 0; i < n; ++i) if (labels[i] > 0) cnt[labels[i]]++;
    for (int c = 1; c <= maxl; ++c) k += cnt[c] > 0;
    if (k < 2 || k > m - 1) { free(cnt); free(sum); return -1; }
    for (int i = 0; i < n; ++i) {
        neighbor[i] = 0;
        width[i] = NAN;
        if (labels[i] <= 0) continue;
        for (int c = 1; c <= maxl; ++c) sum[c] = 0.;
        for (int j = 0; j < n; ++j)
            if (labels[j] > 0 && j != i) sum[labels[j]] += dmat[(size_t)i * ld + j];
        const int A = labels[i];
        double b = INFINITY;
        int nb = 0;
        for (int c = 1; c <= maxl; ++c)
            if (c != A && cnt[c] > 0) {
                double v = sum[c] / cnt[c];
                if (v < b) { b = v; nb = c; } /* first minimum */
            }
        neighbor[i] = nb;
        if (cnt[A] == 1) { width[i] = 0.; continue; }
        double a = sum[A] / (cnt[A] - 1);
        double mx = a > b ? a : b;
        width[i] = mx > 0. ? (b - a) / mx : 0.;
    }
    free(cnt); free(sum);
    return k;
}

/* final partition refresh (RaceID.R:1049-1054): every cell -> nearest medoid column, first minimum
 * (order(x)[1] is stable), then compact empty cluster numbers. labels_out 1-based. */
void orc_refresh(const double *dmat, int n, const int *medoid_idx, int k, int *labels_out)
{
    size_t ld = (size_t)n;
    for (int i = 0; i < n; ++i) {
        int best = 0;
        double bv = dmat[(size_t)medoid_idx[0] * ld + i];
        for (int c = 1; c < k; ++c) {
            double v = dmat[(size_t)medoid_idx[c] * ld + i];
            if (v < bv) { bv = v; best = c; }
        }
        labels_out[i] = best + 1;
    }
    for (int c = k; c >= 1; --c) {
        int cnt = 0;
        for (int i = 0; i < n; ++i) cnt += (labels_out[i] == c);
        if (cnt == 0)
            for (int i = 0; i < n; ++i) if (labels_out[i] > c) labels_out[i]--;
    }
}

/* ---------------------------------------------------------------------------------------------
 * Checker helpers (not restatements): a FastPAM1-style evaluation of the same swap deltas, used by
 * tests to prove the GPU formulation picks the same (h,i) as the literal loop above.
 *   TDnew(h, i) = sum_j min(d(h,j), i==nearest(j) ? dsecond(j) : dnearest(j));  dz = TDnew - TD.
 * out_dz: n x k (row h, medoid slot c in the order of med[]); entries for medoid rows are +inf.
 * ------------------------------------------------------------------------------------------- */
void orc_swap_deltas_colform(const double *dmat, int n_full, const int *subset, int m, const int *med, int k,
                             double *out_dz)
{
    size_t ld = (size_t)n_full;
    int *idx = (int *)malloc((size_t)m * sizeof(int));
    for (int i = 0; i < m; ++i) idx[i] = subset ? subset[i] : i;
    double *dn = (double *)malloc((size_t)m * sizeof(double));
    double *ds = (double *)malloc((size_t)m * sizeof(double));
    int *nr = (int *)malloc((size_t)m * sizeof(int));
    char *ismed = (char *)calloc((size_t)m, 1);
    for (int c = 0; c < k; ++c) ismed[med[c]] = 1;
    double td = 0.;
    for (int j = 0; j < m; ++j) {
        dn[j] = ds[j] = INFINITY; nr[j] = -1;
        for (int c = 0; c < k; ++c) {
            double d = (med[c] == j) ? 0. : dys_at(dmat, ld, idx, med[c], j);
            if (d < dn[j]) { ds[j] = dn[j]; dn[j] = d; nr[j] = c; }
            else if (d < ds[j]) ds[j] = d;
        }
        td += dn[j];
    }
    for (int h = 0; h < m; ++h)
        for (int c = 0; c < k; ++c) {
            if (ismed[h]) { out_dz[(size_t)h * k + c] = INFINITY; continue; }
            double t = 0.;
            for (int j = 0; j < m; ++j) {
                double d = (h == j) ? 0. : dys_at(dmat, ld, idx, h, j);
                double cap = (nr[j] == c) ? ds[j] : dn[j];
                t += d < cap ? d : cap;
            }
            out_dz[(size_t)h * k + c] = t - td;
        }
    free(idx); free(dn); free(ds); free(nr); free(ismed);
}
