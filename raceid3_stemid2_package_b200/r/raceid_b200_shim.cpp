// raceid_b200_shim.cpp — the Rcpp side of the drop-in boundary.  NOT compiled in this repo's image (no R, no Rcpp
// headers here); it is the file a RaceID maintainer adds under src/ next to VarID.cpp, following the package's own
// convention (reference src/RcppExports.cpp:15-24: RcppExport SEXP _RaceID_<fn>(...) with BEGIN_RCPP/END_RCPP).
// Link with:  PKG_LIBS = -L$(RACEID_B200_HOME) -lraceid_b200   (src/Makevars, next to CXX_STD = CXX17)
//
// Every function below is a thin marshalling layer over include/raceid_b200.h: R owns names, seeds, argument
// validation and object packing; the library owns the arithmetic.
#include <Rcpp.h>
#include "raceid_b200.h"
using namespace Rcpp;

// RID_ERR_INTERRUPTED: the user pressed Ctrl-C while a call was running (see rid_poll_user_interrupt): re-raise it as
// R's own interrupt condition, like cluster::pam's R_CheckUserInterrupt does; everything else -> R error condition
static void rid_check(int rc) {
    if (rc == RID_ERR_INTERRUPTED) Rcpp::checkUserInterrupt();
    if (rc != RID_OK) Rcpp::stop(rid_last_error());
}

// The library polls this between waves of GPU work, on the R main thread (the thread that made the .Call).
// R_CheckUserInterrupt longjmps on an interrupt, which must not unwind through the library's frames, so it runs inside
// R_ToplevelExec (FALSE = it was interrupted); the library then drains the GPU, frees its work buffers and returns
// RID_ERR_INTERRUPTED, which rid_check turns back into the interrupt.
static void rid_check_interrupt_fn(void*) { R_CheckUserInterrupt(); }
static int rid_poll_user_interrupt(void*) { return R_ToplevelExec(rid_check_interrupt_fn, nullptr) == FALSE; }

static void ctx_finalizer(rid_ctx* c)   { rid_ctx_destroy(c); }
static void dmat_finalizer(rid_dmat* d) { rid_dmat_free(d); }
typedef XPtr<rid_ctx,  PreserveStorage, ctx_finalizer>  CtxPtr;
typedef XPtr<rid_dmat, PreserveStorage, dmat_finalizer> DmatPtr;

// [[Rcpp::export]]
SEXP rid_context(int device = 0) {
    rid_ctx* c = nullptr;
    rid_check(rid_ctx_create(device, 0, 1, nullptr, &c));
    rid_check(rid_ctx_set_interrupt_poll(c, rid_poll_user_interrupt, nullptr));   // long calls stay interruptible
    return CtxPtr(c, true);
}

// dist.gen(x, method): x is cells x genes (what dist.gen receives, R/RaceID_utils.R:9); cor() sees t(x).
// [[Rcpp::export]]
SEXP rid_dist_gen(SEXP ctx, NumericMatrix x, std::string method) {
    int metric = method == "pearson" ? RID_PEARSON : method == "spearman" ? RID_SPEARMAN
               : method == "logpearson" ? RID_LOGPEARSON : method == "euclidean" ? RID_EUCLIDEAN : -1;
    if (metric < 0) Rcpp::stop("metric has to be one of the following: spearman, pearson, logpearson, euclidean, kendall");
    NumericMatrix tx = transpose(x);                 // genes x cells, column-major: one column per cell
    rid_dmat* dm = nullptr; int nzero = 0;
    rid_check(rid_dist(CtxPtr(ctx), tx.begin(), tx.nrow(), tx.ncol(), metric, 0, &dm, &nzero));
    if (nzero) Rcpp::warning("the standard deviation is zero");
    DmatPtr p(dm, true);
    p.attr("names") = rownames(x);                   // cell names travel with the handle
    return p;
}

// [[Rcpp::export]]
SEXP rid_adopt_matrix(SEXP ctx, NumericMatrix D) {   // clustexp() on a plain @distances matrix
    rid_dmat* dm = nullptr;
    rid_check(rid_dmat_from_host(CtxPtr(ctx), D.begin(), D.nrow(), &dm));
    return DmatPtr(dm, true);
}

// as.matrix(handle) / handle[i, j]
// [[Rcpp::export]]
NumericMatrix rid_as_matrix(SEXP dm, Nullable<IntegerVector> i = R_NilValue, Nullable<IntegerVector> j = R_NilValue) {
    int N = 0; rid_check(rid_dmat_info(DmatPtr(dm), &N, nullptr, nullptr, nullptr, nullptr));
    IntegerVector ri, ci; const int *rp = nullptr, *cp = nullptr; int m = N, n = N;
    if (i.isNotNull()) { ri = IntegerVector(i) - 1; rp = ri.begin(); m = ri.size(); }
    if (j.isNotNull()) { ci = IntegerVector(j) - 1; cp = ci.begin(); n = ci.size(); }
    NumericMatrix out(m, n);
    rid_check(rid_dmat_sub(DmatPtr(dm), rp, m, cp, n, out.begin()));
    return out;
}

// cluster::pam(as.dist(x[n,n]), k) -> list(id.med, clustering, objective, avg.width)
// [[Rcpp::export]]
List rid_pam_call(SEXP dm, int k, Nullable<IntegerVector> subset = R_NilValue) {
    int N = 0; rid_check(rid_dmat_info(DmatPtr(dm), &N, nullptr, nullptr, nullptr, nullptr));
    IntegerVector s; const int* sp = nullptr; int m = N;
    if (subset.isNotNull()) { s = IntegerVector(subset) - 1; sp = s.begin(); m = s.size(); }
    IntegerVector med(k), lab(m); NumericVector obj(2); double sil = 0; int nsw = 0;
    rid_check(rid_pam(DmatPtr(dm), sp, m, k, med.begin(), lab.begin(), obj.begin(), &sil, &nsw));
    obj.names() = CharacterVector::create("build", "swap");
    return List::create(_["id.med"] = med + 1, _["clustering"] = lab, _["objective"] = obj,
                        _["silinfo"] = List::create(_["avg.width"] = sil));
}

// clusGapExt(diM[n,n], FUN = pam, K.max, random = FALSE, diss = TRUE)$Tab[, "logW"]
// [[Rcpp::export]]
NumericVector rid_gap_logW(SEXP dm, int Kmax, Nullable<IntegerVector> subset = R_NilValue) {
    int N = 0; rid_check(rid_dmat_info(DmatPtr(dm), &N, nullptr, nullptr, nullptr, nullptr));
    IntegerVector s; const int* sp = nullptr; int m = N;
    if (subset.isNotNull()) { s = IntegerVector(subset) - 1; sp = s.begin(); m = s.size(); }
    NumericVector logW(Kmax);
    rid_check(rid_gap_sweep(DmatPtr(dm), sp, m, Kmax, logW.begin()));
    return logW;
}

// fpc::clusterboot(diM, B, clustermethod = pamkdCBI, k): `samples` is the list of unique(sample(n,n,TRUE)) vectors
// drawn in R after set.seed(seed) — the library draws no random numbers.
// [[Rcpp::export]]
List rid_clusterboot_call(SEXP dm, int k, List samples) {
    int N = 0; rid_check(rid_dmat_info(DmatPtr(dm), &N, nullptr, nullptr, nullptr, nullptr));
    int B = samples.size(); std::vector<int> cat, len(B);
    for (int b = 0; b < B; ++b) { IntegerVector v = samples[b]; len[b] = v.size(); for (int t : v) cat.push_back(t - 1); }
    IntegerVector part(N), med(k); NumericMatrix res(k, B);
    rid_check(rid_clusterboot(DmatPtr(dm), k, cat.data(), len.data(), B, part.begin(), med.begin(), res.begin()));
    return List::create(_["partition"] = part, _["id.med"] = med + 1, _["bootresult"] = res);
}

// compmedoids(object, part)
// [[Rcpp::export]]
IntegerVector rid_compmedoids(SEXP dm, IntegerVector part) {
    IntegerVector med(part.size()), ids(part.size()); int nc = 0;
    rid_check(rid_medoids(DmatPtr(dm), part.begin(), part.size(), med.begin(), ids.begin(), &nc));
    IntegerVector out(nc); for (int c = 0; c < nc; ++c) out[c] = med[c] + 1;
    return out;
}

// the final-partition refresh of findoutliers (R/RaceID.R:1049-1054)
// [[Rcpp::export]]
IntegerVector rid_refresh_partition(SEXP dm, IntegerVector medoids) {
    int N = 0; rid_check(rid_dmat_info(DmatPtr(dm), &N, nullptr, nullptr, nullptr, nullptr));
    IntegerVector md = medoids - 1, lab(N);
    rid_check(rid_refresh(DmatPtr(dm), md.begin(), md.size(), lab.begin()));
    return lab;
}

// apply(D, 1, function(x) head(order(x), K)) — compdist's knn branch (R/RaceID.R:774), compfr (:1123), pruneKnn, StemID
// [[Rcpp::export]]
IntegerMatrix rid_knn_call(SEXP dm, int K) {
    int N = 0; rid_check(rid_dmat_info(DmatPtr(dm), &N, nullptr, nullptr, nullptr, nullptr));
    IntegerMatrix idx(K, N);                         // K x N, like the matrix apply() returns
    rid_check(rid_knn(DmatPtr(dm), K, idx.begin(), nullptr));
    for (int t = 0; t < K * N; ++t) idx[t] += 1;     // R indices are 1-based
    return idx;
}

// compdist straight from the filtered dgCMatrix (slots i, p, x of ndata[, kept cells]); getfdata is fused on the device
// [[Rcpp::export]]
SEXP rid_dist_csc_call(SEXP ctx, IntegerVector i, IntegerVector p, NumericVector x, int n_rows_all,
                       IntegerVector feature_rows, NumericVector counts, std::string method) {
    int metric = method == "pearson" ? RID_PEARSON : method == "spearman" ? RID_SPEARMAN
               : method == "logpearson" ? RID_LOGPEARSON : method == "euclidean" ? RID_EUCLIDEAN : -1;
    if (metric < 0) Rcpp::stop("metric has to be one of the following: spearman, pearson, logpearson, euclidean, kendall");
    IntegerVector fr = feature_rows - 1;
    rid_dmat* dm = nullptr; int nzero = 0;
    rid_check(rid_dist_csc(CtxPtr(ctx), i.begin(), p.begin(), x.begin(), n_rows_all, counts.size(), fr.begin(), fr.size(),
                           counts.begin(), metric, &dm, &nzero));
    if (nzero) Rcpp::warning("the standard deviation is zero");
    return DmatPtr(dm, true);
}

// silhouette(kpart, as.dist(object@distances)) of plotsilhouette (R/RaceID.R:1283-1284): an n x 3 matrix with the
// columns cluster::silhouette returns (cluster, neighbor, sil_width)
// [[Rcpp::export]]
NumericMatrix rid_silhouette_call(SEXP dm, IntegerVector part) {
    const int N = part.size();
    IntegerVector nb(N);
    NumericVector sw(N);
    rid_check(rid_silhouette(DmatPtr(dm), part.begin(), N, nb.begin(), sw.begin()));
    NumericMatrix out(N, 3);
    for (int i = 0; i < N; ++i) { out(i, 0) = part[i]; out(i, 1) = nb[i]; out(i, 2) = sw[i]; }
    colnames(out) = CharacterVector::create("cluster", "neighbor", "sil_width");
    return out;
}

// cmdscale(di, k = 2) of comptsne (R/RaceID.R:1093): the t-SNE start configuration, an n x k matrix
// [[Rcpp::export]]
NumericMatrix rid_cmdscale_call(SEXP dm, int k) {
    int N = 0; rid_check(rid_dmat_info(DmatPtr(dm), &N, nullptr, nullptr, nullptr, nullptr));
    NumericMatrix pts(N, k);                         // column-major, like cmdscale's result
    NumericVector ev(k);
    rid_check(rid_cmdscale(DmatPtr(dm), k, 1e-9, 1000, pts.begin(), ev.begin(), nullptr));
    int pos = 0; for (int t = 0; t < k; ++t) pos += ev[t] > 0;
    if (pos < k) Rcpp::warning("only %d of the first %d eigenvalues are > 0", pos, k);
    return pts;
}
