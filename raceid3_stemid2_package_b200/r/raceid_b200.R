## raceid_b200.R — R side of the drop-in: replaces the BODIES of dist.gen (R/RaceID_utils.R:9-15), the kmedoids arms of
## clustfun (R/RaceID_utils.R:102-148) and compmedoids (R/RaceID.R:1298-1316).  compdist()/clustexp() keep their
## signatures, validation, seeds and slot packing.  Not executable in this repo's image (no R); see INTEGRATION.md.

.rid <- new.env()
rid_ctx <- function() { if (is.null(.rid$ctx)) .rid$ctx <- rid_context(0L); .rid$ctx }

## S4 handle stored in sc@distances (the slot is "ANY", R/RaceID.R:47): everything callers do to @distances
setClass("RidDist", slots = c(ptr = "externalptr", names = "character"))
setMethod("dim", "RidDist", function(x) rep(length(x@names), 2L))
setMethod("dimnames", "RidDist", function(x) list(x@names, x@names))
as.matrix.RidDist <- function(x, ...) { m <- rid_as_matrix(x@ptr); dimnames(m) <- list(x@names, x@names); m }
setMethod("[", "RidDist", function(x, i, j, ..., drop = TRUE) {
  ii <- if (missing(i)) NULL else if (is.character(i)) match(i, x@names) else seq_along(x@names)[i]
  jj <- if (missing(j)) NULL else if (is.character(j)) match(j, x@names) else seq_along(x@names)[j]
  m <- rid_as_matrix(x@ptr, ii, jj)
  dimnames(m) <- list(if (is.null(ii)) x@names else x@names[ii], if (is.null(jj)) x@names else x@names[jj])
  if (drop) drop(m) else m
})
as.dist.RidDist <- function(m, ...) stats::as.dist(as.matrix(m), ...)   # comptsne / plotsilhouette (R/RaceID.R:1092,1283)

dist.gen <- function(x, method = "euclidean", ...) {
  if (method == "kendall") return(1 - cor(t(x), method = method))       # not on the accelerated path
  new("RidDist", ptr = rid_dist_gen(rid_ctx(), as.matrix(x), method), names = rownames(x))
}

clustfun <- function(diM, clustnr = 20, bootnr = 50, samp = NULL, sat = TRUE, cln = NULL, rseed = 17000,
                     FUNcluster = "kmedoids", verbose = TRUE) {
  if (FUNcluster != "kmedoids") return(clustfun_reference(diM, clustnr, bootnr, samp, sat, cln, rseed, FUNcluster, verbose))
  if (clustnr < 2) stop("Choose clustnr > 1")
  if (!is(diM, "RidDist")) diM <- new("RidDist", ptr = rid_adopt_matrix(rid_ctx(), as.matrix(diM)), names = colnames(diM))
  N <- ncol(diM); if (is.null(cln)) cln <- 0
  if (N - 1 < clustnr) clustnr <- N - 1
  gpr <- NULL; f <- cln == 0
  if (sat) {
    set.seed(rseed)
    n <- if (!is.null(samp)) sample(1:N, min(samp, N)) else NULL
    logW <- rid_gap_logW(diM@ptr, clustnr, n)
    gpr <- structure(class = "clusGap", list(Tab = cbind(logW = logW, E.logW = NA, gap = NA, SE.sim = NA),
                                             n = if (is.null(n)) N else length(n), B = 100L))
    y <- logW[-length(logW)] - logW[-1]
    mm <- nn <- numeric(length(y))
    for (i in seq_along(y)) { mm[i] <- mean(y[i:length(y)]); nn[i] <- sqrt(var(y[i:length(y)])) }
    if (f) cln <- max(min(which(y - (mm + nn) < 0)), 1)
  }
  if (cln <= 1) {
    clb <- list(result = list(partition = rep(1, N)), bootmean = 1); names(clb$result$partition) <- colnames(diM)
    return(list(clb = clb, gpr = gpr))
  }
  set.seed(rseed)                                       # what fpc::clusterboot(seed = rseed) does
  samples <- lapply(seq_len(bootnr), function(b)
    if (is.null(samp)) unique(sample(N, N, replace = TRUE)) else sample(N, min(N, samp), replace = FALSE))
  r <- rid_clusterboot_call(diM@ptr, cln, samples)
  part <- r$partition; names(part) <- colnames(diM)
  clb <- structure(class = "clboot", list(result = list(partition = part, nc = cln, result = list(pamobject = list(id.med = r$id.med, clustering = part))),
                                          partition = part, nc = cln, B = bootnr))
  key <- if (is.null(samp)) "boot" else "subset"
  clb[[paste0(key, "result")]] <- r$bootresult
  clb[[paste0(key, "mean")]] <- rowMeans(r$bootresult)
  clb[[paste0(key, "brd")]] <- rowSums(r$bootresult <= 0.5)
  clb[[paste0(key, "recover")]] <- rowSums(r$bootresult >= 0.75)
  list(clb = clb, gpr = gpr)
}

compmedoids <- function(object, part) {
  if (is.null(object@distances)) return(compmedoids_reference(object, part))
  d <- object@distances
  if (!is(d, "RidDist")) d <- new("RidDist", ptr = rid_adopt_matrix(rid_ctx(), as.matrix(d)), names = colnames(d))
  full <- integer(length(d@names)); full[match(names(part), d@names)] <- as.integer(part)
  d@names[rid_compmedoids(d@ptr, full)]
}

## plotsilhouette (R/RaceID.R:1270-1285): the silhouette object is built from rid_silhouette()'s n x 3 matrix instead
## of cluster::silhouette(kpart, as.dist(object@distances)) on an n x n host matrix; plot.silhouette() is unchanged.
plotsilhouette <- function(object, final = FALSE) {
  if (length(object@cluster$kpart) == 0) stop("run clustexp before plotsilhouette")
  if (length(unique(object@cluster$kpart)) < 2) stop("only a single cluster: no silhouette plot")
  kpart <- if (final) object@cpart else object@cluster$kpart
  if (is(object@distances, "RidDist")) {
    si <- rid_silhouette_call(object@distances@ptr, as.integer(kpart))
    rownames(si) <- names(kpart)
    attr(si, "Ordered") <- FALSE
    attr(si, "call") <- quote(silhouette(x = kpart, dist = distances))
    class(si) <- "silhouette"
  } else {
    si <- cluster::silhouette(kpart, as.dist(object@distances))
  }
  plot(si)
}

## comptsne (R/RaceID.R:1086-1098): only the start configuration changes; Rtsne still receives a dist object, which
## as.dist.RidDist materialises on demand (Rtsne's own O(n^2) input is outside this library's scope).
rid_tsne_initial_config <- function(object, k = 2L) {
  if (is(object@distances, "RidDist")) rid_cmdscale_call(object@distances@ptr, as.integer(k))
  else stats::cmdscale(as.dist(object@distances), k = k)
}
