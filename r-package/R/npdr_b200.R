# npdr_b200.R -- drop-in replacements for the reference's hot-path functions (same formals, same return shapes).
# Everything between argument parsing and order()/p.adjust() runs in libnpdr_b200.so; pt(), pnorm(), order(),
# p.adjust() and the data-frame assembly stay in R, exactly as in R/npdr.R:55,63,484-508 of insilico/npdr.

.diff_code <- c("numeric-abs" = 0L, "manhattan" = 0L, "numeric-sqr" = 1L, "allele-sharing" = 2L, "match-mismatch" = 3L)
.metric_code <- c("manhattan" = 0L, "euclidean" = 1L, "allele-sharing-manhattan" = 2L,
                  "relief-scaled-manhattan" = 3L, "relief-scaled-euclidean" = 4L, "precomputed" = 5L,
                  "hamming" = 6L)                              # hamming.binary, R/utils.R:185-188
.nbd_code <- c("multisurf" = 0L, "surf" = 1L, "relieff" = 2L)
.reg_code <- c("lm" = 0L, "binomial" = 1L)

npdrDistances <- function(attr.mat, metric = "manhattan", fast.dist = FALSE) {
  if (!metric %in% names(.metric_code)) return(NULL)           # the reference falls through to NULL
  .Call("C_npdrb_distances", unname(as.matrix(attr.mat)) + 0, .metric_code[[metric]], PACKAGE = "npdrb200")
}

# R/npdrLearner.R:323-339 (flexclust::dist2): any metric other than euclidean / allele-sharing-manhattan means manhattan
npdrDistances2 <- function(attr.mat1, attr.mat2, metric = "manhattan") {
  m <- if (metric %in% c("euclidean", "allele-sharing-manhattan")) metric else "manhattan"
  .Call("C_npdrb_distances2", unname(as.matrix(attr.mat1)) + 0, unname(as.matrix(attr.mat2)) + 0, .metric_code[[m]],
        PACKAGE = "npdrb200")
}

# R/npdrLearner.R:378-482: list (one element per row of attr.mat2) of neighbour row indices in attr.mat1
nearestNeighbors2 <- function(attr.mat1, attr.mat2, nbd.method = "multisurf", nbd.metric = "manhattan", sd.vec = NULL,
                              sd.frac = 0.5, dopar.nn = FALSE, k = 0) {
  m <- if (nbd.metric %in% c("euclidean", "allele-sharing-manhattan")) nbd.metric else "manhattan"
  .Call("C_npdrb_nearest_neighbors2", unname(as.matrix(attr.mat1)) + 0, unname(as.matrix(attr.mat2)) + 0,
        .nbd_code[[nbd.method]], .metric_code[[m]], sd.vec, sd.frac, k, PACKAGE = "npdrb200")
}

# R/nearestNeighbors.R:15-40
npdrDiff <- function(a, b, diff.type = "manhattan", norm.fac = 1) {
  diff.type <- match.arg(diff.type, c("manhattan", "numeric-abs", "numeric-sqr", "allele-sharing", "match-mismatch", "correlation-data"))
  if (diff.type == "correlation-data") stop("npdr_b200: correlation-data is not accelerated")
  .Call("C_npdrb_diff", as.numeric(a), as.numeric(b), .diff_code[[diff.type]], norm.fac, PACKAGE = "npdrb200")
}

# R/utils.R:12-18
knnSURF <- function(m.samples, sd.frac = .5) {
  erf <- function(x) 2 * pnorm(x * sqrt(2)) - 1
  floor((m.samples - 1) * (1 - erf(sd.frac / sqrt(2))) / 2)
}

# R/utils.R:66-157: per-column lm / glm of the outcome on the raw attribute (+ covariates); the rounding, p.adjust and sort stay in R
uniReg <- function(outcome, dataset, regression.type = "lm", padj.method = "fdr", covars = "none") {
  if (length(outcome) == 1) {
    pheno.vec <- dataset[, outcome]
    attr.mat <- dataset[, setdiff(if (is.character(outcome)) colnames(dataset) else seq_len(ncol(dataset)), outcome), drop = FALSE]
  } else {
    pheno.vec <- outcome
    attr.mat <- dataset
  }
  y <- if (regression.type == "lm") as.numeric(pheno.vec) else as.numeric(as.factor(pheno.vec)) - 1
  C <- if (length(covars) > 1) as.matrix(covars) + 0 else NULL
  s <- t(.Call("C_npdrb_unireg", unname(as.matrix(attr.mat)) + 0, y, C, .reg_code[[regression.type]], PACKAGE = "npdrb200"))
  pval <- if (regression.type == "lm") 2 * pt(-abs(s[, 2]), s[, 5]) else 2 * pnorm(-abs(s[, 2]))
  padj <- p.adjust(pval, method = padj.method)
  betas <- cbind(beta = s[, 1], beta.Z.att = s[, 2], pval = pval, p.adj = padj)
  betas <- apply(betas, 2, function(v) as.numeric(format(v, digits = 5)))       # R/utils.R:143-146
  row.names(betas) <- colnames(attr.mat)
  betas[order(betas[, "pval"], decreasing = FALSE), , drop = FALSE]
}

# shared by nearestNeighbors(), nearestNeighborsSeparateHitMiss() and npdr()'s two-stage branches: the shim's raw list
# (Ri_idx, NN_idx, n_unique, ...).  att_to_remove is applied here as R/nearestNeighbors.R:179-187 does (never for a
# precomputed matrix, R/nearestNeighbors.R:171-176).
.neighbors_raw <- function(attr.mat, pheno.vec = NULL, nbd.method, nbd.metric, sd.vec = NULL, sd.frac = 0.5, k = 0,
                           unique = FALSE, att_to_remove = c()) {
  if (!nbd.metric %in% names(.metric_code)) stop("npdr_b200: unknown nbd.metric '", nbd.metric, "'")
  if (!nbd.method %in% names(.nbd_code)) stop("npdr_b200: unknown nbd.method '", nbd.method, "'")
  X <- as.matrix(attr.mat)
  if (nbd.metric == "precomputed") {
    if (nrow(X) != ncol(X)) stop("npdr_b200: nbd.metric = 'precomputed' needs the square distance matrix")
  } else if (length(att_to_remove)) {
    X <- X[, setdiff(colnames(X), att_to_remove), drop = FALSE]
  }
  if (is.null(pheno.vec)) {
    .Call("C_npdrb_nearest_neighbors", unname(X) + 0, .nbd_code[[nbd.method]], .metric_code[[nbd.metric]],
          sd.vec, sd.frac, k, unique, PACKAGE = "npdrb200")
  } else {
    .Call("C_npdrb_nearest_neighbors_hitmiss", unname(X) + 0, as.numeric(as.factor(pheno.vec)), .nbd_code[[nbd.method]],
          .metric_code[[nbd.metric]], sd.frac, k, unique, PACKAGE = "npdrb200")
  }
}

# R/nearestNeighbors.R:318-551 (npdr(separate.hitmiss.nbds = TRUE))
nearestNeighborsSeparateHitMiss <- function(attr.mat, pheno.vec, nbd.method = "relieff", nbd.metric = "manhattan",
                                            sd.frac = 0.5, k = 0, neighbor.sampling = "none", att_to_remove = c(),
                                            fast.dist = FALSE, dopar.nn = FALSE) {
  r <- .neighbors_raw(attr.mat, pheno.vec, nbd.method, nbd.metric, NULL, sd.frac, k, neighbor.sampling == "unique", att_to_remove)
  data.frame(Ri_idx = r$Ri_idx, NN_idx = r$NN_idx)
}

nearestNeighbors <- function(attr.mat, nbd.method = "multisurf", nbd.metric = "manhattan", sd.vec = NULL,
                             sd.frac = 0.5, k = 0, neighbor.sampling = "none", att_to_remove = c(),
                             fast.dist = FALSE, dopar.nn = FALSE) {
  r <- .neighbors_raw(attr.mat, NULL, nbd.method, nbd.metric, sd.vec, sd.frac, k, neighbor.sampling == "unique", att_to_remove)
  data.frame(Ri_idx = r$Ri_idx, NN_idx = r$NN_idx)
}

# integer option vector of the shim: regression, attr diff, nbd method, nbd metric, knn, n_covars, unique, hit/miss, covar types
.npdr_io <- function(regression.type, attr.diff.type, nbd.method, nbd.metric, knn, ncov, neighbor.sampling,
                     separate.hitmiss.nbds, covar.diff.type) {
  for (t in covar.diff.type) if (!t %in% names(.diff_code)) stop("npdr_b200: unknown covar.diff.type '", t, "'")
  as.integer(c(.reg_code[[regression.type]], .diff_code[[attr.diff.type]], .nbd_code[[nbd.method]], .metric_code[[nbd.metric]],
               as.integer(knn), ncov, as.integer(neighbor.sampling == "unique"), as.integer(separate.hitmiss.nbds),
               if (ncov > 0) .diff_code[covar.diff.type[seq_len(ncov)]] else integer(0)))
}

npdr <- function(outcome, dataset, regression.type = "binomial", attr.diff.type = "numeric-abs",
                 nbd.method = "multisurf", nbd.metric = "manhattan", knn = 0, msurf.sd.frac = 0.5,
                 covars = "none", covar.diff.type = "match-mismatch", padj.method = "bonferroni", verbose = FALSE,
                 use.glmnet = FALSE, glmnet.alpha = 1, glmnet.lower = 0, glmnet.lam = "lambda.1se",
                 rm.attr.from.dist = c(), neighbor.sampling = "none", separate.hitmiss.nbds = FALSE,
                 corr.attr.names = NULL, fast.reg = FALSE, fast.dist = FALSE, dopar.nn = FALSE, dopar.reg = FALSE,
                 unique.dof = FALSE, external.dist = NULL) {
  if (use.glmnet) stop("npdr_b200: use.glmnet = TRUE is not accelerated; call the reference npdr()")
  if (attr.diff.type == "correlation-data") stop("npdr_b200: correlation-data is not accelerated")
  if (!regression.type %in% names(.reg_code)) stop("npdr_b200: unknown regression.type '", regression.type, "'")
  if (!attr.diff.type %in% names(.diff_code)) stop("npdr_b200: unknown attr.diff.type '", attr.diff.type, "'")
  if (length(outcome) == 1) {                                   # R/npdr.R:166-173
    pheno.vec <- dataset[, outcome]
    attr.mat <- dataset[, setdiff(if (is.character(outcome)) colnames(dataset) else seq_len(ncol(dataset)), outcome), drop = FALSE]
  } else {
    pheno.vec <- outcome
    attr.mat <- dataset
  }
  X <- as.matrix(attr.mat)
  y <- if (regression.type == "binomial") as.numeric(as.factor(pheno.vec)) else as.numeric(pheno.vec)
  C <- NULL; ncov <- 0L
  if (length(covars) > 1) {                                     # R/npdr.R:315-345 (num.covs = length(covar.diff.type))
    C <- as.matrix(covars); ncov <- length(covar.diff.type); C <- C[, seq_len(ncov), drop = FALSE] + 0
  }
  io <- .npdr_io(regression.type, attr.diff.type, nbd.method, nbd.metric, knn, ncov, neighbor.sampling,
                 separate.hitmiss.nbds, covar.diff.type)
  if (nbd.metric == "precomputed" || length(rm.attr.from.dist) > 0) {
    # R/npdr.R:215-264: the neighbourhoods come from a user distance matrix (external.dist) or from the attributes minus
    # rm.attr.from.dist, the regression from ALL attributes -> two calls: neighbour list, then the regression loop over it
    src <- attr.mat
    if (nbd.metric == "precomputed") {
      if (is.null(external.dist)) stop("npdr_b200: nbd.metric = 'precomputed' needs external.dist (the N x N distance matrix)")
      src <- external.dist
    }
    nn <- .neighbors_raw(src, if (separate.hitmiss.nbds) pheno.vec else NULL, nbd.method, nbd.metric, NULL, msurf.sd.frac, knn,
                         FALSE, rm.attr.from.dist)
    Ri <- nn$Ri_idx; NN <- nn$NN_idx
    n.pairs <- length(Ri); n.unique <- nn$n_unique
    cnt <- tabulate(Ri); k.ave <- mean(cnt[cnt > 0])            # mean(knnVec(neighbor.pairs.idx)), R/npdr.R:267
    if (neighbor.sampling == "unique") {                        # uniqueNeighbors, R/nearestNeighbors.R:572-583
      keep <- !duplicated(cbind(pmin(Ri, NN), pmax(Ri, NN)))
      Ri <- Ri[keep]; NN <- NN[keep]
    }
    stats <- .Call("C_npdrb_regress", unname(X) + 0, y, C, as.integer(Ri), as.integer(NN), io, PACKAGE = "npdrb200")
    r <- list(stats = stats, n_pairs = n.pairs, n_unique_pairs = n.unique, n_pairs_used = length(Ri), k_ave_empirical = k.ave)
  } else {
    r <- .Call("C_npdrb_npdr", unname(X) + 0, y, C, io, msurf.sd.frac,
               as.integer(getOption("npdr_b200.gpus", 1L)), PACKAGE = "npdrb200")   # > 1: all GPUs of the node, driven in-process
  }
  s <- t(r$stats)                                               # p x 8: beta_a, stat_a, beta_0, stat_0, df, r2|dev, iters, flags
  res.df <- if (unique.dof) r$n_unique_pairs - 2 else s[, 5]    # R/npdr.R:419-423
  pval.att <- pt(s[, 2], res.df, lower.tail = FALSE)            # R/npdr.R:55
  pval.0 <- if (regression.type == "lm") 2 * pt(-abs(s[, 4]), s[, 5]) else 2 * pnorm(-abs(s[, 4]))
  if (verbose) {
    cat(r$n_pairs, "total neighbor pairs (possible repeats).\n")
    cat(r$n_unique_pairs, "unique neighbor pairs.\n")
    cat("Empirical (computed from neighborhood) average neighbors: ", r$k_ave_empirical, ".\n", sep = "")
  }
  o <- order(pval.att, decreasing = FALSE)                      # R/npdr.R:486
  out <- data.frame(att = colnames(X)[o], pval.adj = p.adjust(pval.att[o], method = padj.method),
                    pval.att = pval.att[o], beta.raw.att = s[o, 1], beta.Z.att = s[o, 2], beta.0 = s[o, 3],
                    pval.0 = pval.0[o], stringsAsFactors = FALSE)
  if (regression.type == "lm") out$R.sqr <- s[o, 6]
  out
}

# vwok (R/adaptive.R:39-234): npdr(binomial, numeric-abs, relieff, manhattan, knn = k, padj "bonferroni") for every k of the
# grid, then per attribute the k with the largest standardised beta.  ONE native call serves the whole grid
# (npdrb_npdr_k_grid: one upload, one distance matrix, one per-row (d, row) sort).  The PRROC auPRC scoring
# (signal.names) and correlation-data are not accelerated.
vwok <- function(dats, k.grid = NULL, verbose = FALSE, attr.diff.type = "numeric-abs", corr.attr.names = NULL,
                 signal.names = NULL, separate.hitmiss.nbds = FALSE, label = "class") {
  if (attr.diff.type == "correlation-data") stop("npdr_b200: correlation-data is not accelerated")
  if (!is.null(signal.names)) stop("npdr_b200: the auPRC scoring (signal.names) is not accelerated; call the reference vwok()")
  pheno.vec <- dats[, label]
  attr.mat <- dats[, setdiff(colnames(dats), label), drop = FALSE]
  X <- as.matrix(attr.mat); m <- nrow(X)
  ks <- if (is.null(k.grid)) seq.int(m - 1) else as.integer(k.grid)             # R/adaptive.R:73
  io <- .npdr_io("binomial", "numeric-abs", "relieff", "manhattan", 0L, 0L, "none", separate.hitmiss.nbds, character(0))
  r <- .Call("C_npdrb_npdr_k_grid", unname(X) + 0, as.numeric(as.factor(pheno.vec)), NULL, io, 0.5, as.numeric(ks),
             PACKAGE = "npdrb200")                               # stats: 8 x p x length(ks)
  ord <- order(colnames(X))                                     # arrange(att), R/adaptive.R:112-115
  betas <- matrix(0, ncol(X), length(ks)); pvals <- betas
  for (c in seq_along(ks)) {
    s <- t(r$stats[, , c])
    pval.att <- pt(s[, 2], s[, 5], lower.tail = FALSE)
    betas[, c] <- s[ord, 2]
    pvals[, c] <- p.adjust(pval.att, method = "bonferroni")[ord]
  }
  rownames(betas) <- rownames(pvals) <- colnames(X)[ord]
  colnames(betas) <- colnames(pvals) <- paste0("k.", ks)
  best <- apply(betas, 1, which.max)
  out <- data.frame(att = rownames(betas), best.ks = best, betas = betas[cbind(seq_len(nrow(betas)), best)],   # best.ks: the grid INDEX, as apply(betas, 1, which.max) in R/adaptive.R:203
                    pval.att = pvals[cbind(seq_len(nrow(pvals)), best)], stringsAsFactors = FALSE)
  out <- out[order(out$betas, decreasing = TRUE), ]
  list(vwok.out = out, betas = as.data.frame(betas), pvals = as.data.frame(pvals))
}
