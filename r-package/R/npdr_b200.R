# npdr_b200.R -- drop-in replacements for the reference's hot-path functions (same formals, same return shapes).
# Everything between argument parsing and order()/p.adjust() runs in libnpdr_b200.so; pt(), pnorm(), order(),
# p.adjust() and the data-frame assembly stay in R, exactly as in R/npdr.R:55,63,484-508 of insilico/npdr.

.diff_code <- c("numeric-abs" = 0L, "manhattan" = 0L, "numeric-sqr" = 1L, "allele-sharing" = 2L, "match-mismatch" = 3L)
.metric_code <- c("manhattan" = 0L, "euclidean" = 1L, "allele-sharing-manhattan" = 2L,
                  "relief-scaled-manhattan" = 3L, "relief-scaled-euclidean" = 4L, "precomputed" = 5L)
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

# R/nearestNeighbors.R:318-551 (npdr(separate.hitmiss.nbds = TRUE))
nearestNeighborsSeparateHitMiss <- function(attr.mat, pheno.vec, nbd.method = "relieff", nbd.metric = "manhattan",
                                            sd.frac = 0.5, k = 0, neighbor.sampling = "none", att_to_remove = c(),
                                            fast.dist = FALSE, dopar.nn = FALSE) {
  X <- as.matrix(attr.mat)
  if (length(att_to_remove)) X <- X[, setdiff(colnames(X), att_to_remove), drop = FALSE]
  r <- .Call("C_npdrb_nearest_neighbors_hitmiss", unname(X) + 0, as.numeric(as.character(pheno.vec)), .nbd_code[[nbd.method]],
             .metric_code[[nbd.metric]], sd.frac, k, neighbor.sampling == "unique", PACKAGE = "npdrb200")
  data.frame(Ri_idx = r$Ri_idx, NN_idx = r$NN_idx)
}

nearestNeighbors <- function(attr.mat, nbd.method = "multisurf", nbd.metric = "manhattan", sd.vec = NULL,
                             sd.frac = 0.5, k = 0, neighbor.sampling = "none", att_to_remove = c(),
                             fast.dist = FALSE, dopar.nn = FALSE) {
  X <- as.matrix(attr.mat)
  if (nbd.metric != "precomputed" && length(att_to_remove)) X <- X[, setdiff(colnames(X), att_to_remove), drop = FALSE]
  r <- .Call("C_npdrb_nearest_neighbors", unname(X) + 0, .nbd_code[[nbd.method]], .metric_code[[nbd.metric]],
             sd.vec, sd.frac, k, neighbor.sampling == "unique", PACKAGE = "npdrb200")
  data.frame(Ri_idx = r$Ri_idx, NN_idx = r$NN_idx)
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
  io <- c(.reg_code[[regression.type]], .diff_code[[attr.diff.type]], .nbd_code[[nbd.method]], .metric_code[[nbd.metric]],
          as.integer(knn), ncov, as.integer(neighbor.sampling == "unique"), as.integer(separate.hitmiss.nbds),
          .diff_code[covar.diff.type])
  r <- .Call("C_npdrb_npdr", unname(X) + 0, y, C, as.integer(io), msurf.sd.frac,
             as.integer(getOption("npdr_b200.gpus", 1L)), PACKAGE = "npdrb200")   # > 1: all GPUs of the node, driven in-process
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
