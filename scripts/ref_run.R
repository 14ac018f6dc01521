#!/usr/bin/env Rscript
# ref_run.R -- run the UNMODIFIED reference (insilico/npdr) on BASELINE configs 1-3 and dump what the oracle must reproduce.
#
#   Rscript scripts/ref_run.R <path to the reference checkout> <output directory>
#
# Needs R (>= 3.3.2) with dplyr (>= 1.0.0), magrittr, rlang -- the reference's Imports (DESCRIPTION:11-14).  The reference's R
# files are source()d from the checkout (no install, nothing is written into it).  Neither R nor these packages exist in the
# build container or on the GPU box (probed: `Rscript --version`), so tests/test_reference_r.py SKIPS there; on a machine that
# has them the same test runs this script and pins the oracle (distances, radii, neighbour lists, per-attribute statistics,
# attribute ordering) to the reference's own output -- SURVEY.md 8c.
#
# Output (CSV, full double precision via sprintf("%.17g")): for cfg in cfg1 cfg2 cfg3
#   <out>/<cfg>_dist.csv      N x N distance matrix            npdrDistances            R/nearestNeighbors.R:63-101
#   <out>/<cfg>_pairs.csv     Ri_idx, NN_idx                   nearestNeighbors         R/nearestNeighbors.R:153-285
#   <out>/<cfg>_npdr.csv      att, pval.adj, pval.att, beta.raw.att, beta.Z.att, beta.0, pval.0[, R.sqr]   npdr   R/npdr.R:151-511
args <- commandArgs(trailingOnly = TRUE)
if (length(args) < 2) stop("usage: Rscript scripts/ref_run.R <reference dir> <output dir>")
ref <- normalizePath(args[1]); out <- args[2]
dir.create(out, showWarnings = FALSE, recursive = TRUE)
suppressPackageStartupMessages({ library(dplyr); library(magrittr); library(rlang) })
for (f in list.files(file.path(ref, "R"), pattern = "[.][rR]$", full.names = TRUE)) source(f)
load(file.path(ref, "data", "case.control.3sets.rda"))
load(file.path(ref, "data", "qtrait.3sets.rda"))
load(file.path(ref, "data", "mdd.RNAseq.small.rda"))

w <- function(x, path) {
  x <- as.data.frame(x)
  for (j in seq_along(x)) if (is.double(x[[j]])) x[[j]] <- sprintf("%.17g", x[[j]])
  write.csv(x, path, row.names = FALSE, quote = FALSE)
}

run <- function(name, pheno, attr.mat, regression.type, nbd.method, nbd.metric, ...) {
  w(npdrDistances(unname(as.matrix(attr.mat)), metric = nbd.metric), file.path(out, paste0(name, "_dist.csv")))
  w(nearestNeighbors(attr.mat, nbd.method = nbd.method, nbd.metric = nbd.metric, sd.frac = 0.5, k = 0),
    file.path(out, paste0(name, "_pairs.csv")))
  w(npdr(pheno, attr.mat, regression.type = regression.type, attr.diff.type = "numeric-abs", nbd.method = nbd.method,
         nbd.metric = nbd.metric, msurf.sd.frac = 0.5, padj.method = "bonferroni", ...),
    file.path(out, paste0(name, "_npdr.csv")))
}

cc <- case.control.3sets$train
run("cfg1", cc$class, cc[, colnames(cc) != "class"], "binomial", "multisurf", "manhattan")
qt <- qtrait.3sets$train
run("cfg2", qt$qtrait, qt[, colnames(qt) != "qtrait"], "lm", "multisurf", "euclidean")
mdd <- mdd.RNAseq.small
run("cfg3", mdd$phenos, as.data.frame(mdd$rnaSeq), "binomial", "relieff", "manhattan",
    covars = mdd$covs.short[, c("sex", "age")], covar.diff.type = c("match-mismatch", "numeric-abs"))
cat("reference outputs written to", out, "\n")
