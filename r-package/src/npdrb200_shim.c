/*
 * npdrb200_shim.c -- thin .Call shim between R and libnpdr_b200.so (include/npdr_b200.h).
 *
 * NOT compiled against real R headers in the build container (tests/test_host.py type-checks it against a stub of the R
 * API); it is the binding a maintainer adds to the reference package.  Plain R C API (no Rcpp dependency needed for ten
 * entry points).  Rules followed: inputs (SEXP payloads) are never written; native buffers are copied into R-allocated
 * vectors and freed with npdrb_free(); every native buffer that is alive while R allocates (Rf_allocVector can longjmp) is
 * owned by a guard -- an external pointer with a C finalizer -- so an R-level error or interrupt cannot leak it.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdlib.h>
#include <string.h>
#include "npdr_b200.h"

static npdrb_ctx *g_ctx = NULL;

/* native buffers held across R allocations */
typedef struct { void *p[3]; npdrb_result res; int has_res; } nb_guard;
static void guard_release(SEXP e) {
    nb_guard *g = (nb_guard *)R_ExternalPtrAddr(e);
    if (!g) return;
    for (int i = 0; i < 3; ++i) npdrb_free(g->p[i]);
    if (g->has_res) npdrb_result_release(&g->res);
    free(g);
    R_ClearExternalPtr(e);
}
/* returns a PROTECTed external pointer; *out is zero-initialised */
static SEXP guard_new(nb_guard **out) {
    nb_guard *g = (nb_guard *)calloc(1, sizeof(nb_guard));
    if (!g) Rf_error("npdr_b200: out of memory");
    SEXP e = PROTECT(R_MakeExternalPtr(g, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(e, guard_release, TRUE);
    *out = g;
    return e;
}

static void opts_from_r(npdrb_opts *o, SEXP iopts, SEXP sd_frac) {
    npdrb_opts_default(o);
    const int *io = INTEGER(iopts);     /* regression, attr diff, nbd method, nbd metric, knn, n_covars, unique, hit/miss, covar types... */
    if (XLENGTH(iopts) < 8 || io[5] < 0 || io[5] > NPDRB_MAX_COVARS || XLENGTH(iopts) < 8 + io[5])
        Rf_error("npdr_b200: malformed option vector (0..%d covariates)", NPDRB_MAX_COVARS);
    o->regression_type = io[0]; o->attr_diff_type = io[1]; o->nbd_method = io[2]; o->nbd_metric = io[3];
    o->knn = io[4]; o->n_covars = io[5]; o->neighbor_sampling_unique = io[6]; o->separate_hitmiss_nbds = io[7];
    for (int c = 0; c < o->n_covars; ++c) o->covar_diff_type[c] = io[8 + c];
    if (!Rf_isNull(sd_frac)) o->msurf_sd_frac = Rf_asReal(sd_frac);
}

static npdrb_ctx *ctx_get(void) {
    if (!g_ctx && npdrb_create(0, &g_ctx) != NPDRB_OK) Rf_error("npdr_b200: %s", npdrb_last_error());
    return g_ctx;
}

/* .Call("C_npdrb_distances", X, metric) -> N x N matrix           (npdrDistances, R/nearestNeighbors.R:63) */
SEXP C_npdrb_distances(SEXP X, SEXP metric) {
    SEXP dim = Rf_getAttrib(X, R_DimSymbol);
    const R_xlen_t N = INTEGER(dim)[0], p = INTEGER(dim)[1];
    SEXP D = PROTECT(Rf_allocMatrix(REALSXP, (int)N, (int)N));
    int rc = npdrb_distances(ctx_get(), REAL(X), N, p, Rf_asInteger(metric), 0, REAL(D));
    UNPROTECT(1);
    if (rc) Rf_error("npdr_b200: %s", npdrb_last_error());
    return D;
}

/* .Call("C_npdrb_nearest_neighbors", X, nbd.method, nbd.metric, sd.vec|NULL, sd.frac, k, unique)
 *   -> list(Ri_idx, NN_idx, n_unique, n_empty, radius)             (nearestNeighbors, R/nearestNeighbors.R:153) */
SEXP C_npdrb_nearest_neighbors(SEXP X, SEXP method, SEXP metric, SEXP sd_vec, SEXP sd_frac, SEXP k, SEXP uniq) {
    SEXP dim = Rf_getAttrib(X, R_DimSymbol);
    const R_xlen_t N = INTEGER(dim)[0], p = INTEGER(dim)[1];
    int64_t M = 0, nu = 0, ne = 0;
    nb_guard *g;
    SEXP ge = guard_new(&g);
    SEXP radius = PROTECT(Rf_allocVector(REALSXP, N));
    int rc = npdrb_nearest_neighbors(ctx_get(), REAL(X), N, p, Rf_asInteger(method), Rf_asInteger(metric),
                                     Rf_isNull(sd_vec) ? NULL : REAL(sd_vec), Rf_asReal(sd_frac), (int64_t)Rf_asReal(k),
                                     Rf_asLogical(uniq), (int32_t **)&g->p[0], (int32_t **)&g->p[1], &M, &nu, &ne, REAL(radius));
    if (rc) { guard_release(ge); UNPROTECT(2); Rf_error("npdr_b200: %s", npdrb_last_error()); }
    SEXP ri = PROTECT(Rf_allocVector(INTSXP, M)), nn = PROTECT(Rf_allocVector(INTSXP, M));
    if (M) { memcpy(INTEGER(ri), g->p[0], sizeof(int) * (size_t)M); memcpy(INTEGER(nn), g->p[1], sizeof(int) * (size_t)M); }
    guard_release(ge);
    const char *names[] = {"Ri_idx", "NN_idx", "n_unique", "n_empty", "radius", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, ri); SET_VECTOR_ELT(out, 1, nn);
    SET_VECTOR_ELT(out, 2, Rf_ScalarReal((double)nu)); SET_VECTOR_ELT(out, 3, Rf_ScalarReal((double)ne));
    SET_VECTOR_ELT(out, 4, radius);
    UNPROTECT(5);
    return out;
}

/* .Call("C_npdrb_npdr", X, y, C|NULL, opts(integer vector), sd.frac, n.gpus) -> list(stats = p x 8 (row-major -> t()), counters)
 * the whole path between argument parsing and order()/p.adjust()    (R/npdr.R:256-430) */
SEXP C_npdrb_npdr(SEXP X, SEXP y, SEXP C, SEXP iopts, SEXP sd_frac, SEXP n_gpus) {
    SEXP dim = Rf_getAttrib(X, R_DimSymbol);
    const R_xlen_t N = INTEGER(dim)[0], p = INTEGER(dim)[1];
    npdrb_opts o;
    opts_from_r(&o, iopts, sd_frac);
    const int ng = Rf_asInteger(n_gpus);            /* options(npdr_b200.gpus = 8): one host thread per device inside the call */
    SEXP stats = PROTECT(Rf_allocMatrix(REALSXP, NPDRB_STATS_PER_ATTR, (int)p));     /* 8 x p, column = attribute */
    const char *names[] = {"stats", "n_pairs", "n_unique_pairs", "n_pairs_used", "k_ave_empirical", "n_empty", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    nb_guard *g;
    SEXP ge = guard_new(&g);
    int rc = ng > 1 ? npdrb_npdr_multi(ng, NULL, REAL(X), N, p, REAL(y), Rf_isNull(C) ? NULL : REAL(C), &o, &g->res)
                    : npdrb_npdr(ctx_get(), REAL(X), N, p, REAL(y), Rf_isNull(C) ? NULL : REAL(C), &o, &g->res);
    if (rc) { guard_release(ge); UNPROTECT(3); Rf_error("npdr_b200: %s", npdrb_last_error()); }
    g->has_res = 1;
    memcpy(REAL(stats), g->res.stats, sizeof(double) * (size_t)p * NPDRB_STATS_PER_ATTR);
    SET_VECTOR_ELT(out, 0, stats);
    SET_VECTOR_ELT(out, 1, Rf_ScalarReal((double)g->res.n_pairs));
    SET_VECTOR_ELT(out, 2, Rf_ScalarReal((double)g->res.n_unique_pairs));
    SET_VECTOR_ELT(out, 3, Rf_ScalarReal((double)g->res.n_pairs_used));
    SET_VECTOR_ELT(out, 4, Rf_ScalarReal(g->res.k_ave_empirical));
    SET_VECTOR_ELT(out, 5, Rf_ScalarReal((double)g->res.n_empty));
    guard_release(ge);
    UNPROTECT(3);
    return out;
}

/* .Call("C_npdrb_regress", X, y, C|NULL, Ri, NN, opts) -> 8 x p      the regression loop over a given pair list
 * (npdr()'s branches whose neighbourhoods do not come from attr.mat itself: rm.attr.from.dist, external.dist;
 *  R/npdr.R:215-264, 391-430) */
SEXP C_npdrb_regress(SEXP X, SEXP y, SEXP C, SEXP Ri, SEXP NN, SEXP iopts) {
    SEXP dim = Rf_getAttrib(X, R_DimSymbol);
    const R_xlen_t N = INTEGER(dim)[0], p = INTEGER(dim)[1], M = XLENGTH(Ri);
    if (XLENGTH(NN) != M) Rf_error("npdr_b200: Ri_idx and NN_idx differ in length");
    npdrb_opts o;
    opts_from_r(&o, iopts, R_NilValue);
    SEXP stats = PROTECT(Rf_allocMatrix(REALSXP, NPDRB_STATS_PER_ATTR, (int)p));
    int rc = npdrb_regress(ctx_get(), REAL(X), N, p, INTEGER(Ri), INTEGER(NN), M, REAL(y), Rf_isNull(C) ? NULL : REAL(C),
                           o.n_covars, o.covar_diff_type, o.attr_diff_type, o.regression_type, REAL(stats));
    UNPROTECT(1);
    if (rc) Rf_error("npdr_b200: %s", npdrb_last_error());
    return stats;
}

/* .Call("C_npdrb_npdr_k_grid", X, y, C|NULL, opts, sd.frac, k.grid) -> list(stats = 8 x p x length(k.grid), n_pairs)
 * vwok's loop over knn in one native call                           (R/adaptive.R:85-111) */
SEXP C_npdrb_npdr_k_grid(SEXP X, SEXP y, SEXP C, SEXP iopts, SEXP sd_frac, SEXP k_grid) {
    SEXP dim = Rf_getAttrib(X, R_DimSymbol);
    const R_xlen_t N = INTEGER(dim)[0], p = INTEGER(dim)[1], nk = XLENGTH(k_grid);
    npdrb_opts o;
    opts_from_r(&o, iopts, sd_frac);
    int64_t *ks = (int64_t *)R_alloc((size_t)(nk > 0 ? nk : 1), sizeof(int64_t));          /* R-managed: freed by R on any exit */
    int64_t *np_ = (int64_t *)R_alloc((size_t)(nk > 0 ? nk : 1), sizeof(int64_t));
    for (R_xlen_t c = 0; c < nk; ++c) ks[c] = (int64_t)REAL(k_grid)[c];
    SEXP d3 = PROTECT(Rf_allocVector(INTSXP, 3));
    INTEGER(d3)[0] = NPDRB_STATS_PER_ATTR; INTEGER(d3)[1] = (int)p; INTEGER(d3)[2] = (int)nk;
    SEXP stats = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)NPDRB_STATS_PER_ATTR * p * nk));
    SEXP npairs = PROTECT(Rf_allocVector(REALSXP, nk));
    int rc = npdrb_npdr_k_grid(ctx_get(), REAL(X), N, p, REAL(y), Rf_isNull(C) ? NULL : REAL(C), &o, ks, (int32_t)nk, REAL(stats), np_);
    if (rc) { UNPROTECT(3); Rf_error("npdr_b200: %s", npdrb_last_error()); }
    Rf_setAttrib(stats, R_DimSymbol, d3);
    for (R_xlen_t c = 0; c < nk; ++c) REAL(npairs)[c] = (double)np_[c];
    const char *names[] = {"stats", "n_pairs", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, stats); SET_VECTOR_ELT(out, 1, npairs);
    UNPROTECT(4);
    return out;
}

/* .Call("C_npdrb_diff", a, b, type, norm.fac)                       (npdrDiff, R/nearestNeighbors.R:15) */
SEXP C_npdrb_diff(SEXP a, SEXP b, SEXP type, SEXP norm_fac) {
    const R_xlen_t n = XLENGTH(a);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
    int rc = npdrb_diff(ctx_get(), REAL(a), REAL(b), n, Rf_asInteger(type), Rf_asReal(norm_fac), REAL(out));
    UNPROTECT(1);
    if (rc) Rf_error("npdr_b200: %s", npdrb_last_error());
    return out;
}

/* .Call("C_npdrb_unireg", X, y, C|NULL, regression.type) -> 8 x p   (uniReg, R/utils.R:66) */
SEXP C_npdrb_unireg(SEXP X, SEXP y, SEXP C, SEXP reg) {
    SEXP dim = Rf_getAttrib(X, R_DimSymbol);
    const R_xlen_t N = INTEGER(dim)[0], p = INTEGER(dim)[1];
    int q = Rf_isNull(C) ? 0 : INTEGER(Rf_getAttrib(C, R_DimSymbol))[1];
    SEXP stats = PROTECT(Rf_allocMatrix(REALSXP, NPDRB_STATS_PER_ATTR, (int)p));
    int rc = npdrb_unireg(ctx_get(), REAL(X), N, p, REAL(y), q ? REAL(C) : NULL, q, Rf_asInteger(reg), REAL(stats));
    UNPROTECT(1);
    if (rc) Rf_error("npdr_b200: %s", npdrb_last_error());
    return stats;
}

/* .Call("C_npdrb_distances2", X1, X2, metric) -> m1 x m2 matrix     (npdrDistances2, R/npdrLearner.R:323) */
SEXP C_npdrb_distances2(SEXP X1, SEXP X2, SEXP metric) {
    const int *d1 = INTEGER(Rf_getAttrib(X1, R_DimSymbol)), *d2 = INTEGER(Rf_getAttrib(X2, R_DimSymbol));
    if (d1[1] != d2[1]) Rf_error("npdr_b200: attr.mat1 and attr.mat2 must have the same number of columns");
    SEXP D = PROTECT(Rf_allocMatrix(REALSXP, d1[0], d2[0]));
    int rc = npdrb_distances2(ctx_get(), REAL(X1), d1[0], REAL(X2), d2[0], d1[1], Rf_asInteger(metric), REAL(D));
    UNPROTECT(1);
    if (rc) Rf_error("npdr_b200: %s", npdrb_last_error());
    return D;
}

/* .Call("C_npdrb_nearest_neighbors2", X1, X2, nbd.method, nbd.metric, sd.vec|NULL, sd.frac, k)
 *   -> list of nrow(X2) integer vectors of X1 row indices           (nearestNeighbors2, R/npdrLearner.R:378) */
SEXP C_npdrb_nearest_neighbors2(SEXP X1, SEXP X2, SEXP method, SEXP metric, SEXP sd_vec, SEXP sd_frac, SEXP k) {
    const int *d1 = INTEGER(Rf_getAttrib(X1, R_DimSymbol)), *d2 = INTEGER(Rf_getAttrib(X2, R_DimSymbol));
    if (d1[1] != d2[1]) Rf_error("npdr_b200: attr.mat1 and attr.mat2 must have the same number of columns");
    int64_t M = 0;
    nb_guard *g;
    SEXP ge = guard_new(&g);
    int rc = npdrb_nearest_neighbors2(ctx_get(), REAL(X1), d1[0], REAL(X2), d2[0], d1[1], Rf_asInteger(method),
                                      Rf_asInteger(metric), Rf_isNull(sd_vec) ? NULL : REAL(sd_vec), Rf_asReal(sd_frac),
                                      (int64_t)Rf_asReal(k), (int64_t **)&g->p[0], (int32_t **)&g->p[1], &M, NULL);
    if (rc) { guard_release(ge); UNPROTECT(1); Rf_error("npdr_b200: %s", npdrb_last_error()); }
    const int64_t *off = (const int64_t *)g->p[0];
    const int32_t *idx = (const int32_t *)g->p[1];
    SEXP out = PROTECT(Rf_allocVector(VECSXP, d2[0]));
    for (int i = 0; i < d2[0]; ++i) {
        const int64_t n = off[i + 1] - off[i];
        SEXP v = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)n));
        if (n) memcpy(INTEGER(v), idx + off[i], sizeof(int) * (size_t)n);
        SET_VECTOR_ELT(out, i, v);
        UNPROTECT(1);
    }
    guard_release(ge);
    UNPROTECT(2);
    return out;
}

/* .Call("C_npdrb_nearest_neighbors_hitmiss", X, pheno, nbd.method, nbd.metric, sd.frac, k, unique)
 *   -> list(Ri_idx, NN_idx, n_unique, radius_hit, radius_miss)      (nearestNeighborsSeparateHitMiss, R/nearestNeighbors.R:318) */
SEXP C_npdrb_nearest_neighbors_hitmiss(SEXP X, SEXP pheno, SEXP method, SEXP metric, SEXP sd_frac, SEXP k, SEXP uniq) {
    SEXP dim = Rf_getAttrib(X, R_DimSymbol);
    const R_xlen_t N = INTEGER(dim)[0], p = INTEGER(dim)[1];
    int64_t M = 0, nu = 0;
    nb_guard *g;
    SEXP ge = guard_new(&g);
    SEXP rh = PROTECT(Rf_allocVector(REALSXP, N)), rm = PROTECT(Rf_allocVector(REALSXP, N));
    int rc = npdrb_nearest_neighbors_hitmiss(ctx_get(), REAL(X), N, p, REAL(pheno), Rf_asInteger(method), Rf_asInteger(metric),
                                             Rf_asReal(sd_frac), (int64_t)Rf_asReal(k), Rf_asLogical(uniq), (int32_t **)&g->p[0],
                                             (int32_t **)&g->p[1], &M, &nu, REAL(rh), REAL(rm));
    if (rc) { guard_release(ge); UNPROTECT(3); Rf_error("npdr_b200: %s", npdrb_last_error()); }
    SEXP ri = PROTECT(Rf_allocVector(INTSXP, M)), nn = PROTECT(Rf_allocVector(INTSXP, M));
    if (M) { memcpy(INTEGER(ri), g->p[0], sizeof(int) * (size_t)M); memcpy(INTEGER(nn), g->p[1], sizeof(int) * (size_t)M); }
    guard_release(ge);
    const char *names[] = {"Ri_idx", "NN_idx", "n_unique", "radius_hit", "radius_miss", ""};
    SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, ri); SET_VECTOR_ELT(out, 1, nn); SET_VECTOR_ELT(out, 2, Rf_ScalarReal((double)nu));
    SET_VECTOR_ELT(out, 3, rh); SET_VECTOR_ELT(out, 4, rm);
    UNPROTECT(6);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"C_npdrb_nearest_neighbors_hitmiss", (DL_FUNC)&C_npdrb_nearest_neighbors_hitmiss, 7},
    {"C_npdrb_distances2", (DL_FUNC)&C_npdrb_distances2, 3},
    {"C_npdrb_nearest_neighbors2", (DL_FUNC)&C_npdrb_nearest_neighbors2, 7},
    {"C_npdrb_distances", (DL_FUNC)&C_npdrb_distances, 2},
    {"C_npdrb_nearest_neighbors", (DL_FUNC)&C_npdrb_nearest_neighbors, 7},
    {"C_npdrb_npdr", (DL_FUNC)&C_npdrb_npdr, 6},
    {"C_npdrb_regress", (DL_FUNC)&C_npdrb_regress, 6},
    {"C_npdrb_npdr_k_grid", (DL_FUNC)&C_npdrb_npdr_k_grid, 6},
    {"C_npdrb_diff", (DL_FUNC)&C_npdrb_diff, 4},
    {"C_npdrb_unireg", (DL_FUNC)&C_npdrb_unireg, 4},
    {NULL, NULL, 0}};

void R_init_npdrb200(DllInfo *dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}

void R_unload_npdrb200(DllInfo *dll) {
    (void)dll;
    if (g_ctx) { npdrb_destroy(g_ctx); g_ctx = NULL; }
    npdrb_multi_shutdown();
}
