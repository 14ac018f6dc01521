/*
 * npdr_oracle.c -- CPU restatement of the reference NPDR hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (npdr_b200/, the CUDA
 * library, bench.py's GPU arm) may call into this file; it exists so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs have something to check against and to time.
 *
 * PARITY STATUS: the reference (insilico/npdr, /root/reference) is a pure-R
 * package; its arithmetic lives in base-R `stats` (dist, sd/var, colSums, lm,
 * glm.fit, summary.*) and dplyr (slice_min, filter), neither vendored nor
 * version-pinned (DESCRIPTION:7,11-14), and R is not installed in the build
 * container.  The only known-answer vectors the reference's own tests hold for
 * this path are the five npdrDiff expectations
 * (tests/testthat/test-nearestNeighbors.R:1-12); those are pinned
 * (tests/test_oracle.py).  The lm.fit / glm.fit / summary.* restatement is
 * additionally pinned to what R itself prints for four models on datasets::mtcars
 * (coefficients, standard errors, R^2, deviance, Fisher scoring iterations;
 * tests/golden/r_published_kats.py).  Distances, radii and neighbour lists (and
 * the regression statistics beyond those known answers) are restated from the
 * published base-R / LINPACK / dplyr algorithms and cross-checked against
 * independent numpy/scipy implementations, but are otherwise **parity unpinned**.
 *
 * Each function cites the reference file:line it follows.  Plain C, `long double`
 * is x87 80-bit on x86-64 Linux = R's LDOUBLE.  Build: see oracle/Makefile
 * (gcc -O2 -ffp-contract=off -fopenmp).
 *
 * Conventions: matrices are column-major doubles (as R holds them); instance
 * indices in pair lists are 1-based (as R's Ri_idx / NN_idx).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_EINVAL 1
#define ORC_ENOMEM 2
#define ORC_ESINGLE_LEVEL 3
#define ORC_MAX_DESIGN 64   /* columns of a regression design: intercept + attribute + covariates (R/npdr.R:404-411) */

/* diff types: npdrDiff's diff.type, R/nearestNeighbors.R:15 */
enum { ORC_DIFF_NUMERIC_ABS = 0, ORC_DIFF_NUMERIC_SQR = 1, ORC_DIFF_ALLELE_SHARING = 2,
       ORC_DIFF_MATCH_MISMATCH = 3, ORC_DIFF_RAW = 4 /* uniReg: the raw attribute, R/utils.R:84 */ };
/* metrics: npdrDistances' metric, R/nearestNeighbors.R:63-101 */
enum { ORC_METRIC_MANHATTAN = 0, ORC_METRIC_EUCLIDEAN = 1, ORC_METRIC_ALLELE_SHARING_MANHATTAN = 2,
       ORC_METRIC_RELIEF_SCALED_MANHATTAN = 3, ORC_METRIC_RELIEF_SCALED_EUCLIDEAN = 4 };
enum { ORC_NBD_MULTISURF = 0, ORC_NBD_SURF = 1, ORC_NBD_RELIEFF = 2 };
enum { ORC_REG_LM = 0, ORC_REG_BINOMIAL = 1 };

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py's reference arm: use n host threads regardless of OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1) */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ npdrDiff */
/* R/nearestNeighbors.R:15-40.  `manhattan` == `numeric-abs`. */
static inline double diff1(double a, double b, int type, double norm_fac) {
    switch (type) {
    case ORC_DIFF_NUMERIC_SQR: { double t = fabs(a - b); return t * t / norm_fac; }   /* abs(a-b)^2/norm.fac */
    case ORC_DIFF_ALLELE_SHARING: return fabs(a - b) / 2.0;
    case ORC_DIFF_MATCH_MISMATCH: return (a == b) ? 0.0 : 1.0;
    case ORC_DIFF_RAW: return a;
    default: return fabs(a - b) / norm_fac;
    }
}

int orc_diff(const double *a, const double *b, int64_t n, int type, double norm_fac, double *out) {
    if (type < 0 || type > ORC_DIFF_RAW) return ORC_EINVAL;
    for (int64_t i = 0; i < n; ++i) out[i] = diff1(a[i], b[i], type, norm_fac);
    return ORC_OK;
}

/* `correlation-data`: rowSums(abs(a - b) / norm.fac) on nrow x ncol matrices, R/nearestNeighbors.R:24.
 * rowSums accumulates in long double (base-R do_colsum). */
int orc_diff_corr(const double *a, const double *b, int64_t nrow, int64_t ncol, double norm_fac, double *out) {
    for (int64_t i = 0; i < nrow; ++i) {
        long double s = 0.0L;
        for (int64_t j = 0; j < ncol; ++j) s += fabs(a[i + j * nrow] - b[i + j * nrow]) / norm_fac;
        out[i] = (double)s;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------- npdrDistances */
/* R/nearestNeighbors.R:63-101 -> stats::dist (C R_distance): for each pair, a plain
 * double accumulator over k = 0..p-1 in order; euclidean = sqrt(sum dev*dev), not
 * FMA-contracted (x86-64 gcc).  as.matrix() mirrors to a full symmetric matrix with
 * zero diagonal.  Pre-scalings: allele-sharing-manhattan = X/2 (:76-82);
 * relief-scaled-* = (x - min)/(max - min) per column (:84-95). */
int orc_distances(const double *X, int64_t N, int64_t p, int metric, double *D) {
    if (metric < 0 || metric > ORC_METRIC_RELIEF_SCALED_EUCLIDEAN) return ORC_EINVAL;
    const double *Z = X;
    double *scaled = NULL;
    int euclid = (metric == ORC_METRIC_EUCLIDEAN || metric == ORC_METRIC_RELIEF_SCALED_EUCLIDEAN);
    if (metric >= ORC_METRIC_ALLELE_SHARING_MANHATTAN) {
        scaled = (double *)malloc(sizeof(double) * (size_t)N * (size_t)p);
        if (!scaled) return ORC_ENOMEM;
        for (int64_t k = 0; k < p; ++k) {
            const double *c = X + k * N;
            double *o = scaled + k * N;
            if (metric == ORC_METRIC_ALLELE_SHARING_MANHATTAN) {
                for (int64_t i = 0; i < N; ++i) o[i] = c[i] / 2.0;
            } else {
                double mn = c[0], mx = c[0];
                for (int64_t i = 1; i < N; ++i) { if (c[i] < mn) mn = c[i]; if (c[i] > mx) mx = c[i]; }
                double range = mx - mn;                  /* attr.range, R/utils.R:198-202 */
                for (int64_t i = 0; i < N; ++i) o[i] = (c[i] - mn) / range;
            }
        }
        Z = scaled;
    }
    /* Row-major copy so the k loop is contiguous; arithmetic order is unchanged. */
    double *T = (double *)malloc(sizeof(double) * (size_t)N * (size_t)p);
    if (!T) { free(scaled); return ORC_ENOMEM; }
    for (int64_t k = 0; k < p; ++k)
        for (int64_t i = 0; i < N; ++i) T[i * p + k] = Z[k * N + i];
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t j = 0; j < N; ++j) {
        D[j * N + j] = 0.0;
        const double *xj = T + j * p;
        for (int64_t i = j + 1; i < N; ++i) {
            const double *xi = T + i * p;
            double dist = 0.0;
            if (euclid) {
                for (int64_t k = 0; k < p; ++k) { double dev = xi[k] - xj[k]; dist += dev * dev; }
                dist = sqrt(dist);
            } else {
                for (int64_t k = 0; k < p; ++k) { double dev = fabs(xi[k] - xj[k]); dist += dev; }
            }
            D[j * N + i] = dist;
            D[i * N + j] = dist;
        }
    }
    free(T);
    free(scaled);
    return ORC_OK;
}

/* hamming.binary, R/utils.R:185-188: D <- t(1 - X) %*% X; D + t(D).  npdrDistances calls it on the m x p
 * attribute matrix as it stands (R/nearestNeighbors.R:72-73), for which that expression is a p x p
 * attribute-by-attribute matrix -- it is an instance distance only for an attributes x instances input.
 * Restated here as the N x N INSTANCE distance the neighbour search needs (the formula applied to t(X)):
 * the number of attributes on which two 0/1 instances differ.  A documented deviation from the reference
 * as written (deviation 8). */
int orc_hamming(const double *X, int64_t N, int64_t p, double *D) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i)
        for (int64_t j = 0; j < N; ++j) {
            double s = 0.0;                                   /* diagonal = 2 sum x(1-x): 0 for 0/1 data, as in R */
            for (int64_t k = 0; k < p; ++k)
                s += (1.0 - X[k * N + i]) * X[k * N + j] + (1.0 - X[k * N + j]) * X[k * N + i];
            D[j * N + i] = s;
        }
    return ORC_OK;
}

/* ------------------------------------------------------------------- radii */
/* base-R var()/sd() on one vector (stats cov.c, cov_na_1 + MEAN_): long double two-pass
 * mean, then long double sum of (x-xm)*(x-xm) with xm the mean rounded to double. */
static double r_var(const double *x, int64_t n, int64_t skip /* index to leave out or -1 */) {
    int64_t nobs = (skip >= 0) ? n - 1 : n;
    long double sum = 0.0L, tmp;
    for (int64_t k = 0; k < n; ++k) if (k != skip) sum += x[k];
    tmp = sum / nobs;
    if (isfinite((double)tmp)) {
        sum = 0.0L;
        for (int64_t k = 0; k < n; ++k) if (k != skip) sum += (x[k] - tmp);
        tmp = tmp + sum / nobs;
    }
    long double xm = (double)tmp;
    sum = 0.0L;
    for (int64_t k = 0; k < n; ++k) if (k != skip) sum += (x[k] - xm) * (x[k] - xm);
    int64_t n1 = nobs - 1;
    return (double)(sum / n1);
}

/* multiSURF radius, R/nearestNeighbors.R:242-247:
 *   sd.vec[x] = sd(dist.mat[-x, x]);  Ri.radius = colSums(dist.mat)/(N-1) - sd.frac*sd.vec
 * colSums: long double sequential sum per column, includes the zero diagonal. */
int orc_radius_multisurf(const double *D, int64_t N, double sd_frac, const double *sd_vec_in,
                         double *radius, double *sd_vec_out) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < N; ++c) {
        const double *col = D + c * N;
        double sdv = sd_vec_in ? sd_vec_in[c] : sqrt(r_var(col, N, c));
        long double s = 0.0L;
        for (int64_t i = 0; i < N; ++i) s += col[i];
        double colsum = (double)s;
        radius[c] = colsum / (double)(N - 1) - sd_frac * sdv;
        if (sd_vec_out) sd_vec_out[c] = sdv;
    }
    return ORC_OK;
}

/* SURF radius, R/nearestNeighbors.R:234-241: one constant,
 *   sum(dist.mat)/(2*num.pair) - sd.frac*sd(dist.mat[upper.tri(dist.mat)])
 * sum(): one long double accumulator over all N*N entries, column-major order.
 * upper.tri extraction is column-major: for j, for i<j: D[i,j]. */
int orc_radius_surf(const double *D, int64_t N, double sd_frac, double *radius) {
    long double s = 0.0L;
    for (int64_t e = 0; e < N * N; ++e) s += D[e];
    double total = (double)s;
    double num_pair = (double)N * (double)(N - 1) / 2.0;
    int64_t nu = N * (N - 1) / 2;
    double *u = (double *)malloc(sizeof(double) * (size_t)(nu > 0 ? nu : 1));
    if (!u) return ORC_ENOMEM;
    int64_t t = 0;
    for (int64_t j = 0; j < N; ++j) for (int64_t i = 0; i < j; ++i) u[t++] = D[j * N + i];
    double sdc = sqrt(r_var(u, nu, -1));
    free(u);
    *radius = total / (2.0 * num_pair) - sd_frac * sdc;
    return ORC_OK;
}

/* knnSURF, R/utils.R:12-18 and the relieff default, R/nearestNeighbors.R:197-203:
 *   erf(x) = 2*pnorm(x*sqrt(2)) - 1;  k = floor((m-1)*(1 - erf(sd.frac/sqrt(2)))/2)
 * pnorm(z) evaluated as 0.5*erfc(-z/sqrt(2)) (libm); R uses Cody's algorithm, equal to
 * within an ulp, which can matter only if the argument of floor() is within ~1e-13 of an integer. */
int64_t orc_knn_surf(int64_t m_samples, double sd_frac) {
    double x = sd_frac / sqrt(2.0);
    double z = x * sqrt(2.0);
    double pn = 0.5 * erfc(-z / sqrt(2.0));
    double erfv = 2.0 * pn - 1.0;
    return (int64_t)floor((double)(m_samples - 1) * (1.0 - erfv) / 2.0);
}

/* --------------------------------------------------------- nearestNeighbors */
/* get_Ri_nearest + assembly, R/nearestNeighbors.R:221-231,262-273,622-634.
 * surf/multisurf: rows j (ascending) with D[j,Ri] < radius[Ri] & D[j,Ri] > 0.
 * relieff: dplyr slice_min(n=k+1, with_ties=TRUE) = rows with min_rank <= k+1, ordered
 * by distance ascending (ties in row order), then filter(> 0).
 * Instances with an empty neighbourhood contribute no rows (the reference errors there,
 * SURVEY 8a quirk 4; the count is returned in n_empty).
 * Output arrays are malloc'ed here; free with orc_free. */
typedef struct { double d; int32_t j; } dj_t;
static int cmp_dj(const void *a, const void *b) {
    const dj_t *x = (const dj_t *)a, *y = (const dj_t *)b;
    if (x->d < y->d) return -1;
    if (x->d > y->d) return 1;
    return (x->j > y->j) - (x->j < y->j);
}

int orc_neighbors(const double *D, int64_t N, int method, int64_t k, const double *radius,
                  int32_t **Ri_out, int32_t **NN_out, int64_t *M_out, int64_t *n_empty) {
    int64_t cap = 1024, M = 0, empty = 0;
    int32_t *Ri = (int32_t *)malloc(sizeof(int32_t) * cap), *NN = (int32_t *)malloc(sizeof(int32_t) * cap);
    dj_t *buf = (dj_t *)malloc(sizeof(dj_t) * (size_t)(N > 0 ? N : 1));
    if (!Ri || !NN || !buf) return ORC_ENOMEM;
    for (int64_t c = 0; c < N; ++c) {
        const double *col = D + c * N;
        int64_t before = M;
        if (method == ORC_NBD_RELIEFF) {
            for (int64_t j = 0; j < N; ++j) { buf[j].d = col[j]; buf[j].j = (int32_t)j; }
            qsort(buf, (size_t)N, sizeof(dj_t), cmp_dj);
            int64_t n = k + 1;
            if (n > N) n = N;
            if (n <= 0) { empty++; continue; }
            double cut = buf[n - 1].d;                     /* (k+1)-th smallest: min_rank <= k+1  <=>  d <= cut */
            for (int64_t t = 0; t < N && buf[t].d <= cut; ++t) {
                if (!(buf[t].d > 0)) continue;
                if (M == cap) { cap *= 2; Ri = realloc(Ri, sizeof(int32_t) * cap); NN = realloc(NN, sizeof(int32_t) * cap); if (!Ri || !NN) return ORC_ENOMEM; }
                Ri[M] = (int32_t)(c + 1); NN[M] = buf[t].j + 1; ++M;
            }
        } else {
            double r = radius[c];
            for (int64_t j = 0; j < N; ++j) {
                if ((col[j] < r) && (col[j] > 0)) {
                    if (M == cap) { cap *= 2; Ri = realloc(Ri, sizeof(int32_t) * cap); NN = realloc(NN, sizeof(int32_t) * cap); if (!Ri || !NN) return ORC_ENOMEM; }
                    Ri[M] = (int32_t)(c + 1); NN[M] = (int32_t)(j + 1); ++M;
                }
            }
        }
        if (M == before) empty++;
    }
    free(buf);
    *Ri_out = Ri; *NN_out = NN; *M_out = M;
    if (n_empty) *n_empty = empty;
    return ORC_OK;
}

/* ------------------------------------------------ npdrDistances2 / nearestNeighbors2 (R/npdrLearner.R:323-482) */
static double r_mean(const double *x, int64_t n);
/* npdrDistances2, R/npdrLearner.R:323-339 -> flexclust::dist2(x, y, method) (third-party, unvendored and unpinned:
 * DESCRIPTION Suggests flexclust).  dist2 is flexclust's two-matrix generalisation of stats::dist and is restated with
 * dist()'s arithmetic: plain double accumulation over k = 0..p-1 in order, euclidean = sqrt(sum dev*dev).
 * metric: MANHATTAN, EUCLIDEAN, or ALLELE_SHARING_MANHATTAN (= both matrices / 2, then manhattan, :330-332).
 * D: m1 x m2 column-major (rows = instances of X1, columns = instances of X2), as the reference returns it. */
int orc_distances2(const double *X1, int64_t m1, const double *X2, int64_t m2, int64_t p, int metric, double *D) {
    if (metric != ORC_METRIC_MANHATTAN && metric != ORC_METRIC_EUCLIDEAN && metric != ORC_METRIC_ALLELE_SHARING_MANHATTAN)
        return ORC_EINVAL;
    const int euclid = metric == ORC_METRIC_EUCLIDEAN;
    const double scale = metric == ORC_METRIC_ALLELE_SHARING_MANHATTAN ? 2.0 : 1.0;
    double *T1 = (double *)malloc(sizeof(double) * (size_t)m1 * (size_t)p), *T2 = (double *)malloc(sizeof(double) * (size_t)m2 * (size_t)p);
    if (!T1 || !T2) { free(T1); free(T2); return ORC_ENOMEM; }
    for (int64_t k = 0; k < p; ++k) {
        for (int64_t i = 0; i < m1; ++i) T1[i * p + k] = X1[k * m1 + i] / scale;
        for (int64_t i = 0; i < m2; ++i) T2[i * p + k] = X2[k * m2 + i] / scale;
    }
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < m2; ++j) {
        const double *xj = T2 + j * p;
        for (int64_t i = 0; i < m1; ++i) {
            const double *xi = T1 + i * p;
            double dist = 0.0;
            if (euclid) { for (int64_t k = 0; k < p; ++k) { double dev = xi[k] - xj[k]; dist += dev * dev; } dist = sqrt(dist); }
            else { for (int64_t k = 0; k < p; ++k) { double dev = fabs(xi[k] - xj[k]); dist += dev; } }
            D[j * m1 + i] = dist;
        }
    }
    free(T1); free(T2);
    return ORC_OK;
}

/* nearestNeighbors2 radii, R/npdrLearner.R:439-450.  D is the m1 x m2 matrix of orc_distances2.
 *  surf:      radius = mean(dist.mat) - sd.frac * sd(dist.mat), replicated m2 times.  mean(): long-double two-pass over all
 *             m1*m2 entries in storage order; sd(<matrix>) = sqrt(var(as.double(x))) -> r_var on the flattened matrix.
 *  multisurf: sd.vec[i] = sd(dist.mat[, i]) (all m1 entries: a test instance has no self distance);
 *             radius = colMeans(dist.mat) - sd.frac * sd.vec; colMeans: long-double sum / m1, rounded once. */
int orc_radius2(const double *D, int64_t m1, int64_t m2, int method, double sd_frac, const double *sd_vec_in, double *radius) {
    if (method == ORC_NBD_SURF) {
        const double r = r_mean(D, m1 * m2) - sd_frac * sqrt(r_var(D, m1 * m2, -1));
        for (int64_t c = 0; c < m2; ++c) radius[c] = r;
        return ORC_OK;
    }
    if (method != ORC_NBD_MULTISURF) return ORC_EINVAL;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < m2; ++c) {
        const double *col = D + c * m1;
        const double sdv = sd_vec_in ? sd_vec_in[c] : sqrt(r_var(col, m1, -1));
        long double s = 0.0L;
        for (int64_t i = 0; i < m1; ++i) s += col[i];
        s /= m1;
        radius[c] = (double)s - sd_frac * sdv;
    }
    return ORC_OK;
}

/* nearestNeighbors2 selection, R/npdrLearner.R:421-433,470-478 -> get_Ri_nearest (R/nearestNeighbors.R:622-634) on
 * column Ri of the m1 x m2 matrix: the same dplyr verbs as nearestNeighbors, but no row is the test instance itself,
 * so relieff's slice_min(n = k + 1) keeps k + 1 training instances (and boundary ties) unless a distance is exactly 0.
 * Returns CSR: offsets[m2 + 1] (caller-allocated), idx malloc'ed here (1-based training indices). */
int orc_neighbors2(const double *D, int64_t m1, int64_t m2, int method, int64_t k, const double *radius,
                   int64_t *offsets, int32_t **idx_out) {
    int64_t cap = 1024, M = 0;
    int32_t *NN = (int32_t *)malloc(sizeof(int32_t) * cap);
    dj_t *buf = (dj_t *)malloc(sizeof(dj_t) * (size_t)(m1 > 0 ? m1 : 1));
    if (!NN || !buf) return ORC_ENOMEM;
    for (int64_t c = 0; c < m2; ++c) {
        const double *col = D + c * m1;
        offsets[c] = M;
        if (method == ORC_NBD_RELIEFF) {
            for (int64_t j = 0; j < m1; ++j) { buf[j].d = col[j]; buf[j].j = (int32_t)j; }
            qsort(buf, (size_t)m1, sizeof(dj_t), cmp_dj);
            int64_t n = k + 1;
            if (n > m1) n = m1;
            if (n <= 0) continue;
            const double cut = buf[n - 1].d;
            for (int64_t t = 0; t < m1 && buf[t].d <= cut; ++t) {
                if (!(buf[t].d > 0)) continue;
                if (M == cap) { cap *= 2; NN = realloc(NN, sizeof(int32_t) * cap); if (!NN) return ORC_ENOMEM; }
                NN[M++] = buf[t].j + 1;
            }
        } else {
            const double r = radius[c];
            for (int64_t j = 0; j < m1; ++j)
                if ((col[j] < r) && (col[j] > 0)) {
                    if (M == cap) { cap *= 2; NN = realloc(NN, sizeof(int32_t) * cap); if (!NN) return ORC_ENOMEM; }
                    NN[M++] = (int32_t)(j + 1);
                }
        }
    }
    offsets[m2] = M;
    free(buf);
    *idx_out = NN;
    return ORC_OK;
}

/* nearestNeighborsSeparateHitMiss, R/nearestNeighbors.R:318-551 (npdr(separate.hitmiss.nbds = TRUE), R/npdr.R:215-251).
 * Same distance matrix; hit and miss neighbourhoods are formed separately and listed hits first, then misses.
 *  multisurf (:483-500): per Ri, radius_hit = mean(d[hits, not self]) - sd.frac*sd(same), radius_miss likewise over misses
 *                        (base-R mean(): long double two-pass; sd(): cov.c as in r_var);
 *  surf (:456-481):      one hit radius from the concatenation over i of d[i, hits(i) without self], one miss radius;
 *  selection (:529-533): hits = same class & d < r_hit & d > 0; misses = other class & d < r_miss (no d > 0 test), each ascending j;
 *  relieff (:385-427):   order(d[Ri, ]) (stable), hits = that order restricted to the class of Ri, misses = the rest;
 *                        balanced classes: k.hits = k.miss = floor(0.5*knnSURF(N-1, sd.frac)); else knnSURF(m.hits), knnSURF(m.miss)
 *                        with m.hits = #hits incl. self - 1; neighbours = hits[2:(k.hits+1)] then misses[1:k.miss].
 *                        (k is only used through sd.frac; positions past the end of a list are dropped here, R would yield NA.)
 * pheno: class labels as doubles (as.numeric(as.character(pheno.vec)) in the reference). */
static double r_mean_skip(const double *x, const uint8_t *use, int64_t n, int64_t *cnt_out) {
    long double s = 0.0L; int64_t c = 0;
    for (int64_t i = 0; i < n; ++i) if (use[i]) { s += x[i]; c++; }
    s /= c;
    long double t = 0.0L;
    for (int64_t i = 0; i < n; ++i) if (use[i]) t += (x[i] - s);
    s += t / c;
    if (cnt_out) *cnt_out = c;
    return (double)s;
}
static double r_var_mask(const double *x, const uint8_t *use, int64_t n) {
    int64_t c = 0;
    long double xm = r_mean_skip(x, use, n, &c);           /* identical recipe to cov.c MEAN_ */
    long double sum = 0.0L;
    for (int64_t k = 0; k < n; ++k) if (use[k]) sum += (x[k] - xm) * (x[k] - xm);
    return (double)(sum / (c - 1));
}

int orc_radius_hitmiss_multisurf(const double *D, int64_t N, const double *pheno, double sd_frac,
                                 double *r_hit, double *r_miss) {
#pragma omp parallel
    {
        uint8_t *use = (uint8_t *)malloc((size_t)N);
#pragma omp for schedule(static)
        for (int64_t c = 0; c < N; ++c) {
            const double *col = D + c * N;                 /* dist.mat[i, ] == column i (symmetric) */
            for (int64_t j = 0; j < N; ++j) use[j] = (uint8_t)(j != c && pheno[j] == pheno[c]);
            r_hit[c] = r_mean_skip(col, use, N, NULL) - sd_frac * sqrt(r_var_mask(col, use, N));
            for (int64_t j = 0; j < N; ++j) use[j] = (uint8_t)(pheno[j] != pheno[c]);
            r_miss[c] = r_mean_skip(col, use, N, NULL) - sd_frac * sqrt(r_var_mask(col, use, N));
        }
        free(use);
    }
    return ORC_OK;
}

int orc_radius_hitmiss_surf(const double *D, int64_t N, const double *pheno, double sd_frac, double *r_hit, double *r_miss) {
    double *h = (double *)malloc(sizeof(double) * (size_t)N * (size_t)N), *m = (double *)malloc(sizeof(double) * (size_t)N * (size_t)N);
    uint8_t *use = (uint8_t *)malloc((size_t)N * (size_t)N);
    if (!h || !m || !use) return ORC_ENOMEM;
    int64_t nh = 0, nm = 0;
    for (int64_t i = 0; i < N; ++i)                         /* unlist(list of rows): row i, ascending j */
        for (int64_t j = 0; j < N; ++j) {
            if (j != i && pheno[j] == pheno[i]) h[nh++] = D[i * N + j];
            if (pheno[j] != pheno[i]) m[nm++] = D[i * N + j];
        }
    memset(use, 1, (size_t)N * (size_t)N);
    *r_hit = r_mean_skip(h, use, nh, NULL) - sd_frac * sqrt(r_var_mask(h, use, nh));
    *r_miss = r_mean_skip(m, use, nm, NULL) - sd_frac * sqrt(r_var_mask(m, use, nm));
    free(h); free(m); free(use);
    return ORC_OK;
}

int orc_neighbors_hitmiss(const double *D, int64_t N, const double *pheno, int method, double sd_frac,
                          const double *r_hit, const double *r_miss,
                          int32_t **Ri_out, int32_t **NN_out, int64_t *M_out) {
    int64_t cap = 1024, M = 0;
    int32_t *Ri = (int32_t *)malloc(sizeof(int32_t) * cap), *NN = (int32_t *)malloc(sizeof(int32_t) * cap);
    dj_t *buf = (dj_t *)malloc(sizeof(dj_t) * (size_t)(N > 0 ? N : 1));
    if (!Ri || !NN || !buf) return ORC_ENOMEM;
#define ORC_PUSH(c_, j_) do { if (M == cap) { cap *= 2; Ri = realloc(Ri, sizeof(int32_t) * cap); NN = realloc(NN, sizeof(int32_t) * cap); \
        if (!Ri || !NN) return ORC_ENOMEM; } Ri[M] = (int32_t)((c_) + 1); NN[M] = (int32_t)((j_) + 1); ++M; } while (0)
    /* class balance: table(as.character(pheno.vec)) with two classes */
    int64_t n_first = 0;
    for (int64_t i = 0; i < N; ++i) if (pheno[i] == pheno[0]) n_first++;
    const int balanced = (2 * n_first == N);
    for (int64_t c = 0; c < N; ++c) {
        const double *col = D + c * N;
        if (method == ORC_NBD_RELIEFF) {
            for (int64_t j = 0; j < N; ++j) { buf[j].d = col[j]; buf[j].j = (int32_t)j; }
            qsort(buf, (size_t)N, sizeof(dj_t), cmp_dj);            /* order(): ascending, ties by index */
            int64_t n_hits = 0, n_miss = 0;
            for (int64_t j = 0; j < N; ++j) { if (pheno[j] == pheno[c]) n_hits++; else n_miss++; }
            int64_t k_hits, k_miss;
            if (balanced) { k_hits = (int64_t)floor(0.5 * (double)orc_knn_surf(N - 1, sd_frac)); k_miss = k_hits; }
            else { k_hits = orc_knn_surf(n_hits - 1, sd_frac); k_miss = orc_knn_surf(n_miss, sd_frac); }
            int64_t seen = 0;
            for (int64_t t = 0; t < N && seen < k_hits + 1; ++t)
                if (pheno[buf[t].j] == pheno[c]) { seen++; if (seen >= 2) ORC_PUSH(c, buf[t].j); }     /* hits[2:(k.hits+1)] */
            seen = 0;
            for (int64_t t = 0; t < N && seen < k_miss; ++t)
                if (pheno[buf[t].j] != pheno[c]) { seen++; ORC_PUSH(c, buf[t].j); }
        } else {
            const double rh = (method == ORC_NBD_SURF) ? r_hit[0] : r_hit[c];
            const double rm = (method == ORC_NBD_SURF) ? r_miss[0] : r_miss[c];
            for (int64_t j = 0; j < N; ++j) if (pheno[j] == pheno[c] && col[j] < rh && col[j] > 0) ORC_PUSH(c, j);
            for (int64_t j = 0; j < N; ++j) if (pheno[j] != pheno[c] && col[j] < rm) ORC_PUSH(c, j);
        }
    }
#undef ORC_PUSH
    free(buf);
    *Ri_out = Ri; *NN_out = NN; *M_out = M;
    return ORC_OK;
}

void orc_free(void *p) { free(p); }

/* uniqueNeighbors, R/nearestNeighbors.R:572-583: key = "min,max"; keep rows that are
 * not duplicated() (first occurrence in list order).  keep[m] = 1 for kept rows. */
int64_t orc_unique_neighbors(const int32_t *Ri, const int32_t *NN, int64_t M, int64_t N, uint8_t *keep) {
    /* seen bitmap over unordered pairs (lo, hi), N x N bits */
    size_t nbits = (size_t)N * (size_t)N;
    uint8_t *seen = (uint8_t *)calloc((nbits + 7) / 8, 1);
    if (!seen) return -1;
    int64_t cnt = 0;
    for (int64_t m = 0; m < M; ++m) {
        int64_t a = Ri[m] - 1, b = NN[m] - 1;
        int64_t lo = a < b ? a : b, hi = a < b ? b : a;
        size_t bit = (size_t)lo * (size_t)N + (size_t)hi;
        int was = (seen[bit >> 3] >> (bit & 7)) & 1;
        if (!was) { seen[bit >> 3] |= (uint8_t)(1u << (bit & 7)); cnt++; }
        if (keep) keep[m] = (uint8_t)!was;
    }
    free(seen);
    return cnt;
}

/* knnVec, R/nearestNeighbors.R:603-607: count per Ri_idx for the Ri that appear (ascending).
 * counts must have room for N entries; returns the number of distinct Ri. */
int64_t orc_knn_vec(const int32_t *Ri, int64_t M, int64_t N, int64_t *counts) {
    int64_t *c = (int64_t *)calloc((size_t)N + 1, sizeof(int64_t));
    if (!c) return -1;
    for (int64_t m = 0; m < M; ++m) c[Ri[m]]++;
    int64_t n = 0;
    for (int64_t i = 1; i <= N; ++i) if (c[i] > 0) counts[n++] = c[i];
    free(c);
    return n;
}

/* ----------------------------------------------------- LINPACK / BLAS pieces */
/* reference-BLAS dnrm2 (scale / ssq form), ddot and daxpy in plain sequential double. */
static double b_dnrm2(int64_t n, const double *x) {
    if (n < 1) return 0.0;
    if (n == 1) return fabs(x[0]);
    double scale = 0.0, ssq = 1.0;
    for (int64_t i = 0; i < n; ++i) {
        if (x[i] != 0.0) {
            double absxi = fabs(x[i]);
            if (scale < absxi) { double t = scale / absxi; ssq = 1.0 + ssq * t * t; scale = absxi; }
            else { double t = absxi / scale; ssq += t * t; }
        }
    }
    return scale * sqrt(ssq);
}
static double b_ddot(int64_t n, const double *x, const double *y) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}
static void b_daxpy(int64_t n, double a, const double *x, double *y) {
    if (a == 0.0) return;
    for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
}

/* dqrdc2 (R's modification of LINPACK dqrdc: limited column pivoting; a column whose
 * norm falls below tol x its original norm is cycled to the end).  x is n x p column-major,
 * overwritten with the QR; returns rank k.  jpvt is 0-based on entry (0..p-1) and exit.
 * work must hold 2*p doubles. */
static int dqrdc2(double *x, int64_t n, int p, double tol, double *qraux, int *jpvt, double *work) {
    double *w1 = work, *w2 = work + p;
    for (int j = 0; j < p; ++j) {
        qraux[j] = b_dnrm2(n, x + (int64_t)j * n);
        w1[j] = qraux[j];
        w2[j] = qraux[j];
        if (w2[j] == 0.0) w2[j] = 1.0;
    }
    int lup = (n < p) ? (int)n : p;
    int k = p + 1;                                  /* 1-based bookkeeping as in the Fortran */
    for (int l = 1; l <= lup; ++l) {
        while (l < k && qraux[l - 1] < w2[l - 1] * tol) {
            /* move column l to the end */
            for (int64_t i = 0; i < n; ++i) {
                double t = x[i + (int64_t)(l - 1) * n];
                for (int j = l + 1; j <= p; ++j) x[i + (int64_t)(j - 2) * n] = x[i + (int64_t)(j - 1) * n];
                x[i + (int64_t)(p - 1) * n] = t;
            }
            int ii = jpvt[l - 1];
            double t = qraux[l - 1], tt = w1[l - 1], ttt = w2[l - 1];
            for (int j = l + 1; j <= p; ++j) {
                jpvt[j - 2] = jpvt[j - 1]; qraux[j - 2] = qraux[j - 1];
                w1[j - 2] = w1[j - 1]; w2[j - 2] = w2[j - 1];
            }
            jpvt[p - 1] = ii; qraux[p - 1] = t; w1[p - 1] = tt; w2[p - 1] = ttt;
            k--;
        }
        if (l == n) continue;
        double *xl = x + (l - 1) + (int64_t)(l - 1) * n;
        int64_t len = n - l + 1;
        double nrmxl = b_dnrm2(len, xl);
        if (nrmxl == 0.0) continue;
        if (xl[0] != 0.0) nrmxl = copysign(nrmxl, xl[0]);
        { double s = 1.0 / nrmxl; for (int64_t i = 0; i < len; ++i) xl[i] *= s; }
        xl[0] = 1.0 + xl[0];
        for (int j = l + 1; j <= p; ++j) {
            double *xj = x + (l - 1) + (int64_t)(j - 1) * n;
            double t = -b_ddot(len, xl, xj) / xl[0];
            b_daxpy(len, t, xl, xj);
            if (qraux[j - 1] != 0.0) {
                double r = fabs(xj[0]) / qraux[j - 1];
                double tt = 1.0 - r * r;
                if (tt < 0.0) tt = 0.0;
                t = tt;
                if (fabs(t) < 1e-6) {
                    qraux[j - 1] = b_dnrm2(n - l, xj + 1);
                    w1[j - 1] = qraux[j - 1];
                } else {
                    qraux[j - 1] = qraux[j - 1] * sqrt(t);
                }
            }
        }
        qraux[l - 1] = xl[0];
        xl[0] = -nrmxl;
    }
    k = (k - 1 < n) ? k - 1 : (int)n;
    return k;
}

/* dqrls core: given the dqrdc2 factorisation, compute qty = Q'y, solve R b = qty[0:k]
 * (dqrsl job 1100) and residuals (job 10).  b has p entries (pivoted order, 0 beyond rank). */
static void qr_solve(const double *qr, int64_t n, int p, int k, const double *qraux,
                     const double *y, double *qty, double *b, double *rsd) {
    memcpy(qty, y, sizeof(double) * (size_t)n);
    int ju = (k < n - 1) ? k : (int)(n - 1);
    for (int j = 0; j < ju; ++j) {
        if (qraux[j] != 0.0) {
            const double *col = qr + j + (int64_t)j * n;
            double t = qraux[j] * qty[j];                    /* diag holds -nrmxl; the Householder head is qraux */
            for (int64_t i = 1; i < n - j; ++i) t += col[i] * qty[j + i];
            t = -t / qraux[j];
            qty[j] += t * qraux[j];
            for (int64_t i = 1; i < n - j; ++i) qty[j + i] += t * col[i];
        }
    }
    for (int j = 0; j < p; ++j) b[j] = 0.0;
    for (int j = 0; j < k; ++j) b[j] = qty[j];
    for (int j = k - 1; j >= 0; --j) {                       /* back substitution */
        b[j] = b[j] / qr[j + (int64_t)j * n];
        if (j > 0) { double t = -b[j]; for (int i = 0; i < j; ++i) b[i] += t * qr[i + (int64_t)j * n]; }
    }
    if (rsd) {
        /* residuals = Q * [0; qty[k:]] */
        for (int64_t i = 0; i < n; ++i) rsd[i] = (i < k) ? 0.0 : qty[i];
        for (int j = ju - 1; j >= 0; --j) {
            if (qraux[j] != 0.0) {
                const double *col = qr + j + (int64_t)j * n;
                double t = qraux[j] * rsd[j];
                for (int64_t i = 1; i < n - j; ++i) t += col[i] * rsd[j + i];
                t = -t / qraux[j];
                rsd[j] += t * qraux[j];
                for (int64_t i = 1; i < n - j; ++i) rsd[j + i] += t * col[i];
            }
        }
    }
}

/* chol2inv on the leading k x k upper triangle of qr: diag of (R'R)^-1. Also full inverse if cov != NULL (k x k col-major). */
static void chol2inv_diag(const double *qr, int64_t n, int k, double *diag, double *cov) {
    double Rinv[ORC_MAX_DESIGN * ORC_MAX_DESIGN];
    if (k > ORC_MAX_DESIGN) k = ORC_MAX_DESIGN;
    for (int j = 0; j < k; ++j) {                             /* invert upper-triangular R column by column */
        for (int i = 0; i < k; ++i) Rinv[i + j * k] = 0.0;
        Rinv[j + j * k] = 1.0 / qr[j + (int64_t)j * n];
        for (int i = j - 1; i >= 0; --i) {
            double s = 0.0;
            for (int t = i + 1; t <= j; ++t) s += qr[i + (int64_t)t * n] * Rinv[t + j * k];
            Rinv[i + j * k] = -s / qr[i + (int64_t)i * n];
        }
    }
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            double s = 0.0;
            int t0 = i > j ? i : j;
            for (int t = t0; t < k; ++t) s += Rinv[i + t * k] * Rinv[j + t * k];
            if (cov) cov[i + j * k] = s;
            if (i == j) diag[i] = s;
        }
}

/* R's mean(): long double sum / n, then one long double correction pass (summary.c real_mean). */
static double r_mean(const double *x, int64_t n) {
    long double s = 0.0L;
    for (int64_t i = 0; i < n; ++i) s += x[i];
    s /= n;
    long double t = 0.0L;
    for (int64_t i = 0; i < n; ++i) t += (x[i] - s);
    s += t / n;
    return (double)s;
}

/* -------------------------------------------------------------- lm + summary */
typedef struct {
    int rank;            /* numerical rank found by dqrdc2 (tol 1e-7) */
    int64_t df_resid;    /* n - rank */
    double rss, mss, r_squared, deviance;
    int iters;           /* glm: IRLS iterations used */
    int converged;       /* glm: 1 if the deviance criterion was met */
    int boundary;        /* unused (no step-halving to the boundary in this family) */
} orc_fit_info;

/* lm(y ~ X) + summary.lm, R/npdr.R:34,41-42,67 (base-R lm.fit -> Cdqrls -> dqrdc2, tol 1e-7;
 * summary.lm: resvar = rss/rdf, se = sqrt(diag(chol2inv(R)) * resvar), r.squared = mss/(mss+rss)).
 * Xd: n x p design (column 0 must be the intercept), column-major.  coef/se have p entries in
 * ORIGINAL column order; aliased coefficients are NaN. */
int orc_lm(const double *Xd, const double *y, int64_t n, int p, double *coef, double *se, orc_fit_info *info) {
    if (p > ORC_MAX_DESIGN || p < 1) return ORC_EINVAL;
    double *qr = (double *)malloc(sizeof(double) * (size_t)n * (size_t)p);
    double *qty = (double *)malloc(sizeof(double) * (size_t)n);
    double *rsd = (double *)malloc(sizeof(double) * (size_t)n);
    double *fit = (double *)malloc(sizeof(double) * (size_t)n);
    if (!qr || !qty || !rsd || !fit) { free(qr); free(qty); free(rsd); free(fit); return ORC_ENOMEM; }
    memcpy(qr, Xd, sizeof(double) * (size_t)n * (size_t)p);
    double qraux[ORC_MAX_DESIGN], work[2 * ORC_MAX_DESIGN], b[ORC_MAX_DESIGN], diag[ORC_MAX_DESIGN];
    int jpvt[ORC_MAX_DESIGN];
    for (int j = 0; j < p; ++j) jpvt[j] = j;
    int k = dqrdc2(qr, n, p, 1e-7, qraux, jpvt, work);
    qr_solve(qr, n, p, k, qraux, y, qty, b, rsd);
    for (int64_t i = 0; i < n; ++i) fit[i] = y[i] - rsd[i];
    long double rss = 0.0L;
    for (int64_t i = 0; i < n; ++i) rss += rsd[i] * rsd[i];
    double mf = r_mean(fit, n);
    long double mss = 0.0L;
    for (int64_t i = 0; i < n; ++i) { double t = fit[i] - mf; mss += t * t; }
    int64_t rdf = n - k;
    double resvar = (double)rss / (double)rdf;
    chol2inv_diag(qr, n, k, diag, NULL);
    for (int j = 0; j < p; ++j) { coef[j] = NAN; se[j] = NAN; }
    for (int j = 0; j < k; ++j) { coef[jpvt[j]] = b[j]; se[jpvt[j]] = sqrt(diag[j] * resvar); }
    if (info) {
        info->rank = k; info->df_resid = rdf; info->rss = (double)rss; info->mss = (double)mss;
        info->r_squared = (double)mss / ((double)mss + (double)rss);
        info->deviance = (double)rss; info->iters = 0; info->converged = 1; info->boundary = 0;
    }
    free(qr); free(qty); free(rsd); free(fit);
    return ORC_OK;
}

/* --------------------------------------------------- glm.fit binomial(logit) */
/* binomial()$linkinv / mu.eta (C logit_linkinv / logit_mu_eta in stats family.c):
 * THRESH = 30, MTHRESH = -30, INVEPS = 1/DBL_EPSILON. */
static inline double logit_linkinv(double eta) {
    double t = (eta < -30.0) ? DBL_EPSILON : ((eta > 30.0) ? 1.0 / DBL_EPSILON : exp(eta));
    return t / (1.0 + t);
}
static inline double logit_mu_eta(double eta) {
    double opexp = 1.0 + exp(eta);
    return (eta > 30.0 || eta < -30.0) ? DBL_EPSILON : exp(eta) / (opexp * opexp);
}
static inline double y_log_y(double y, double mu) { return (y != 0.0) ? (y * log(y / mu)) : 0.0; }
static inline double binomial_dev_resid(double y, double mu) {   /* weights 1 */
    return 2.0 * (y_log_y(y, mu) + y_log_y(1.0 - y, 1.0 - mu));
}

/* glm(y ~ X, binomial(logit)) + summary.glm, R/npdr.R:36,41-42 (base-R glm.fit):
 * mustart = (y + 0.5)/2; IRLS with Cdqrls (tol = min(1e-7, epsilon/1000) = 1e-11);
 * stop when |dev - devold|/(|dev| + 0.1) < 1e-8, maxit 25; step-halving on non-finite deviance;
 * summary: dispersion 1, cov.unscaled = chol2inv(R) of the LAST weighted QR.
 * y must be 0/1.  coef/se in original column order; aliased -> NaN. */
int orc_glm_binomial(const double *Xd, const double *y, int64_t n, int p, double *coef, double *se, orc_fit_info *info) {
    if (p > ORC_MAX_DESIGN || p < 1) return ORC_EINVAL;
    size_t nn = (size_t)n;
    double *qr = (double *)malloc(sizeof(double) * nn * (size_t)p);
    double *eta = (double *)malloc(sizeof(double) * nn), *mu = (double *)malloc(sizeof(double) * nn);
    double *zw = (double *)malloc(sizeof(double) * nn), *qty = (double *)malloc(sizeof(double) * nn);
    if (!qr || !eta || !mu || !zw || !qty) { free(qr); free(eta); free(mu); free(zw); free(qty); return ORC_ENOMEM; }
    double qraux[ORC_MAX_DESIGN], work[2 * ORC_MAX_DESIGN], b[ORC_MAX_DESIGN], start[ORC_MAX_DESIGN], coefold[ORC_MAX_DESIGN], diag[ORC_MAX_DESIGN];
    int jpvt[ORC_MAX_DESIGN], have_old = 0, k = 0, conv = 0, iter;
    long double acc = 0.0L;
    for (int64_t i = 0; i < n; ++i) {
        double ms = (y[i] + 0.5) / 2.0;
        eta[i] = log(ms / (1.0 - ms));
        mu[i] = logit_linkinv(eta[i]);
        acc += binomial_dev_resid(y[i], mu[i]);
    }
    double devold = (double)acc, dev = devold;
    for (int j = 0; j < p; ++j) { start[j] = 0.0; coefold[j] = 0.0; }
    for (iter = 1; iter <= 25; ++iter) {
        /* all weights are 1 and varmu > 0 here; `good` = mu.eta != 0 (always true: floor is DBL_EPSILON) */
        for (int64_t i = 0; i < n; ++i) {
            double varmu = mu[i] * (1.0 - mu[i]);
            double me = logit_mu_eta(eta[i]);
            double z = eta[i] + (y[i] - mu[i]) / me;
            double w = sqrt(me * me / varmu);
            zw[i] = z * w;
            for (int j = 0; j < p; ++j) qr[i + (int64_t)j * n] = Xd[i + (int64_t)j * n] * w;
        }
        for (int j = 0; j < p; ++j) jpvt[j] = j;
        k = dqrdc2(qr, n, p, 1e-11, qraux, jpvt, work);
        qr_solve(qr, n, p, k, qraux, zw, qty, b, NULL);
        int finite = 1;
        for (int j = 0; j < p; ++j) if (!isfinite(b[j])) finite = 0;
        if (!finite) { conv = 0; break; }
        for (int j = 0; j < p; ++j) start[jpvt[j]] = b[j];
        acc = 0.0L;
        for (int64_t i = 0; i < n; ++i) {
            double e = 0.0;
            for (int j = 0; j < p; ++j) e += Xd[i + (int64_t)j * n] * start[j];
            eta[i] = e; mu[i] = logit_linkinv(e);
            acc += binomial_dev_resid(y[i], mu[i]);
        }
        dev = (double)acc;
        if (!isfinite(dev)) {
            if (!have_old) { conv = 0; break; }
            int ii = 1;
            while (!isfinite(dev)) {
                if (ii > 25) break;
                ii++;
                for (int j = 0; j < p; ++j) start[j] = (start[j] + coefold[j]) / 2.0;
                acc = 0.0L;
                for (int64_t i = 0; i < n; ++i) {
                    double e = 0.0;
                    for (int j = 0; j < p; ++j) e += Xd[i + (int64_t)j * n] * start[j];
                    eta[i] = e; mu[i] = logit_linkinv(e);
                    acc += binomial_dev_resid(y[i], mu[i]);
                }
                dev = (double)acc;
            }
        }
        if (fabs(dev - devold) / (fabs(dev) + 0.1) < 1e-8) { conv = 1; break; }
        devold = dev;
        for (int j = 0; j < p; ++j) coefold[j] = start[j];
        have_old = 1;
    }
    if (iter > 25) iter = 25;
    chol2inv_diag(qr, n, k, diag, NULL);
    for (int j = 0; j < p; ++j) { coef[j] = NAN; se[j] = NAN; }
    for (int j = 0; j < k; ++j) { coef[jpvt[j]] = start[jpvt[j]]; se[jpvt[j]] = sqrt(diag[j]); }
    if (info) {
        info->rank = k; info->df_resid = n - k; info->rss = NAN; info->mss = NAN; info->r_squared = NAN;
        info->deviance = dev; info->iters = iter; info->converged = conv; info->boundary = 0;
    }
    free(qr); free(eta); free(mu); free(zw); free(qty);
    return ORC_OK;
}

/* ------------------------------------------------------------ diffRegression */
/* One attribute: design = [1, attr.diff, covar.diffs...] (R/npdr.R:404-411), fit
 * (R/npdr.R:22-40), pull coefficient rows (R/npdr.R:41-68).
 * out[8] = { beta_a, stat_a, beta_0, stat_0, df_resid, r_squared (lm) or deviance (glm), iters, flags }
 *   stat = beta/se (t for lm, z for glm); p-values (pt / pnorm) are formed by the caller.
 *   flags bit0 = attribute aliased, bit1 = glm not converged.
 * Quirk: summary()$coefficients drops aliased rows, so with covariates present
 * coeffs[2,] is the first non-aliased row after the intercept; without covariates
 * nrow(coeffs) < 2 -> NA.  Reproduced here. */
int orc_diff_regression(const double *design /* M x k, col 0 = 1s */, const double *y, int64_t M, int k,
                        int regression_type, double *out) {
    double coef[ORC_MAX_DESIGN], se[ORC_MAX_DESIGN];
    orc_fit_info info;
    int rc = (regression_type == ORC_REG_LM) ? orc_lm(design, y, M, k, coef, se, &info)
                                             : orc_glm_binomial(design, y, M, k, coef, se, &info);
    if (rc) return rc;
    int flags = 0, row2 = -1;
    if (isnan(coef[1])) flags |= 1;
    for (int j = 1; j < k; ++j) if (!isnan(coef[j])) { row2 = j; break; }
    if (regression_type == ORC_REG_BINOMIAL && !info.converged) flags |= 2;
    out[0] = (row2 >= 0) ? coef[row2] : NAN;
    out[1] = (row2 >= 0) ? coef[row2] / se[row2] : NAN;
    out[2] = coef[0];
    out[3] = coef[0] / se[0];
    out[4] = (double)info.df_resid;
    out[5] = (regression_type == ORC_REG_LM) ? info.r_squared : info.deviance;
    out[6] = (double)info.iters;
    out[7] = (double)flags;
    return ORC_OK;
}

/* Regression loop, R/npdr.R:391-430: for every attribute column build the projected diff over the
 * pair list and fit.  ydiff: M (lm: |y_i - y_j|, binomial: 0/1 miss indicator, R/npdr.R:299-308);
 * cdiff: M x q column-major covariate diffs (R/npdr.R:315-345).  out: p x 8 row-major (see above).
 * OpenMP over attributes = the reference's dopar.reg (R/npdr.R:348-388). */
int orc_regress_attrs(const double *X, int64_t N, int64_t p, const int32_t *Ri, const int32_t *NN, int64_t M,
                      const double *ydiff, const double *cdiff, int q, int attr_diff_type, int regression_type,
                      double *out) {
    int k = 2 + q;
    int err = 0;
#pragma omp parallel
    {
        double *design = (double *)malloc(sizeof(double) * (size_t)M * (size_t)k);
        if (!design) {
#pragma omp atomic write
            err = ORC_ENOMEM;
        } else {
            for (int64_t m = 0; m < M; ++m) design[m] = 1.0;
            for (int c = 0; c < q; ++c) memcpy(design + (int64_t)(2 + c) * M, cdiff + (int64_t)c * M, sizeof(double) * (size_t)M);
#pragma omp for schedule(dynamic, 1)
            for (int64_t a = 0; a < p; ++a) {
                const double *col = X + a * N;
                double *d = design + M;
                for (int64_t m = 0; m < M; ++m) d[m] = diff1(col[Ri[m] - 1], col[NN[m] - 1], attr_diff_type, 1.0);
                int rc = orc_diff_regression(design, ydiff, M, k, regression_type, out + a * 8);
                if (rc) {
#pragma omp atomic write
                    err = rc;
                }
            }
            free(design);
        }
    }
    return err;
}

/* Phenotype / covariate diffs over a pair list, R/npdr.R:299-345. */
int orc_pair_diffs(const double *v, const int32_t *Ri, const int32_t *NN, int64_t M, int type, double *out) {
    for (int64_t m = 0; m < M; ++m) out[m] = diff1(v[Ri[m] - 1], v[NN[m] - 1], type, 1.0);
    return ORC_OK;
}
