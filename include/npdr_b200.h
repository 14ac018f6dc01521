/*
 * npdr_b200.h -- C ABI of libnpdr_b200.so, the B200-native engine for NPDR's hot path
 * (distances -> neighbour pairs -> per-attribute GLM on projected pair differences).
 *
 * The reference (insilico/npdr) has no FFI boundary: its operator API is a set of exported
 * R closures (NAMESPACE:11-24,29-30).  Each entry point below replaces one of them and is
 * what an R `.Call` shim binds (see INTEGRATION.md for the shim and the R wrapper):
 *
 *   npdrb_diff               <- npdrDiff            R/nearestNeighbors.R:15-40
 *   npdrb_distances          <- npdrDistances       R/nearestNeighbors.R:63-101
 *   npdrb_nearest_neighbors  <- nearestNeighbors    R/nearestNeighbors.R:153-285 (+ get_Ri_nearest :622-634,
 *                                                    uniqueNeighbors :572-583, knnVec :603-607)
 *   npdrb_nearest_neighbors_hitmiss <- nearestNeighborsSeparateHitMiss   R/nearestNeighbors.R:318-551
 *   npdrb_distances2         <- npdrDistances2      R/npdrLearner.R:323-339 (flexclust::dist2)
 *   npdrb_nearest_neighbors2 <- nearestNeighbors2   R/npdrLearner.R:378-482 (+ get_Ri_nearest R/nearestNeighbors.R:622-634)
 *   npdrb_knn_surf           <- knnSURF             R/utils.R:12-18
 *   npdrb_regress            <- the regression loop + diffRegression   R/npdr.R:391-430, 18-71
 *   npdrb_npdr               <- npdr() between argument parsing and order()/p.adjust()
 *                                                    R/npdr.R:256-270, 299-345, 391-430
 *   npdrb_npdr_multi         <- the same on several GPUs from one process (the R session's multi-GPU entry point)
 *   npdrb_unireg             <- uniReg's per-column fits   R/utils.R:66-157
 *   npdrb_npdr_k_grid        <- vwok's loop of npdr(nbd.method = "relieff", knn = k) over a k grid   R/adaptive.R:73,85-111
 *
 * What stays on the R side (per BASELINE north_star): argument parsing, factor coding,
 * pt()/pnorm() p-values, order(), p.adjust(), data-frame assembly.
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns 0 on success, non-zero on error;
 *     npdrb_last_error() returns a thread-local message for the last failure.
 *   - matrices are column-major doubles, exactly as R holds them; inputs are never written.
 *   - instance indices in pair lists are 1-based int32 (R's Ri_idx / NN_idx).
 *   - there is NO CPU fallback: every compute entry point fails with NPDRB_ENODEVICE
 *     when no sm_100 device is usable.
 *   - buffers returned through pointer-to-pointer arguments are malloc'ed by the library
 *     and released with npdrb_free().
 *
 * The `npdrb_session_*` functions expose the same stages on DEVICE-resident data so a host
 * (R with several GPUs, or the Python/torch.distributed harness in this repo) can shard the
 * path: row blocks for distance+neighbours, attribute columns for the regression.
 */
#ifndef NPDR_B200_H
#define NPDR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPDRB_OK 0
#define NPDRB_EINVAL 1        /* bad argument / unsupported option combination */
#define NPDRB_ENODEVICE 2     /* no usable CUDA device (never falls back to the CPU) */
#define NPDRB_ECUDA 3         /* CUDA runtime error, see npdrb_last_error() */
#define NPDRB_ENOMEM 4
#define NPDRB_EDATA 5         /* NA/NaN/Inf in the input, single-level binomial outcome, ... */

/* npdrDiff diff.type, R/nearestNeighbors.R:15 ("manhattan" == "numeric-abs") */
enum { NPDRB_DIFF_NUMERIC_ABS = 0, NPDRB_DIFF_NUMERIC_SQR = 1, NPDRB_DIFF_ALLELE_SHARING = 2,
       NPDRB_DIFF_MATCH_MISMATCH = 3, NPDRB_DIFF_RAW = 4 /* uniReg: the attribute itself */ };
/* npdrDistances metric, R/nearestNeighbors.R:63-101; PRECOMPUTED: R/nearestNeighbors.R:171-176 */
enum { NPDRB_METRIC_MANHATTAN = 0, NPDRB_METRIC_EUCLIDEAN = 1, NPDRB_METRIC_ALLELE_SHARING_MANHATTAN = 2,
       NPDRB_METRIC_RELIEF_SCALED_MANHATTAN = 3, NPDRB_METRIC_RELIEF_SCALED_EUCLIDEAN = 4,
       NPDRB_METRIC_PRECOMPUTED = 5, NPDRB_METRIC_HAMMING = 6 /* hamming.binary, R/utils.R:185-188 */ };
/* nbd.method, R/nearestNeighbors.R:153-154 */
enum { NPDRB_NBD_MULTISURF = 0, NPDRB_NBD_SURF = 1, NPDRB_NBD_RELIEFF = 2 };
/* regression.type, R/npdr.R:152 */
enum { NPDRB_REG_LM = 0, NPDRB_REG_BINOMIAL = 1 };

/* R/npdr.R:315-345 has no limit; up to 4 covariates run in the streaming kernels, 5..60 through the wide-design kernel */
#define NPDRB_MAX_COVARS 60

/* per-attribute statistics, one row per attribute (8 doubles):
 *   [0] beta_a   coefficient of the attribute diff       (coeffs[2,1], R/npdr.R:52)
 *   [1] stat_a   beta_a / se (t for lm, z for binomial)  (coeffs[2,3], R/npdr.R:53)
 *   [2] beta_0   intercept                                (coeffs[1,1], R/npdr.R:62)
 *   [3] stat_0   intercept / se
 *   [4] df_resid M - rank                                 (mod$df.residual, R/npdr.R:39)
 *   [5] r_squared (lm, R/npdr.R:67) or residual deviance (binomial)
 *   [6] iters    IRLS iterations (binomial), 0 for lm
 *   [7] flags    bit0: attribute diff aliased (constant column); with covariates the row then holds the
 *                first non-aliased covariate, as summary()$coefficients would;
 *                bit1: IRLS did not converge in 25 iterations; bit2: non-finite deviance/coefficients.
 * pt()/pnorm() of stat_a, stat_0 are formed by the caller. */
#define NPDRB_STATS_PER_ATTR 8
#define NPDRB_FLAG_ALIASED 1
#define NPDRB_FLAG_NOT_CONVERGED 2
#define NPDRB_FLAG_NONFINITE 4

typedef struct {
    int32_t regression_type;                       /* NPDRB_REG_* */
    int32_t attr_diff_type;                        /* NPDRB_DIFF_* */
    int32_t nbd_method;                            /* NPDRB_NBD_* */
    int32_t nbd_metric;                            /* NPDRB_METRIC_* */
    int64_t knn;                                   /* relieff k; 0 -> knnSURF(N, sd_frac) */
    double  msurf_sd_frac;                         /* sd.frac, default 0.5 */
    int32_t n_covars;                              /* number of covariate columns used (= length(covar.diff.type), R/npdr.R:321) */
    int32_t covar_diff_type[NPDRB_MAX_COVARS];     /* NPDRB_DIFF_* per covariate */
    int32_t neighbor_sampling_unique;              /* neighbor.sampling == "unique", R/npdr.R:271-277 */
    int32_t fast_euclidean;                        /* 1: allow FMA contraction in the euclidean sum (<=1e-12 rel) */
    int32_t return_pairs;                          /* 1: also return the (Ri, NN) list */
    int32_t separate_hitmiss_nbds;                 /* separate.hitmiss.nbds, R/npdr.R:160,215-251 (binomial outcome = class labels) */
    int32_t reserved[7];
} npdrb_opts;

typedef struct {
    int64_t n_attr;
    double *stats;                                 /* n_attr x NPDRB_STATS_PER_ATTR, row-major; npdrb_free() */
    int64_t n_pairs;                               /* num.neighbor.pairs, R/npdr.R:266 (before unique sampling) */
    int64_t n_unique_pairs;                        /* num.unique.neighbors, R/npdr.R:269 */
    int64_t n_pairs_used;                          /* rows of the regression (differs when neighbor.sampling == "unique") */
    double  k_ave_empirical;                       /* mean(knnVec(pairs)), R/npdr.R:267 */
    int64_t n_empty;                               /* instances with an empty neighbourhood (skipped) */
    int64_t knn_used;                              /* relieff k actually used */
    int32_t *Ri, *NN;                              /* if requested: pair list, 1-based; npdrb_free() */
    double  ms_h2d, ms_distance, ms_neighbors, ms_prepare, ms_regress, ms_total; /* stage times (CUDA events) */
    int32_t irls_passes;                           /* binomial: number of full passes over the pair stream */
    int32_t reserved[7];
} npdrb_result;

typedef struct npdrb_ctx npdrb_ctx;                /* one per device */
typedef struct npdrb_session npdrb_session;        /* one NPDR problem resident on one device */

const char *npdrb_version(void);
const char *npdrb_last_error(void);
void npdrb_free(void *p);
void npdrb_opts_default(npdrb_opts *o);
void npdrb_result_release(npdrb_result *r);        /* frees the buffers inside *r */

int npdrb_create(int device, npdrb_ctx **ctx);     /* NPDRB_ENODEVICE if the device is absent or not sm_100 */
void npdrb_destroy(npdrb_ctx *ctx);
/* use a caller-owned CUDA stream (cudaStream_t as void*) for all subsequent work, e.g. torch's current stream */
int npdrb_set_stream(npdrb_ctx *ctx, void *cuda_stream);
int npdrb_synchronize(npdrb_ctx *ctx);
/* measured FP64 issue peaks of this device (DFMA and DADD micro-kernels): TFLOP/s counting FMA = 2, add = 1 */
int npdrb_measure_fp64_peak(npdrb_ctx *ctx, double *dfma_tflops, double *dadd_tflops);
/* return the library's cached device memory on this device to the driver (buffers are otherwise kept for reuse) */
int npdrb_release_cache(npdrb_ctx *ctx);
/* Copy `ncols` contiguous host columns of `rows` doubles (src, column-major, leading dimension rows) to device memory
 * (dst, leading dimension ld_dst doubles) and wait for it.  A pageable src (R's heap, numpy) goes through the context's
 * pinned staging ring (see npdrb_session_upload); used by hosts that shard the upload of X by attribute columns. */
int npdrb_upload_columns(npdrb_ctx *ctx, void *dst_device, int64_t ld_dst, const double *src_host, int64_t rows, int64_t ncols);
/* number of this library's kernel launches since the context was created (bench.py's gpu_launches) */
int64_t npdrb_launch_count(npdrb_ctx *ctx);

/* ---- host-buffer operator API (inputs/outputs in host memory; copies included) ---- */

/* npdrDiff on vectors of length n */
int npdrb_diff(npdrb_ctx *ctx, const double *a, const double *b, int64_t n, int32_t diff_type, double norm_fac,
               double *out);
/* N x N distance matrix (column-major, symmetric, zero diagonal) from N x p attributes */
int npdrb_distances(npdrb_ctx *ctx, const double *X, int64_t N, int64_t p, int32_t metric, int32_t fast_euclidean,
                    double *D_out);
/* npdrDistances2: m1 x m2 matrix (column-major) of distances between the rows of X1 (m1 x p) and of X2 (m2 x p).
 * metric: MANHATTAN, EUCLIDEAN or ALLELE_SHARING_MANHATTAN (both matrices halved first). */
int npdrb_distances2(npdrb_ctx *ctx, const double *X1, int64_t m1, const double *X2, int64_t m2, int64_t p, int32_t metric,
                     double *D_out);
/* nearestNeighbors2: for every instance of X2 (test) its neighbours among the instances of X1 (train).
 * The reference returns a list of m2 integer vectors; here: offsets (m2 + 1, malloc'ed) into idx (M train indices,
 * 1-based, malloc'ed).  sd_vec (length m2) may be NULL; radius_out (length m2) may be NULL (surf / multisurf). */
int npdrb_nearest_neighbors2(npdrb_ctx *ctx, const double *X1, int64_t m1, const double *X2, int64_t m2, int64_t p,
                             int32_t nbd_method, int32_t nbd_metric, const double *sd_vec, double sd_frac, int64_t k,
                             int64_t **offsets, int32_t **idx, int64_t *M, double *radius_out);
int64_t npdrb_knn_surf(int64_t m_samples, double sd_frac);
/* nearestNeighbors: X is N x p attributes, or the N x N distance matrix when nbd_metric == PRECOMPUTED.
 * sd_vec (length N) may be NULL.  radius_out (length N) may be NULL; filled for surf/multisurf. */
int npdrb_nearest_neighbors(npdrb_ctx *ctx, const double *X, int64_t N, int64_t p, int32_t nbd_method,
                            int32_t nbd_metric, const double *sd_vec, double sd_frac, int64_t k,
                            int32_t unique_sampling, int32_t **Ri, int32_t **NN, int64_t *M,
                            int64_t *n_unique, int64_t *n_empty, double *radius_out);
/* nearestNeighborsSeparateHitMiss: hit and miss neighbourhoods formed separately (class-conditional radii for
 * surf/multisurf, class-conditional k for relieff), listed hits first.  pheno: N class labels (two classes).
 * r_hit_out / r_miss_out (length N) may be NULL. */
int npdrb_nearest_neighbors_hitmiss(npdrb_ctx *ctx, const double *X, int64_t N, int64_t p, const double *pheno,
                                    int32_t nbd_method, int32_t nbd_metric, double sd_frac, int64_t k,
                                    int32_t unique_sampling, int32_t **Ri, int32_t **NN, int64_t *M, int64_t *n_unique,
                                    double *r_hit_out, double *r_miss_out);
/* regression loop over all p attributes for a given pair list.
 * y: length N outcome (lm: numeric; binomial: class labels, any two distinct doubles); pheno diff is formed
 * here as in R/npdr.R:299-308.  C: N x q covariates or NULL.  stats_out: p x 8 row-major. */
int npdrb_regress(npdrb_ctx *ctx, const double *X, int64_t N, int64_t p, const int32_t *Ri, const int32_t *NN,
                  int64_t M, const double *y, const double *C, int32_t q, const int32_t *covar_diff_type,
                  int32_t attr_diff_type, int32_t regression_type, double *stats_out);
/* the whole path on one device */
int npdrb_npdr(npdrb_ctx *ctx, const double *X, int64_t N, int64_t p, const double *y, const double *C,
               const npdrb_opts *opts, npdrb_result *res);
/* The whole path on n_devices GPUs of one node from ONE host process (what an R session calls: one host thread per
 * device for the duration of the call).  devices: CUDA ordinals or NULL for 0..n_devices-1.  Row blocks of the distance /
 * neighbour stage, symmetric distance halves pulled from the peers over NVLink, full pair list on every device, attribute
 * columns of the regression; same options, same result layout and the same pair list (bit for bit) as npdrb_npdr. */
int npdrb_npdr_multi(int32_t n_devices, const int32_t *devices, const double *X, int64_t N, int64_t p, const double *y,
                     const double *C, const npdrb_opts *opts, npdrb_result *res);
void npdrb_multi_shutdown(void);                   /* releases the per-device contexts npdrb_npdr_multi keeps between calls */
/* Batched k (vwok, R/adaptive.R:85-111): npdr(nbd.method = "relieff", knn = k_grid[c]) for every c on ONE resident problem --
 * one upload, one distance matrix, one stable (d, row) sort of every distance row (a ReliefF list is a prefix of the sorted
 * row, so each k costs two binary searches per row instead of a selection and a sort), one instance-major attribute copy.
 * opts->nbd_method must be RELIEFF; opts->knn is ignored.  stats_out: n_k x p x 8 (k-major); n_pairs_out (length n_k) may be
 * NULL.  Results are identical to n_k separate npdrb_npdr calls. */
int npdrb_npdr_k_grid(npdrb_ctx *ctx, const double *X, int64_t N, int64_t p, const double *y, const double *C,
                      const npdrb_opts *opts, const int64_t *k_grid, int32_t n_k, double *stats_out, int64_t *n_pairs_out);
/* uniReg: univariate lm/glm of y on each raw column (+ covariates C as additional regressors) */
int npdrb_unireg(npdrb_ctx *ctx, const double *X, int64_t N, int64_t p, const double *y, const double *C,
                 int32_t q, int32_t regression_type, double *stats_out);

/* ---- device-resident session API (sharding across GPUs; bench "value" with inputs already in HBM) ---- */

int npdrb_session_create(npdrb_ctx *ctx, int64_t N, int64_t p, int32_t q, npdrb_session **s);
void npdrb_session_destroy(npdrb_session *s);
/* copy host inputs in; C may be NULL when q == 0.
 * A matrix of 64 MB or more is uploaded as 8 column slabs on a second stream and the distance kernel starts on the first
 * slab while the others cross PCIe (bit-identical results).  A pinned (cudaHostAlloc / cudaHostRegister) X is enqueued
 * at once; a PAGEABLE X (R's heap) is staged slab by slab through the context's ring of three 32 MB pinned buffers
 * (helper threads memcpy, async H2D), each slab right before the kernel that consumes it is launched, so the staging of
 * slab i+1 runs under the kernel of slab i.  X must therefore stay valid until the first call that consumes it
 * (distances / neighbours / regress) has returned.  NPDRB_NO_PIPELINED_UPLOAD=1 restores one blocking copy;
 * NPDRB_STAGE_THREADS sets the number of copier threads (default: host cores / GPUs of the box, between 4 and 8). */
int npdrb_session_upload(npdrb_session *s, const double *X, const double *y, const double *C);
/* or adopt device buffers: dX column-major with leading dimension ldx >= N (ldx % 2 == 0), dy[N], dC N x q (ld N) */
int npdrb_session_set_device_inputs(npdrb_session *s, const double *dX, int64_t ldx, const double *dy,
                                    const double *dC);
/* The adopted device matrix may still be arriving (one NCCL broadcast per rank's attribute shard, a peer copy, ...): column
 * slab i = attributes [i * slab_cols, (i+1) * slab_cols) is complete once events[i] (cudaEvent_t, owned by the caller, already
 * recorded on the caller's stream) fires.  The next distance call runs one launch per slab, each waiting for its own event
 * only, so the transfer of slab i+1 hides behind the kernel of slab i; results are bit-identical.  slab_cols % 16 == 0.
 * ready (may be NULL): host flags for transfers that ANOTHER host thread is still issuing while the distance call runs --
 * events[i] is waited for only after ready[i] != 0 (the producer records the event, then sets the flag); the events must
 * already exist (have been recorded once), and both arrays must stay valid until the distance call has returned. */
int npdrb_session_set_device_slab_events(npdrb_session *s, int32_t n_slabs, int64_t slab_cols, void *const *events,
                                         const int32_t *ready);
/* the same with slabs of different widths: slab i = attributes [slab_starts[i], slab_starts[i+1]), slab_starts[0] = 0,
 * slab_starts[n_slabs] = p, every bound but the last a multiple of 16 */
int npdrb_session_set_device_slab_events_v(npdrb_session *s, int32_t n_slabs, const int64_t *slab_starts, void *const *events,
                                           const int32_t *ready);
/* rows [row0, row0+nrows) of the distance matrix against all N instances (nrows == N: symmetric mode) */
int npdrb_session_distances(npdrb_session *s, int32_t metric, int32_t fast_euclidean, int64_t row0, int64_t nrows);
/* adopt a precomputed distance block (device, nrows x N row-major = the columns Ri of a symmetric D) */
/* Balanced symmetric distances for `world` row blocks (row_starts: world + 1 instance bounds, multiples of 128 except the
 * last = N): this rank computes its diagonal square and a checkerboard half of every off-diagonal square (~N^2/(2 world)
 * instance pairs instead of N^2/world); the other half arrives from the peers: _balanced_counts gives the per-peer message
 * sizes in doubles, _balanced_pack fills a device send buffer (peers in rank order), the host runs an all-to-all
 * (NCCL over NVLink), _balanced_unpack drops the received tiles into the block.  Bit-identical to the unsharded matrix. */
int npdrb_session_distances_balanced(npdrb_session *s, int32_t metric, int32_t fast_euclidean, const int64_t *row_starts,
                                     int32_t rank, int32_t world);
int npdrb_session_balanced_counts(npdrb_session *s, int64_t *send_doubles, int64_t *recv_doubles);   /* length world each */
int npdrb_session_balanced_pack(npdrb_session *s, double *d_send);
int npdrb_session_balanced_unpack(npdrb_session *s, const double *d_recv);
int npdrb_session_set_device_distances(npdrb_session *s, const double *dD, int64_t ldd, int64_t row0, int64_t nrows);
int npdrb_session_set_sd_vec(npdrb_session *s, const double *sd_vec_host);   /* user sd.vec, length N */
/* neighbour lists for the owned rows; *M_local = number of pairs found */
int npdrb_session_neighbors(npdrb_session *s, int32_t nbd_method, double sd_frac, int64_t k, int64_t *M_local,
                            int64_t *n_empty_local);
/* separate hit/miss neighbourhoods for the owned rows, class labels = the session's outcome vector
 * (multisurf and relieff shard by rows; surf needs the whole matrix on one device) */
int npdrb_session_neighbors_hitmiss(npdrb_session *s, int32_t nbd_method, double sd_frac, int64_t k, int64_t *M_local,
                                    int64_t *n_empty_local);
int npdrb_session_copy_radius2(npdrb_session *s, double *radius_host);      /* owned rows: miss radius */
/* k grid: sort the owned distance rows by (d, row) once (_sort_rows), then emit the ReliefF list of any k from the sorted rows
 * (_neighbors_sorted: the same list, bit for bit, as npdrb_session_neighbors(RELIEFF, k)). */
int npdrb_session_sort_rows(npdrb_session *s);
int npdrb_session_neighbors_sorted(npdrb_session *s, int64_t k, int64_t *M_local, int64_t *n_empty_local);
/* SURF needs one global mean/sd (R/nearestNeighbors.R:234-241).  Partial moments of the owned rows:
 * out4[0] = sum of all owned entries, out4[1] = sum over the owned part of the strict upper triangle,
 * out4[2] = its count, out4[3] = sum over that part of (d - centre)^2.  Multi-rank hosts call it with
 * centre = 0, combine to the global mean, call again with that mean, then set the radius
 *   sum_all/(2*num.pair) - sd.frac*sqrt(sumsq/(count-1))   with npdrb_session_set_surf_radius(). */
int npdrb_session_surf_partial(npdrb_session *s, double centre, double *out4);
int npdrb_session_set_surf_radius(npdrb_session *s, double radius);
/* The pair list must be the reference's bit for bit, and R's radius is one sequential long-double chain over N^2 terms.
 * Protocol for row blocks / large N: combine the partial moments (m0 at centre 0, m1 at centre m0[1]/m0[2]) into a radius
 * and a rigorous bound eps on its distance from R's value (_from_moments); count the distances of every block inside
 * [radius - eps, radius + eps] (_band_count); if the total is 0 every radius in the band selects the same pairs and the
 * moment radius is used; otherwise gather the matrix on the host and replay R's chains exactly (_exact_host: D_host is the
 * full symmetric N x N matrix, leading dimension ldd). */
int npdrb_surf_radius_from_moments(const double *m0, const double *m1, int64_t N, double sd_frac, double *radius, double *eps);
int npdrb_session_surf_band_count(npdrb_session *s, double lo, double hi, int64_t *count);
int npdrb_surf_radius_exact_host(const double *D_host, int64_t ldd, int64_t N, double sd_frac, double *radius);
/* device pointers to the local pair segment (int32, 1-based), valid until the next neighbours/import call */
int npdrb_session_local_pairs(npdrb_session *s, const int32_t **dRi, const int32_t **dNN, int64_t *M_local);
/* copy the local pair segment into caller-owned DEVICE buffers (e.g. torch tensors that NCCL will all-gather) */
int npdrb_session_copy_local_pairs_device(npdrb_session *s, int32_t *dRi_dst, int32_t *dNN_dst);
int npdrb_session_copy_radius(npdrb_session *s, double *radius_host);       /* owned rows */
int npdrb_session_copy_distances(npdrb_session *s, double *D_host);         /* nrows x N, row-major */
/* the full pair list the regression will use (device pointers, e.g. the all-gathered segments) */
int npdrb_session_import_pairs(npdrb_session *s, const int32_t *dRi, const int32_t *dNN, int64_t M);
int npdrb_session_import_pairs_host(npdrb_session *s, const int32_t *Ri, const int32_t *NN, int64_t M);
/* uniqueNeighbors on the imported list: count, and optionally replace the list by its unique subset.
 * Precondition (checked on the device, NPDRB_EINVAL otherwise): rows grouped by ascending Ri_idx, no ordered pair twice --
 * the shape nearestNeighbors() emits; only then is "drop (i, j) when i > j and (j, i) is present" the reference's
 * keep-the-first-occurrence rule (R/nearestNeighbors.R:576-582). */
int npdrb_session_unique_pairs(npdrb_session *s, int32_t keep_unique_only, int64_t *n_unique);
/* regression over attribute columns [attr0, attr0+nattr); stats_out host, nattr x 8 row-major */
int npdrb_session_regress(npdrb_session *s, int64_t attr0, int64_t nattr, int32_t regression_type,
                          int32_t attr_diff_type, const int32_t *covar_diff_type, double *stats_out,
                          int32_t *irls_passes);
int npdrb_session_copy_pairs(npdrb_session *s, int32_t *Ri_host, int32_t *NN_host);  /* imported list */
/* milliseconds of the last call of each stage on this session (CUDA events on the context stream) */
int npdrb_session_stage_ms(npdrb_session *s, double *ms_distance, double *ms_neighbors, double *ms_prepare,
                           double *ms_regress);

/* kernel-level timings of the last session_regress call, measured with CUDA events on the launching stream:
 * out[0] ms of the first pass kernel (lm: the only pass; binomial: the pass at the mustart weights),
 * out[1] summed ms of the full IRLS pass kernels, out[2] their number, out[3] attribute x record evaluations
 * they executed (active attribute tiles x 128 x records), out[4] records in the pair stream. */
int npdrb_session_kernel_ms(npdrb_session *s, double *out5);
/* design the pass kernels of the last session_regress call evaluated: *q_eff = covariates per record (q, or q - 1 when a
 * match-mismatch covariate was folded into the intercept: its diff is 0 / 1, so the pair list is split by its value),
 * *n_streams = pair streams ({one-directional, mutual pairs} x {covariate value 0, 1}), *n_exact_stops = attributes whose
 * glm.fit stopping test fell inside the streamed deviance's error band and was decided from exactly re-evaluated deviances */
int npdrb_session_regress_design(npdrb_session *s, int32_t *q_eff, int32_t *n_streams, int32_t *n_exact_stops);

#ifdef __cplusplus
}
#endif
#endif /* NPDR_B200_H */
