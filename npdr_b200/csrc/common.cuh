// common.cuh -- shared plumbing of libnpdr_b200.so (context, session, error handling).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/npdr_b200.h"

void npdrb_set_error(const char *fmt, ...);

// Device-memory cache (api.cu).  npdr() is called in loops (vwok's k grid, benchmarks) with identical buffer sizes;
// raw cudaMalloc/cudaFree of multi-hundred-MB buffers cost ~100+ ms per call in page mapping and implicit device
// synchronisation.  All library work on a device is issued to ONE stream, so a freed block may be handed out again
// immediately: the next user is ordered after the previous one on that stream.  Every cudaMalloc/cudaFree in this
// library goes through the cache; npdrb_release_cache() / npdrb_destroy() return the memory to the driver.
cudaError_t nb_cached_malloc(void **p, size_t bytes);
cudaError_t nb_cached_free(void *p);
#define cudaMalloc(p, n) nb_cached_malloc((void **)(p), (n))
#define cudaFree(p) nb_cached_free((void *)(p))

#define NB_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            npdrb_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return NPDRB_ECUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define NB_CHECK(call)                 \
    do {                               \
        int _rc = (call);              \
        if (_rc != NPDRB_OK) return _rc; \
    } while (0)

#define NB_LAUNCH_CHECK(ctx)                                                                   \
    do {                                                                                       \
        (ctx)->launches++;                                                                     \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            npdrb_set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return NPDRB_ECUDA;                                                                \
        }                                                                                      \
    } while (0)

// The streaming regression kernels (regress.cu) hold q <= NB_FAST_COVARS covariates in registers; wider designs go through
// regress_wide.cu (up to NPDRB_MAX_COVARS).
#define NB_FAST_COVARS 4
// pair streams of one regression: {weight 1, weight 2 (folded mutual pairs)} x {binary covariate 0, 1} (regress.cu)
#define NB_MAX_STREAMS 4

// Device buffers that live for one call: everything allocated through a scope is returned to the cache when the scope ends,
// early error returns included.
struct nb_scope {
    std::vector<void *> bufs;
    nb_scope() = default;
    nb_scope(const nb_scope &) = delete;
    nb_scope &operator=(const nb_scope &) = delete;
    template <typename T>
    int alloc(T **p, size_t n) {
        const size_t bytes = sizeof(T) * (n ? n : 1);
        cudaError_t e = cudaMalloc((void **)p, bytes);
        if (e != cudaSuccess) { npdrb_set_error("cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e)); *p = nullptr; return NPDRB_ENOMEM; }
        bufs.push_back((void *)*p);
        return NPDRB_OK;
    }
    ~nb_scope() { for (void *b : bufs) cudaFree(b); }
};

static inline int64_t nb_round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }
static inline int64_t nb_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

struct npdrb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;   // host->device slabs of a large X travel here while the distance kernel consumes earlier ones
    int sm_count = 148;
    int64_t launches = 0;
    size_t smem_optin = 0;
    struct nb_stage_ring *ring = nullptr; // pinned staging ring + memcpy helpers for pageable host sources (api.cu)
};

// One bucketed pair stream for the regression kernels (see regress.cu).
struct nb_stream {
    int64_t n_rec = 0;            // records in this stream
    double weight = 1.0;          // multiplicity applied to every record's contribution
    double fold_val = 0.0;        // value (0 / 1) of the binary covariate folded into the intercept for this stream
    uint32_t *rec = nullptr;      // il | jl << 8 | ybit << 16, bucketed by (I-block, J-block) tile
    double *yv = nullptr;         // lm: |y_i - y_j| per record
    double *cd[NB_FAST_COVARS] = {nullptr, nullptr, nullptr, nullptr};  // covariate diffs per record
    int32_t n_tiles = 0;          // non-empty tiles
    int64_t *tile_ptr = nullptr;  // [n_tiles + 1] record offsets
    int64_t *tile_mid = nullptr;  // [n_tiles] offset of the first miss (outcome bit 1) record of each tile
    int32_t *tile_ib = nullptr, *tile_jb = nullptr;
    int32_t n_chunks = 0;
    int32_t *chunk_tile = nullptr;  // [n_chunks + 1] tile ranges per work chunk
};

struct npdrb_session {
    npdrb_ctx *ctx = nullptr;
    int64_t N = 0, p = 0;
    int32_t q = 0;
    // inputs
    double *dX = nullptr; int64_t ldx = 0; bool own_X = false;          // column-major, N x p, leading dim ldx
    double *dXs = nullptr;                                               // pre-scaled copy for scaled metrics
    // pipelined upload (npdrb_session_upload of a large X): column slabs in flight on ctx->copy_stream; slab i covers
    // attributes [i * slab_cols, ...) and is complete once slab_ev[i] has fired.  nb_wait_upload() joins them.
    // A PAGEABLE host matrix (what R hands over) is not copied by npdrb_session_upload at all: hX_lazy keeps the caller's
    // pointer and nb_issue_slab() stages slab i through the context's pinned ring (helper threads memcpy, async H2D) right
    // before the consumer of slab i is launched, so the copy of slab i+1 runs under the distance kernel of slab i.
    std::vector<cudaEvent_t> slab_ev; int64_t slab_cols = 0; int n_slabs_pending = 0;
    std::vector<int64_t> slab_start;   // n_slabs_pending + 1 column bounds of the slabs (multiples of 16 except the last = p)
    const volatile int32_t *slab_ready = nullptr;   // optional host flags: slab_ev[i] may only be waited for once slab_ready[i] != 0
    bool slab_ev_external = false;     // slab_ev are the caller's events (npdrb_session_set_device_slab_events): never issued or destroyed here
    const double *hX_lazy = nullptr; int n_slabs_issued = 0;
    double *dy = nullptr; bool own_y = false;
    double *dC = nullptr; bool own_C = false;                            // N x q, ld N
    // distance block: nrows x N, row-major with leading dim ldd (row r = column row0+r of the symmetric D)
    double *dD = nullptr; int64_t ldd = 0; int64_t row0 = 0, nrows = 0; bool own_D = false;
    std::vector<int64_t> bal_tile_starts; int bal_rank = -1;            // partition of the last balanced distance pass
    bool rect = false;                                                   // rows are a second instance set (nearestNeighbors2)
    double *d_sd_vec = nullptr;                                          // optional user sd.vec (length N)
    double *d_radius = nullptr;                                          // owned rows
    double *d_radius2 = nullptr;                                         // owned rows: miss radius (separate hit/miss neighbourhoods)
    bool have_surf_radius = false; double surf_radius = 0.0;
    // local neighbour segment
    int32_t *dRiL = nullptr, *dNNL = nullptr; int64_t M_local = 0; int64_t n_empty_local = 0;
    int64_t *d_rowptr = nullptr;
    // k grid (kgrid.cu): the owned rows sorted once by (d, row), same leading dimension as the distance block
    double *d_sortD = nullptr; int32_t *d_sortJ = nullptr; bool sorted_valid = false; double ms_sort = 0;
    // list used by the regression
    int32_t *dRi = nullptr, *dNN = nullptr; int64_t M = 0; bool own_pairs = false;
    // regression workspace
    double *dXT = nullptr; int64_t ldt = 0; int64_t xt_attr0 = -1, xt_nattr = 0;
    nb_stream streams[NB_MAX_STREAMS]; int n_streams = 0;
    int32_t streams_reg_type = -1; int32_t streams_cov_sig = -1; bool streams_valid = false;
    // timing
    cudaEvent_t ev[2] = {nullptr, nullptr};
    double ms_distance = 0, ms_neighbors = 0, ms_prepare = 0, ms_regress = 0;
    // kernel-level timings of the last regress call (CUDA events around the pass-kernel launches)
    double ms_pass_first = 0;      // lm: the single pass; binomial: the pass at the mustart weights
    double ms_pass_full = 0;       // binomial: sum over the full IRLS passes
    int n_pass_full = 0;
    double evals_pass_full = 0;    // attribute x record evaluations executed by those passes (active tiles only)
    int n_exact_stops = 0;         // attributes whose stopping test was decided from exactly re-evaluated deviances (last regress call)
    int q_eff = 0;                 // covariates of the design the pass kernels evaluated (q, or q - 1 with a folded binary covariate)
};

// stage implementations (one .cu each)
int nb_wait_upload(npdrb_session *s);     // main stream waits for every slab; NA/NaN/Inf check of X (NPDRB_EDATA)
int nb_issue_slab(npdrb_session *s, int slab);   // make sure the H2D copy of column slab `slab` has been enqueued (slab_ev[slab] recorded)
// staged host -> device copy of `ncols` contiguous columns of `rows` doubles (pageable source) through the context's pinned
// ring on ctx->copy_stream; returns once every chunk has been ENQUEUED (the host source may not be touched until the
// copy stream has drained, which the slab event / a stream synchronize guarantees to the callers)
bool nb_host_pointer_is_pinned(const void *p);
int nb_staged_h2d(npdrb_ctx *ctx, double *dst, int64_t ld_dst, const double *src, int64_t rows, int64_t ncols);
int nb_check_finite(npdrb_ctx *ctx, const double *d, int64_t n, const char *what);
int nb_distances(npdrb_session *s, int32_t metric, int32_t fast_euclidean, int64_t row0, int64_t nrows);
int nb_distances_balanced(npdrb_session *s, int32_t metric, int32_t fast_euclidean, const int64_t *row_starts, int32_t rank,
                          int32_t world);
int nb_balanced_counts(npdrb_session *s, int64_t *send_doubles, int64_t *recv_doubles);
int nb_balanced_pack(npdrb_session *s, double *d_send);
int nb_balanced_unpack(npdrb_session *s, const double *d_recv);
int nb_distances2(npdrb_ctx *ctx, const double *dX1, int64_t ld1, int64_t m1, const double *dX2, int64_t ld2, int64_t m2,
                  int64_t p, int32_t metric, int32_t fast_euclidean, double *dD, int64_t ldd);
int nb_neighbors(npdrb_session *s, int32_t nbd_method, double sd_frac, int64_t k, int64_t *M_local, int64_t *n_empty);
int nb_neighbors_hitmiss(npdrb_session *s, int32_t nbd_method, double sd_frac, int64_t k, int64_t *M_local, int64_t *n_empty);
int nb_sort_rows(npdrb_session *s);
int nb_neighbors_sorted(npdrb_session *s, int64_t k, int64_t *M_local, int64_t *n_empty);
int nb_unique_pairs(npdrb_session *s, int32_t keep_unique_only, int64_t *n_unique);
int nb_regress(npdrb_session *s, int64_t attr0, int64_t nattr, int32_t regression_type, int32_t attr_diff_type,
               const int32_t *covar_diff_type, double *stats_out_host, int32_t *irls_passes);
int nb_diff_vectors(npdrb_ctx *ctx, const double *da, const double *db, int64_t n, int32_t type, double norm_fac,
                    double *dout);
int nb_measure_fp64_peak(npdrb_ctx *ctx, double *dfma_tflops, double *dadd_tflops);
void nb_free_streams(npdrb_session *s);
int nb_regress_wide(npdrb_session *s, int64_t attr0, int64_t nattr, int32_t regression_type, int32_t attr_diff_type,
                    const int32_t *covar_diff_type, double *stats_out_host, int32_t *irls_passes);

// diff function shared by the pair kernels (npdrDiff, R/nearestNeighbors.R:15-40, norm.fac = 1)
__device__ __forceinline__ double nb_diff1(double a, double b, int type) {
    switch (type) {
    case NPDRB_DIFF_NUMERIC_SQR: { double t = fabs(a - b); return t * t; }
    case NPDRB_DIFF_ALLELE_SHARING: return fabs(a - b) * 0.5;
    case NPDRB_DIFF_MATCH_MISMATCH: return (a == b) ? 0.0 : 1.0;
    case NPDRB_DIFF_RAW: return a;
    default: return fabs(a - b);
    }
}
