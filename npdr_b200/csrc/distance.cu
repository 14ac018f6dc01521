// distance.cu -- K1: pairwise manhattan / euclidean distances in FP64 (npdrDistances,
// R/nearestNeighbors.R:63-101 -> stats::dist).
//
// Bound: FP64 issue rate (2 DADD per element pair for manhattan, no FMA/MMA form exists).
// Design (sm_100a):
//   * X stays in R's column-major layout: for a fixed attribute k the 128 instances of a tile
//     are 1 KB contiguous, which is exactly the k-major shared-memory layout the register tile
//     wants, so TMA (cp.async.bulk.tensor.2d) drops [BK x 128] boxes straight into place;
//     rows >= N and attributes >= p are zero-filled by the tensor map (|0-0| adds nothing).
//   * 128x128 output tile per CTA, 256 threads, 8x8 accumulators per thread (interleaved
//     2-wide chunks so every LDS.128 is conflict free), 4-stage mbarrier pipeline.
//   * every accumulator sums k = 0..p-1 strictly in order in double, with no FMA contraction
//     (unless fast_euclidean): the result is bit-identical to the sequential C loop of dist(),
//     which is what makes the downstream `d < radius` neighbour test exact.
//   * persistent CTAs (one per SM) walk the tile list; in symmetric mode only tiles on or
//     below the diagonal are computed and each is written twice (D and its mirror).
#include <cuda.h>
#include "common.cuh"

namespace {

constexpr int TM = 128;       // tile rows (i)
constexpr int TN = 128;       // tile cols (j)
constexpr int BK = 16;        // attributes per pipeline stage
constexpr int STAGES = 4;
constexpr int DTHREADS = 256;
constexpr int STAGE_DOUBLES = BK * TM;                 // one operand, one stage
constexpr size_t DIST_SMEM = (size_t)STAGES * 2 * STAGE_DOUBLES * sizeof(double) + 128 /*barriers*/ + 128 /*align*/;

enum { MODE_MANHATTAN = 0, MODE_EUCLID_EXACT = 1, MODE_EUCLID_FMA = 2, MODE_HAMMING = 3 };

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ double2 lds_v2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a TMA that never lands (bad tensor map, wrong byte count) traps after ~2 s instead of
// hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

template <int MODE>
__device__ __forceinline__ void accumulate(double &acc, double a, double b) {
    if (MODE == MODE_HAMMING) {
        // hamming.binary (R/utils.R:185-188): t(1 - X) %*% X + its transpose = sum (1-a) b + (1-b) a
        // (= the number of differing attributes for 0/1 data; every partial sum is then an exact integer)
        acc = __dadd_rn(acc, __dadd_rn(__dmul_rn(__dsub_rn(1.0, a), b), __dmul_rn(__dsub_rn(1.0, b), a)));
        return;
    }
    double d = __dsub_rn(a, b);
    if (MODE == MODE_MANHATTAN) acc = __dadd_rn(acc, fabs(d));
    else if (MODE == MODE_EUCLID_EXACT) acc = __dadd_rn(acc, __dmul_rn(d, d));
    else acc = __fma_rn(d, d, acc);
}

// tile index -> (it, jt)
__device__ __forceinline__ void tile_coords(int64_t t, int symmetric, int n_jt, int &it, int &jt) {
    if (symmetric) {
        int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while ((int64_t)(r + 1) * (r + 2) / 2 <= t) ++r;
        while ((int64_t)r * (r + 1) / 2 > t) --r;
        it = r;
        jt = (int)(t - (int64_t)r * (r + 1) / 2);
    } else {
        it = (int)(t / n_jt);
        jt = (int)(t % n_jt);
    }
}

// D block layout: Dout[(i - row0) * ldd + j] for owned rows i in [row0, row0 + nrows), all j in [0, N).
template <int MODE>
__global__ void __launch_bounds__(DTHREADS, 1)
dist_tma_kernel(const __grid_constant__ CUtensorMap tmap, int64_t N, int64_t p, int64_t row0, int64_t nrows,
                int symmetric, int64_t n_tiles, int n_jt, double *__restrict__ Dout, int64_t ldd) {
    extern __shared__ unsigned char smem_raw[];
    // explicit 32-bit shared-window addresses (keeps every access an LDS / SYNCS, never a generic LD)
    constexpr uint32_t STAGE_BYTES = STAGE_DOUBLES * sizeof(double);
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t sA = base;                                      // [STAGES][BK][TM]
    const uint32_t sB = base + STAGES * STAGE_BYTES;               // [STAGES][BK][TN]
    const uint32_t full = base + 2 * STAGES * STAGE_BYTES;         // STAGES mbarriers

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(full + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_chunks = (int)((p + BK - 1) / BK);
    uint32_t g = 0;                                                // chunks consumed so far (all tiles)

    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        int it, jt;
        tile_coords(t, symmetric, n_jt, it, jt);
        const int i0 = (int)row0 + it * TM;                        // global instance index of the tile's first row
        const int j0 = jt * TN;
        const bool diag = (i0 == j0);
        const uint32_t tx_bytes = (uint32_t)(STAGE_DOUBLES * sizeof(double)) * (diag ? 1u : 2u);

        // prologue: fill STAGES-1 stages (all threads passed the __syncthreads that ended the previous tile)
        if (tid == 0) {
            for (int c = 0; c < STAGES - 1 && c < n_chunks; ++c) {
                int s = (int)((g + c) % STAGES);
                mbar_expect_tx(full + 8 * s, tx_bytes);
                tma_load_2d(sA + s * STAGE_BYTES, &tmap, i0, c * BK, full + 8 * s);
                if (!diag) tma_load_2d(sB + s * STAGE_BYTES, &tmap, j0, c * BK, full + 8 * s);
            }
        }

        double acc[8][8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.0;

        for (int c = 0; c < n_chunks; ++c) {
            // refill the stage freed by the previous iteration
            if (tid == 0 && c + STAGES - 1 < n_chunks) {
                int cn = c + STAGES - 1;
                int s = (int)((g + cn) % STAGES);
                mbar_expect_tx(full + 8 * s, tx_bytes);
                tma_load_2d(sA + s * STAGE_BYTES, &tmap, i0, cn * BK, full + 8 * s);
                if (!diag) tma_load_2d(sB + s * STAGE_BYTES, &tmap, j0, cn * BK, full + 8 * s);
            }
            const uint32_t gc = g + (uint32_t)c;
            const int s = (int)(gc % STAGES);
            mbar_wait(full + 8 * s, (gc / STAGES) & 1u);
            const uint32_t As = sA + s * STAGE_BYTES + ty * 16;
            const uint32_t Bs = (diag ? sA : sB) + s * STAGE_BYTES + tx * 16;
#pragma unroll 4
            for (int kk = 0; kk < BK; ++kk) {
                double a[8], b[8];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    double2 av = lds_v2(As + kk * (TM * 8) + h * 256);
                    double2 bv = lds_v2(Bs + kk * (TN * 8) + h * 256);
                    a[2 * h] = av.x; a[2 * h + 1] = av.y;
                    b[2 * h] = bv.x; b[2 * h + 1] = bv.y;
                }
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) accumulate<MODE>(acc[r][cc], a[r], b[cc]);
            }
            __syncthreads();                                       // stage s is free for the refill of the next iteration
        }
        g += (uint32_t)n_chunks;

        // epilogue: rows i = i0 + (r>>1)*32 + ty*2 + (r&1), cols j = j0 + (c>>1)*32 + tx*2 + (c&1)
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int64_t i = (int64_t)i0 + (r >> 1) * 32 + ty * 2 + (r & 1);
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                const int64_t j = (int64_t)j0 + (cc >> 1) * 32 + tx * 2 + (cc & 1);
                double v = acc[r][cc];
                if (MODE == MODE_EUCLID_EXACT || MODE == MODE_EUCLID_FMA) v = sqrt(v);
                if (i < row0 + nrows && i < N && j < N) {
                    Dout[(i - row0) * ldd + j] = v;
                    if (symmetric && !diag) Dout[j * ldd + i] = v;      // row0 == 0 in symmetric mode
                }
            }
        }
    }
}

// Reference-simple variant (no TMA, no pipeline): same per-accumulator arithmetic order, used to
// validate the TMA kernel on hardware (NPDRB_DIST_KERNEL=simple) and as a diagnostic fallback.
template <int MODE>
__global__ void __launch_bounds__(256)
dist_simple_kernel(const double *__restrict__ X, int64_t ldx, int64_t N, int64_t p, int64_t row0, int64_t nrows,
                   double *__restrict__ Dout, int64_t ldd) {
    __shared__ double sA[16][64];
    __shared__ double sB[16][64];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t i0 = row0 + (int64_t)blockIdx.y * 64, j0 = (int64_t)blockIdx.x * 64;
    double acc[4][4] = {};
    for (int64_t k0 = 0; k0 < p; k0 += 16) {
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            int kk = e >> 6, r = e & 63;
            int64_t k = k0 + kk;
            sA[kk][r] = (k < p && i0 + r < N) ? X[k * ldx + i0 + r] : 0.0;
            sB[kk][r] = (k < p && j0 + r < N) ? X[k * ldx + j0 + r] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk)
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) accumulate<MODE>(acc[r][c], sA[kk][ty * 4 + r], sB[kk][tx * 4 + c]);
        __syncthreads();
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            int64_t i = i0 + ty * 4 + r, j = j0 + tx * 4 + c;
            double v = acc[r][c];
            if (MODE == MODE_EUCLID_EXACT || MODE == MODE_EUCLID_FMA) v = sqrt(v);
            if (i < row0 + nrows && i < N && j < N) Dout[(i - row0) * ldd + j] = v;
        }
}

// column pre-scalings of npdrDistances (R/nearestNeighbors.R:76-95): one CTA per attribute column
__global__ void prescale_kernel(const double *__restrict__ X, int64_t ldx, int64_t N, int metric,
                                double *__restrict__ Xs) {
    const int64_t k = blockIdx.x;
    const double *c = X + k * ldx;
    double *o = Xs + k * ldx;
    if (metric == NPDRB_METRIC_ALLELE_SHARING_MANHATTAN) {
        for (int64_t i = threadIdx.x; i < ldx; i += blockDim.x) o[i] = (i < N) ? c[i] / 2.0 : 0.0;
        return;
    }
    __shared__ double smin[256], smax[256];
    double mn = c[0], mx = c[0];
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) { mn = fmin(mn, c[i]); mx = fmax(mx, c[i]); }
    smin[threadIdx.x] = mn; smax[threadIdx.x] = mx;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            smin[threadIdx.x] = fmin(smin[threadIdx.x], smin[threadIdx.x + s]);
            smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + s]);
        }
        __syncthreads();
    }
    mn = smin[0];
    const double range = __dsub_rn(smax[0], mn);
    for (int64_t i = threadIdx.x; i < ldx; i += blockDim.x) o[i] = (i < N) ? __ddiv_rn(__dsub_rn(c[i], mn), range) : 0.0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tensor_map(CUtensorMap *map, const double *X, int64_t N, int64_t p, int64_t ldx) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) {
            npdrb_set_error("cuTensorMapEncodeTiled entry point unavailable");
            return NPDRB_ECUDA;
        }
        fn = (PFN_encodeTiled)ptr;
    }
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)p};
    cuuint64_t strides[1] = {(cuuint64_t)ldx * sizeof(double)};
    cuuint32_t box[2] = {(cuuint32_t)TM, (cuuint32_t)BK};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)X, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        npdrb_set_error("cuTensorMapEncodeTiled failed (CUresult %d; N=%lld p=%lld ldx=%lld)", (int)r, (long long)N,
                        (long long)p, (long long)ldx);
        return NPDRB_ECUDA;
    }
    return NPDRB_OK;
}

template <int MODE>
int launch_tma(npdrb_session *s, const double *X, int64_t row0, int64_t nrows, int symmetric) {
    npdrb_ctx *ctx = s->ctx;
    CUtensorMap map;
    NB_CHECK(make_tensor_map(&map, X, s->N, s->p, s->ldx));
    const int n_it = (int)nb_cdiv(nrows, TM), n_jt = (int)nb_cdiv(s->N, TN);
    const int64_t n_tiles = symmetric ? (int64_t)n_it * (n_it + 1) / 2 : (int64_t)n_it * n_jt;
    NB_CUDA(cudaFuncSetAttribute(dist_tma_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIST_SMEM));
    int grid = (int)((n_tiles < ctx->sm_count) ? n_tiles : ctx->sm_count);
    if (grid < 1) grid = 1;
    dist_tma_kernel<MODE><<<grid, DTHREADS, DIST_SMEM, ctx->stream>>>(map, s->N, s->p, row0, nrows, symmetric, n_tiles,
                                                                      n_jt, s->dD, s->ldd);
    NB_LAUNCH_CHECK(ctx);
    return NPDRB_OK;
}

template <int MODE>
int launch_simple(npdrb_session *s, const double *X, int64_t row0, int64_t nrows) {
    dim3 grid((unsigned)nb_cdiv(s->N, 64), (unsigned)nb_cdiv(nrows, 64));
    dist_simple_kernel<MODE><<<grid, 256, 0, s->ctx->stream>>>(X, s->ldx, s->N, s->p, row0, nrows, s->dD, s->ldd);
    NB_LAUNCH_CHECK(s->ctx);
    return NPDRB_OK;
}

}  // namespace

int nb_distances(npdrb_session *s, int32_t metric, int32_t fast_euclidean, int64_t row0, int64_t nrows) {
    npdrb_ctx *ctx = s->ctx;
    if (!s->dX) { npdrb_set_error("distances: no attribute matrix in the session"); return NPDRB_EINVAL; }
    if ((metric < 0 || metric > NPDRB_METRIC_RELIEF_SCALED_EUCLIDEAN) && metric != NPDRB_METRIC_HAMMING) {
        npdrb_set_error("distances: unsupported metric %d", metric);
        return NPDRB_EINVAL;
    }
    if (row0 < 0 || nrows < 1 || row0 + nrows > s->N) { npdrb_set_error("distances: bad row block"); return NPDRB_EINVAL; }
    if (row0 % TM != 0) { npdrb_set_error("distances: row0 must be a multiple of %d", TM); return NPDRB_EINVAL; }
    const int symmetric = (row0 == 0 && nrows == s->N) ? 1 : 0;

    if (s->dD && s->own_D && (s->nrows != nrows)) { cudaFree(s->dD); s->dD = nullptr; }
    if (!s->dD || !s->own_D) {
        s->ldd = nb_round_up(s->N, 2);
        NB_CUDA(cudaMalloc(&s->dD, sizeof(double) * (size_t)nrows * (size_t)s->ldd));
        s->own_D = true;
    }
    s->row0 = row0; s->nrows = nrows;

    NB_CUDA(cudaEventRecord(s->ev[0], ctx->stream));
    const double *X = s->dX;
    if (metric >= NPDRB_METRIC_ALLELE_SHARING_MANHATTAN && metric <= NPDRB_METRIC_RELIEF_SCALED_EUCLIDEAN) {
        if (!s->dXs) NB_CUDA(cudaMalloc(&s->dXs, sizeof(double) * (size_t)s->ldx * (size_t)s->p));
        prescale_kernel<<<(unsigned)s->p, 256, 0, ctx->stream>>>(s->dX, s->ldx, s->N, metric, s->dXs);
        NB_LAUNCH_CHECK(ctx);
        X = s->dXs;
    }
    const bool euclid = (metric == NPDRB_METRIC_EUCLIDEAN || metric == NPDRB_METRIC_RELIEF_SCALED_EUCLIDEAN);
    const int mode = metric == NPDRB_METRIC_HAMMING ? MODE_HAMMING
                   : !euclid ? MODE_MANHATTAN : (fast_euclidean ? MODE_EUCLID_FMA : MODE_EUCLID_EXACT);
    const char *kenv = getenv("NPDRB_DIST_KERNEL");
    const bool simple = kenv && strcmp(kenv, "simple") == 0;
    int rc;
    if (simple) {
        rc = mode == MODE_MANHATTAN ? launch_simple<MODE_MANHATTAN>(s, X, row0, nrows)
           : mode == MODE_HAMMING ? launch_simple<MODE_HAMMING>(s, X, row0, nrows)
           : mode == MODE_EUCLID_EXACT ? launch_simple<MODE_EUCLID_EXACT>(s, X, row0, nrows)
                                       : launch_simple<MODE_EUCLID_FMA>(s, X, row0, nrows);
    } else {
        rc = mode == MODE_MANHATTAN ? launch_tma<MODE_MANHATTAN>(s, X, row0, nrows, symmetric)
           : mode == MODE_HAMMING ? launch_tma<MODE_HAMMING>(s, X, row0, nrows, symmetric)
           : mode == MODE_EUCLID_EXACT ? launch_tma<MODE_EUCLID_EXACT>(s, X, row0, nrows, symmetric)
                                       : launch_tma<MODE_EUCLID_FMA>(s, X, row0, nrows, symmetric);
    }
    NB_CHECK(rc);
    NB_CUDA(cudaEventRecord(s->ev[1], ctx->stream));
    NB_CUDA(cudaEventSynchronize(s->ev[1]));
    float ms = 0;
    NB_CUDA(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
    s->ms_distance = ms;
    return NPDRB_OK;
}
