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
constexpr int TN = 128;       // tile cols (j) of the wide kernel; the narrow variant uses 64 (twice the tiles: fills the SMs of
                              // a row block that has fewer than ~2 waves of 128 x 128 tiles)
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

// The host hands the kernel an explicit tile list: .x = row tile (relative to row0), .y = column tile, bit 30 of .y set
// when the tile's transpose also belongs to this block (symmetric matrix: D[j, i] = D[i, j] bit for bit, since
// |a - b| = |b - a| and (a - b)^2 = (b - a)^2 exactly).  Three list shapes are used (build_tiles below):
//   full symmetric matrix  : tiles on or below the diagonal, mirrored              (one device)
//   row block, every column: no mirroring                                         (row-sharded, or rectangular A x B)
//   balanced row block     : the block's own diagonal square (lower half, mirrored) plus HALF of every off-diagonal
//                            square by a parity rule; the owner of the other half sends its tiles over NVLink
//                            (pack/unpack kernels below) -- every device computes ~N^2/(2G) pairs instead of N^2/G.
constexpr int TILE_MIRROR = 1 << 30;

// D block layout: Dout[(i - row0) * ldd + j] for owned rows i in [row0, row0 + nrows), all j in [0, N).
// same_ab: the row and column operands come from the same matrix (tile i0 == j0 is then a diagonal tile and is
// loaded once); otherwise tmapA describes the row-side matrix and tmapB the column-side one (npdrDistances2).
template <int MODE, int TNV>
__global__ void __launch_bounds__(DTHREADS, 1)
dist_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmapB, int same_ab,
                int64_t Nrows_total, int64_t N, int64_t p, int64_t row0, int64_t nrows,
                const int2 *__restrict__ tiles, int64_t n_tiles, double *__restrict__ Dout, int64_t ldd,
                int chunk_lo, int chunk_hi, int resume, int finalize) {
    // Attribute range [chunk_lo, chunk_hi) * BK of this launch.  resume: the accumulators start from the partial sums a
    // previous launch left in Dout (same per-element order k = 0..p-1, so slab-wise launches are bit-identical to one
    // launch); finalize: last range -- euclidean takes its square root only then.
    extern __shared__ unsigned char smem_raw[];
    // explicit 32-bit shared-window addresses (keeps every access an LDS / SYNCS, never a generic LD)
    constexpr uint32_t STAGE_BYTES = STAGE_DOUBLES * sizeof(double);
    constexpr uint32_t STAGE_BYTES_B = BK * TNV * sizeof(double);
    constexpr int NC = TNV / 16;                                   // accumulator columns per thread (8 or 4)
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t sA = base;                                      // [STAGES][BK][TM]
    const uint32_t sB = base + STAGES * STAGE_BYTES;               // [STAGES][BK][TNV]
    const uint32_t full = base + 2 * STAGES * STAGE_BYTES;         // STAGES mbarriers

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(full + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_chunks = chunk_hi - chunk_lo;
    uint32_t g = 0;                                                // chunks consumed so far (all tiles)

    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int2 tc = tiles[t];
        const int it = tc.x, jt = tc.y & (TILE_MIRROR - 1);
        const bool mirror = (tc.y & TILE_MIRROR) != 0;
        const int i0 = (int)row0 + it * TM;                        // global instance index of the tile's first row
        const int j0 = jt * TNV;
        const bool diag = same_ab && (j0 >= i0) && (j0 + TNV <= i0 + TM);     // the column range lies inside the row range: one load
        const uint32_t tx_bytes = STAGE_BYTES + (diag ? 0u : STAGE_BYTES_B);

        // prologue: fill STAGES-1 stages (all threads passed the __syncthreads that ended the previous tile)
        if (tid == 0) {
            for (int c = 0; c < STAGES - 1 && c < n_chunks; ++c) {
                int s = (int)((g + c) % STAGES);
                mbar_expect_tx(full + 8 * s, tx_bytes);
                tma_load_2d(sA + s * STAGE_BYTES, &tmap, i0, (chunk_lo + c) * BK, full + 8 * s);
                if (!diag) tma_load_2d(sB + s * STAGE_BYTES_B, &tmapB, j0, (chunk_lo + c) * BK, full + 8 * s);
            }
        }

        double acc[8][NC];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int64_t i = (int64_t)i0 + (r >> 1) * 32 + ty * 2 + (r & 1);
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int64_t j = (int64_t)j0 + (c >> 1) * 32 + tx * 2 + (c & 1);
                acc[r][c] = (resume && i < row0 + nrows && i < Nrows_total && j < N) ? Dout[(i - row0) * ldd + j] : 0.0;
            }
        }

        for (int c = 0; c < n_chunks; ++c) {
            // refill the stage freed by the previous iteration
            if (tid == 0 && c + STAGES - 1 < n_chunks) {
                int cn = c + STAGES - 1;
                int s = (int)((g + cn) % STAGES);
                mbar_expect_tx(full + 8 * s, tx_bytes);
                tma_load_2d(sA + s * STAGE_BYTES, &tmap, i0, (chunk_lo + cn) * BK, full + 8 * s);
                if (!diag) tma_load_2d(sB + s * STAGE_BYTES_B, &tmapB, j0, (chunk_lo + cn) * BK, full + 8 * s);
            }
            const uint32_t gc = g + (uint32_t)c;
            const int s = (int)(gc % STAGES);
            mbar_wait(full + 8 * s, (gc / STAGES) & 1u);
            const uint32_t As = sA + s * STAGE_BYTES + ty * 16;
            const uint32_t Bs = (diag ? sA + s * STAGE_BYTES + (uint32_t)(j0 - i0) * 8u : sB + s * STAGE_BYTES_B) + tx * 16;
            const uint32_t bstride = diag ? (uint32_t)(TM * 8) : (uint32_t)(TNV * 8);
#pragma unroll 4
            for (int kk = 0; kk < BK; ++kk) {
                double a[8], b[NC];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    double2 av = lds_v2(As + kk * (TM * 8) + h * 256);
                    a[2 * h] = av.x; a[2 * h + 1] = av.y;
                }
#pragma unroll
                for (int h = 0; h < NC / 2; ++h) {
                    double2 bv = lds_v2(Bs + kk * bstride + h * 256);
                    b[2 * h] = bv.x; b[2 * h + 1] = bv.y;
                }
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc) accumulate<MODE>(acc[r][cc], a[r], b[cc]);
            }
            __syncthreads();                                       // stage s is free for the refill of the next iteration
        }
        g += (uint32_t)n_chunks;

        // epilogue: rows i = i0 + (r>>1)*32 + ty*2 + (r&1), cols j = j0 + (c>>1)*32 + tx*2 + (c&1)
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int64_t i = (int64_t)i0 + (r >> 1) * 32 + ty * 2 + (r & 1);
#pragma unroll
            for (int cc = 0; cc < NC; ++cc) {
                const int64_t j = (int64_t)j0 + (cc >> 1) * 32 + tx * 2 + (cc & 1);
                double v = acc[r][cc];
                if ((MODE == MODE_EUCLID_EXACT || MODE == MODE_EUCLID_FMA) && finalize) v = sqrt(v);
                if (i < row0 + nrows && i < Nrows_total && j < N) {
                    Dout[(i - row0) * ldd + j] = v;
                    if (mirror && !diag && j >= row0 && j < row0 + nrows) Dout[(j - row0) * ldd + i] = v;
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

int make_tensor_map(CUtensorMap *map, const double *X, int64_t N, int64_t p, int64_t ldx, int box_rows = TM) {
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
    cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)BK};
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

// Which device computes tile (I, J) of a symmetric matrix whose tile rows are partitioned over ranks (I and J owned by
// different ranks)?  The owner of the smaller tile index when I + J is even, the owner of the larger one otherwise:
// every off-diagonal square is split checkerboard-fashion between its two owners.
inline bool balanced_row_owner_computes(int I, int J) { return ((I < J) == (((I + J) & 1) == 0)); }

struct TilePlan {
    std::vector<int2> tiles;       // what this block computes
    int tn = TN;                   // tile width the list was built for
};

// list shapes: 0 = full symmetric, 1 = row block x all columns, 2 = balanced row block (tile_starts: G + 1 tile-row bounds)
// tn: tile width (128 or 64); ownership rules are stated on 128 x 128 squares, each square is 128 / tn column tiles.
int build_tiles(int shape, int64_t row0, int64_t nrows, int64_t ncols, const std::vector<int64_t> *tile_starts, int rank,
                int tn, TilePlan &plan) {
    const int n_it = (int)nb_cdiv(nrows, TM), n_J = (int)nb_cdiv(ncols, TM), sub = TM / tn;
    const int it0 = (int)(row0 / TM);
    plan.tiles.clear();
    plan.tn = tn;
    auto push = [&](int it, int J, int flag) {
        for (int h = 0; h < sub; ++h)
            if ((int64_t)(J * sub + h) * tn < ncols) plan.tiles.push_back(make_int2(it, (J * sub + h) | flag));
    };
    if (shape == 0) {
        for (int it = 0; it < n_it; ++it)
            for (int J = 0; J <= it; ++J) push(it, J, TILE_MIRROR);
    } else if (shape == 1) {
        for (int it = 0; it < n_it; ++it)
            for (int J = 0; J < n_J; ++J) push(it, J, 0);
    } else {
        const int a = (int)(*tile_starts)[rank], b = (int)(*tile_starts)[rank + 1];
        if (a != it0 || b - a != n_it) { npdrb_set_error("balanced distances: row block does not match the partition"); return NPDRB_EINVAL; }
        for (int I = a; I < b; ++I)
            for (int J = 0; J < n_J; ++J) {
                if (J >= a && J < b) { if (J <= I) push(I - a, J, TILE_MIRROR); }
                else if (balanced_row_owner_computes(I, J)) push(I - a, J, 0);
            }
    }
    return NPDRB_OK;
}

// 128 x 64 tiles when they shorten the last wave (NPDRB_DIST_TILE=64|128 forces a shape)
int pick_tile_width(int64_t n_squares, int sm_count) {
    const char *e = getenv("NPDRB_DIST_TILE");
    if (e && atoi(e) == 64) return 64;
    if (e && atoi(e) == 128) return 128;
    // makespan in units of one 128 x 128 tile: whole waves of wide tiles vs half-cost waves of twice as many narrow ones
    // (the narrow kernel measures ~3 % slower per element pair)
    const double wide = (double)nb_cdiv(n_squares, sm_count);
    const double narrow = 0.5 * 1.03 * (double)nb_cdiv(2 * n_squares, sm_count);
    return narrow < wide ? 64 : 128;
}

template <int MODE>
int launch_tma(npdrb_ctx *ctx, const double *XA, int64_t NA, int64_t ldA, const double *XB, int64_t NB, int64_t ldB, int64_t p,
               int64_t row0, int64_t nrows, const TilePlan &plan, double *dD, int64_t ldd, npdrb_session *slab_s = nullptr) {
    // slab_s: the session whose X is still arriving as column slabs (pipelined upload): one launch per slab, each issued
    // right after its H2D copy has been enqueued (pageable sources are staged by this thread while the previous slab's
    // kernel runs)
    const int n_slabs = slab_s ? slab_s->n_slabs_pending : 0;
    CUtensorMap mapA, mapB;
    NB_CHECK(make_tensor_map(&mapA, XA, NA, p, ldA));
    NB_CHECK(make_tensor_map(&mapB, XB, NB, p, ldB, plan.tn));
    const int64_t n_tiles = (int64_t)plan.tiles.size();
    if (n_tiles == 0) return NPDRB_OK;
    int2 *d_tiles = nullptr;
    NB_CUDA(cudaMalloc(&d_tiles, sizeof(int2) * (size_t)n_tiles));
    NB_CUDA(cudaMemcpyAsync(d_tiles, plan.tiles.data(), sizeof(int2) * (size_t)n_tiles, cudaMemcpyHostToDevice, ctx->stream));
    int grid = (int)((n_tiles < ctx->sm_count) ? n_tiles : ctx->sm_count);
    if (grid < 1) grid = 1;
    NB_CUDA(cudaFuncSetAttribute(dist_tma_kernel<MODE, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIST_SMEM));
    NB_CUDA(cudaFuncSetAttribute(dist_tma_kernel<MODE, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIST_SMEM));
    // one launch over all attributes, or one per uploaded column slab as it arrives (pipelined npdrb_session_upload)
    const int n_chunks = (int)nb_cdiv(p, BK);
    const int n_launch = (n_slabs > 1) ? n_slabs : 1;
    for (int l = 0; l < n_launch; ++l) {
        int c_lo = 0, c_hi = n_chunks;
        if (n_launch > 1) {
            NB_CHECK(nb_issue_slab(slab_s, l));
            NB_CUDA(cudaStreamWaitEvent(ctx->stream, slab_s->slab_ev[(size_t)l], 0));
            c_lo = (int)(slab_s->slab_start[(size_t)l] / BK);
            c_hi = (int)nb_cdiv(slab_s->slab_start[(size_t)l + 1], BK);
        }
        const int acc = l > 0, fin = l == n_launch - 1;
        if (plan.tn == 64)
            dist_tma_kernel<MODE, 64><<<grid, DTHREADS, DIST_SMEM, ctx->stream>>>(mapA, mapB, XA == XB ? 1 : 0, NA, NB, p, row0, nrows,
                                                                                  d_tiles, n_tiles, dD, ldd, c_lo, c_hi, acc, fin);
        else
            dist_tma_kernel<MODE, TN><<<grid, DTHREADS, DIST_SMEM, ctx->stream>>>(mapA, mapB, XA == XB ? 1 : 0, NA, NB, p, row0, nrows,
                                                                                  d_tiles, n_tiles, dD, ldd, c_lo, c_hi, acc, fin);
        NB_LAUNCH_CHECK(ctx);
    }
    NB_CUDA(cudaStreamSynchronize(ctx->stream));        // the pinned-less H2D source (plan.tiles) must outlive the copy
    cudaFree(d_tiles);
    return NPDRB_OK;
}
template <int MODE>
int launch_tma(npdrb_session *s, const double *X, int64_t row0, int64_t nrows, const TilePlan &plan) {
    const bool slabs = s->n_slabs_pending > 1 && X == s->dX;
    int rc = launch_tma<MODE>(s->ctx, X, s->N, s->ldx, X, s->N, s->ldx, s->p, row0, nrows, plan, s->dD, s->ldd, slabs ? s : nullptr);
    if (rc == NPDRB_OK && slabs) {                     // every slab has been waited for on the main stream
        s->n_slabs_pending = 0; s->n_slabs_issued = 0; s->hX_lazy = nullptr;
        rc = nb_check_finite(s->ctx, s->dX, s->ldx * s->p, "the attribute matrix");
    }
    return rc;
}

template <int MODE>
int launch_simple(npdrb_session *s, const double *X, int64_t row0, int64_t nrows) {
    dim3 grid((unsigned)nb_cdiv(s->N, 64), (unsigned)nb_cdiv(nrows, 64));
    dist_simple_kernel<MODE><<<grid, 256, 0, s->ctx->stream>>>(X, s->ldx, s->N, s->p, row0, nrows, s->dD, s->ldd);
    NB_LAUNCH_CHECK(s->ctx);
    return NPDRB_OK;
}

// pack: tile (x = local row tile, y = column tile) of this block, TRANSPOSED, into buf[tile][c][r] (128 x 128 doubles)
__global__ void __launch_bounds__(256)
pack_tiles_kernel(const double *__restrict__ D, int64_t ldd, int64_t nrows, int64_t N, const int2 *__restrict__ desc,
                  double *__restrict__ buf) {
    __shared__ double sm[32][33];
    const int2 d = desc[blockIdx.x];
    double *out = buf + (size_t)blockIdx.x * TM * TN;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8
    for (int sb = 0; sb < 16; ++sb) {
        const int r0 = (sb >> 2) * 32, c0 = (sb & 3) * 32;
        for (int rr = ty; rr < 32; rr += 8) {
            const int64_t row = (int64_t)d.x * TM + r0 + rr, col = (int64_t)d.y * TN + c0 + tx;
            sm[rr][tx] = (row < nrows && col < N) ? D[row * ldd + col] : 0.0;
        }
        __syncthreads();
        for (int cc = ty; cc < 32; cc += 8) out[(size_t)(c0 + cc) * TM + r0 + tx] = sm[tx][cc];
        __syncthreads();
    }
}
// unpack: buf[tile][c][r] -> D[(x * 128 + c) * ldd + y * 128 + r]   (x = local row tile, y = column tile of THIS block)
__global__ void __launch_bounds__(256)
unpack_tiles_kernel(double *__restrict__ D, int64_t ldd, int64_t nrows, int64_t N, const int2 *__restrict__ desc,
                    const double *__restrict__ buf) {
    const int2 d = desc[blockIdx.x];
    const double *in = buf + (size_t)blockIdx.x * TM * TN;
    for (int e = threadIdx.x; e < TM * TN; e += 256) {
        const int c = e >> 7, r = e & 127;
        const int64_t row = (int64_t)d.x * TM + c, col = (int64_t)d.y * TN + r;
        if (row < nrows && col < N) D[row * ldd + col] = in[e];
    }
}

int metric_mode(int32_t metric, int32_t fast_euclidean) {
    const bool euclid = (metric == NPDRB_METRIC_EUCLIDEAN || metric == NPDRB_METRIC_RELIEF_SCALED_EUCLIDEAN);
    return metric == NPDRB_METRIC_HAMMING ? MODE_HAMMING
         : !euclid ? MODE_MANHATTAN : (fast_euclidean ? MODE_EUCLID_FMA : MODE_EUCLID_EXACT);
}

int dist_core(npdrb_session *s, int32_t metric, int32_t fast_euclidean, int64_t row0, int64_t nrows, int shape,
              const std::vector<int64_t> *tile_starts, int rank) {
    npdrb_ctx *ctx = s->ctx;
    if (!s->dX) { npdrb_set_error("distances: no attribute matrix in the session"); return NPDRB_EINVAL; }
    if ((metric < 0 || metric > NPDRB_METRIC_RELIEF_SCALED_EUCLIDEAN) && metric != NPDRB_METRIC_HAMMING) {
        npdrb_set_error("distances: unsupported metric %d", metric);
        return NPDRB_EINVAL;
    }
    if (row0 < 0 || nrows < 1 || row0 + nrows > s->N) { npdrb_set_error("distances: bad row block"); return NPDRB_EINVAL; }
    if (row0 % TM != 0) { npdrb_set_error("distances: row0 must be a multiple of %d", TM); return NPDRB_EINVAL; }
    if (s->n_slabs_pending == 1) NB_CHECK(nb_wait_upload(s));     // a single pending slab: nothing to pipeline against

    if (s->dD && s->own_D && (s->nrows != nrows)) { cudaFree(s->dD); s->dD = nullptr; }
    if (!s->dD || !s->own_D) {
        s->ldd = nb_round_up(s->N, 2);
        NB_CUDA(cudaMalloc(&s->dD, sizeof(double) * (size_t)nrows * (size_t)s->ldd));
        s->own_D = true;
    }
    s->row0 = row0; s->nrows = nrows;

    NB_CUDA(cudaEventRecord(s->ev[0], ctx->stream));
    const double *X = s->dX;
    const char *kenv0 = getenv("NPDRB_DIST_KERNEL");
    if ((metric >= NPDRB_METRIC_ALLELE_SHARING_MANHATTAN && metric <= NPDRB_METRIC_RELIEF_SCALED_EUCLIDEAN) ||
        (kenv0 && strcmp(kenv0, "simple") == 0))
        NB_CHECK(nb_wait_upload(s));                    // pre-scaling / the diagnostic kernel need the whole matrix
    if (metric >= NPDRB_METRIC_ALLELE_SHARING_MANHATTAN && metric <= NPDRB_METRIC_RELIEF_SCALED_EUCLIDEAN) {
        if (!s->dXs) NB_CUDA(cudaMalloc(&s->dXs, sizeof(double) * (size_t)s->ldx * (size_t)s->p));
        prescale_kernel<<<(unsigned)s->p, 256, 0, ctx->stream>>>(s->dX, s->ldx, s->N, metric, s->dXs);
        NB_LAUNCH_CHECK(ctx);
        X = s->dXs;
    }
    const int mode = metric_mode(metric, fast_euclidean);
    const char *kenv = getenv("NPDRB_DIST_KERNEL");
    const bool simple = kenv && strcmp(kenv, "simple") == 0 && shape != 2;
    int rc;
    if (simple) {
        rc = mode == MODE_MANHATTAN ? launch_simple<MODE_MANHATTAN>(s, X, row0, nrows)
           : mode == MODE_HAMMING ? launch_simple<MODE_HAMMING>(s, X, row0, nrows)
           : mode == MODE_EUCLID_EXACT ? launch_simple<MODE_EUCLID_EXACT>(s, X, row0, nrows)
                                       : launch_simple<MODE_EUCLID_FMA>(s, X, row0, nrows);
    } else {
        TilePlan plan;
        NB_CHECK(build_tiles(shape, row0, nrows, s->N, tile_starts, rank, TN, plan));
        const int tn = pick_tile_width((int64_t)plan.tiles.size(), ctx->sm_count);
        if (tn != TN) NB_CHECK(build_tiles(shape, row0, nrows, s->N, tile_starts, rank, tn, plan));
        rc = mode == MODE_MANHATTAN ? launch_tma<MODE_MANHATTAN>(s, X, row0, nrows, plan)
           : mode == MODE_HAMMING ? launch_tma<MODE_HAMMING>(s, X, row0, nrows, plan)
           : mode == MODE_EUCLID_EXACT ? launch_tma<MODE_EUCLID_EXACT>(s, X, row0, nrows, plan)
                                       : launch_tma<MODE_EUCLID_FMA>(s, X, row0, nrows, plan);
    }
    NB_CHECK(rc);
    NB_CUDA(cudaEventRecord(s->ev[1], ctx->stream));
    NB_CUDA(cudaEventSynchronize(s->ev[1]));
    float ms = 0;
    NB_CUDA(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
    s->ms_distance = ms;
    return NPDRB_OK;
}

// tiles exchanged with the other ranks after a balanced distance pass, in the canonical order both sides enumerate:
// peer h ascending; within a peer the SENDER's row tile ascending, then its column tile ascending
int exchange_lists(npdrb_session *s, std::vector<int2> &send, std::vector<int2> &recv, std::vector<int64_t> &send_cnt,
                   std::vector<int64_t> &recv_cnt) {
    if (s->bal_rank < 0) { npdrb_set_error("balanced exchange: run npdrb_session_distances_balanced first"); return NPDRB_EINVAL; }
    const std::vector<int64_t> &ts = s->bal_tile_starts;
    const int G = (int)ts.size() - 1, g = s->bal_rank;
    const int a = (int)ts[g], b = (int)ts[g + 1];
    send.clear(); recv.clear(); send_cnt.assign(G, 0); recv_cnt.assign(G, 0);
    for (int h = 0; h < G; ++h) {
        if (h == g) continue;
        const int c = (int)ts[h], d = (int)ts[h + 1];
        for (int I = a; I < b; ++I)                // I send what I computed in the square (my rows, h's columns) ...
            for (int J = c; J < d; ++J)
                if (balanced_row_owner_computes(I, J)) { send.push_back(make_int2(I - a, J)); send_cnt[h] += (int64_t)TM * TN; }
        for (int J = c; J < d; ++J)                // ... and receive what h computed there, transposed into my rows
            for (int I = a; I < b; ++I)
                if (balanced_row_owner_computes(J, I)) { recv.push_back(make_int2(I - a, J)); recv_cnt[h] += (int64_t)TM * TN; }
    }
    return NPDRB_OK;
}

int run_tile_copy(npdrb_session *s, const std::vector<int2> &desc, double *buf, bool pack) {
    npdrb_ctx *ctx = s->ctx;
    if (desc.empty()) return NPDRB_OK;
    int2 *d_desc = nullptr;
    NB_CUDA(cudaMalloc(&d_desc, sizeof(int2) * desc.size()));
    NB_CUDA(cudaMemcpyAsync(d_desc, desc.data(), sizeof(int2) * desc.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (pack) pack_tiles_kernel<<<(unsigned)desc.size(), 256, 0, ctx->stream>>>(s->dD, s->ldd, s->nrows, s->N, d_desc, buf);
    else unpack_tiles_kernel<<<(unsigned)desc.size(), 256, 0, ctx->stream>>>(s->dD, s->ldd, s->nrows, s->N, d_desc, buf);
    NB_LAUNCH_CHECK(ctx);
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_desc);
    return NPDRB_OK;
}

}  // namespace

int nb_distances(npdrb_session *s, int32_t metric, int32_t fast_euclidean, int64_t row0, int64_t nrows) {
    const int symmetric = (row0 == 0 && nrows == s->N) ? 1 : 0;
    s->bal_rank = -1;
    return dist_core(s, metric, fast_euclidean, row0, nrows, symmetric ? 0 : 1, nullptr, 0);
}

// Balanced symmetric distances over `world` row blocks (row_starts: world + 1 instance bounds, tile aligned).  After
// this call the block holds the tiles this rank computed; nb_balanced_pack / all-to-all / nb_balanced_unpack complete it.
int nb_distances_balanced(npdrb_session *s, int32_t metric, int32_t fast_euclidean, const int64_t *row_starts, int32_t rank,
                          int32_t world) {
    if (!row_starts || world < 1 || rank < 0 || rank >= world) { npdrb_set_error("balanced distances: bad partition"); return NPDRB_EINVAL; }
    std::vector<int64_t> ts((size_t)world + 1);
    for (int r = 0; r <= world; ++r) {
        if (row_starts[r] % TM != 0 && row_starts[r] != s->N) { npdrb_set_error("balanced distances: row bounds must be multiples of %d", TM); return NPDRB_EINVAL; }
        ts[r] = nb_cdiv(row_starts[r], TM);
    }
    if (row_starts[0] != 0 || row_starts[world] != s->N) { npdrb_set_error("balanced distances: partition must cover 0..N"); return NPDRB_EINVAL; }
    const int64_t row0 = row_starts[rank], nrows = row_starts[rank + 1] - row0;
    if (nrows < 1) { npdrb_set_error("balanced distances: empty row block"); return NPDRB_EINVAL; }
    NB_CHECK(dist_core(s, metric, fast_euclidean, row0, nrows, 2, &ts, rank));
    s->bal_tile_starts = ts; s->bal_rank = rank;
    return NPDRB_OK;
}

int nb_balanced_counts(npdrb_session *s, int64_t *send_doubles, int64_t *recv_doubles) {
    std::vector<int2> sd, rv; std::vector<int64_t> sc, rc;
    NB_CHECK(exchange_lists(s, sd, rv, sc, rc));
    for (size_t h = 0; h < sc.size(); ++h) { send_doubles[h] = sc[h]; recv_doubles[h] = rc[h]; }
    return NPDRB_OK;
}
int nb_balanced_pack(npdrb_session *s, double *d_send) {
    std::vector<int2> sd, rv; std::vector<int64_t> sc, rc;
    NB_CHECK(exchange_lists(s, sd, rv, sc, rc));
    return run_tile_copy(s, sd, d_send, true);
}
int nb_balanced_unpack(npdrb_session *s, const double *d_recv) {
    std::vector<int2> sd, rv; std::vector<int64_t> sc, rc;
    NB_CHECK(exchange_lists(s, sd, rv, sc, rc));
    return run_tile_copy(s, rv, (double *)d_recv, false);
}

// Rectangular distances (npdrDistances2, R/npdrLearner.R:323-339): rows = instances of X2 (m2), columns = instances of
// X1 (m1); dD is m2 x ldd row-major, i.e. column i of the reference's m1 x m2 matrix is row i here.
int nb_distances2(npdrb_ctx *ctx, const double *dX1, int64_t ld1, int64_t m1, const double *dX2, int64_t ld2, int64_t m2,
                  int64_t p, int32_t metric, int32_t fast_euclidean, double *dD, int64_t ldd) {
    if (metric != NPDRB_METRIC_MANHATTAN && metric != NPDRB_METRIC_EUCLIDEAN) {
        npdrb_set_error("distances2: metric must be manhattan or euclidean (allele-sharing is a pre-scaling by the caller)");
        return NPDRB_EINVAL;
    }
    TilePlan plan;
    NB_CHECK(build_tiles(1, 0, m2, m1, nullptr, 0, TN, plan));
    const int tn = pick_tile_width((int64_t)plan.tiles.size(), ctx->sm_count);
    if (tn != TN) NB_CHECK(build_tiles(1, 0, m2, m1, nullptr, 0, tn, plan));
    const int mode = metric_mode(metric, fast_euclidean);
    return mode == MODE_MANHATTAN ? launch_tma<MODE_MANHATTAN>(ctx, dX2, m2, ld2, dX1, m1, ld1, p, 0, m2, plan, dD, ldd)
         : mode == MODE_EUCLID_EXACT ? launch_tma<MODE_EUCLID_EXACT>(ctx, dX2, m2, ld2, dX1, m1, ld1, p, 0, m2, plan, dD, ldd)
                                     : launch_tma<MODE_EUCLID_FMA>(ctx, dX2, m2, ld2, dX1, m1, ld1, p, 0, m2, plan, dD, ldd);
}
