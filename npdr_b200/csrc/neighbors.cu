// neighbors.cu -- K2..K5: neighbourhood radii, threshold / fixed-k selection, ordered pair
// emission, uniqueNeighbors (nearestNeighbors + get_Ri_nearest, R/nearestNeighbors.R:153-285,
// 622-634; uniqueNeighbors :572-583).
//
// Bound: one or two HBM scans of the distance block (N^2 doubles) -- negligible next to K1/K7.
// Exactness: radii are computed with xf80 (bit-exact emulation of base-R's long-double
// colSums()/sd()), distances are bit-identical to dist() (distance.cu), so the test
// `d < radius & d > 0` selects exactly the reference's neighbours, in the reference's order.
#include "common.cuh"
#include "xf80.cuh"

namespace {

// ---------------------------------------------------------------- multiSURF radius
// One lane per owned row; a warp walks 32 rows through 32x32 transposed tiles so that global
// reads stay coalesced while every lane runs its own strictly sequential extended-precision sums.
// PASSES: 1 = colSums only (user sd.vec), 3 = colSums + two-pass var().
__global__ void __launch_bounds__(128)
radius_multisurf_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, int64_t row0, int64_t nrows,
                        double sd_frac, const double *__restrict__ sd_vec, double *__restrict__ radius) {
    __shared__ double tile[4][32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t rbase = ((int64_t)blockIdx.x * 4 + w) * 32;
    if (rbase >= nrows) return;
    const int64_t r = rbase + lane;             // local row
    const int64_t g = row0 + r;                 // global instance (its own distance is skipped)
    const bool valid = r < nrows;
    const int64_t nobs = N - 1;

    xf80 S = xf_zero(), T = xf_zero(), XM = xf_zero();
    const int n_pass = sd_vec ? 1 : 3;
    for (int pass = 0; pass < n_pass; ++pass) {
        xf80 acc = xf_zero();
        for (int64_t c0 = 0; c0 < N; c0 += 32) {
            for (int rr = 0; rr < 32; ++rr) {
                int64_t row = rbase + rr, col = c0 + lane;
                tile[w][rr][lane] = (row < nrows && col < N) ? D[row * ldd + col] : 0.0;
            }
            __syncwarp();
            if (valid) {
                const int lim = (int)((N - c0 < 32) ? (N - c0) : 32);
                for (int cc = 0; cc < lim; ++cc) {
                    if (c0 + cc == g) continue;
                    xf80 x = xf_from_double(tile[w][lane][cc]);
                    if (pass == 0) acc = xf_add(acc, x);
                    else if (pass == 1) acc = xf_add(acc, xf_sub(x, T));
                    else { xf80 d = xf_sub(x, XM); acc = xf_add(acc, xf_mul(d, d)); }
                }
            }
            __syncwarp();
        }
        if (pass == 0) { S = acc; T = xf_div(S, xf_from_int(nobs)); }
        else if (pass == 1) { T = xf_add(T, xf_div(acc, xf_from_int(nobs))); XM = xf_from_double(xf_to_double(T)); }
        else { S = S; XM = acc; }                // XM now holds the sum of squares
    }
    if (!valid) return;
    const double colsum = xf_to_double(S);       // colSums() includes the zero diagonal: adding 0 changes nothing
    double sdv;
    if (sd_vec) sdv = sd_vec[g];
    else sdv = sqrt(xf_to_double(xf_div(XM, xf_from_int(nobs - 1))));
    // Ri.radius = colSums(dist.mat)/(N-1) - sd.frac*sd.vec   (R/nearestNeighbors.R:246), plain double ops
    radius[r] = __dsub_rn(__ddiv_rn(colsum, (double)(N - 1)), __dmul_rn(sd_frac, sdv));
}

// ---------------------------------------------------------------- SURF radius
// R/nearestNeighbors.R:234-241.  Exact (N <= 1024): one thread replays sum(dist.mat) and
// sd(dist.mat[upper.tri]) in their sequential long-double order.
__global__ void surf_radius_serial_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, double sd_frac,
                                          double *__restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    xf80 s = xf_zero();
    for (int64_t c = 0; c < N; ++c)
        for (int64_t i = 0; i < N; ++i) s = xf_add(s, xf_from_double(D[c * ldd + i]));
    const double total = xf_to_double(s);
    const int64_t nu = N * (N - 1) / 2;
    xf80 sum = xf_zero();
    for (int64_t j = 0; j < N; ++j) for (int64_t i = 0; i < j; ++i) sum = xf_add(sum, xf_from_double(D[j * ldd + i]));
    xf80 tmp = xf_div(sum, xf_from_int(nu));
    sum = xf_zero();
    for (int64_t j = 0; j < N; ++j) for (int64_t i = 0; i < j; ++i) sum = xf_add(sum, xf_sub(xf_from_double(D[j * ldd + i]), tmp));
    tmp = xf_add(tmp, xf_div(sum, xf_from_int(nu)));
    const xf80 xm = xf_from_double(xf_to_double(tmp));
    sum = xf_zero();
    for (int64_t j = 0; j < N; ++j)
        for (int64_t i = 0; i < j; ++i) { xf80 d = xf_sub(xf_from_double(D[j * ldd + i]), xm); sum = xf_add(sum, xf_mul(d, d)); }
    const double var = xf_to_double(xf_div(sum, xf_from_int(nu - 1)));
    const double num_pair = __ddiv_rn(__dmul_rn((double)N, (double)(N - 1)), 2.0);
    out[0] = __dsub_rn(__ddiv_rn(total, __dmul_rn(2.0, num_pair)), __dmul_rn(sd_frac, sqrt(var)));
}

// Large N (or row blocks): per-row moments (sum over all j, sum and sum of squares over j > global row)
// in double-double; combined on the host.  Not bit-identical to R's single sequential chain
// (<= 1 ulp on the radius); documented in DESIGN.md.
__global__ void surf_moments_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, int64_t row0, int64_t nrows,
                                    double *__restrict__ out /* nrows x 3: sum_all, sum_upper, count_upper */) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= nrows) return;
    const int64_t g = row0 + r;
    double sa = 0, su = 0;
    for (int64_t j = lane; j < N; j += 32) { double v = D[r * ldd + j]; sa += v; if (j > g) su += v; }
    for (int o = 16; o; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); su += __shfl_xor_sync(0xffffffffu, su, o); }
    if (lane == 0) { out[r * 3 + 0] = sa; out[r * 3 + 1] = su; out[r * 3 + 2] = (double)(N - 1 - g); }
}
__global__ void surf_sumsq_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, int64_t row0, int64_t nrows,
                                  double mean, double *__restrict__ out /* nrows */) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= nrows) return;
    const int64_t g = row0 + r;
    double ss = 0;
    for (int64_t j = g + 1 + lane; j < N; j += 32) { double v = D[r * ldd + j] - mean; ss += v * v; }
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) out[r] = ss;
}

// ---------------------------------------------------------------- threshold selection (surf / multisurf)
__global__ void count_threshold_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, int64_t nrows,
                                       const double *__restrict__ radius, int radius_is_scalar,
                                       int64_t *__restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= nrows) return;
    const double rad = radius_is_scalar ? radius[0] : radius[r];
    const double *row = D + r * ldd;
    int cnt = 0;
    for (int64_t j = lane; j < N; j += 32) { double v = row[j]; cnt += ((v < rad) && (v > 0.0)) ? 1 : 0; }
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) counts[r] = cnt;
}

// exclusive scan of n int64 counts into ptr[0..n]; also counts empty rows.  Single block.
__global__ void __launch_bounds__(1024)
scan_counts_kernel(const int64_t *__restrict__ counts, int64_t n, int64_t *__restrict__ ptr, int64_t *__restrict__ n_empty) {
    __shared__ int64_t wsum[32];
    __shared__ int64_t carry_s;
    __shared__ int64_t empty_s;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) { carry_s = 0; empty_s = 0; }
    __syncthreads();
    int64_t my_empty = 0;
    for (int64_t base = 0; base < n; base += 1024) {
        int64_t i = base + tid;
        int64_t v = (i < n) ? counts[i] : 0;
        if (i < n && v == 0) my_empty++;
        int64_t x = v;
        for (int o = 1; o < 32; o <<= 1) { int64_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        if (w == 0) {
            int64_t s = wsum[lane];
            for (int o = 1; o < 32; o <<= 1) { int64_t y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
            wsum[lane] = s;
        }
        __syncthreads();
        int64_t incl = x + (w ? wsum[w - 1] : 0) + carry_s;
        if (i < n) ptr[i] = incl - v;
        __syncthreads();
        if (tid == 1023) carry_s = incl;
        __syncthreads();
    }
    for (int o = 16; o; o >>= 1) my_empty += __shfl_xor_sync(0xffffffffu, my_empty, o);
    if (lane == 0 && my_empty) atomicAdd((unsigned long long *)&empty_s, (unsigned long long)my_empty);
    __syncthreads();
    if (tid == 0) { ptr[n] = carry_s; n_empty[0] = empty_s; }
}

__global__ void write_threshold_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, int64_t row0, int64_t nrows,
                                       const double *__restrict__ radius, int radius_is_scalar,
                                       const int64_t *__restrict__ ptr, int32_t *__restrict__ Ri, int32_t *__restrict__ NN) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= nrows) return;
    const double rad = radius_is_scalar ? radius[0] : radius[r];
    const double *row = D + r * ldd;
    int64_t off = ptr[r];
    const int32_t ri = (int32_t)(row0 + r + 1);
    for (int64_t j0 = 0; j0 < N; j0 += 32) {
        int64_t j = j0 + lane;
        double v = (j < N) ? row[j] : 0.0;
        bool sel = (j < N) && (v < rad) && (v > 0.0);
        unsigned m = __ballot_sync(0xffffffffu, sel);
        if (sel) {
            int64_t pos = off + __popc(m & ((1u << lane) - 1u));
            Ri[pos] = ri;
            NN[pos] = (int32_t)(j + 1);                        // ascending j: dplyr::filter keeps row order
        }
        off += __popc(m);
    }
}

// ---------------------------------------------------------------- relieff (fixed k)
// slice_min(n = k+1, with_ties = TRUE) keeps rows with min_rank <= k+1, i.e. d <= (k+1)-th smallest
// (self included), ordered by (d, row); filter(d > 0) then drops self and zero-distance duplicates.
__device__ __forceinline__ uint64_t dkey(double v) {
    uint64_t b = (uint64_t)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// One CTA per owned row: 8-pass radix select of the rank-(k+1) key, then count {d <= cut, d > 0}.
__global__ void __launch_bounds__(256)
relieff_select_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, int64_t nrows, int64_t kplus1,
                      uint64_t *__restrict__ cut_keys, int64_t *__restrict__ counts) {
    __shared__ unsigned hist[256];
    __shared__ uint64_t prefix_s, mask_s;
    __shared__ int64_t rank_s;
    __shared__ int cnt_s;
    const int64_t r = blockIdx.x;
    const double *row = D + r * ldd;
    int64_t want = kplus1 < N ? kplus1 : N;
    if (threadIdx.x == 0) { prefix_s = 0; mask_s = 0; rank_s = want; cnt_s = 0; }
    __syncthreads();
    for (int pass = 7; pass >= 0; --pass) {
        hist[threadIdx.x] = 0;
        __syncthreads();
        const uint64_t prefix = prefix_s, mask = mask_s;
        const int shift = pass * 8;
        for (int64_t j = threadIdx.x; j < N; j += 256) {
            uint64_t key = dkey(row[j]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int64_t rank = rank_s, cum = 0;
            int b = 0;
            for (; b < 256; ++b) { if (cum + hist[b] >= rank) break; cum += hist[b]; }
            rank_s = rank - cum;
            prefix_s = prefix | ((uint64_t)b << shift);
            mask_s = mask | (0xFFull << shift);
        }
        __syncthreads();
    }
    const uint64_t cut = prefix_s;
    int c = 0;
    for (int64_t j = threadIdx.x; j < N; j += 256) { double v = row[j]; c += (dkey(v) <= cut && v > 0.0) ? 1 : 0; }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&cnt_s, c);
    __syncthreads();
    if (threadIdx.x == 0) { cut_keys[r] = cut; counts[r] = cnt_s; }
}

// One CTA per owned row: gather the selected (key, j), sort by (key, j) with a normalised bitonic
// network (virtual +inf padding), write NN in that order.  Segment lives in shared memory when it
// fits (<= SORT_CAP entries), else it is sorted in place in a global scratch segment.
constexpr int SORT_CAP = 8192;
template <typename KeyPtr, typename IdxPtr>
__device__ __forceinline__ void bitonic_sort(KeyPtr keys, IdxPtr idx, int n) {
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (np2 >> 1); t += blockDim.x) {
                int lo, hi;
                if (j == (k >> 1)) {                          // first sub-stage of a merge: mirror partner
                    int blk = t / j, off = t % j;
                    lo = blk * k + off;
                    hi = blk * k + k - 1 - off;
                } else {
                    int blk = t / j, off = t % j;
                    lo = blk * (j << 1) + off;
                    hi = lo + j;
                }
                if (hi < n) {                                  // elements >= n are +inf and never move
                    uint64_t ka = keys[lo], kb = keys[hi];
                    int32_t ia = idx[lo], ib = idx[hi];
                    if (ka > kb || (ka == kb && ia > ib)) { keys[lo] = kb; keys[hi] = ka; idx[lo] = ib; idx[hi] = ia; }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(256)
relieff_write_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, int64_t row0, int64_t nrows,
                     const uint64_t *__restrict__ cut_keys, const int64_t *__restrict__ ptr,
                     uint64_t *__restrict__ scratch_keys, int32_t *__restrict__ Ri, int32_t *__restrict__ NN) {
    extern __shared__ unsigned char sm_raw[];
    uint64_t *skeys = (uint64_t *)sm_raw;
    int32_t *sidx = (int32_t *)(skeys + SORT_CAP);
    __shared__ int fill;
    const int64_t r = blockIdx.x;
    const double *row = D + r * ldd;
    const uint64_t cut = cut_keys[r];
    const int64_t off = ptr[r];
    const int n = (int)(ptr[r + 1] - off);
    if (n == 0) return;
    const bool in_smem = n <= SORT_CAP;
    uint64_t *gkeys = scratch_keys + off;
    int32_t *gidx = NN + off;
    if (threadIdx.x == 0) fill = 0;
    __syncthreads();
    // gather (order irrelevant: (key, j) is a total order)
    for (int64_t j = threadIdx.x; j < N; j += 256) {
        double v = row[j];
        uint64_t key = dkey(v);
        if (key <= cut && v > 0.0) {
            int pos = atomicAdd(&fill, 1);
            if (in_smem) { skeys[pos] = key; sidx[pos] = (int32_t)(j + 1); }
            else { gkeys[pos] = key; gidx[pos] = (int32_t)(j + 1); }
        }
    }
    __syncthreads();
    if (in_smem) {
        bitonic_sort(skeys, sidx, n);
        for (int t = threadIdx.x; t < n; t += 256) NN[off + t] = sidx[t];
    } else {
        bitonic_sort(gkeys, gidx, n);
    }
    const int32_t ri = (int32_t)(row0 + r + 1);
    for (int t = threadIdx.x; t < n; t += 256) Ri[off + t] = ri;
}

// ---------------------------------------------------------------- uniqueNeighbors
// R/nearestNeighbors.R:572-583: key "min,max", keep first occurrence in list order.  Equivalent
// (any list order): row m is a duplicate iff an earlier row holds the same unordered pair.  For a
// list without repeated ordered pairs that is: (j,i) also present and it comes first.  We mark
// all ordered pairs in an N x N bitmap, then row m = (i,j) is dropped iff bit(j,i) is set and the
// (j,i) row precedes it; with rows grouped by ascending Ri that is simply i > j.  For arbitrary
// order we compare positions through a first-position table when the list is small enough;
// lists produced by this library are always Ri-grouped, which the caller asserts.
__global__ void mark_pairs_kernel(const int32_t *__restrict__ Ri, const int32_t *__restrict__ NN, int64_t M, int64_t N,
                                  unsigned *__restrict__ bits) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    uint64_t b = (uint64_t)(Ri[m] - 1) * (uint64_t)N + (uint64_t)(NN[m] - 1);
    atomicOr(&bits[b >> 5], 1u << (b & 31));
}
__global__ void flag_dups_kernel(const int32_t *__restrict__ Ri, const int32_t *__restrict__ NN, int64_t M, int64_t N,
                                 const unsigned *__restrict__ bits, uint8_t *__restrict__ keep,
                                 unsigned long long *__restrict__ n_dup) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int dup = 0;
    if (m < M) {
        int64_t i = Ri[m] - 1, j = NN[m] - 1;
        if (i > j) {
            uint64_t b = (uint64_t)j * (uint64_t)N + (uint64_t)i;
            dup = (bits[b >> 5] >> (b & 31)) & 1u;
        }
        keep[m] = (uint8_t)!dup;
    }
    unsigned mask = __ballot_sync(0xffffffffu, dup);
    if ((threadIdx.x & 31) == 0 && mask) atomicAdd(n_dup, (unsigned long long)__popc(mask));
}
// stable compaction of kept rows: per-block counts -> scan (host-free: reuse scan_counts_kernel) -> scatter
__global__ void block_keep_count_kernel(const uint8_t *__restrict__ keep, int64_t M, int64_t *__restrict__ counts) {
    __shared__ int c;
    if (threadIdx.x == 0) c = 0;
    __syncthreads();
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int k = (m < M) ? keep[m] : 0;
    unsigned mask = __ballot_sync(0xffffffffu, k);
    if ((threadIdx.x & 31) == 0 && mask) atomicAdd(&c, __popc(mask));
    __syncthreads();
    if (threadIdx.x == 0) counts[blockIdx.x] = c;
}
__global__ void __launch_bounds__(256)
compact_pairs_kernel(const int32_t *__restrict__ Ri, const int32_t *__restrict__ NN, const uint8_t *__restrict__ keep,
                     int64_t M, const int64_t *__restrict__ block_ptr, int32_t *__restrict__ Ro, int32_t *__restrict__ No) {
    __shared__ int wcount[8];
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int k = (m < M) ? keep[m] : 0;
    unsigned mask = __ballot_sync(0xffffffffu, k);
    if (lane == 0) wcount[w] = __popc(mask);
    __syncthreads();
    int base = 0;
    for (int i = 0; i < w; ++i) base += wcount[i];
    if (k) {
        int64_t pos = block_ptr[blockIdx.x] + base + __popc(mask & ((1u << lane) - 1u));
        Ro[pos] = Ri[m];
        No[pos] = NN[m];
    }
}

__global__ void diff_vectors_kernel(const double *__restrict__ a, const double *__restrict__ b, int64_t n, int type,
                                    double norm_fac, double *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = a[i], y = b[i], v;
    switch (type) {                                            // npdrDiff, R/nearestNeighbors.R:20-27
    case NPDRB_DIFF_NUMERIC_SQR: { double t = fabs(__dsub_rn(x, y)); v = __ddiv_rn(__dmul_rn(t, t), norm_fac); break; }
    case NPDRB_DIFF_ALLELE_SHARING: v = __ddiv_rn(fabs(__dsub_rn(x, y)), 2.0); break;
    case NPDRB_DIFF_MATCH_MISMATCH: v = (x == y) ? 0.0 : 1.0; break;
    case NPDRB_DIFF_RAW: v = x; break;
    default: v = __ddiv_rn(fabs(__dsub_rn(x, y)), norm_fac);
    }
    out[i] = v;
}

}  // namespace

int nb_diff_vectors(npdrb_ctx *ctx, const double *da, const double *db, int64_t n, int32_t type, double norm_fac,
                    double *dout) {
    if (n <= 0) return NPDRB_OK;
    diff_vectors_kernel<<<(unsigned)nb_cdiv(n, 256), 256, 0, ctx->stream>>>(da, db, n, type, norm_fac, dout);
    NB_LAUNCH_CHECK(ctx);
    return NPDRB_OK;
}

// Partial SURF moments of the owned rows (see npdr_b200.h).  Each unordered pair {g, j} is counted once,
// by the owner of row g = min index, through the entries j > g of that row.
extern "C" int npdrb_session_surf_partial(npdrb_session *s, double centre, double *out4) {
    npdrb_ctx *ctx = s->ctx;
    if (!s->dD) { npdrb_set_error("surf partial: no distance block"); return NPDRB_EINVAL; }
    const int64_t N = s->N, nrows = s->nrows;
    double *d_out = nullptr, *d_ss = nullptr;
    NB_CUDA(cudaMalloc(&d_out, sizeof(double) * 3 * (size_t)nrows));
    NB_CUDA(cudaMalloc(&d_ss, sizeof(double) * (size_t)nrows));
    surf_moments_kernel<<<(unsigned)nb_cdiv(nrows, 8), 256, 0, ctx->stream>>>(s->dD, s->ldd, N, s->row0, nrows, d_out);
    NB_LAUNCH_CHECK(ctx);
    surf_sumsq_kernel<<<(unsigned)nb_cdiv(nrows, 8), 256, 0, ctx->stream>>>(s->dD, s->ldd, N, s->row0, nrows, centre, d_ss);
    NB_LAUNCH_CHECK(ctx);
    std::vector<double> h(3 * (size_t)nrows), hs((size_t)nrows);
    NB_CUDA(cudaMemcpyAsync(h.data(), d_out, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaMemcpyAsync(hs.data(), d_ss, sizeof(double) * hs.size(), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_out); cudaFree(d_ss);
    long double sa = 0, su = 0, cu = 0, ss = 0;
    for (int64_t r = 0; r < nrows; ++r) { sa += h[3 * r]; su += h[3 * r + 1]; cu += h[3 * r + 2]; ss += hs[r]; }
    out4[0] = (double)sa; out4[1] = (double)su; out4[2] = (double)cu; out4[3] = (double)ss;
    return NPDRB_OK;
}

// SURF radius into d_rad_scalar[0] (R/nearestNeighbors.R:234-241).
static int surf_radius(npdrb_session *s, double sd_frac, double *d_rad_scalar) {
    npdrb_ctx *ctx = s->ctx;
    const int64_t N = s->N;
    double radius;
    if (s->have_surf_radius) {
        radius = s->surf_radius;
    } else if (s->nrows != N) {
        npdrb_set_error("surf on a row block needs npdrb_session_set_surf_radius (combine npdrb_session_surf_partial across ranks)");
        return NPDRB_EINVAL;
    } else if (N <= 1024) {
        surf_radius_serial_kernel<<<1, 32, 0, ctx->stream>>>(s->dD, s->ldd, N, sd_frac, d_rad_scalar);
        NB_LAUNCH_CHECK(ctx);
        return NPDRB_OK;
    } else {
        // large N: parallel moments (<= 1 ulp from R's single sequential long-double chain)
        double m0[4], m1[4];
        NB_CHECK(npdrb_session_surf_partial(s, 0.0, m0));
        const double mean_u = m0[1] / m0[2];
        NB_CHECK(npdrb_session_surf_partial(s, mean_u, m1));
        const double num_pair = (double)N * (double)(N - 1) / 2.0;
        const double sdv = sqrt(m1[3] / (m1[2] - 1.0));
        radius = m0[0] / (2.0 * num_pair) - sd_frac * sdv;
    }
    NB_CUDA(cudaMemcpyAsync(d_rad_scalar, &radius, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    return NPDRB_OK;
}

int nb_neighbors(npdrb_session *s, int32_t nbd_method, double sd_frac, int64_t k, int64_t *M_local, int64_t *n_empty) {
    npdrb_ctx *ctx = s->ctx;
    if (!s->dD) { npdrb_set_error("neighbors: compute or set distances first"); return NPDRB_EINVAL; }
    if (nbd_method < 0 || nbd_method > NPDRB_NBD_RELIEFF) { npdrb_set_error("neighbors: bad nbd_method"); return NPDRB_EINVAL; }
    const int64_t N = s->N, nrows = s->nrows;
    if (N < 3) { npdrb_set_error("neighbors: need at least 3 instances"); return NPDRB_EINVAL; }
    NB_CUDA(cudaEventRecord(s->ev[0], ctx->stream));

    if (s->dRiL) { cudaFree(s->dRiL); s->dRiL = nullptr; }
    if (s->dNNL) { cudaFree(s->dNNL); s->dNNL = nullptr; }
    if (!s->d_radius) NB_CUDA(cudaMalloc(&s->d_radius, sizeof(double) * (size_t)N));
    if (!s->d_rowptr) NB_CUDA(cudaMalloc(&s->d_rowptr, sizeof(int64_t) * (size_t)(N + 2)));
    int64_t *d_counts = nullptr, *d_nempty = nullptr;
    NB_CUDA(cudaMalloc(&d_counts, sizeof(int64_t) * (size_t)nrows));
    NB_CUDA(cudaMalloc(&d_nempty, sizeof(int64_t)));
    int64_t tot[2] = {0, 0};
    const unsigned wgrid = (unsigned)nb_cdiv(nrows, 8);

    if (nbd_method == NPDRB_NBD_RELIEFF) {
        if (k <= 0) k = npdrb_knn_surf(N, sd_frac);             // R/nearestNeighbors.R:197-203
        uint64_t *d_cut = nullptr, *d_scratch = nullptr;
        NB_CUDA(cudaMalloc(&d_cut, sizeof(uint64_t) * (size_t)nrows));
        relieff_select_kernel<<<(unsigned)nrows, 256, 0, ctx->stream>>>(s->dD, s->ldd, N, nrows, k + 1, d_cut, d_counts);
        NB_LAUNCH_CHECK(ctx);
        scan_counts_kernel<<<1, 1024, 0, ctx->stream>>>(d_counts, nrows, s->d_rowptr, d_nempty);
        NB_LAUNCH_CHECK(ctx);
        NB_CUDA(cudaMemcpyAsync(&tot[0], s->d_rowptr + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        NB_CUDA(cudaMemcpyAsync(&tot[1], d_nempty, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        NB_CUDA(cudaStreamSynchronize(ctx->stream));
        const int64_t M = tot[0];
        NB_CUDA(cudaMalloc(&s->dRiL, sizeof(int32_t) * (size_t)(M > 0 ? M : 1)));
        NB_CUDA(cudaMalloc(&s->dNNL, sizeof(int32_t) * (size_t)(M > 0 ? M : 1)));
        NB_CUDA(cudaMalloc(&d_scratch, sizeof(uint64_t) * (size_t)(M > 0 ? M : 1)));
        const size_t smem = (size_t)SORT_CAP * (sizeof(uint64_t) + sizeof(int32_t));
        NB_CUDA(cudaFuncSetAttribute(relieff_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        relieff_write_kernel<<<(unsigned)nrows, 256, smem, ctx->stream>>>(s->dD, s->ldd, N, s->row0, nrows, d_cut,
                                                                          s->d_rowptr, d_scratch, s->dRiL, s->dNNL);
        NB_LAUNCH_CHECK(ctx);
        NB_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_cut); cudaFree(d_scratch);
    } else {
        int scalar = 0;
        if (nbd_method == NPDRB_NBD_MULTISURF) {
            radius_multisurf_kernel<<<(unsigned)nb_cdiv(nrows, 128), 128, 0, ctx->stream>>>(
                s->dD, s->ldd, N, s->row0, nrows, sd_frac, s->d_sd_vec, s->d_radius);
            NB_LAUNCH_CHECK(ctx);
        } else {
            scalar = 1;
            NB_CHECK(surf_radius(s, sd_frac, s->d_radius));
        }
        count_threshold_kernel<<<wgrid, 256, 0, ctx->stream>>>(s->dD, s->ldd, N, nrows, s->d_radius, scalar, d_counts);
        NB_LAUNCH_CHECK(ctx);
        scan_counts_kernel<<<1, 1024, 0, ctx->stream>>>(d_counts, nrows, s->d_rowptr, d_nempty);
        NB_LAUNCH_CHECK(ctx);
        NB_CUDA(cudaMemcpyAsync(&tot[0], s->d_rowptr + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        NB_CUDA(cudaMemcpyAsync(&tot[1], d_nempty, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        NB_CUDA(cudaStreamSynchronize(ctx->stream));
        const int64_t M = tot[0];
        NB_CUDA(cudaMalloc(&s->dRiL, sizeof(int32_t) * (size_t)(M > 0 ? M : 1)));
        NB_CUDA(cudaMalloc(&s->dNNL, sizeof(int32_t) * (size_t)(M > 0 ? M : 1)));
        write_threshold_kernel<<<wgrid, 256, 0, ctx->stream>>>(s->dD, s->ldd, N, s->row0, nrows, s->d_radius, scalar,
                                                               s->d_rowptr, s->dRiL, s->dNNL);
        NB_LAUNCH_CHECK(ctx);
        if (scalar) {
            // replicate the scalar radius over the owned rows so copy_radius() reports Ri.radius as R does
            double rad;
            NB_CUDA(cudaMemcpyAsync(&rad, s->d_radius, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            NB_CUDA(cudaStreamSynchronize(ctx->stream));
            std::vector<double> rep((size_t)nrows, rad);
            NB_CUDA(cudaMemcpyAsync(s->d_radius, rep.data(), sizeof(double) * rep.size(), cudaMemcpyHostToDevice, ctx->stream));
            NB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
    cudaFree(d_counts); cudaFree(d_nempty);
    s->M_local = tot[0];
    s->n_empty_local = tot[1];
    if (M_local) *M_local = tot[0];
    if (n_empty) *n_empty = tot[1];
    NB_CUDA(cudaEventRecord(s->ev[1], ctx->stream));
    NB_CUDA(cudaEventSynchronize(s->ev[1]));
    float ms = 0;
    NB_CUDA(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
    s->ms_neighbors = ms;
    return NPDRB_OK;
}

int nb_unique_pairs(npdrb_session *s, int32_t keep_unique_only, int64_t *n_unique) {
    npdrb_ctx *ctx = s->ctx;
    const int64_t M = s->M, N = s->N;
    if (!s->dRi || M < 0) { npdrb_set_error("unique_pairs: import a pair list first"); return NPDRB_EINVAL; }
    if (M == 0) { if (n_unique) *n_unique = 0; return NPDRB_OK; }
    const size_t words = ((size_t)N * (size_t)N + 31) / 32;
    unsigned *d_bits = nullptr; uint8_t *d_keep = nullptr; unsigned long long *d_ndup = nullptr;
    NB_CUDA(cudaMalloc(&d_bits, sizeof(unsigned) * words));
    NB_CUDA(cudaMalloc(&d_keep, (size_t)M));
    NB_CUDA(cudaMalloc(&d_ndup, sizeof(unsigned long long)));
    NB_CUDA(cudaMemsetAsync(d_bits, 0, sizeof(unsigned) * words, ctx->stream));
    NB_CUDA(cudaMemsetAsync(d_ndup, 0, sizeof(unsigned long long), ctx->stream));
    const unsigned grid = (unsigned)nb_cdiv(M, 256);
    mark_pairs_kernel<<<grid, 256, 0, ctx->stream>>>(s->dRi, s->dNN, M, N, d_bits);
    NB_LAUNCH_CHECK(ctx);
    flag_dups_kernel<<<grid, 256, 0, ctx->stream>>>(s->dRi, s->dNN, M, N, d_bits, d_keep, d_ndup);
    NB_LAUNCH_CHECK(ctx);
    unsigned long long ndup = 0;
    NB_CUDA(cudaMemcpyAsync(&ndup, d_ndup, sizeof(ndup), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    const int64_t nu = M - (int64_t)ndup;
    if (n_unique) *n_unique = nu;
    if (keep_unique_only && ndup > 0) {
        int64_t *d_bc = nullptr, *d_bp = nullptr, *d_ne = nullptr;
        int32_t *nRi = nullptr, *nNN = nullptr;
        NB_CUDA(cudaMalloc(&d_bc, sizeof(int64_t) * grid));
        NB_CUDA(cudaMalloc(&d_bp, sizeof(int64_t) * ((size_t)grid + 1)));
        NB_CUDA(cudaMalloc(&d_ne, sizeof(int64_t)));
        NB_CUDA(cudaMalloc(&nRi, sizeof(int32_t) * (size_t)nu));
        NB_CUDA(cudaMalloc(&nNN, sizeof(int32_t) * (size_t)nu));
        block_keep_count_kernel<<<grid, 256, 0, ctx->stream>>>(d_keep, M, d_bc);
        NB_LAUNCH_CHECK(ctx);
        scan_counts_kernel<<<1, 1024, 0, ctx->stream>>>(d_bc, grid, d_bp, d_ne);
        NB_LAUNCH_CHECK(ctx);
        compact_pairs_kernel<<<grid, 256, 0, ctx->stream>>>(s->dRi, s->dNN, d_keep, M, d_bp, nRi, nNN);
        NB_LAUNCH_CHECK(ctx);
        NB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (s->own_pairs) { cudaFree(s->dRi); cudaFree(s->dNN); }
        s->dRi = nRi; s->dNN = nNN; s->M = nu; s->own_pairs = true;
        s->streams_valid = false;
        cudaFree(d_bc); cudaFree(d_bp); cudaFree(d_ne);
    }
    cudaFree(d_bits); cudaFree(d_keep); cudaFree(d_ndup);
    return NPDRB_OK;
}
