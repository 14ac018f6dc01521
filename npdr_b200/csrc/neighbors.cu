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
// RECT (nearestNeighbors2, R/npdrLearner.R:446-450): the rows are a second instance set, so there is no own distance to
// skip, the statistics run over all N entries, and the mean is colMeans() (long-double sum / n, rounded once).
template <bool RECT>
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
    const int64_t nobs = RECT ? N : N - 1;

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
                    if (!RECT && c0 + cc == g) continue;
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
    if (RECT) {                                  // colMeans(dist.mat) - sd.frac * sd.vec
        radius[r] = __dsub_rn(xf_to_double(xf_div(S, xf_from_int(nobs))), __dmul_rn(sd_frac, sdv));
        return;
    }
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

// nearestNeighbors2 SURF radius (R/npdrLearner.R:439-444): mean(dist.mat) - sd.frac * sd(dist.mat) over all nrows x N
// entries in storage order (column-major m1 x m2 in R = row-major here).  mean(): long-double sum / n plus one
// correction pass; sd(): sqrt(var(as.double(x))) = cov_complete1's two-pass mean and the sum of squared deviations.
__global__ void surf_rect_serial_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, int64_t nrows, double sd_frac,
                                        double *__restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int64_t n = N * nrows;
    xf80 s = xf_zero();
    for (int64_t r = 0; r < nrows; ++r) for (int64_t j = 0; j < N; ++j) s = xf_add(s, xf_from_double(D[r * ldd + j]));
    xf80 m = xf_div(s, xf_from_int(n));
    xf80 t = xf_zero();
    for (int64_t r = 0; r < nrows; ++r) for (int64_t j = 0; j < N; ++j) t = xf_add(t, xf_sub(xf_from_double(D[r * ldd + j]), m));
    const xf80 mean_ld = xf_add(m, xf_div(t, xf_from_int(n)));     // mean() and cov_complete1 compute the same two passes
    const double mean_d = xf_to_double(mean_ld);
    const xf80 xm = xf_from_double(mean_d);
    xf80 ss = xf_zero();
    for (int64_t r = 0; r < nrows; ++r)
        for (int64_t j = 0; j < N; ++j) { xf80 d = xf_sub(xf_from_double(D[r * ldd + j]), xm); ss = xf_add(ss, xf_mul(d, d)); }
    const double var = xf_to_double(xf_div(ss, xf_from_int(n - 1)));
    out[0] = __dsub_rn(mean_d, __dmul_rn(sd_frac, sqrt(var)));
}
// large blocks: per-row sums (pass 0) / sums of squared deviations (pass 1) in double, combined on the host
__global__ void rect_row_moment_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, int64_t nrows, int pass, double mean,
                                       double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= nrows) return;
    double a = 0;
    for (int64_t j = lane; j < N; j += 32) { double v = D[r * ldd + j]; if (pass == 0) a += v; else { v -= mean; a += v * v; } }
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) out[r] = a;
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

// ---- SURF, large N / row blocks: provably the reference's pair list -------------------------------------------------
// R's radius is ONE sequential long-double chain over N^2 terms; it cannot be parallelised, but the pair list only depends on
// which side of the radius every distance falls.  So: (1) radius r~ from parallel moments, with a rigorous bound eps on
// |r~ - r_R| (npdrb_surf_radius_from_moments); (2) count the distances inside [r~ - eps, r~ + eps] (surf_band_count_kernel);
// none -> every radius in the band, r_R included, selects the same pairs, done; (3) otherwise replay R's chains exactly on the
// host (native x87 long double on x86-64, the software xf80 elsewhere) -- rare: the band is ~1e-10 wide relative to r.
__global__ void surf_band_count_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, int64_t nrows, double lo, double hi,
                                       unsigned long long *__restrict__ count) {
    unsigned long long c = 0;
    const int64_t total = nrows * N;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const double v = D[(e / N) * ldd + (e % N)];
        c += (v >= lo && v <= hi) ? 1u : 0u;
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

#if defined(__x86_64__) && defined(__LDBL_MANT_DIG__) && (__LDBL_MANT_DIG__ == 64)
#define NB_HOST_X87 1
#else
#define NB_HOST_X87 0
#endif
// sum(dist.mat) / (2 num.pair) - sd.frac * sd(dist.mat[upper.tri(dist.mat)]) exactly as R evaluates it:
// D is the full symmetric matrix on the host, element (i, c) at D[c * ldd + i].
static double surf_radius_exact_host_impl(const double *D, int64_t ldd, int64_t N, double sd_frac) {
    const int64_t nu = N * (N - 1) / 2;
    double total, var;
#if NB_HOST_X87
    long double s = 0.0L;
    for (int64_t c = 0; c < N; ++c) { const double *col = D + c * ldd; for (int64_t i = 0; i < N; ++i) s += col[i]; }
    total = (double)s;
    long double sum = 0.0L;
    for (int64_t j = 0; j < N; ++j) { const double *col = D + j * ldd; for (int64_t i = 0; i < j; ++i) sum += col[i]; }
    long double tmp = sum / (long double)nu;
    sum = 0.0L;
    for (int64_t j = 0; j < N; ++j) { const double *col = D + j * ldd; for (int64_t i = 0; i < j; ++i) sum += (col[i] - tmp); }
    tmp = tmp + sum / (long double)nu;
    const long double xm = (double)tmp;
    sum = 0.0L;
    for (int64_t j = 0; j < N; ++j) { const double *col = D + j * ldd; for (int64_t i = 0; i < j; ++i) { const long double d = col[i] - xm; sum += d * d; } }
    var = (double)(sum / (long double)(nu - 1));
#else
    xf80 s = xf_zero();
    for (int64_t c = 0; c < N; ++c) for (int64_t i = 0; i < N; ++i) s = xf_add(s, xf_from_double(D[c * ldd + i]));
    total = xf_to_double(s);
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
    var = xf_to_double(xf_div(sum, xf_from_int(nu - 1)));
#endif
    const double num_pair = (double)N * (double)(N - 1) / 2.0;
    return total / (2.0 * num_pair) - sd_frac * sqrt(var);
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
        if (m > 0 && Ri[m] < Ri[m - 1]) atomicOr((unsigned *)(n_dup + 1), 1u);   // precondition: grouped by ascending Ri
    }
    unsigned mask = __ballot_sync(0xffffffffu, dup);
    if ((threadIdx.x & 31) == 0 && mask) atomicAdd(n_dup, (unsigned long long)__popc(mask));
}
__global__ void popcount_words_kernel(const unsigned *__restrict__ bits, size_t words, unsigned long long *__restrict__ out) {
    unsigned long long c = 0;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (size_t)gridDim.x * blockDim.x) c += __popc(bits[w]);
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
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

// ================================================================ separate hit / miss neighbourhoods
// nearestNeighborsSeparateHitMiss (R/nearestNeighbors.R:318-551).  Same distance block; per owned instance the hits
// (same class, not self) and the misses (other class) get their own radius / their own k, and are listed hits first.

// multisurf radii (:483-500): radius = mean(d[subset]) - sd.frac * sd(d[subset]), base-R mean() and sd() both start from
// the same long-double two-pass mean; xf80 reproduces them bit for bit.  One lane per owned row, 32x32 transposed tiles.
__global__ void __launch_bounds__(128)
radius_hitmiss_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, int64_t row0, int64_t nrows,
                      const double *__restrict__ y, double sd_frac, double *__restrict__ r_hit, double *__restrict__ r_miss) {
    __shared__ double tile[4][32][33];
    __shared__ double ycol[4][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t rbase = ((int64_t)blockIdx.x * 4 + w) * 32;
    if (rbase >= nrows) return;
    const int64_t r = rbase + lane, g = row0 + r;
    const bool valid = r < nrows;
    const double yg = valid ? y[g] : 0.0;
    xf80 T[2] = {xf_zero(), xf_zero()}, XM[2] = {xf_zero(), xf_zero()}, SS[2] = {xf_zero(), xf_zero()};
    int64_t cnt[2] = {0, 0};
    for (int pass = 0; pass < 3; ++pass) {
        xf80 acc[2] = {xf_zero(), xf_zero()};
        int64_t c2[2] = {0, 0};
        for (int64_t c0 = 0; c0 < N; c0 += 32) {
            for (int rr = 0; rr < 32; ++rr) {
                int64_t row = rbase + rr, col = c0 + lane;
                tile[w][rr][lane] = (row < nrows && col < N) ? D[row * ldd + col] : 0.0;
            }
            ycol[w][lane] = (c0 + lane < N) ? y[c0 + lane] : 0.0;
            __syncwarp();
            if (valid) {
                const int lim = (int)((N - c0 < 32) ? (N - c0) : 32);
                for (int cc = 0; cc < lim; ++cc) {
                    const int64_t j = c0 + cc;
                    const bool same = (ycol[w][cc] == yg);
                    if (same && j == g) continue;
                    const int k = same ? 0 : 1;
                    xf80 x = xf_from_double(tile[w][lane][cc]);
                    if (pass == 0) { acc[k] = xf_add(acc[k], x); c2[k]++; }
                    else if (pass == 1) acc[k] = xf_add(acc[k], xf_sub(x, T[k]));
                    else { xf80 d = xf_sub(x, XM[k]); acc[k] = xf_add(acc[k], xf_mul(d, d)); }
                }
            }
            __syncwarp();
        }
        for (int k = 0; k < 2; ++k) {
            if (pass == 0) { cnt[k] = c2[k]; T[k] = cnt[k] > 0 ? xf_div(acc[k], xf_from_int(cnt[k])) : xf_zero(); }
            else if (pass == 1) { if (cnt[k] > 0) T[k] = xf_add(T[k], xf_div(acc[k], xf_from_int(cnt[k]))); XM[k] = xf_from_double(xf_to_double(T[k])); }
            else SS[k] = acc[k];
        }
    }
    if (!valid) return;
    const double nanv = __longlong_as_double(0x7ff8000000000000ll);
    double out[2];
    for (int k = 0; k < 2; ++k) {
        if (cnt[k] < 2) { out[k] = nanv; continue; }                        // sd() of < 2 values is NA in R
        const double mean = xf_to_double(XM[k]);
        const double sdv = sqrt(xf_to_double(xf_div(SS[k], xf_from_int(cnt[k] - 1))));
        out[k] = __dsub_rn(mean, __dmul_rn(sd_frac, sdv));
    }
    r_hit[r] = out[0];
    r_miss[r] = out[1];
}

// surf radii (:456-481), exact for N <= 1024: one thread replays mean()/sd() of the concatenated hit and miss vectors
__global__ void surf_hitmiss_serial_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, const double *__restrict__ y,
                                           double sd_frac, double *__restrict__ out2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int k = 0; k < 2; ++k) {
        xf80 s = xf_zero(); int64_t c = 0;
        for (int64_t i = 0; i < N; ++i) for (int64_t j = 0; j < N; ++j) {
            const bool same = y[j] == y[i];
            if (k == 0 ? (same && j != i) : !same) { s = xf_add(s, xf_from_double(D[i * ldd + j])); c++; }
        }
        xf80 t = xf_div(s, xf_from_int(c));
        s = xf_zero();
        for (int64_t i = 0; i < N; ++i) for (int64_t j = 0; j < N; ++j) {
            const bool same = y[j] == y[i];
            if (k == 0 ? (same && j != i) : !same) s = xf_add(s, xf_sub(xf_from_double(D[i * ldd + j]), t));
        }
        t = xf_add(t, xf_div(s, xf_from_int(c)));
        const double mean = xf_to_double(t);
        const xf80 xm = xf_from_double(mean);
        s = xf_zero();
        for (int64_t i = 0; i < N; ++i) for (int64_t j = 0; j < N; ++j) {
            const bool same = y[j] == y[i];
            if (k == 0 ? (same && j != i) : !same) { xf80 d = xf_sub(xf_from_double(D[i * ldd + j]), xm); s = xf_add(s, xf_mul(d, d)); }
        }
        const double var = xf_to_double(xf_div(s, xf_from_int(c - 1)));
        out2[k] = __dsub_rn(mean, __dmul_rn(sd_frac, sqrt(var)));
    }
}
// large N: per-row moments in double (<= few ulp from the sequential chain); out: nrows x 6 =
// sum_hit, cnt_hit, ss_hit(centre_hit), sum_miss, cnt_miss, ss_miss(centre_miss)
__global__ void surf_hitmiss_moments_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, int64_t row0, int64_t nrows,
                                            const double *__restrict__ y, double ch, double cm, double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= nrows) return;
    const int64_t g = row0 + r;
    const double yg = y[g];
    double v[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t j = lane; j < N; j += 32) {
        const double d = D[r * ldd + j];
        const bool same = y[j] == yg;
        if (same && j == g) continue;
        if (same) { v[0] += d; v[1] += 1.0; v[2] += (d - ch) * (d - ch); }
        else { v[3] += d; v[4] += 1.0; v[5] += (d - cm) * (d - cm); }
    }
    for (int k = 0; k < 6; ++k) for (int o = 16; o; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) for (int k = 0; k < 6; ++k) out[r * 6 + k] = v[k];
}

// threshold selection (:529-533): hits = same class & d < r_hit & d > 0, then misses = other class & d < r_miss
__global__ void count_hitmiss_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, int64_t row0, int64_t nrows,
                                     const double *__restrict__ y, const double *__restrict__ r_hit, const double *__restrict__ r_miss,
                                     int scalar, int64_t *__restrict__ counts, int32_t *__restrict__ n_hits) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= nrows) return;
    const double rh = scalar ? r_hit[0] : r_hit[r], rm = scalar ? r_miss[0] : r_miss[r], yg = y[row0 + r];
    const double *row = D + r * ldd;
    int ch = 0, cm = 0;
    for (int64_t j = lane; j < N; j += 32) {
        const double v = row[j];
        const bool same = y[j] == yg;
        ch += (same && v < rh && v > 0.0) ? 1 : 0;
        cm += (!same && v < rm) ? 1 : 0;
    }
    for (int o = 16; o; o >>= 1) { ch += __shfl_xor_sync(0xffffffffu, ch, o); cm += __shfl_xor_sync(0xffffffffu, cm, o); }
    if (lane == 0) { counts[r] = ch + cm; n_hits[r] = ch; }
}
__global__ void write_hitmiss_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, int64_t row0, int64_t nrows,
                                     const double *__restrict__ y, const double *__restrict__ r_hit, const double *__restrict__ r_miss,
                                     int scalar, const int64_t *__restrict__ ptr, const int32_t *__restrict__ n_hits,
                                     int32_t *__restrict__ Ri, int32_t *__restrict__ NN) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= nrows) return;
    const double rh = scalar ? r_hit[0] : r_hit[r], rm = scalar ? r_miss[0] : r_miss[r], yg = y[row0 + r];
    const double *row = D + r * ldd;
    int64_t oh = ptr[r], om = ptr[r] + n_hits[r];
    const int32_t ri = (int32_t)(row0 + r + 1);
    for (int64_t j0 = 0; j0 < N; j0 += 32) {
        const int64_t j = j0 + lane;
        const double v = (j < N) ? row[j] : 0.0;
        const bool same = (j < N) && (y[j] == yg);
        const bool sh = (j < N) && same && v < rh && v > 0.0;
        const bool sm = (j < N) && !same && v < rm;
        const unsigned mh = __ballot_sync(0xffffffffu, sh), mm = __ballot_sync(0xffffffffu, sm);
        const unsigned lt = (1u << lane) - 1u;
        if (sh) { int64_t pos = oh + __popc(mh & lt); Ri[pos] = ri; NN[pos] = (int32_t)(j + 1); }
        if (sm) { int64_t pos = om + __popc(mm & lt); Ri[pos] = ri; NN[pos] = (int32_t)(j + 1); }
        oh += __popc(mh); om += __popc(mm);
    }
}

// relieff (:385-427): order(d[Ri, ]) restricted to the hits / to the misses; neighbours = hits[2:(k.hits+1)], misses[1:k.miss].
// One CTA per owned row.  For each class subset: 8-pass radix select of the rank-n key, gather everything below the cut,
// then as many cut-valued entries as still needed in ascending index order (order() breaks ties by index), bitonic sort
// by (key, index), write.  n <= RH_CAP entries per subset.
constexpr int RH_CAP = 8192;
__device__ void hm_select_sorted(const double *__restrict__ row, const double *__restrict__ y, double yg, bool want_same,
                                 int64_t N, int n_want, uint64_t *skeys, int32_t *sidx, unsigned *hist, int *n_out) {
    __shared__ uint64_t prefix_s, mask_s;
    __shared__ int64_t rank_s;
    __shared__ int fill_s, need_s, wsum[8];
    if (threadIdx.x == 0) { prefix_s = 0; mask_s = 0; rank_s = n_want; fill_s = 0; }
    __syncthreads();
    if (n_want <= 0) { if (threadIdx.x == 0) *n_out = 0; __syncthreads(); return; }
    for (int pass = 7; pass >= 0; --pass) {
        hist[threadIdx.x] = 0;
        __syncthreads();
        const uint64_t prefix = prefix_s, mask = mask_s;
        const int shift = pass * 8;
        for (int64_t j = threadIdx.x; j < N; j += 256) {
            if ((y[j] == yg) != want_same) continue;
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
    for (int64_t j = threadIdx.x; j < N; j += 256) {               // strictly below the cut: all of them (< n_want)
        if ((y[j] == yg) != want_same) continue;
        uint64_t key = dkey(row[j]);
        if (key < cut) { int pos = atomicAdd(&fill_s, 1); skeys[pos] = key; sidx[pos] = (int32_t)(j + 1); }
    }
    __syncthreads();
    if (threadIdx.x == 0) need_s = n_want - fill_s;
    __syncthreads();
    // cut-valued entries in ascending index order until n_want is reached (ordered block compaction)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t base = 0; base < N && need_s > 0; base += 256) {
        const int64_t j = base + threadIdx.x;
        const bool f = (j < N) && ((y[j] == yg) == want_same) && (dkey(row[j]) == cut);
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) wsum[w] = __popc(m);
        __syncthreads();
        int off = 0, tot = 0;
        for (int i = 0; i < 8; ++i) { if (i < w) off += wsum[i]; tot += wsum[i]; }
        const int pos = off + __popc(m & ((1u << lane) - 1u));
        const int need = need_s, fill = fill_s;
        if (f && pos < need) { skeys[fill + pos] = cut; sidx[fill + pos] = (int32_t)(j + 1); }
        __syncthreads();
        if (threadIdx.x == 0) { const int take = tot < need ? tot : need; fill_s = fill + take; need_s = need - take; }
        __syncthreads();
    }
    const int n = fill_s;
    bitonic_sort(skeys, sidx, n);
    if (threadIdx.x == 0) *n_out = n;
    __syncthreads();
}

__global__ void __launch_bounds__(256)
relieff_hitmiss_kernel(const double *__restrict__ D, int64_t ldd, int64_t N, int64_t row0, int64_t nrows,
                       const double *__restrict__ y, double y_first, int kh_first, int km_first, int kh_other, int km_other,
                       const int64_t *__restrict__ ptr, int32_t *__restrict__ Ri, int32_t *__restrict__ NN) {
    extern __shared__ unsigned char sm_raw[];
    uint64_t *skeys = (uint64_t *)sm_raw;
    int32_t *sidx = (int32_t *)(skeys + RH_CAP);
    __shared__ unsigned hist[256];
    __shared__ int n_sel;
    const int64_t r = blockIdx.x;
    const double *row = D + r * ldd;
    const double yg = y[row0 + r];
    const int kh = (yg == y_first) ? kh_first : kh_other, km = (yg == y_first) ? km_first : km_other;
    int64_t off = ptr[r];
    const int32_t ri = (int32_t)(row0 + r + 1);
    hm_select_sorted(row, y, yg, true, N, kh + 1, skeys, sidx, hist, &n_sel);       // hits incl. the first of the order (self)
    int n = n_sel;
    for (int t = threadIdx.x + 1; t < n; t += 256) { Ri[off + t - 1] = ri; NN[off + t - 1] = sidx[t]; }   // hits[2:(k.hits+1)]
    off += (n > 0 ? n - 1 : 0);
    __syncthreads();
    hm_select_sorted(row, y, yg, false, N, km, skeys, sidx, hist, &n_sel);          // misses[1:k.miss]
    n = n_sel;
    for (int t = threadIdx.x; t < n; t += 256) { Ri[off + t] = ri; NN[off + t] = sidx[t]; }
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

extern "C" int npdrb_surf_radius_exact_host(const double *D_host, int64_t ldd, int64_t N, double sd_frac, double *radius) {
    if (!D_host || !radius || N < 3 || ldd < N) { npdrb_set_error("npdrb_surf_radius_exact_host: bad arguments"); return NPDRB_EINVAL; }
    *radius = surf_radius_exact_host_impl(D_host, ldd, N, sd_frac);
    return NPDRB_OK;
}

// radius from the combined partial moments of npdrb_session_surf_partial (m0: centre 0, m1: centre = m0[1]/m0[2]) and a
// rigorous bound on its distance from R's value.  R side: three sequential x87 chains of <= N^2 non-negative terms, each
// with relative error <= n 2^-64, plus a handful of double roundings; this side: per-row sums in double over 32 strided
// lanes + a shuffle tree (relative error <= (N/32 + 8) 2^-53 per moment), combined in long double.
extern "C" int npdrb_surf_radius_from_moments(const double *m0, const double *m1, int64_t N, double sd_frac, double *radius, double *eps) {
    if (!m0 || !m1 || !radius || N < 3) { npdrb_set_error("npdrb_surf_radius_from_moments: bad arguments"); return NPDRB_EINVAL; }
    const double num_pair = (double)N * (double)(N - 1) / 2.0;
    const double mean = m0[0] / (2.0 * num_pair);
    const double sdv = sqrt(m1[3] / (m1[2] - 1.0));
    *radius = mean - sd_frac * sdv;
    if (eps) {
        const double n2 = (double)N * (double)N;
        const double rel = 2.0 * n2 * 5.421010862427522e-20 /* 2^-64 */ + ((double)N / 32.0 + 24.0) * 2.220446049250313e-16 /* 2^-52 */;
        *eps = 1.25 * rel * (fabs(mean) + fabs(sd_frac) * sdv);
    }
    return NPDRB_OK;
}

// number of distances of the owned block inside [lo, hi]
extern "C" int npdrb_session_surf_band_count(npdrb_session *s, double lo, double hi, int64_t *count) {
    npdrb_ctx *ctx = s->ctx;
    if (!s->dD || !count) { npdrb_set_error("surf band count: no distance block"); return NPDRB_EINVAL; }
    NB_CUDA(cudaSetDevice(ctx->device));
    unsigned long long *d_c = nullptr, h_c = 0;
    nb_scope scope;
    NB_CHECK(scope.alloc(&d_c, 1));
    NB_CUDA(cudaMemsetAsync(d_c, 0, sizeof(unsigned long long), ctx->stream));
    surf_band_count_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(s->dD, s->ldd, s->N, s->nrows, lo, hi, d_c);
    NB_LAUNCH_CHECK(ctx);
    NB_CUDA(cudaMemcpyAsync(&h_c, d_c, sizeof(h_c), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    *count = (int64_t)h_c;
    return NPDRB_OK;
}

// Partial SURF moments of the owned rows (see npdr_b200.h).  Each unordered pair {g, j} is counted once,
// by the owner of row g = min index, through the entries j > g of that row.
extern "C" int npdrb_session_surf_partial(npdrb_session *s, double centre, double *out4) {
    npdrb_ctx *ctx = s->ctx;
    if (!s->dD) { npdrb_set_error("surf partial: no distance block"); return NPDRB_EINVAL; }
    const int64_t N = s->N, nrows = s->nrows;
    double *d_out = nullptr, *d_ss = nullptr;
    nb_scope scope;
    NB_CHECK(scope.alloc(&d_out, 3 * (size_t)nrows));
    NB_CHECK(scope.alloc(&d_ss, (size_t)nrows));
    surf_moments_kernel<<<(unsigned)nb_cdiv(nrows, 8), 256, 0, ctx->stream>>>(s->dD, s->ldd, N, s->row0, nrows, d_out);
    NB_LAUNCH_CHECK(ctx);
    surf_sumsq_kernel<<<(unsigned)nb_cdiv(nrows, 8), 256, 0, ctx->stream>>>(s->dD, s->ldd, N, s->row0, nrows, centre, d_ss);
    NB_LAUNCH_CHECK(ctx);
    std::vector<double> h(3 * (size_t)nrows), hs((size_t)nrows);
    NB_CUDA(cudaMemcpyAsync(h.data(), d_out, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaMemcpyAsync(hs.data(), d_ss, sizeof(double) * hs.size(), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
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
    } else if (s->rect) {
        const int64_t nrows = s->nrows;
        if (N * nrows <= (1ll << 18)) {
            surf_rect_serial_kernel<<<1, 32, 0, ctx->stream>>>(s->dD, s->ldd, N, nrows, sd_frac, d_rad_scalar);
            NB_LAUNCH_CHECK(ctx);
            return NPDRB_OK;
        }
        // large blocks: parallel moments (a few ulp from R's sequential long-double chains)
        double *d_m = nullptr;
        nb_scope scope;
        NB_CHECK(scope.alloc(&d_m, (size_t)nrows));
        std::vector<double> h((size_t)nrows);
        long double tot = 0, ss = 0;
        rect_row_moment_kernel<<<(unsigned)nb_cdiv(nrows, 8), 256, 0, ctx->stream>>>(s->dD, s->ldd, N, nrows, 0, 0.0, d_m);
        NB_LAUNCH_CHECK(ctx);
        NB_CUDA(cudaMemcpyAsync(h.data(), d_m, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
        NB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (double v : h) tot += v;
        const double mean = (double)(tot / (long double)(N * nrows));
        rect_row_moment_kernel<<<(unsigned)nb_cdiv(nrows, 8), 256, 0, ctx->stream>>>(s->dD, s->ldd, N, nrows, 1, mean, d_m);
        NB_LAUNCH_CHECK(ctx);
        NB_CUDA(cudaMemcpyAsync(h.data(), d_m, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
        NB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (double v : h) ss += v;
        radius = mean - sd_frac * sqrt((double)(ss / (long double)(N * nrows - 1)));
    } else if (s->nrows != N) {
        npdrb_set_error("surf on a row block needs npdrb_session_set_surf_radius (combine npdrb_session_surf_partial across ranks)");
        return NPDRB_EINVAL;
    } else if (N <= 1024) {
        surf_radius_serial_kernel<<<1, 32, 0, ctx->stream>>>(s->dD, s->ldd, N, sd_frac, d_rad_scalar);
        NB_LAUNCH_CHECK(ctx);
        return NPDRB_OK;
    } else {
        // large N: radius from parallel moments + guard band; exact host replay only if a distance falls inside the band
        double m0[4], m1[4], eps = 0.0;
        NB_CHECK(npdrb_session_surf_partial(s, 0.0, m0));
        const double mean_u = m0[1] / m0[2];
        NB_CHECK(npdrb_session_surf_partial(s, mean_u, m1));
        NB_CHECK(npdrb_surf_radius_from_moments(m0, m1, N, sd_frac, &radius, &eps));
        int64_t in_band = 0;
        const char *force = getenv("NPDRB_SURF_FORCE_EXACT");           // tests: always take the replay
        NB_CHECK(npdrb_session_surf_band_count(s, radius - eps, radius + eps, &in_band));
        if (in_band > 0 || (force && force[0] == '1')) {
            std::vector<double> hD((size_t)N * (size_t)N);
            NB_CHECK(npdrb_session_copy_distances(s, hD.data()));
            NB_CHECK(npdrb_surf_radius_exact_host(hD.data(), N, N, sd_frac, &radius));
        }
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
    const int64_t nmax = N > nrows ? N : nrows;                 // rectangular blocks may have more rows than columns
    if (!s->d_radius) NB_CUDA(cudaMalloc(&s->d_radius, sizeof(double) * (size_t)nmax));
    if (!s->d_rowptr) NB_CUDA(cudaMalloc(&s->d_rowptr, sizeof(int64_t) * (size_t)(nmax + 2)));
    int64_t *d_counts = nullptr, *d_nempty = nullptr;
    nb_scope scope;                                     // temporaries: released on every return path
    NB_CHECK(scope.alloc(&d_counts, (size_t)nrows));
    NB_CHECK(scope.alloc(&d_nempty, 1));
    int64_t tot[2] = {0, 0};
    const unsigned wgrid = (unsigned)nb_cdiv(nrows, 8);

    if (nbd_method == NPDRB_NBD_RELIEFF) {
        if (k <= 0) k = npdrb_knn_surf(N, sd_frac);             // R/nearestNeighbors.R:197-203
        uint64_t *d_cut = nullptr, *d_scratch = nullptr;
        NB_CHECK(scope.alloc(&d_cut, (size_t)nrows));
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
        NB_CHECK(scope.alloc(&d_scratch, (size_t)(M > 0 ? M : 1)));
        const size_t smem = (size_t)SORT_CAP * (sizeof(uint64_t) + sizeof(int32_t));
        NB_CUDA(cudaFuncSetAttribute(relieff_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        relieff_write_kernel<<<(unsigned)nrows, 256, smem, ctx->stream>>>(s->dD, s->ldd, N, s->row0, nrows, d_cut,
                                                                          s->d_rowptr, d_scratch, s->dRiL, s->dNNL);
        NB_LAUNCH_CHECK(ctx);
        NB_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
        int scalar = 0;
        if (nbd_method == NPDRB_NBD_MULTISURF) {
            if (s->rect)
                radius_multisurf_kernel<true><<<(unsigned)nb_cdiv(nrows, 128), 128, 0, ctx->stream>>>(
                    s->dD, s->ldd, N, s->row0, nrows, sd_frac, s->d_sd_vec, s->d_radius);
            else
                radius_multisurf_kernel<false><<<(unsigned)nb_cdiv(nrows, 128), 128, 0, ctx->stream>>>(
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

// nearestNeighborsSeparateHitMiss on the owned rows (uses the session's outcome vector as class labels)
int nb_neighbors_hitmiss(npdrb_session *s, int32_t nbd_method, double sd_frac, int64_t k, int64_t *M_local, int64_t *n_empty) {
    npdrb_ctx *ctx = s->ctx;
    (void)k;                                                     // the reference uses k only through sd.frac (:385-393)
    if (!s->dD) { npdrb_set_error("neighbors_hitmiss: compute or set distances first"); return NPDRB_EINVAL; }
    if (!s->dy) { npdrb_set_error("neighbors_hitmiss: the outcome (class labels) is not set"); return NPDRB_EINVAL; }
    if (nbd_method < 0 || nbd_method > NPDRB_NBD_RELIEFF) { npdrb_set_error("neighbors_hitmiss: bad nbd_method"); return NPDRB_EINVAL; }
    const int64_t N = s->N, nrows = s->nrows;
    if (N < 4) { npdrb_set_error("neighbors_hitmiss: need at least 4 instances"); return NPDRB_EINVAL; }
    NB_CUDA(cudaEventRecord(s->ev[0], ctx->stream));
    // class sizes (two classes; the reference's table(pheno.vec))
    std::vector<double> hy((size_t)N);
    NB_CUDA(cudaMemcpyAsync(hy.data(), s->dy, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    const double y_first = hy[0];
    int64_t n_first = 0;
    for (int64_t i = 0; i < N; ++i) n_first += (hy[(size_t)i] == y_first) ? 1 : 0;
    const int64_t n_other = N - n_first;
    if (n_first == 0 || n_other == 0) { npdrb_set_error("neighbors_hitmiss: the outcome has a single class"); return NPDRB_EDATA; }

    if (s->dRiL) { cudaFree(s->dRiL); s->dRiL = nullptr; }
    if (s->dNNL) { cudaFree(s->dNNL); s->dNNL = nullptr; }
    if (!s->d_radius) NB_CUDA(cudaMalloc(&s->d_radius, sizeof(double) * (size_t)N));
    if (!s->d_radius2) NB_CUDA(cudaMalloc(&s->d_radius2, sizeof(double) * (size_t)N));
    if (!s->d_rowptr) NB_CUDA(cudaMalloc(&s->d_rowptr, sizeof(int64_t) * (size_t)(N + 2)));
    int64_t tot[2] = {0, 0};
    if (nbd_method == NPDRB_NBD_RELIEFF) {
        int64_t kh[2], km[2];                                    // [class of y_first, other class]
        if (n_first == n_other) {
            kh[0] = kh[1] = km[0] = km[1] = (int64_t)floor(0.5 * (double)npdrb_knn_surf(N - 1, sd_frac));
        } else {
            kh[0] = npdrb_knn_surf(n_first - 1, sd_frac); km[0] = npdrb_knn_surf(n_other, sd_frac);
            kh[1] = npdrb_knn_surf(n_other - 1, sd_frac); km[1] = npdrb_knn_surf(n_first, sd_frac);
        }
        for (int c = 0; c < 2; ++c) {
            const int64_t nh = c == 0 ? n_first : n_other, nm = N - nh;
            if (kh[c] + 1 > nh) kh[c] = nh - 1;                  // positions past the end of the list are dropped (R: NA)
            if (km[c] > nm) km[c] = nm;
            if (kh[c] < 0) kh[c] = 0;
            if (kh[c] + 1 > RH_CAP || km[c] > RH_CAP) { npdrb_set_error("neighbors_hitmiss: k.hits/k.miss above %d not supported", RH_CAP - 1); return NPDRB_EINVAL; }
        }
        std::vector<int64_t> hptr((size_t)nrows + 1);
        int64_t run = 0, empty = 0;
        for (int64_t r = 0; r < nrows; ++r) {
            const int c = (hy[(size_t)(s->row0 + r)] == y_first) ? 0 : 1;
            hptr[(size_t)r] = run;
            const int64_t cnt = kh[c] + km[c];
            if (cnt == 0) empty++;
            run += cnt;
        }
        hptr[(size_t)nrows] = run;
        tot[0] = run; tot[1] = empty;
        NB_CUDA(cudaMemcpyAsync(s->d_rowptr, hptr.data(), sizeof(int64_t) * hptr.size(), cudaMemcpyHostToDevice, ctx->stream));
        NB_CUDA(cudaMalloc(&s->dRiL, sizeof(int32_t) * (size_t)(run > 0 ? run : 1)));
        NB_CUDA(cudaMalloc(&s->dNNL, sizeof(int32_t) * (size_t)(run > 0 ? run : 1)));
        const size_t smem = (size_t)RH_CAP * (sizeof(uint64_t) + sizeof(int32_t));
        NB_CUDA(cudaFuncSetAttribute(relieff_hitmiss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        relieff_hitmiss_kernel<<<(unsigned)nrows, 256, smem, ctx->stream>>>(s->dD, s->ldd, N, s->row0, nrows, s->dy, y_first,
                                                                            (int)kh[0], (int)km[0], (int)kh[1], (int)km[1],
                                                                            s->d_rowptr, s->dRiL, s->dNNL);
        NB_LAUNCH_CHECK(ctx);
        NB_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
        int scalar = 0;
        if (nbd_method == NPDRB_NBD_MULTISURF) {
            radius_hitmiss_kernel<<<(unsigned)nb_cdiv(nrows, 128), 128, 0, ctx->stream>>>(s->dD, s->ldd, N, s->row0, nrows, s->dy,
                                                                                        sd_frac, s->d_radius, s->d_radius2);
            NB_LAUNCH_CHECK(ctx);
        } else {
            scalar = 1;
            if (nrows != N) { npdrb_set_error("surf hit/miss radii on a row block are not supported"); return NPDRB_EINVAL; }
            double *d_two = nullptr;
            nb_scope surf_scope;
            NB_CHECK(surf_scope.alloc(&d_two, 2));
            double two[2];
            if (N <= 1024) {
                surf_hitmiss_serial_kernel<<<1, 32, 0, ctx->stream>>>(s->dD, s->ldd, N, s->dy, sd_frac, d_two);
                NB_LAUNCH_CHECK(ctx);
                NB_CUDA(cudaMemcpyAsync(two, d_two, sizeof(two), cudaMemcpyDeviceToHost, ctx->stream));
                NB_CUDA(cudaStreamSynchronize(ctx->stream));
            } else {                                             // parallel moments (documented deviation, <= few ulp)
                double *d_m = nullptr;
                NB_CHECK(surf_scope.alloc(&d_m, 6 * (size_t)nrows));
                std::vector<double> hm(6 * (size_t)nrows);
                double centre[2] = {0.0, 0.0};
                long double acc[6];
                for (int round = 0; round < 2; ++round) {
                    surf_hitmiss_moments_kernel<<<(unsigned)nb_cdiv(nrows, 8), 256, 0, ctx->stream>>>(s->dD, s->ldd, N, s->row0, nrows,
                                                                                                    s->dy, centre[0], centre[1], d_m);
                    NB_LAUNCH_CHECK(ctx);
                    NB_CUDA(cudaMemcpyAsync(hm.data(), d_m, sizeof(double) * hm.size(), cudaMemcpyDeviceToHost, ctx->stream));
                    NB_CUDA(cudaStreamSynchronize(ctx->stream));
                    for (int q6 = 0; q6 < 6; ++q6) acc[q6] = 0;
                    for (int64_t r = 0; r < nrows; ++r) for (int q6 = 0; q6 < 6; ++q6) acc[q6] += hm[(size_t)r * 6 + q6];
                    centre[0] = (double)(acc[0] / acc[1]); centre[1] = (double)(acc[3] / acc[4]);
                }
                two[0] = centre[0] - sd_frac * sqrt((double)(acc[2] / (acc[1] - 1)));
                two[1] = centre[1] - sd_frac * sqrt((double)(acc[5] / (acc[4] - 1)));
            }
            std::vector<double> rep((size_t)nrows, two[0]), rep2((size_t)nrows, two[1]);
            NB_CUDA(cudaMemcpyAsync(s->d_radius, rep.data(), sizeof(double) * rep.size(), cudaMemcpyHostToDevice, ctx->stream));
            NB_CUDA(cudaMemcpyAsync(s->d_radius2, rep2.data(), sizeof(double) * rep2.size(), cudaMemcpyHostToDevice, ctx->stream));
            NB_CUDA(cudaStreamSynchronize(ctx->stream));
            scalar = 0;                                          // radii are replicated per row
        }
        int64_t *d_counts = nullptr, *d_nempty = nullptr; int32_t *d_nh = nullptr;
        nb_scope scope;
        NB_CHECK(scope.alloc(&d_counts, (size_t)nrows));
        NB_CHECK(scope.alloc(&d_nempty, 1));
        NB_CHECK(scope.alloc(&d_nh, (size_t)nrows));
        const unsigned wgrid = (unsigned)nb_cdiv(nrows, 8);
        count_hitmiss_kernel<<<wgrid, 256, 0, ctx->stream>>>(s->dD, s->ldd, N, s->row0, nrows, s->dy, s->d_radius, s->d_radius2, scalar,
                                                             d_counts, d_nh);
        NB_LAUNCH_CHECK(ctx);
        scan_counts_kernel<<<1, 1024, 0, ctx->stream>>>(d_counts, nrows, s->d_rowptr, d_nempty);
        NB_LAUNCH_CHECK(ctx);
        NB_CUDA(cudaMemcpyAsync(&tot[0], s->d_rowptr + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        NB_CUDA(cudaMemcpyAsync(&tot[1], d_nempty, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        NB_CUDA(cudaStreamSynchronize(ctx->stream));
        const int64_t M = tot[0];
        NB_CUDA(cudaMalloc(&s->dRiL, sizeof(int32_t) * (size_t)(M > 0 ? M : 1)));
        NB_CUDA(cudaMalloc(&s->dNNL, sizeof(int32_t) * (size_t)(M > 0 ? M : 1)));
        write_hitmiss_kernel<<<wgrid, 256, 0, ctx->stream>>>(s->dD, s->ldd, N, s->row0, nrows, s->dy, s->d_radius, s->d_radius2, scalar,
                                                             s->d_rowptr, d_nh, s->dRiL, s->dNNL);
        NB_LAUNCH_CHECK(ctx);
        NB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
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
    nb_scope scope;                                     // temporaries: released on every return path
    NB_CHECK(scope.alloc(&d_bits, words));
    NB_CHECK(scope.alloc(&d_keep, (size_t)M));
    NB_CHECK(scope.alloc(&d_ndup, 3));      // [0] duplicates, [1] "not grouped by Ri" flag, [2] distinct ordered pairs
    NB_CUDA(cudaMemsetAsync(d_bits, 0, sizeof(unsigned) * words, ctx->stream));
    NB_CUDA(cudaMemsetAsync(d_ndup, 0, 3 * sizeof(unsigned long long), ctx->stream));
    const unsigned grid = (unsigned)nb_cdiv(M, 256);
    mark_pairs_kernel<<<grid, 256, 0, ctx->stream>>>(s->dRi, s->dNN, M, N, d_bits);
    NB_LAUNCH_CHECK(ctx);
    flag_dups_kernel<<<grid, 256, 0, ctx->stream>>>(s->dRi, s->dNN, M, N, d_bits, d_keep, d_ndup);
    NB_LAUNCH_CHECK(ctx);
    popcount_words_kernel<<<1024, 256, 0, ctx->stream>>>(d_bits, words, d_ndup + 2);
    NB_LAUNCH_CHECK(ctx);
    unsigned long long h3[3] = {0, 0, 0};
    NB_CUDA(cudaMemcpyAsync(h3, d_ndup, sizeof(h3), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    // The "i > j and (j, i) present" rule equals uniqueNeighbors' keep-the-first-occurrence (R/nearestNeighbors.R:576-582)
    // only for lists grouped by ascending Ri_idx without repeated ordered pairs -- what nearestNeighbors() emits.  Anything
    // else (a hand-made list through npdrb_session_import_pairs*) is rejected rather than answered differently from R.
    if (h3[1] != 0 || (int64_t)h3[2] != M) {
        npdrb_set_error("unique_pairs: the pair list must be grouped by ascending Ri_idx and must not repeat an ordered pair "
                        "(%lld rows, %llu distinct pairs%s)", (long long)M, h3[2], h3[1] ? ", Ri_idx decreases somewhere" : "");
        return NPDRB_EINVAL;
    }
    const unsigned long long ndup = h3[0];
    const int64_t nu = M - (int64_t)ndup;
    if (n_unique) *n_unique = nu;
    if (keep_unique_only && ndup > 0) {
        int64_t *d_bc = nullptr, *d_bp = nullptr, *d_ne = nullptr;
        int32_t *nRi = nullptr, *nNN = nullptr;
        NB_CHECK(scope.alloc(&d_bc, (size_t)grid));
        NB_CHECK(scope.alloc(&d_bp, (size_t)grid + 1));
        NB_CHECK(scope.alloc(&d_ne, 1));
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
    }
    return NPDRB_OK;
}
