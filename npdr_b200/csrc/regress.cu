// regress.cu -- K6/K7: per-attribute GLM on projected pair differences (the regression loop of
// npdr(), R/npdr.R:391-430, and diffRegression, R/npdr.R:18-71 -> base-R lm / glm.fit).
//
// The reference gathers 2M values per attribute and runs a QR (lm) or IRLS-with-QR (binomial) on an
// M x (2+q) design, p times.  Here the pair list is re-bucketed once into (I-block x J-block) tiles
// of 64 x 64 instances; a CTA owns 128 attributes x 4 pair slices, stages the two 64 x 128 attribute
// tiles (from an instance-major copy of X) plus the tile's pair records in shared memory, and every
// thread streams the records through FP64 registers, accumulating its attribute's sufficient
// statistics:
//   lm        one pass:  sum d, d^2, d*y, d*c_r            -> closed-form OLS / t / R^2
//   binomial  pass 0 (weights at mustart) + one pass per IRLS iteration:
//             X'WX (upper triangle), X'Wz, deviance        -> (2+q)^2 Cholesky per attribute
// Partials are reduced in a fixed order (deterministic), the small solves run one thread per
// attribute, and attributes that have converged stop contributing work at tile granularity.
//
// Bound: FP64 issue rate (binomial: ~52 FP64 instr per attribute x pair x pass at q = 2, exp/log via
// dmath.cuh); staged bytes per evaluation are < 1 B, so HBM is far from binding.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <math.h>
#include <mutex>
#include "common.cuh"
#include "dmath.cuh"

namespace {

constexpr int BI = 64, BJ = 64;        // instance block sizes (tile = BI x BJ)
constexpr int TA = 128;                // attributes per CTA
constexpr int NSLICE = 4;              // pair slices per attribute
constexpr int RTHREADS = TA * NSLICE;  // 512
constexpr int RCHUNK = 1024;           // records staged per step
constexpr int KMAX = 2 + NB_FAST_COVARS;

// MODE_LM: moments with the outcome diff from the yv stream; MODE_SIGN: the same moments with the outcome
// taken as s = +1 (miss) / -1 (hit) from the record bit (closed-form first IRLS pass); MODE_IRLS_FULL: one IRLS pass.
enum { MODE_LM = 0, MODE_SIGN = 1, MODE_IRLS_FULL = 2 };

__host__ __device__ constexpr int nacc_for(int mode, int q) {
    return mode != MODE_IRLS_FULL ? (3 + q) : ((2 + q) * (3 + q) / 2 + (2 + q) + 2);
}
// IRLS pass statistics per attribute (NA = K(K+1)/2 + K + 2, K = 2 + q):
//   [0 .. K(K+1)/2)   X'WX upper triangle, row-major over r <= s, features v = [1, d, c_1..c_q], W = mu(1-mu)
//   next K            sum of (1 - mu) * v over ALL records.  The score X'(y - mu) = that minus the sum of v over the HIT
//                     records (y - mu = (1 - mu) - 1 for a hit, 1 - mu for a miss); the hit moments are known from pass 0,
//                     so the pass does not need to look at the outcome.  (Newton form: beta_new = beta +
//                     (X'WX)^-1 X'(y - mu), identical to glm.fit's weighted least squares on z = eta + (y - mu)/mu.eta.)
//   NA-2              sum of log(1 + e^eta)
//   NA-1              clamp correction: sum over clamped MISS records of (log t - eta) (zero unless |eta| > 30)
//   residual deviance = 2 * ([NA-2] - beta . (sum of v over the MISS records) - [NA-1])

// ------------------------------------------------------------------ layout helpers
// instance-major copy of X for attribute columns [attr0, attr0+nattr): XT[i * ldt + (a - attr0)]
__global__ void transpose_kernel(const double *__restrict__ X, int64_t ldx, int64_t N, int64_t attr0, int64_t nattr,
                                 double *__restrict__ XT, int64_t ldt) {
    __shared__ double t[32][33];
    const int64_t i0 = (int64_t)blockIdx.x * 32, a0 = (int64_t)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        int64_t a = a0 + r, i = i0 + threadIdx.x;
        t[r][threadIdx.x] = (a < nattr && i < N) ? X[(attr0 + a) * ldx + i] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        int64_t i = i0 + r, a = a0 + threadIdx.x;
        if (i < N && a < nattr) XT[i * ldt + a] = t[threadIdx.x][r];
    }
}

// sort key = 2 * tile + outcome bit: inside every tile the hit records (bit 0) precede the miss records, so the
// IRLS kernel runs each half with the outcome known at compile time.  ybits: 0 = lm (no bit), 1 = match-mismatch
// of y over the pair, 2 = raw (uniReg: y of the instance itself)
__global__ void tile_keys_kernel(const int32_t *__restrict__ Ri, const int32_t *__restrict__ NN, int64_t M, int n_jb,
                                 const double *__restrict__ y, int ybits, uint32_t *__restrict__ keys,
                                 uint32_t *__restrict__ idx) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int i = Ri[m] - 1, j = NN[m] - 1;
    uint32_t bit = 0;
    if (ybits == 1) bit = (y[i] != y[j]);
    else if (ybits == 2) bit = (y[i] != 0.0);
    keys[m] = 2u * ((uint32_t)(i / BI) * (uint32_t)n_jb + (uint32_t)(j / BJ)) + bit;
    idx[m] = (uint32_t)m;
}

// per attribute: an upper bound of the projected diff over any pair (max - min of the column, mapped through the diff type);
// used by the IRLS step kernel to bound |x'delta| and |eta| (see irls_step_kernel)
__global__ void __launch_bounds__(256)
column_dmax_kernel(const double *__restrict__ X, int64_t ldx, int64_t N, int64_t attr0, int64_t nattr, int diff_type,
                   double *__restrict__ dmax) {
    __shared__ double smin[256], smax[256];
    const int64_t a = blockIdx.x;
    if (a >= nattr) return;
    const double *c = X + (attr0 + a) * ldx;
    double mn = c[0], mx = c[0];
    for (int64_t i = threadIdx.x; i < N; i += 256) { mn = fmin(mn, c[i]); mx = fmax(mx, c[i]); }
    smin[threadIdx.x] = mn; smax[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { smin[threadIdx.x] = fmin(smin[threadIdx.x], smin[threadIdx.x + o]); smax[threadIdx.x] = fmax(smax[threadIdx.x], smax[threadIdx.x + o]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double range = smax[0] - smin[0];
        double v;
        switch (diff_type) {
        case NPDRB_DIFF_NUMERIC_SQR: v = range * range; break;
        case NPDRB_DIFF_ALLELE_SHARING: v = 0.5 * range; break;
        case NPDRB_DIFF_MATCH_MISMATCH: v = 1.0; break;
        case NPDRB_DIFF_RAW: v = fmax(fabs(smax[0]), fabs(smin[0])); break;
        default: v = range;
        }
        dmax[a] = v;
    }
}

__global__ void tile_hist_kernel(const uint32_t *__restrict__ keys, int64_t M, unsigned *__restrict__ hist) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < M) atomicAdd(&hist[keys[m]], 1u);
}

// records in tile order; phenotype diff (R/npdr.R:299-308) and covariate diffs (R/npdr.R:315-345) formed here
struct GatherArgs {
    const int32_t *Ri, *NN;
    const uint32_t *perm;
    const double *y, *C;
    int64_t M, N;
    int q, reg_type;           // q: covariates stored per record (the folded binary covariate is not)
    int cov_type[NB_FAST_COVARS];
    int cov_src[NB_FAST_COVARS];   // column of C behind stored covariate c
    int y_raw;                 // uniReg: outcome of instance Ri itself, no diff
    uint32_t *rec;
    double *yv;
    double *cd[NB_FAST_COVARS];
    unsigned long long *n_miss;
};
__global__ void gather_records_kernel(GatherArgs g) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int miss = 0;
    if (t < g.M) {
        const int64_t m = g.perm[t];
        const int64_t i = g.Ri[m] - 1, j = g.NN[m] - 1;
        // record = byte offsets of the two instance rows inside a staged [64][128] double tile, plus the outcome bit:
        // (il << 10) | (jl << 26) | y
        uint32_t r = ((uint32_t)(i % BI) << 10) | ((uint32_t)(j % BJ) << 26);
        const double yi = g.y[i], yj = g.y[j];
        if (g.reg_type == NPDRB_REG_BINOMIAL) {
            miss = g.y_raw ? (yi != 0.0) : (yi != yj);     // match-mismatch: hit = 0, miss = 1
            r |= (uint32_t)miss;
        } else {
            g.yv[t] = g.y_raw ? yi : fabs(yi - yj);         // always numeric-abs (R/npdr.R:303)
        }
        g.rec[t] = r;
        for (int c = 0; c < g.q; ++c) {
            const double *col = g.C + (int64_t)g.cov_src[c] * g.N;
            g.cd[c][t] = g.y_raw ? col[i] : nb_diff1(col[i], col[j], g.cov_type[c]);
        }
    }
    unsigned mask = __ballot_sync(0xffffffffu, miss);
    if ((threadIdx.x & 31) == 0 && mask) atomicAdd(g.n_miss, (unsigned long long)__popc(mask));
}

// ---- mutual-pair folding -------------------------------------------------------------------------
// Rows (i,j) and (j,i) of the pair list give the SAME regression row (same |x_i - x_j|, same phenotype and
// covariate diffs), so every sufficient statistic is sum over unordered pairs of multiplicity x term.  The list
// is split into a weight-1 stream (one-directional pairs) and a weight-2 stream (the i < j half of the mutual
// pairs); the reverse halves are dropped.  Valid when no ordered pair is repeated (checked via popcount).
__global__ void mark_bits_kernel(const int32_t *__restrict__ Ri, const int32_t *__restrict__ NN, int64_t M, int64_t N,
                                 unsigned *__restrict__ bits) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    uint64_t b = (uint64_t)(Ri[m] - 1) * (uint64_t)N + (uint64_t)(NN[m] - 1);
    atomicOr(&bits[b >> 5], 1u << (b & 31));
}
__global__ void popcount_kernel(const unsigned *__restrict__ bits, size_t words, unsigned long long *__restrict__ out) {
    unsigned long long c = 0;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (size_t)gridDim.x * blockDim.x) c += __popc(bits[w]);
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
__global__ void classify_pairs_kernel(const int32_t *__restrict__ Ri, const int32_t *__restrict__ NN, int64_t M, int64_t N,
                                      const unsigned *__restrict__ bits, unsigned char *__restrict__ f_single,
                                      unsigned char *__restrict__ f_double) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int64_t i = Ri[m] - 1, j = NN[m] - 1;
    int mutual = 0;
    if (i != j) {
        uint64_t b = (uint64_t)j * (uint64_t)N + (uint64_t)i;
        mutual = (bits[b >> 5] >> (b & 31)) & 1u;
    }
    f_single[m] = (unsigned char)!mutual;
    f_double[m] = (unsigned char)(mutual && i < j);
}

// shared (attribute-independent) lm moments of the record stream: M, Sy, Syy, Sc_r, Scy_r, Scc_rs
// flags of the pairs whose folded covariate matches (c0) / differs (c1)
__global__ void cov_flags_kernel(const int32_t *__restrict__ Ri, const int32_t *__restrict__ NN, int64_t M,
                                 const double *__restrict__ col, unsigned char *__restrict__ c0, unsigned char *__restrict__ c1) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const unsigned char d = (col[Ri[m] - 1] != col[NN[m] - 1]) ? 1 : 0;
    c1[m] = d; c0[m] = 1 - d;
}

constexpr int NSHARED_MAX = 3 + 2 * NB_FAST_COVARS + NB_FAST_COVARS * (NB_FAST_COVARS + 1) / 2;
// cd[r]: diffs of covariate r over the stream's records, or NULL for the covariate that is constant (fold_val) on this stream
struct SharedArgs { const double *yv; const uint32_t *rec; const double *cd[NB_FAST_COVARS]; double fold_val; int64_t n; int q; double *out; };
__global__ void __launch_bounds__(256)
shared_moments_kernel(SharedArgs a) {
    __shared__ double red[256];
    double acc[NSHARED_MAX];
    const int ns = 3 + 2 * a.q + a.q * (a.q + 1) / 2;
    for (int s = 0; s < NSHARED_MAX; ++s) acc[s] = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < a.n; t += (int64_t)gridDim.x * 256) {
        const double y = a.yv ? a.yv[t] : ((a.rec[t] & 1u) ? 1.0 : -1.0);
        double c[NB_FAST_COVARS];
        for (int r = 0; r < a.q; ++r) c[r] = a.cd[r] ? a.cd[r][t] : a.fold_val;
        int s = 0;
        acc[s++] += 1.0; acc[s++] += y; acc[s++] += y * y;
        for (int r = 0; r < a.q; ++r) acc[s++] += c[r];
        for (int r = 0; r < a.q; ++r) acc[s++] += c[r] * y;
        for (int r = 0; r < a.q; ++r) for (int u = r; u < a.q; ++u) acc[s++] += c[r] * c[u];
    }
    for (int s = 0; s < ns; ++s) {
        red[threadIdx.x] = acc[s];
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
        if (threadIdx.x == 0) a.out[(int64_t)blockIdx.x * NSHARED_MAX + s] = red[0];
        __syncthreads();
    }
}

// ------------------------------------------------------------------ the pass kernel
struct StreamArgs {                // one bucketed pair stream (see nb_stream)
    const uint32_t *rec; const double *yv; const double *cd[NB_FAST_COVARS];
    const int64_t *tile_ptr; const int64_t *tile_mid; const int32_t *tile_ib; const int32_t *tile_jb; const int32_t *chunk_tile;
    double *partials;              // [n_chunks][NACC][nattr_pad]
    const double *beta;            // [2 + Q][nattr_pad] coefficients this stream evaluates with (IRLS_FULL; see fold_beta_kernel)
};
// All streams of a pass go into ONE launch (grid.x = the chunks of stream 0, then those of stream 1, ...), so the pass pays
// one wave-quantisation tail instead of one per stream.
struct PassArgs {
    const double *XT; int64_t ldt; int64_t N; int nattr; int nattr_pad;
    StreamArgs st[NB_MAX_STREAMS]; int chunk_start[NB_MAX_STREAMS + 1];   // first chunk of every stream (unused: the total)
    const unsigned char *tile_active;  // per attribute tile, or NULL
    const nb_tables *tables;       // device copy of the exp/log tables (v1 kernel)
    const nb_tables1024 *tables1024; // 1024-entry exp table (v2 kernel)
    const double *dmax;            // [nattr_pad] bound on the attribute diff (v2 IRLS pass: clamp pre-check)
    double cmax[NB_FAST_COVARS]; // bounds on the covariate diffs
    int diff_type;
    int dbg;                       // development switches (NPDRB_V3_DBG)
};

template <int DIFF>
__device__ __forceinline__ double proj_diff(double xi, double xj, int rt_type) {
    if (DIFF == 0) return fabs(xi - xj);              // numeric-abs / manhattan
    return nb_diff1(xi, xj, rt_type);
}

template <int MODE, int Q, int DIFF>
__global__ void __launch_bounds__(RTHREADS, 1)
pass_kernel(PassArgs P) {
    constexpr int K = 2 + Q;
    constexpr int NA = (MODE != MODE_IRLS_FULL) ? (3 + Q) : (K * (K + 1) / 2 + K + 2);
    extern __shared__ unsigned char smem_raw[];
    double *sXI = (double *)smem_raw;                   // [BI][TA]
    double *sXJ = sXI + BI * TA;                        // [BJ][TA]
    double *sCD = sXJ + BJ * TA;                        // [max(Q,1)][RCHUNK]  (lm: slot Q holds yv)
    uint32_t *sREC = (uint32_t *)(sCD + (Q + 1) * RCHUNK);
    nb_tables *sT = (nb_tables *)(sREC + RCHUNK);

    const int tid = threadIdx.x;
    const int al = tid % TA, slice = tid / TA;
    // chunk index is the fast grid dimension: the CTAs resident at any moment share a few attribute tiles, whose XT slabs
    // (N x 128 doubles each) then stay in L2 while every chunk of the pair stream sweeps over them
    const int at = blockIdx.y;
    int sidx = 0;
#pragma unroll
    for (int t = 1; t < NB_MAX_STREAMS; ++t) sidx += ((int)blockIdx.x >= P.chunk_start[t]) ? 1 : 0;
    const int chunk = (int)blockIdx.x - P.chunk_start[sidx];
    const StreamArgs &A = P.st[sidx];
    if (P.tile_active && !P.tile_active[at]) return;
    const int64_t a0 = (int64_t)at * TA;

    if (MODE == MODE_IRLS_FULL) {
        const double *src = (const double *)P.tables;
        double *dst = (double *)sT;
        for (int e = tid; e < (int)(sizeof(nb_tables) / sizeof(double)); e += RTHREADS) dst[e] = src[e];
    }
    double beta[K];
    if (MODE == MODE_IRLS_FULL) {
#pragma unroll
        for (int r = 0; r < K; ++r) beta[r] = A.beta[(int64_t)r * P.nattr_pad + a0 + al];
    }
    double acc[NA];
#pragma unroll
    for (int n = 0; n < NA; ++n) acc[n] = 0.0;

    const int t_begin = A.chunk_tile[chunk], t_end = A.chunk_tile[chunk + 1];
    int cur_ib = -1;
    for (int t = t_begin; t < t_end; ++t) {
        const int ib = A.tile_ib[t], jb = A.tile_jb[t];
        const int64_t r0 = A.tile_ptr[t], r1 = A.tile_ptr[t + 1];
        __syncthreads();                                  // previous tile fully consumed
        // stage attribute tiles: 16-byte loads, rows are TA contiguous doubles of XT (zero padded)
        {
            const double2 *gJ = reinterpret_cast<const double2 *>(P.XT + ((int64_t)jb * BJ) * P.ldt + a0);
            double2 *sJ = reinterpret_cast<double2 *>(sXJ);
            for (int e = tid; e < BJ * TA / 2; e += RTHREADS) {
                int row = e / (TA / 2), c2 = e % (TA / 2);
                sJ[e] = __ldg(gJ + (int64_t)row * (P.ldt / 2) + c2);
            }
            if (ib != cur_ib) {
                const double2 *gI = reinterpret_cast<const double2 *>(P.XT + ((int64_t)ib * BI) * P.ldt + a0);
                double2 *sI = reinterpret_cast<double2 *>(sXI);
                for (int e = tid; e < BI * TA / 2; e += RTHREADS) {
                    int row = e / (TA / 2), c2 = e % (TA / 2);
                    sI[e] = __ldg(gI + (int64_t)row * (P.ldt / 2) + c2);
                }
                cur_ib = ib;
            }
        }
        for (int64_t rc = r0; rc < r1; rc += RCHUNK) {
            const int n = (int)((r1 - rc < RCHUNK) ? (r1 - rc) : RCHUNK);
            if (rc != r0) __syncthreads();                // previous record chunk consumed
            for (int e = tid; e < n; e += RTHREADS) {
                sREC[e] = A.rec[rc + e];
#pragma unroll
                for (int c = 0; c < Q; ++c) sCD[c * RCHUNK + e] = A.cd[c][rc + e];
                if (MODE == MODE_LM) sCD[Q * RCHUNK + e] = A.yv[rc + e];
            }
            __syncthreads();
#pragma unroll 2
            for (int m = slice; m < n; m += NSLICE) {
                const uint32_t r = sREC[m];
                const double xi = sXI[((r >> 10) & 0x3Fu) * TA + al];
                const double xj = sXJ[(r >> 26) * TA + al];
                const double d = proj_diff<DIFF>(xi, xj, P.diff_type);
                double c[Q > 0 ? Q : 1];
#pragma unroll
                for (int u = 0; u < Q; ++u) c[u] = sCD[u * RCHUNK + m];
                if (MODE != MODE_IRLS_FULL) {
                    const double yv = (MODE == MODE_LM) ? sCD[Q * RCHUNK + m] : ((r & 1u) ? 1.0 : -1.0);
                    acc[0] += d;
                    acc[1] = fma(d, d, acc[1]);
                    acc[2] = fma(d, yv, acc[2]);
#pragma unroll
                    for (int u = 0; u < Q; ++u) acc[3 + u] = fma(d, c[u], acc[3 + u]);
                } else {
                    const bool yb = r & 1u;
                    double w, wz;
                    {
                        double eta = beta[0];
                        eta = fma(beta[1], d, eta);
#pragma unroll
                        for (int u = 0; u < Q; ++u) eta = fma(beta[2 + u], c[u], eta);
                        // logit linkinv clamps (stats family.c): eta < -30 -> DBL_EPSILON, eta > 30 -> 1/DBL_EPSILON
                        const int hi = __double2hiint(eta) & 0x7fffffff;
                        double t, eta_dev = eta;
                        if (hi >= 0x403E0000) {
                            if (eta > 30.0) { t = 4503599627370496.0; eta_dev = NB_LOG_INV_EPS; }
                            else if (eta < -30.0) { t = 2.220446049250313e-16; eta_dev = -NB_LOG_INV_EPS; }
                            else t = nb_exp(eta, sT->exp_t);
                        } else {
                            t = nb_exp(eta, sT->exp_t);
                        }
                        const double u1 = 1.0 + t;
                        double inv;
                        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(inv) : "d"(u1));
                        double e1 = fma(-u1, inv, 1.0);
                        inv = fma(inv, e1, inv);
                        e1 = fma(-u1, inv, 1.0);
                        inv = fma(inv, e1, inv);            // 1/(1+t) = 1 - mu
                        const double mu = t * inv;
                        w = mu * inv;                       // mu(1-mu) = mu.eta (logit)
                        const double lg = nb_log_ge1(u1, sT->log_rc, sT->log_lrc, sT->log_eln2);
                        // deviance residual / 2 = -log(mu) (y=1) or -log(1-mu) (y=0) = log(1+t) - y*log(t)
                        acc[NA - 2] += lg;
                        if (yb && eta_dev != eta) acc[NA - 1] += (eta_dev - eta);
                        wz = inv;                           // (1 - mu); the hit moments are subtracted in the solve
                    }
                    // X'WX upper triangle (row-major over r <= s), features v = [1, d, c...]
                    double wv[K];
                    wv[0] = w;
                    wv[1] = w * d;
#pragma unroll
                    for (int u = 0; u < Q; ++u) wv[2 + u] = w * c[u];
                    int n_ = 0;
#pragma unroll
                    for (int rr = 0; rr < K; ++rr) {
#pragma unroll
                        for (int ss = rr; ss < K; ++ss) {
                            if (rr == 0) acc[n_] += wv[ss];
                            else acc[n_] = fma(wv[rr], (ss == 1) ? d : c[ss - 2 < 0 ? 0 : ss - 2], acc[n_]);
                            ++n_;
                        }
                    }
                    acc[n_] += wz;
                    acc[n_ + 1] = fma(wz, d, acc[n_ + 1]);
#pragma unroll
                    for (int u = 0; u < Q; ++u) acc[n_ + 2 + u] = fma(wz, c[u], acc[n_ + 2 + u]);
                }
            }
        }
    }
    // cross-slice reduction in a fixed order, then one coalesced store per statistic
    __syncthreads();
    double *red = (double *)smem_raw;                      // [NSLICE][NA][TA] <= 2 * BI * TA doubles
#pragma unroll
    for (int n = 0; n < NA; ++n) red[(slice * NA + n) * TA + al] = acc[n];
    __syncthreads();
    for (int e = tid; e < NA * TA; e += RTHREADS) {
        const int n = e / TA, a = e % TA;
        double s = red[(0 * NA + n) * TA + a];
#pragma unroll
        for (int sl = 1; sl < NSLICE; ++sl) s += red[(sl * NA + n) * TA + a];
        A.partials[((int64_t)chunk * NA + n) * P.nattr_pad + a0 + a] = s;
    }
}

// ------------------------------------------------------------------ IRLS pass kernel, v2 (q <= 2)
// Same statistics as pass_kernel<MODE_IRLS_FULL>, restructured for the FP64 issue port (an FP64 warp
// instruction holds the dispatch port for 2 cycles, so cycles per evaluation ~ 2 * FP64 + other instructions):
//   * each thread owns TWO adjacent attributes: one LDS.128 per instance row, one record fetch and one covariate fetch
//     per pair for both; two records are in flight, i.e. four independent exp -> reciprocal chains per thread;
//   * records are ordered hits-then-misses inside each tile and the accumulation is outcome independent: the score is
//     X'(1 - mu) over ALL records plus the hit/miss moments of the closed-form first pass (no per-record select);
//   * record fields are pre-shifted into byte offsets when the chunk is staged; record pairs come in one 8-byte load;
//   * constants come from constant memory (constant-bank operands, no IMAD.MOV/UMOV);
//   * u = 1 + exp(eta) in 7 FP64 instructions (table + economised degree-2 polynomial, below), reciprocal = rcp.approx +
//     ONE Newton step (1.4e-14), weight = (1-mu) - (1-mu)^2, deviance = running product of u (v2_tail / v2_renorm);
//   * R's |eta| > 30 clamps are ruled out per work item from bounds on |x.beta| (may_clamp), so the common path carries no
//     per-record check.
// The v2 kernel carries the linear predictor in units of ln2/1024: s = eta * 1024/ln2 (the coefficients are pre-scaled
// when they are loaded), so k = rint(s) and the reduced argument frac = s - k (exact) need no multiplication, and
// exp(eta) = 2^(k/1024) * exp(frac * ln2/1024) with an economised degree-2 polynomial (1.6e-12 relative; 1024 table
// entries buy one FMA per evaluation over the degree-3 / 256-entry form).  The shared table image stores 2^(j/1024) with
// j * 1024 pre-subtracted from its high word, so one IMAD (k * 1024 + high word) applies the 2^(k >> 10) scale without
// masking k (nb_one_plus_exp1024 in dmath.cuh is the host-tested statement of the same recurrence).
#define NB_C NB_LN2_1024
__constant__ double c_k[10] = {
    6755399441055744.0,                                                  // 0     1.5 * 2^52 (round-to-integer magic)
    0.0, NB_EXP2_C2, NB_EXP2_C1,                                         // 1..3  exp polynomial in frac (degree 2)
    NB_LN2, 0.0, 0.0,                                                    // 4     ln2 (5, 6 unused)
    NB_LOG_INV_EPS * NB_1024_OVER_LN2,                                   // 7     log(1/DBL_EPSILON) in units of ln2/1024
    30.0 * NB_1024_OVER_LN2,                                             // 8     clamp threshold in the same units
    NB_1024_OVER_LN2                                                     // 9
};

struct nb_tables2 {                 // shared-memory image used by v2
    double exp_t[1024];             // 2^(j/1024), high words pre-biased
};

constexpr int V2_TPS = TA / 2;      // threads per slice (64)
// pair slices per CTA: 8 (512 threads, <= 128 registers) except the q = 2 IRLS pass, whose four interleaved
// chains per thread need ~170 registers -> 6 slices (384 threads)
__host__ __device__ constexpr int v2_nslice(int mode, int q) { return (mode == MODE_IRLS_FULL && q >= 2) ? 6 : 8; }
constexpr int V2_RCHUNK = 512;      // records per staged chunk (two buffers)

// u = 1 + exp(eta) (dmath.cuh: nb_one_plus_exp1024 is the host-tested statement of the same recurrence)
__device__ __forceinline__ double v2_u1_fast(double sv, uint32_t sT) {             // sv = eta * 1024/ln2, |eta| < 30
    const double tt = sv + c_k[0];
    const int ki = __double2loint(tt);                           // rint(eta * 1024/ln2)
    const double kf = tt - c_k[0];
    const double fr = sv - kf;                                   // exact, |fr| <= 1/2
    double qq = fma(fr, c_k[2], c_k[3]);
    qq = fma(qq, fr, 1.0);                                       // exp(fr * ln2/1024)
    double tj;
    asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(sT + (uint32_t)((ki << 3) & 0x1FF8)));
    int thi;
    asm("mad.lo.s32 %0, %1, 1024, %2;" : "=r"(thi) : "r"(ki), "r"(__double2hiint(tj)));    // * 2^(k >> 10) (entry is pre-biased by -j * 1024)
    return fma(__hiloint2double(thi, __double2loint(tj)), qq, 1.0);
}
// logit linkinv with R's clamps (stats family.c): eta < -30 -> DBL_EPSILON, eta > 30 -> 1/DBL_EPSILON.
// -> (1 + t, log t - eta): the second component is the deviance correction of a clamped record (0 otherwise)
__device__ __noinline__ double2 v2_exp_clamped(double sv, uint32_t sT) {
    if (sv > c_k[8]) return make_double2(1.0 + 4503599627370496.0, (c_k[7] - sv) * NB_C);
    if (sv < -c_k[8]) return make_double2(1.0 + 2.220446049250313e-16, (-c_k[7] - sv) * NB_C);
    return make_double2(v2_u1_fast(sv, sT), 0.0);
}

// everything after u1 = 1 + exp(eta): weight, deviance, accumulation -- independent of the outcome.
// Deviance: sum log(u1) = log(prod u1).  acc[NA - 2] is a running PRODUCT of the u1 (one DMUL per evaluation instead of
// a table-driven log: 6 FP64 + 3 other instructions); v2_renorm moves its binary exponent into an integer register
// every other record and the kernel takes one log() per thread at the end.  Error: one rounding (2^-53 relative) per
// factor, i.e. <= n * 1.1e-16 ABSOLUTE on a sum of order 0.7 n -- tighter than any summed log.
template <int Q>
__device__ __forceinline__ void v2_tail(const double d, const double (&c)[Q > 0 ? Q : 1], const double u1,
                                        double (&acc)[(2 + Q) * (3 + Q) / 2 + (2 + Q) + 2]) {
    constexpr int K = 2 + Q;
    constexpr int NA = K * (K + 1) / 2 + K + 2;
    double inv;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(inv) : "d"(u1));
    inv = fma(inv, fma(-u1, inv, 1.0), inv);                    // 1/(1+t) = 1 - mu
    const double w = fma(-inv, inv, inv);                       // mu(1-mu) = (1-mu) - (1-mu)^2 = mu.eta for the logit link
    acc[NA - 2] *= u1;                                          // prod (1 + t);  log(1 + t) = -log(1 - mu)
    double wv[K];
    wv[0] = w;
    wv[1] = w * d;
#pragma unroll
    for (int u = 0; u < Q; ++u) wv[2 + u] = w * c[u];
    int n_ = 0;
#pragma unroll
    for (int rr = 0; rr < K; ++rr) {
#pragma unroll
        for (int ss = rr; ss < K; ++ss) {
            if (rr == 0) acc[n_] += wv[ss];
            else acc[n_] = fma(wv[rr], (ss == 1) ? d : c[ss - 2 < 0 ? 0 : ss - 2], acc[n_]);
            ++n_;
        }
    }
    acc[n_] += inv;                                             // sum (1 - mu) v
    acc[n_ + 1] = fma(inv, d, acc[n_ + 1]);
#pragma unroll
    for (int u = 0; u < Q; ++u) acc[n_ + 2 + u] = fma(inv, c[u], acc[n_ + 2 + u]);
}

// running product in [1, 2^107) -> [1, 2): the biased exponent goes to esum (wrapping; 1023 x (calls) comes off at the end)
__device__ __forceinline__ void v2_renorm(double &prod, unsigned &esum) {
    const int hi = __double2hiint(prod);
    esum += (unsigned)(hi >> 20);
    int mhi;
    asm("lop3.b32 %0, %1, 0x000FFFFF, 0x3FF00000, 0xEA;" : "=r"(mhi) : "r"(hi));         // (hi & 0xFFFFF) | 0x3FF00000
    prod = __hiloint2double(mhi, __double2loint(prod));
}

// Per-attribute constants of the q = 0 fast path ("folded intercept").  With s = b0 + b1 d (units of ln2/1024) the range
// reduction k = rint(s), fr = s - k costs an FMA and three adds.  Split b0 = B0i + B0f (integer part, |B0f| <= 1/2) instead:
//   tt = fma(b1, d, M + B0i)   = M + B0i + rint(b1 d) exactly (M = 1.5 * 2^52, ulp 1)  -> low word = the table index k
//   kd = tt - (M + B0i)        = rint(b1 d)
//   r1 = fma(b1, d, -kd)       in [-1/2, 1/2], exact up to one rounding
//   exp(eta) = 2^(k/1024) * exp(B0f c) * exp(r1 c),  c = ln2/1024
// and exp(B0f c) =: E0 is a per-attribute constant folded into the polynomial coefficients (p0, p1, p2) = E0 * (1, C1, C2).
// Six FP64 instructions to u = 1 + exp(eta) instead of seven; the pass executes 17 FP64 instructions per evaluation.
struct V2Fold { double b1, mb, p0, p1, p2; };
__device__ __forceinline__ V2Fold v2_fold(double b0s, double b1s) {
    V2Fold f;
    const double b0i = rint(b0s), b0f = b0s - b0i;
    const double e0 = exp(b0f * NB_C);
    f.b1 = b1s; f.mb = c_k[0] + b0i; f.p0 = e0; f.p1 = e0 * c_k[3]; f.p2 = e0 * c_k[2];
    return f;
}
__device__ __forceinline__ double v2_u1_fold(double d, const V2Fold &f, uint32_t sT) {
    const double tt = fma(f.b1, d, f.mb);
    const int ki = __double2loint(tt);
    const double kd = tt - f.mb;
    const double r1 = fma(f.b1, d, -kd);
    double qq = fma(r1, f.p2, f.p1);
    qq = fma(qq, r1, f.p0);
    double tj;
    asm("ld.shared.f64 %0, [%1];" : "=d"(tj) : "r"(sT + (uint32_t)((ki << 3) & 0x1FF8)));
    int thi;
    asm("mad.lo.s32 %0, %1, 1024, %2;" : "=r"(thi) : "r"(ki), "r"(__double2hiint(tj)));
    return fma(__hiloint2double(thi, __double2loint(tj)), qq, 1.0);
}

// records [m_begin, m_end) of the staged chunk, this thread's slice.  Two records x two attributes are in flight per
// thread: four independent exp -> reciprocal chains (eight at q = 0, where two record pairs share one straight-line block:
// predicated row reloads instead of branches keep it one basic block, which is what lets the scheduler interleave them).
// The running products of u (deviance) are renormalised once per V2_RENORM records: without clamps |eta| < 30, so
// u < 2^43.3 and 16 factors times a mantissa stay below 2^694.
#ifndef NB_EXP_FOLD
#define NB_EXP_FOLD 0
#endif
#ifndef NB_EXP_HOIST
#define NB_EXP_HOIST 1
#endif
constexpr int V2_RENORM = NB_EXP_HOIST ? 16 : 4;
template <int Q, int DIFF, int NSL, bool CHECK>
__device__ __forceinline__ void v2_range(int m_begin, int m_end, int slice, uint32_t sREC_addr, uint32_t sCD_addr,
                                         uint32_t sXI_addr, uint32_t sXJ_addr, uint32_t sT, int diff_type,
                                         const double (&beta0)[2 + Q], const double (&beta1)[2 + Q], const V2Fold &f0, const V2Fold &f1,
                                         double (&acc0)[(2 + Q) * (3 + Q) / 2 + (2 + Q) + 2],
                                         double (&acc1)[(2 + Q) * (3 + Q) / 2 + (2 + Q) + 2], unsigned &esum0, unsigned &esum1, int &nren) {
    constexpr int NA = (2 + Q) * (3 + Q) / 2 + (2 + Q) + 2;
    constexpr bool FOLD = (Q == 0 && !CHECK && NB_EXP_FOLD);
    // each slice walks a CONTIGUOUS part of the range: records are in list order (grouped by Ri), so the I row
    // repeats for several records in a row and stays in registers (halves the shared-memory tile traffic).
    // m_begin is even and slices are an even number of records long: record pairs are fetched with one 8-byte load.
    const int len = m_end - m_begin, per = (((len + NSL - 1) / NSL) + 1) & ~1;
    const int m_lo = m_begin + slice * per, m_hi = (m_lo + per < m_end) ? (m_lo + per) : m_end;
    constexpr int UNR = (Q == 0) ? 2 : 1;
    uint32_t io_prev = 0xFFFFFFFFu;
    double xi0 = 0.0, xi1 = 0.0;                                 // cached I row (two attributes)
    int m = m_lo;
    auto two_records = [&](int m) {
        uint32_t ra, rb;                                         // (il << 10) | (jl << 26) | y: byte offsets of the two rows
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ra), "=r"(rb) : "r"(sREC_addr + 4u * m));
        const uint32_t ioa = ra & 0xFC00u, iob = rb & 0xFC00u;
        double ja0, ja1, jb0, jb1;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(ja0), "=d"(ja1) : "r"(sXJ_addr + (ra >> 16)));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(jb0), "=d"(jb1) : "r"(sXJ_addr + (rb >> 16)));
        if (ioa != io_prev) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(xi0), "=d"(xi1) : "r"(sXI_addr + ioa));
        const double da0 = proj_diff<DIFF>(xi0, ja0, diff_type), da1 = proj_diff<DIFF>(xi1, ja1, diff_type);
        if (iob != ioa) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(xi0), "=d"(xi1) : "r"(sXI_addr + iob));
        const double db0 = proj_diff<DIFF>(xi0, jb0, diff_type), db1 = proj_diff<DIFF>(xi1, jb1, diff_type);
        io_prev = iob;
        double ca[Q > 0 ? Q : 1], cb[Q > 0 ? Q : 1];
#pragma unroll
        for (int u = 0; u < Q; ++u) {
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(ca[u]) : "r"(sCD_addr + 8u * (u * V2_RCHUNK + m)));
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(cb[u]) : "r"(sCD_addr + 8u * (u * V2_RCHUNK + m) + 8u));
        }
        double ta0, ta1, tb0, tb1;
        if (FOLD) {                                              // q = 0, no clamps possible: folded intercept, 17 FP64 per evaluation
            ta0 = v2_u1_fold(da0, f0, sT); ta1 = v2_u1_fold(da1, f1, sT);
            tb0 = v2_u1_fold(db0, f0, sT); tb1 = v2_u1_fold(db1, f1, sT);
            v2_tail<Q>(da0, ca, ta0, acc0);
            v2_tail<Q>(da1, ca, ta1, acc1);
            v2_tail<Q>(db0, cb, tb0, acc0);
            v2_tail<Q>(db1, cb, tb1, acc1);
            return;
        }
        double ea0 = fma(beta0[1], da0, beta0[0]), ea1 = fma(beta1[1], da1, beta1[0]);
        double eb0 = fma(beta0[1], db0, beta0[0]), eb1 = fma(beta1[1], db1, beta1[0]);
#pragma unroll
        for (int u = 0; u < Q; ++u) {
            ea0 = fma(beta0[2 + u], ca[u], ea0); ea1 = fma(beta1[2 + u], ca[u], ea1);
            eb0 = fma(beta0[2 + u], cb[u], eb0); eb1 = fma(beta1[2 + u], cb[u], eb1);
        }
        bool fast = true;
        if (CHECK) {                                             // !CHECK: the CTA proved |eta| < 30 for all its records up front
            const int hmax = max(max(__double2hiint(ea0) & 0x7fffffff, __double2hiint(ea1) & 0x7fffffff),
                                 max(__double2hiint(eb0) & 0x7fffffff, __double2hiint(eb1) & 0x7fffffff));
            fast = hmax < 0x40E5A380;                            // |s| < 44316 < 30 * 1024/ln2 (= 44319.6) for all four
        }
        if (fast) {
            ta0 = v2_u1_fast(ea0, sT); ta1 = v2_u1_fast(ea1, sT);
            tb0 = v2_u1_fast(eb0, sT); tb1 = v2_u1_fast(eb1, sT);
        } else {                                                 // rare: R's eta clamps; misses carry a deviance correction
            double2 v;
            v = v2_exp_clamped(ea0, sT); ta0 = v.x; if (ra & 1u) acc0[NA - 1] += v.y;
            v = v2_exp_clamped(ea1, sT); ta1 = v.x; if (ra & 1u) acc1[NA - 1] += v.y;
            v = v2_exp_clamped(eb0, sT); tb0 = v.x; if (rb & 1u) acc0[NA - 1] += v.y;
            v = v2_exp_clamped(eb1, sT); tb1 = v.x; if (rb & 1u) acc1[NA - 1] += v.y;
        }
        v2_tail<Q>(da0, ca, ta0, acc0);
        v2_tail<Q>(da1, ca, ta1, acc1);
        v2_tail<Q>(db0, cb, tb0, acc0);
        v2_tail<Q>(db1, cb, tb1, acc1);
    };
    auto renorm = [&]() {
        v2_renorm(acc0[NA - 2], esum0);
        v2_renorm(acc1[NA - 2], esum1);
        ++nren;
    };
    // one loop body for (almost) all records -- a second copy of it for remainders would run them with half the chains in
    // flight -- with the renormalisation on a counter: every V2_RENORM records, or after every step on the checked path
    // (a clamped factor can be 2^52)
    constexpr int STEP = 2 * UNR;                                // records per loop trip
    constexpr int TRIPS = CHECK ? 1 : (NB_EXP_HOIST ? V2_RENORM / STEP : 1);
    int trips = 0;
#pragma unroll 1
    for (; m + STEP <= m_hi; m += STEP) {
        two_records(m);
        if (UNR == 2) two_records(m + 2);
        if (++trips == TRIPS) { renorm(); trips = 0; }
    }
    if (UNR == 2 && m + 1 < m_hi) { two_records(m); m += 2; }
    if (m < m_hi) {                                              // odd tail: one record
        uint32_t r;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(sREC_addr + 4u * m));
        const uint32_t io = r & 0xFC00u;
        if (io != io_prev) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(xi0), "=d"(xi1) : "r"(sXI_addr + io));
        double xj0, xj1;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(xj0), "=d"(xj1) : "r"(sXJ_addr + (r >> 16)));
        double c[Q > 0 ? Q : 1];
#pragma unroll
        for (int u = 0; u < Q; ++u) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(c[u]) : "r"(sCD_addr + 8u * (u * V2_RCHUNK + m)));
        const double d0 = proj_diff<DIFF>(xi0, xj0, diff_type), d1 = proj_diff<DIFF>(xi1, xj1, diff_type);
        double e0 = fma(beta0[1], d0, beta0[0]), e1 = fma(beta1[1], d1, beta1[0]);
#pragma unroll
        for (int u = 0; u < Q; ++u) { e0 = fma(beta0[2 + u], c[u], e0); e1 = fma(beta1[2 + u], c[u], e1); }
        double2 a, b;
        if (CHECK) { a = v2_exp_clamped(e0, sT); b = v2_exp_clamped(e1, sT); }
        else { a = make_double2(v2_u1_fast(e0, sT), 0.0); b = make_double2(v2_u1_fast(e1, sT), 0.0); }
        if (CHECK && (r & 1u)) { acc0[NA - 1] += a.y; acc1[NA - 1] += b.y; }
        v2_tail<Q>(d0, c, a.x, acc0);
        v2_tail<Q>(d1, c, b.x, acc1);
    }
    renorm();
}

// Moment passes (lm pass / closed-form first IRLS pass).  MODE_LM: S(d), S(d^2), S(d*y), S(d*c_r) with y from the yv
// stream.  MODE_SIGN: the outcome is s = -1 (hit) / +1 (miss); with hits-then-misses ordering it is enough to keep S(d)
// of the hits and of the misses apart: S(d) = hits + misses, S(d*s) = misses - hits (one add + one FMA per evaluation).
template <int MODE, int Q, int DIFF, int Y>
__device__ __forceinline__ void m2_one(uint32_t r, int m, double xj0, double xj1, uint32_t &io_prev, double &xi0, double &xi1,
                                       uint32_t sCD_addr, uint32_t sXI_addr, int diff_type, double (&acc0)[3 + Q], double (&acc1)[3 + Q]) {
    const uint32_t io = r & 0xFC00u;
    if (io != io_prev) {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(xi0), "=d"(xi1) : "r"(sXI_addr + io));
        io_prev = io;
    }
    const double d0 = proj_diff<DIFF>(xi0, xj0, diff_type);
    const double d1 = proj_diff<DIFF>(xi1, xj1, diff_type);
    if (MODE == MODE_LM) {
        double yv;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(yv) : "r"(sCD_addr + 8u * (Q * V2_RCHUNK + m)));
        acc0[0] += d0; acc1[0] += d1;
        acc0[2] = fma(d0, yv, acc0[2]); acc1[2] = fma(d1, yv, acc1[2]);
    } else if (Y == 0) {
        acc0[0] += d0; acc1[0] += d1;
    } else {
        acc0[2] += d0; acc1[2] += d1;
    }
    acc0[1] = fma(d0, d0, acc0[1]); acc1[1] = fma(d1, d1, acc1[1]);
#pragma unroll
    for (int u = 0; u < Q; ++u) {
        double cu;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(cu) : "r"(sCD_addr + 8u * (u * V2_RCHUNK + m)));
        acc0[3 + u] = fma(d0, cu, acc0[3 + u]); acc1[3 + u] = fma(d1, cu, acc1[3 + u]);
    }
}

// The loop is latency bound (record -> row address -> row -> 3 FP64 ops), so four records are fetched with one 16-byte
// load and their J rows are requested together before any of them is consumed.
template <int MODE, int Q, int DIFF, int Y, int NSL>
__device__ __forceinline__ void m2_range(int m_begin, int m_end, int slice, uint32_t sREC_addr, uint32_t sCD_addr,
                                         uint32_t sXI_addr, uint32_t sXJ_addr, int diff_type,
                                         double (&acc0)[3 + Q], double (&acc1)[3 + Q]) {
    const int len = m_end - m_begin, per = (len + NSL - 1) / NSL;
    const int m_lo = m_begin + slice * per, m_hi = (m_lo + per < m_end) ? (m_lo + per) : m_end;
    uint32_t io_prev = 0xFFFFFFFFu;
    double xi0 = 0.0, xi1 = 0.0;
    int m = m_lo;
    for (; m < m_hi && (m & 3); ++m) {                           // peel to a 16-byte boundary of the record buffer
        uint32_t r;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(sREC_addr + 4u * m));
        double xj0, xj1;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(xj0), "=d"(xj1) : "r"(sXJ_addr + (r >> 16)));
        m2_one<MODE, Q, DIFF, Y>(r, m, xj0, xj1, io_prev, xi0, xi1, sCD_addr, sXI_addr, diff_type, acc0, acc1);
    }
    for (; m + 3 < m_hi; m += 4) {
        uint32_t r[4];
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(sREC_addr + 4u * m));
        double xj[4][2];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(xj[k][0]), "=d"(xj[k][1]) : "r"(sXJ_addr + (r[k] >> 16)));
#pragma unroll
        for (int k = 0; k < 4; ++k)
            m2_one<MODE, Q, DIFF, Y>(r[k], m + k, xj[k][0], xj[k][1], io_prev, xi0, xi1, sCD_addr, sXI_addr, diff_type, acc0, acc1);
    }
    for (; m < m_hi; ++m) {
        uint32_t r;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(sREC_addr + 4u * m));
        double xj0, xj1;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(xj0), "=d"(xj1) : "r"(sXJ_addr + (r >> 16)));
        m2_one<MODE, Q, DIFF, Y>(r, m, xj0, xj1, io_prev, xi0, xi1, sCD_addr, sXI_addr, diff_type, acc0, acc1);
    }
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// One skeleton for the three v2 passes.  Shared memory: I tile [64][128], two J tiles (double buffered), two record
// chunk buffers (records + covariate diffs (+ yv for lm)), the exp/log tables.  While tile t is being consumed the
// J tile and the first record chunk of tile t+1 are in flight (cp.async), so the FP64 pipe does not idle on staging.
template <int MODE, int Q>
struct V2Layout {
    static constexpr int NCD = (MODE == MODE_LM) ? (Q + 1) : (Q > 0 ? Q : 0);       // double streams per record
    static constexpr int NA = (MODE != MODE_IRLS_FULL) ? (3 + Q) : ((2 + Q) * (3 + Q) / 2 + (2 + Q) + 2);
    static constexpr uint32_t XI = 0;
    static constexpr uint32_t XJ = BI * TA * 8;                                      // two buffers of BJ*TA*8
    static constexpr uint32_t REC = XJ + 2 * BJ * TA * 8;                            // two buffers
    static constexpr uint32_t REC_BYTES = V2_RCHUNK * 4;
    static constexpr uint32_t CD = REC + 2 * REC_BYTES;                              // two buffers
    static constexpr uint32_t CD_BYTES = (NCD > 0 ? NCD : 1) * V2_RCHUNK * 8;
    static constexpr uint32_t TAB = CD + 2 * CD_BYTES;
    static constexpr uint32_t TOTAL = TAB + (MODE == MODE_IRLS_FULL ? sizeof(nb_tables2) : 0);
    static constexpr int NSL = v2_nslice(MODE, Q);
    static constexpr int THREADS = V2_TPS * NSL;
    static constexpr uint32_t RED = NSL * NA * TA * 8;
    static constexpr uint32_t SMEM = TOTAL > RED ? TOTAL : RED;
};

template <int MODE, int Q, int DIFF>
__global__ void __launch_bounds__(V2Layout<MODE, Q>::THREADS, 1)
pass2_kernel(PassArgs P) {
    using L = V2Layout<MODE, Q>;
    constexpr int K = 2 + Q;
    constexpr int NA = L::NA;
    constexpr int NSL = L::NSL;
    constexpr int V2_THREADS = L::THREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
    asm volatile("" : "+r"(sbase));                      // opaque: keep the window base in a register, not rematerialised

    const int tid = threadIdx.x;
    const int tl = tid % V2_TPS, slice = tid / V2_TPS;
    // chunk index is the fast grid dimension: the CTAs resident at any moment share a few attribute tiles, whose XT slabs
    // (N x 128 doubles each) then stay in L2 while every chunk of the pair stream sweeps over them
    const int at = blockIdx.y;
    int sidx = 0;
#pragma unroll
    for (int t = 1; t < NB_MAX_STREAMS; ++t) sidx += ((int)blockIdx.x >= P.chunk_start[t]) ? 1 : 0;
    const int chunk = (int)blockIdx.x - P.chunk_start[sidx];
    const StreamArgs &A = P.st[sidx];
    if (MODE == MODE_IRLS_FULL && P.tile_active && !P.tile_active[at]) return;
    const int64_t a0 = (int64_t)at * TA;

    if (MODE == MODE_IRLS_FULL) {   // table image
        nb_tables2 *sT = (nb_tables2 *)(smem_raw + L::TAB);
        const nb_tables1024 *g = P.tables1024;
        for (int e = tid; e < 1024; e += V2_THREADS) {         // high word pre-biased by -j * 1024 (see v2_u1_fast)
            const double tj = g->exp_t[e];
            sT->exp_t[e] = __hiloint2double(__double2hiint(tj) - e * 1024, __double2loint(tj));
        }
    }
    const uint32_t sT_addr = sbase + L::TAB;
    const uint32_t sXI_addr = sbase + L::XI + (uint32_t)tl * 16u;

    double beta0[K], beta1[K];
    if (MODE == MODE_IRLS_FULL) {
#pragma unroll
        for (int r = 0; r < K; ++r) {
            beta0[r] = A.beta[(int64_t)r * P.nattr_pad + a0 + 2 * tl] * c_k[9];         // units of ln2/1024
            beta1[r] = A.beta[(int64_t)r * P.nattr_pad + a0 + 2 * tl + 1] * c_k[9];
        }
    }
    V2Fold f0 = {0, 0, 0, 0, 0}, f1 = {0, 0, 0, 0, 0};
    if (MODE == MODE_IRLS_FULL && Q == 0) { f0 = v2_fold(beta0[0], beta0[1]); f1 = v2_fold(beta1[0], beta1[1]); }
    double acc0[NA], acc1[NA];
#pragma unroll
    for (int n = 0; n < NA; ++n) { acc0[n] = 0.0; acc1[n] = 0.0; }
    unsigned esum0 = 0, esum1 = 0; int nren = 0;   // wrapping sums of the biased exponents taken out of the running products
                                          // of u1 = 1 + exp(eta); the true sum (<= 53 per record, < 2^25 records per thread) fits
    if (MODE == MODE_IRLS_FULL) { acc0[NA - 2] = 1.0; acc1[NA - 2] = 1.0; }
    bool may_clamp = false;
    if (MODE == MODE_IRLS_FULL) {         // can any record of this attribute tile reach R's |eta| > 30 clamps?
        double b0 = fabs(beta0[0]) + fabs(beta0[1]) * P.dmax[a0 + 2 * tl];
        double b1 = fabs(beta1[0]) + fabs(beta1[1]) * P.dmax[a0 + 2 * tl + 1];
#pragma unroll
        for (int u = 0; u < Q; ++u) { b0 += fabs(beta0[2 + u]) * P.cmax[u]; b1 += fabs(beta1[2 + u]) * P.cmax[u]; }
        const double lim = c_k[8] * 0.999;
        may_clamp = __syncthreads_or(!(b0 < lim) || !(b1 < lim)) != 0;    // NaN coefficients take the checked path
    }

    const int t_begin = A.chunk_tile[chunk], t_end = A.chunk_tile[chunk + 1];

    // async stage of a [64][128] double tile (rows are TA contiguous doubles of XT, zero padded)
    auto stage_tile = [&](uint32_t dst, int block) {
        const double *g = P.XT + ((int64_t)block * 64) * P.ldt + a0;
        for (int e = tid; e < 64 * TA / 2; e += V2_THREADS) {
            const int row = e / (TA / 2), c2 = e % (TA / 2);
            cp_async16(dst + (uint32_t)e * 16u, g + (int64_t)row * P.ldt + 2 * c2);
        }
    };
    // async stage of records [rc, rc + n) (+ their covariate diffs / yv)
    auto stage_records = [&](int buf, int64_t rc, int n) {
        const uint32_t rdst = sbase + L::REC + buf * L::REC_BYTES, cdst = sbase + L::CD + buf * L::CD_BYTES;
        for (int e = tid; e < n; e += V2_THREADS) {
            cp_async4(rdst + 4u * e, A.rec + rc + e);
#pragma unroll
            for (int c = 0; c < Q; ++c) cp_async8(cdst + 8u * (c * V2_RCHUNK + e), A.cd[c] + rc + e);
            if (MODE == MODE_LM) cp_async8(cdst + 8u * (Q * V2_RCHUNK + e), A.yv + rc + e);
        }
    };
    auto first_chunk_len = [&](int t) -> int {
        const int64_t len = A.tile_ptr[t + 1] - A.tile_ptr[t];
        return (int)(len < V2_RCHUNK ? len : V2_RCHUNK);
    };

    int cur_ib = -1;
    if (t_begin < t_end) {                                 // prologue: tile t_begin in flight
        stage_tile(sbase + L::XJ, A.tile_jb[t_begin]);
        stage_records(0, A.tile_ptr[t_begin], first_chunk_len(t_begin));
        cp_async_commit();
    }
    for (int t = t_begin; t < t_end; ++t) {
        const int buf = (t - t_begin) & 1;
        const int ib = A.tile_ib[t];
        const int64_t r0 = A.tile_ptr[t], r1 = A.tile_ptr[t + 1], rmid = A.tile_mid[t];
        if (ib != cur_ib) {                                // new I block (once per row of tiles): nobody reads XI now? ->
            __syncthreads();                               // all threads are done with the previous tile's compute
            stage_tile(sbase + L::XI, ib);
            cp_async_commit();
            cur_ib = ib;
        }
        cp_async_wait_all();
        __syncthreads();                                   // tile t staged; previous tile's buffers are free
        if (t + 1 < t_end) {                               // prefetch tile t+1 into the other buffers
            stage_tile(sbase + L::XJ + (buf ^ 1) * (BJ * TA * 8), A.tile_jb[t + 1]);
            stage_records(buf ^ 1, A.tile_ptr[t + 1], first_chunk_len(t + 1));
            cp_async_commit();
        }
        const uint32_t sXJ_addr = sbase + L::XJ + buf * (BJ * TA * 8) + (uint32_t)tl * 16u;
        const uint32_t sREC_addr = sbase + L::REC + buf * L::REC_BYTES;
        const uint32_t sCD_addr = sbase + L::CD + buf * L::CD_BYTES;
        for (int64_t rc = r0; rc < r1; rc += V2_RCHUNK) {
            const int n = (int)((r1 - rc < V2_RCHUNK) ? (r1 - rc) : V2_RCHUNK);
            if (rc != r0) {                                // rare: more than one record chunk in this tile
                __syncthreads();                           // previous chunk consumed
                stage_records(buf, rc, n);
                cp_async_commit();
                cp_async_wait_all();                       // (also completes the prefetch; correctness only)
                __syncthreads();
            }
            int ms = (int)(rmid - rc);                      // first miss record of this chunk
            ms = ms < 0 ? 0 : (ms > n ? n : ms);
            if constexpr (MODE == MODE_IRLS_FULL) {
                if (may_clamp) v2_range<Q, DIFF, NSL, true>(0, n, slice, sREC_addr, sCD_addr, sXI_addr, sXJ_addr, sT_addr, P.diff_type, beta0, beta1, f0, f1, acc0, acc1, esum0, esum1, nren);
                else v2_range<Q, DIFF, NSL, false>(0, n, slice, sREC_addr, sCD_addr, sXI_addr, sXJ_addr, sT_addr, P.diff_type, beta0, beta1, f0, f1, acc0, acc1, esum0, esum1, nren);
            } else if constexpr (MODE == MODE_LM) {
                m2_range<MODE_LM, Q, DIFF, 0, NSL>(0, n, slice, sREC_addr, sCD_addr, sXI_addr, sXJ_addr, P.diff_type, acc0, acc1);
            } else {
                m2_range<MODE_SIGN, Q, DIFF, 0, NSL>(0, ms, slice, sREC_addr, sCD_addr, sXI_addr, sXJ_addr, P.diff_type, acc0, acc1);
                m2_range<MODE_SIGN, Q, DIFF, 1, NSL>(ms, n, slice, sREC_addr, sCD_addr, sXI_addr, sXJ_addr, P.diff_type, acc0, acc1);
            }
        }
    }
    if (MODE == MODE_IRLS_FULL) {                          // integer exponent sums -> log(u1) total
        acc0[NA - 2] = fma((double)(int)(esum0 - 1023u * (unsigned)nren), c_k[4], log(acc0[NA - 2]));
        acc1[NA - 2] = fma((double)(int)(esum1 - 1023u * (unsigned)nren), c_k[4], log(acc1[NA - 2]));
    }
    if (MODE == MODE_SIGN) {                               // (hits, misses) -> (S d, S d*s)
        const double h0 = acc0[0], m0 = acc0[2], h1 = acc1[0], m1 = acc1[2];
        acc0[0] = h0 + m0; acc0[2] = m0 - h0; acc1[0] = h1 + m1; acc1[2] = m1 - h1;
    }
    cp_async_wait_all();
    __syncthreads();
    double *red = (double *)smem_raw;                      // [NSL][NA][TA]
#pragma unroll
    for (int n = 0; n < NA; ++n) {
        red[(slice * NA + n) * TA + 2 * tl] = acc0[n];
        red[(slice * NA + n) * TA + 2 * tl + 1] = acc1[n];
    }
    __syncthreads();
    for (int e = tid; e < NA * TA; e += V2_THREADS) {
        const int n = e / TA, a = e % TA;
        double sacc = red[(0 * NA + n) * TA + a];
#pragma unroll
        for (int sl = 1; sl < NSL; ++sl) sacc += red[(sl * NA + n) * TA + a];
        A.partials[((int64_t)chunk * NA + n) * P.nattr_pad + a0 + a] = sacc;
    }
}

// ------------------------------------------------------------------ pass kernels, v3 (q <= 2): FOUR attributes per thread
// The issue-slot micro-benchmark (scripts/dev/ubench_fp64.cu) confirms the model of DESIGN.md 4: an FP64 instruction takes two
// issue cycles and integer / LDS / MUFU instructions do NOT issue underneath it, so the only way to raise the FP64 pipe's
// utilisation is to execute fewer non-FP64 instructions per evaluation.  v3 does that structurally:
//   * a thread owns FOUR attributes -- lane l of a warp the attributes {2l, 2l+1, 64+2l, 64+2l+1} of the 128-attribute tile, so
//     a row is fetched with two conflict-free LDS.128 -- and everything that is per RECORD (record word, row addresses, the
//     I-row change test, covariate fetches, loop control) is paid once per four evaluations instead of once per two;
//   * one warp = one pair slice (32 lanes x 4 attributes = the tile): record handling is warp-uniform, no divergence;
//   * the I-row change test of two records is one LOP3 + one branch (rows change every ~16-25 records);
//   * the running products of u = 1 + exp(eta) (deviance) are renormalised once per 16 records instead of every 2-4:
//     without clamps |eta| < 30, so u < 2^43.3 and 16 factors times a mantissa stay below 2^694.
// Per evaluation at q = 0: 18 FP64 + ~7.6 other instructions (v2: 18 + 10.6).
constexpr int V3_APT = 4;                  // attributes per thread
// pair slices = warps per CTA, as many as the register file allows: the q = 0 IRLS pass (two records x four attributes in
// flight) fits 168 registers -> 12 warps; with covariates ~250 registers -> 8 warps; the moment passes 16 (12 at q = 2)
__host__ __device__ constexpr int v3_nslice(int mode, int q) {
    return mode == MODE_IRLS_FULL ? (q == 0 ? 12 : 8) : (q <= 1 ? 16 : 12);
}
constexpr int V3_RENORM = 16;              // records between renormalisations of the running products

template <int MODE, int Q>
struct V3Layout {
    static constexpr int NCD = (MODE == MODE_LM) ? (Q + 1) : (Q > 0 ? Q : 0);       // double streams per record
    static constexpr int NA = (MODE != MODE_IRLS_FULL) ? (3 + Q) : ((2 + Q) * (3 + Q) / 2 + (2 + Q) + 2);
    static constexpr uint32_t XI = 0;
    static constexpr uint32_t XJ = BI * TA * 8;                                      // two buffers of BJ*TA*8
    static constexpr uint32_t REC = XJ + 2 * BJ * TA * 8;                            // two buffers
    static constexpr uint32_t REC_BYTES = V2_RCHUNK * 4;
    static constexpr uint32_t CD = REC + 2 * REC_BYTES;                              // two buffers
    static constexpr uint32_t CD_BYTES = (NCD > 0 ? NCD : 1) * V2_RCHUNK * 8;
    static constexpr uint32_t TAB = CD + 2 * CD_BYTES;
    static constexpr uint32_t TOTAL = TAB + (MODE == MODE_IRLS_FULL ? sizeof(nb_tables2) : 0);
    static constexpr int NSL = v3_nslice(MODE, Q);
    static constexpr int THREADS = 32 * NSL;
    static constexpr uint32_t RED = NSL * NA * TA * 8;
    static constexpr uint32_t SMEM = TOTAL > RED ? TOTAL : RED;
};

// row of four attributes: {2l, 2l+1} at row + 16 l, {64+2l, 64+2l+1} at row + 512 + 16 l (addr already includes 16 l)
__device__ __forceinline__ void v3_ld_row(uint32_t addr, double (&x)[4]) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x[0]), "=d"(x[1]) : "r"(addr));
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+512];" : "=d"(x[2]), "=d"(x[3]) : "r"(addr));
}

// IRLS pass over records [m_lo, m_hi) of the staged chunk (this warp's slice).  NR records are in flight per step
// (2 at q = 0, 1 with covariates: register budget); CHECK = the work item may reach R's |eta| > 30 clamps.
template <int Q, int DIFF, bool CHECK>
__device__ __forceinline__ void v3_irls_range(int m_lo, int m_hi, uint32_t sREC_addr, uint32_t sCD_addr, uint32_t sXI_addr,
                                              uint32_t sXJ_addr, uint32_t sT, int diff_type, const double (&beta)[V3_APT][2 + Q],
                                              double (&acc)[V3_APT][(2 + Q) * (3 + Q) / 2 + (2 + Q) + 2], unsigned (&esum)[V3_APT],
                                              int &nren, int dbg = 0) {
    constexpr int NA = (2 + Q) * (3 + Q) / 2 + (2 + Q) + 2;
    constexpr int NR = (Q == 0 && !CHECK) ? 2 : 1;
    if (m_lo >= m_hi) return;
    uint32_t prev;                                               // previous record word (its I-row field is what matters)
    double xi[V3_APT];
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(prev) : "r"(sREC_addr + 4u * m_lo));
    v3_ld_row(sXI_addr + (prev & 0xFC00u), xi);                  // the range starts on the I row of its first record
    auto renorm_all = [&]() {
#pragma unroll
        for (int k = 0; k < V3_APT; ++k) v2_renorm(acc[k][NA - 2], esum[k]);
        ++nren;
    };
    auto eta_of = [&](int k, double d, const double (&c)[Q > 0 ? Q : 1]) {
        double e = fma(beta[k][1], d, beta[k][0]);
#pragma unroll
        for (int u = 0; u < Q; ++u) e = fma(beta[k][2 + u], c[u], e);
        return e;
    };
    auto one_record = [&](int m) {                               // NR == 1 form (covariates, checked path, odd tails)
        uint32_t r;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(sREC_addr + 4u * m));
        double xj[V3_APT];
        v3_ld_row(sXJ_addr + (r >> 16), xj);
        if ((r ^ prev) & 0xFC00u) v3_ld_row(sXI_addr + (r & 0xFC00u), xi);
        prev = r;
        double c[Q > 0 ? Q : 1];
#pragma unroll
        for (int u = 0; u < Q; ++u) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(c[u]) : "r"(sCD_addr + 8u * (u * V2_RCHUNK + m)));
        double d[V3_APT], u1[V3_APT];
#pragma unroll
        for (int k = 0; k < V3_APT; ++k) d[k] = proj_diff<DIFF>(xi[k], xj[k], diff_type);
#pragma unroll
        for (int k = 0; k < V3_APT; ++k) {
            const double e = eta_of(k, d[k], c);
            if (CHECK) {
                const double2 v = v2_exp_clamped(e, sT);
                u1[k] = v.x;
                if (r & 1u) acc[k][NA - 1] += v.y;
            } else {
                u1[k] = v2_u1_fast(e, sT);
            }
        }
#pragma unroll
        for (int k = 0; k < V3_APT; ++k) v2_tail<Q>(d[k], c, u1[k], acc[k]);
    };
    auto two_records = [&](int m) {                              // q = 0, unchecked: eight exp -> reciprocal chains in flight
        uint32_t ra, rb;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ra), "=r"(rb) : "r"(sREC_addr + 4u * m));
        double ja[V3_APT], jb[V3_APT], da[V3_APT], db[V3_APT];
        v3_ld_row(sXJ_addr + (ra >> 16), ja);
        v3_ld_row(sXJ_addr + (rb >> 16), jb);
        if (((ra ^ prev) | (rb ^ prev)) & 0xFC00u) {             // an I row changes here (every ~16-25 records)
            if ((ra ^ rb) & 0xFC00u) { one_record(m); one_record(m + 1); return; }   // between the two records: one at a time
            v3_ld_row(sXI_addr + (ra & 0xFC00u), xi);
        }
#pragma unroll
        for (int k = 0; k < V3_APT; ++k) { da[k] = proj_diff<DIFF>(xi[k], ja[k], diff_type); db[k] = proj_diff<DIFF>(xi[k], jb[k], diff_type); }
        prev = rb;
        const double c0[1] = {0.0};
        double ua[V3_APT], ub[V3_APT];
#pragma unroll
        for (int k = 0; k < V3_APT; ++k) { ua[k] = v2_u1_fast(fma(beta[k][1], da[k], beta[k][0]), sT); ub[k] = v2_u1_fast(fma(beta[k][1], db[k], beta[k][0]), sT); }
#pragma unroll
        for (int k = 0; k < V3_APT; ++k) { v2_tail<0>(da[k], c0, ua[k], reinterpret_cast<double (&)[7]>(acc[k])); v2_tail<0>(db[k], c0, ub[k], reinterpret_cast<double (&)[7]>(acc[k])); }
    };
    int m = m_lo;
    if (CHECK) {                                                 // rare: a factor can be 2^52, renormalise after every record
        for (; m < m_hi; ++m) { one_record(m); renorm_all(); }
        return;
    }
    if (NR == 2 && (dbg & 1)) {
        for (; m < m_hi; ++m) { one_record(m); if (((m - m_lo) & 15) == 15) renorm_all(); }
        renorm_all();
        return;
    }
    if (NR == 2) {
        if ((m & 1) && m < m_hi) { one_record(m); ++m; }         // 8-byte alignment of the record pairs
        while (m + V3_RENORM <= m_hi) {
#pragma unroll 2
            for (int i = 0; i < V3_RENORM; i += 2) two_records(m + i);
            m += V3_RENORM;
            renorm_all();
        }
        for (; m + 1 < m_hi; m += 2) two_records(m);
        if (m < m_hi) { one_record(m); ++m; }
        renorm_all();
    } else {
        while (m + V3_RENORM <= m_hi) {
#pragma unroll 2
            for (int i = 0; i < V3_RENORM; ++i) one_record(m + i);
            m += V3_RENORM;
            renorm_all();
        }
        for (; m < m_hi; ++m) one_record(m);
        renorm_all();
    }
}

// moment passes, four attributes per thread, four records requested together (the loop is bound by the shared-memory pipe)
template <int MODE, int Q, int DIFF, int Y>
__device__ __forceinline__ void v3_moment_range(int m_lo, int m_hi, uint32_t sREC_addr, uint32_t sCD_addr, uint32_t sXI_addr,
                                                uint32_t sXJ_addr, int diff_type, double (&acc)[V3_APT][3 + Q]) {
    if (m_lo >= m_hi) return;
    uint32_t prev;
    double xi[V3_APT];
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(prev) : "r"(sREC_addr + 4u * m_lo));
    v3_ld_row(sXI_addr + (prev & 0xFC00u), xi);                  // the range starts on the I row of its first record
    auto consume = [&](uint32_t r, int m, const double (&xj)[V3_APT]) {
        if ((r ^ prev) & 0xFC00u) v3_ld_row(sXI_addr + (r & 0xFC00u), xi);
        prev = r;
        double d[V3_APT];
#pragma unroll
        for (int k = 0; k < V3_APT; ++k) d[k] = proj_diff<DIFF>(xi[k], xj[k], diff_type);
        if (MODE == MODE_LM) {
            double yv;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(yv) : "r"(sCD_addr + 8u * (Q * V2_RCHUNK + m)));
#pragma unroll
            for (int k = 0; k < V3_APT; ++k) { acc[k][0] += d[k]; acc[k][2] = fma(d[k], yv, acc[k][2]); }
        } else {
#pragma unroll
            for (int k = 0; k < V3_APT; ++k) acc[k][Y == 0 ? 0 : 2] += d[k];
        }
#pragma unroll
        for (int k = 0; k < V3_APT; ++k) acc[k][1] = fma(d[k], d[k], acc[k][1]);
#pragma unroll
        for (int u = 0; u < Q; ++u) {
            double cu;
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(cu) : "r"(sCD_addr + 8u * (u * V2_RCHUNK + m)));
#pragma unroll
            for (int k = 0; k < V3_APT; ++k) acc[k][3 + u] = fma(d[k], cu, acc[k][3 + u]);
        }
    };
    auto single = [&](int m) {
        uint32_t r;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(sREC_addr + 4u * m));
        double xj[V3_APT];
        v3_ld_row(sXJ_addr + (r >> 16), xj);
        consume(r, m, xj);
    };
    int m = m_lo;
    for (; m < m_hi && (m & 3); ++m) single(m);                  // peel to a 16-byte boundary of the record buffer
    for (; m + 3 < m_hi; m += 4) {
        uint32_t r[4];
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(sREC_addr + 4u * m));
        double xj[4][V3_APT];
#pragma unroll
        for (int k = 0; k < 4; ++k) v3_ld_row(sXJ_addr + (r[k] >> 16), xj[k]);
#pragma unroll
        for (int k = 0; k < 4; ++k) consume(r[k], m + k, xj[k]);
    }
    for (; m < m_hi; ++m) single(m);
}

template <int MODE, int Q, int DIFF>
__global__ void __launch_bounds__(V3Layout<MODE, Q>::THREADS, 1)
pass3_kernel(PassArgs P) {
    using L = V3Layout<MODE, Q>;
    constexpr int K = 2 + Q;
    constexpr int NA = L::NA;
    constexpr int V3_NSL = L::NSL;
    constexpr int V3_THREADS = L::THREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
    asm volatile("" : "+r"(sbase));                      // opaque: keep the window base in a register, not rematerialised

    const int tid = threadIdx.x;
    const int lane = tid & 31, slice = tid >> 5;
    // chunk index is the fast grid dimension: the CTAs resident at any moment share a few attribute tiles, whose XT slabs
    // (N x 128 doubles each) then stay in L2 while every chunk of the pair stream sweeps over them
    const int at = blockIdx.y;
    int sidx = 0;
#pragma unroll
    for (int t = 1; t < NB_MAX_STREAMS; ++t) sidx += ((int)blockIdx.x >= P.chunk_start[t]) ? 1 : 0;
    const int chunk = (int)blockIdx.x - P.chunk_start[sidx];
    const StreamArgs &A = P.st[sidx];
    if (MODE == MODE_IRLS_FULL && P.tile_active && !P.tile_active[at]) return;
    const int64_t a0 = (int64_t)at * TA;
    // local attribute index of this thread's k-th attribute
    auto attr_of = [&](int k) { return (k >> 1) * 64 + 2 * lane + (k & 1); };

    if (MODE == MODE_IRLS_FULL) {   // table image
        nb_tables2 *sT = (nb_tables2 *)(smem_raw + L::TAB);
        const nb_tables1024 *g = P.tables1024;
        for (int e = tid; e < 1024; e += V3_THREADS) {         // high word pre-biased by -j * 1024 (see v2_u1_fast)
            const double tj = g->exp_t[e];
            sT->exp_t[e] = __hiloint2double(__double2hiint(tj) - e * 1024, __double2loint(tj));
        }
    }
    const uint32_t sT_addr = sbase + L::TAB;
    const uint32_t sXI_addr = sbase + L::XI + (uint32_t)lane * 16u;

    double beta[V3_APT][K];
    if (MODE == MODE_IRLS_FULL) {
#pragma unroll
        for (int k = 0; k < V3_APT; ++k)
#pragma unroll
            for (int r = 0; r < K; ++r) beta[k][r] = A.beta[(int64_t)r * P.nattr_pad + a0 + attr_of(k)] * c_k[9];      // units of ln2/1024
    }
    double acc[V3_APT][NA];
#pragma unroll
    for (int k = 0; k < V3_APT; ++k)
#pragma unroll
        for (int n = 0; n < NA; ++n) acc[k][n] = 0.0;
    unsigned esum[V3_APT] = {0, 0, 0, 0};
    int nren = 0;        // wrapping sums of the biased exponents taken out of the running products of u1 = 1 + exp(eta)
    if (MODE == MODE_IRLS_FULL) {
#pragma unroll
        for (int k = 0; k < V3_APT; ++k) acc[k][NA - 2] = 1.0;
    }
    bool may_clamp = false;
    if (MODE == MODE_IRLS_FULL) {         // can any record of this attribute tile reach R's |eta| > 30 clamps?
        bool over = false;
        const double lim = c_k[8] * 0.999;
#pragma unroll
        for (int k = 0; k < V3_APT; ++k) {
            double b = fabs(beta[k][0]) + fabs(beta[k][1]) * P.dmax[a0 + attr_of(k)];
#pragma unroll
            for (int u = 0; u < Q; ++u) b += fabs(beta[k][2 + u]) * P.cmax[u];
            over = over || !(b < lim);                           // NaN coefficients take the checked path
        }
        may_clamp = __syncthreads_or(over) != 0;
    }

    const int t_begin = A.chunk_tile[chunk], t_end = A.chunk_tile[chunk + 1];

    // async stage of a [64][128] double tile (rows are TA contiguous doubles of XT, zero padded)
    auto stage_tile = [&](uint32_t dst, int block) {
        const double *g = P.XT + ((int64_t)block * 64) * P.ldt + a0;
        for (int e = tid; e < 64 * TA / 2; e += V3_THREADS) {
            const int row = e / (TA / 2), c2 = e % (TA / 2);
            cp_async16(dst + (uint32_t)e * 16u, g + (int64_t)row * P.ldt + 2 * c2);
        }
    };
    // async stage of records [rc, rc + n) (+ their covariate diffs / yv)
    auto stage_records = [&](int buf, int64_t rc, int n) {
        const uint32_t rdst = sbase + L::REC + buf * L::REC_BYTES, cdst = sbase + L::CD + buf * L::CD_BYTES;
        for (int e = tid; e < n; e += V3_THREADS) {
            cp_async4(rdst + 4u * e, A.rec + rc + e);
#pragma unroll
            for (int c = 0; c < Q; ++c) cp_async8(cdst + 8u * (c * V2_RCHUNK + e), A.cd[c] + rc + e);
            if (MODE == MODE_LM) cp_async8(cdst + 8u * (Q * V2_RCHUNK + e), A.yv + rc + e);
        }
    };
    auto first_chunk_len = [&](int t) -> int {
        const int64_t len = A.tile_ptr[t + 1] - A.tile_ptr[t];
        return (int)(len < V2_RCHUNK ? len : V2_RCHUNK);
    };
    // this warp's contiguous part of records [lo, hi): even length, so record pairs stay 8-byte aligned
    auto slice_of = [&](int lo, int hi, int &s_lo, int &s_hi) {
        const int len = hi - lo, per = (((len + V3_NSL - 1) / V3_NSL) + 3) & ~3;
        s_lo = lo + slice * per;
        s_hi = (s_lo + per < hi) ? (s_lo + per) : hi;
        if (s_lo > hi) s_lo = hi;
    };

    int cur_ib = -1;
    if (t_begin < t_end) {                                 // prologue: tile t_begin in flight
        stage_tile(sbase + L::XJ, A.tile_jb[t_begin]);
        stage_records(0, A.tile_ptr[t_begin], first_chunk_len(t_begin));
        cp_async_commit();
    }
    for (int t = t_begin; t < t_end; ++t) {
        const int buf = (t - t_begin) & 1;
        const int ib = A.tile_ib[t];
        const int64_t r0 = A.tile_ptr[t], r1 = A.tile_ptr[t + 1], rmid = A.tile_mid[t];
        if (ib != cur_ib) {                                // new I block (once per row of tiles)
            __syncthreads();                               // all threads are done with the previous tile's compute
            stage_tile(sbase + L::XI, ib);
            cp_async_commit();
            cur_ib = ib;
        }
        cp_async_wait_all();
        __syncthreads();                                   // tile t staged; previous tile's buffers are free
        if (t + 1 < t_end) {                               // prefetch tile t+1 into the other buffers
            stage_tile(sbase + L::XJ + (buf ^ 1) * (BJ * TA * 8), A.tile_jb[t + 1]);
            stage_records(buf ^ 1, A.tile_ptr[t + 1], first_chunk_len(t + 1));
            cp_async_commit();
        }
        const uint32_t sXJ_addr = sbase + L::XJ + buf * (BJ * TA * 8) + (uint32_t)lane * 16u;
        const uint32_t sREC_addr = sbase + L::REC + buf * L::REC_BYTES;
        const uint32_t sCD_addr = sbase + L::CD + buf * L::CD_BYTES;
        for (int64_t rc = r0; rc < r1; rc += V2_RCHUNK) {
            const int n = (int)((r1 - rc < V2_RCHUNK) ? (r1 - rc) : V2_RCHUNK);
            if (rc != r0) {                                // rare: more than one record chunk in this tile
                __syncthreads();                           // previous chunk consumed
                stage_records(buf, rc, n);
                cp_async_commit();
                cp_async_wait_all();                       // (also completes the prefetch; correctness only)
                __syncthreads();
            }
            int s_lo, s_hi;
            if constexpr (MODE == MODE_IRLS_FULL) {
                slice_of(0, n, s_lo, s_hi);
                if (s_lo < s_hi) {
                    if (may_clamp) v3_irls_range<Q, DIFF, true>(s_lo, s_hi, sREC_addr, sCD_addr, sXI_addr, sXJ_addr, sT_addr, P.diff_type, beta, acc, esum, nren);
                    else v3_irls_range<Q, DIFF, false>(s_lo, s_hi, sREC_addr, sCD_addr, sXI_addr, sXJ_addr, sT_addr, P.diff_type, beta, acc, esum, nren, P.dbg);
                }
            } else if constexpr (MODE == MODE_LM) {
                slice_of(0, n, s_lo, s_hi);
                v3_moment_range<MODE_LM, Q, DIFF, 0>(s_lo, s_hi, sREC_addr, sCD_addr, sXI_addr, sXJ_addr, P.diff_type, acc);
            } else {
                int ms = (int)(rmid - rc);                 // first miss record of this chunk
                ms = ms < 0 ? 0 : (ms > n ? n : ms);
                slice_of(0, ms, s_lo, s_hi);
                v3_moment_range<MODE_SIGN, Q, DIFF, 0>(s_lo, s_hi, sREC_addr, sCD_addr, sXI_addr, sXJ_addr, P.diff_type, acc);
                slice_of(ms, n, s_lo, s_hi);
                v3_moment_range<MODE_SIGN, Q, DIFF, 1>(s_lo, s_hi, sREC_addr, sCD_addr, sXI_addr, sXJ_addr, P.diff_type, acc);
            }
        }
    }
    if (MODE == MODE_IRLS_FULL) {                          // integer exponent sums -> log(u1) total
#pragma unroll
        for (int k = 0; k < V3_APT; ++k)
            acc[k][NA - 2] = fma((double)(int)(esum[k] - 1023u * (unsigned)nren), c_k[4], log(acc[k][NA - 2]));
    }
    if (MODE == MODE_SIGN) {                               // (hits, misses) -> (S d, S d*s)
#pragma unroll
        for (int k = 0; k < V3_APT; ++k) { const double h = acc[k][0], ms = acc[k][2]; acc[k][0] = h + ms; acc[k][2] = ms - h; }
    }
    cp_async_wait_all();
    __syncthreads();
    double *red = (double *)smem_raw;                      // [NSL][NA][TA]
#pragma unroll
    for (int k = 0; k < V3_APT; ++k)
#pragma unroll
        for (int n = 0; n < NA; ++n) red[(slice * NA + n) * TA + attr_of(k)] = acc[k][n];
    __syncthreads();
    for (int e = tid; e < NA * TA; e += V3_THREADS) {
        const int n = e / TA, a = e % TA;
        double sacc = red[(0 * NA + n) * TA + a];
#pragma unroll
        for (int sl = 1; sl < V3_NSL; ++sl) sacc += red[(sl * NA + n) * TA + a];
        A.partials[((int64_t)chunk * NA + n) * P.nattr_pad + a0 + a] = sacc;
    }
}

// ------------------------------------------------------------------ small dense solves (per attribute)
// Cholesky of the symmetric K x K matrix A (full storage) after Jacobi scaling; columns whose pivot
// falls below tol * (scaled) diagonal are treated as aliased, like dqrdc2's limited pivoting moves
// them last and lm/glm report NA for them.  Returns the solution of A x = b (0 for aliased) and
// diag(A^-1) over the non-aliased set.
template <int K>
__device__ void chol_solve(const double (&A)[K][K], const double (&b)[K], double tol, double (&x)[K],
                           double (&dinv)[K], bool (&aliased)[K]) {
    double s[K], L[K][K];
    for (int i = 0; i < K; ++i) { aliased[i] = !(A[i][i] > 0.0); s[i] = aliased[i] ? 0.0 : 1.0 / sqrt(A[i][i]); }
    for (int j = 0; j < K; ++j) {
        for (int i = 0; i < K; ++i) L[i][j] = 0.0;
        if (aliased[j]) continue;
        double v = 1.0;                                    // scaled diagonal
        for (int t = 0; t < j; ++t) if (!aliased[t]) v -= L[j][t] * L[j][t];
        if (!(v > tol)) { aliased[j] = true; continue; }
        const double ljj = sqrt(v);
        L[j][j] = ljj;
        for (int i = j + 1; i < K; ++i) {
            if (aliased[i]) continue;
            double u = A[i][j] * s[i] * s[j];
            for (int t = 0; t < j; ++t) if (!aliased[t]) u -= L[i][t] * L[j][t];
            L[i][j] = u / ljj;
        }
    }
    // forward / backward substitution on the scaled system
    double yv[K];
    for (int i = 0; i < K; ++i) {
        if (aliased[i]) { yv[i] = 0.0; continue; }
        double u = b[i] * s[i];
        for (int t = 0; t < i; ++t) if (!aliased[t]) u -= L[i][t] * yv[t];
        yv[i] = u / L[i][i];
    }
    for (int i = K - 1; i >= 0; --i) {
        if (aliased[i]) { x[i] = 0.0; continue; }
        double u = yv[i];
        for (int t = i + 1; t < K; ++t) if (!aliased[t]) u -= L[t][i] * x[t];
        x[i] = u / L[i][i];
    }
    for (int i = 0; i < K; ++i) x[i] *= s[i];
    // diag of the inverse: (A^-1)_cc = s_c^2 * sum_{i >= c} (L^-1)_{ic}^2, with L^-1 formed once, row by row
    // (forward substitution on the rows: Li[i][c] = (delta_ic - sum_{t=c}^{i-1} L[i][t] Li[t][c]) / L[i][i])
    double Li[K][K];
    for (int i = 0; i < K; ++i)
        for (int c = 0; c < K; ++c) Li[i][c] = 0.0;
    for (int i = 0; i < K; ++i) {
        if (aliased[i]) continue;
        const double rl = 1.0 / L[i][i];
        for (int c = 0; c <= i; ++c) {
            if (aliased[c]) continue;
            double u = (i == c) ? 1.0 : 0.0;
            for (int t = c; t < i; ++t) u -= L[i][t] * Li[t][c];      // rows / columns of aliased t are zero
            Li[i][c] = u * rl;
        }
    }
    for (int c = 0; c < K; ++c) {
        double d = 0.0;
        for (int i = c; i < K; ++i) d += Li[i][c] * Li[i][c];
        dinv[c] = d * s[c] * s[c];
    }
}

struct SolveArgs {
    const double *partials; int n_chunks[NB_MAX_STREAMS]; double weight[NB_MAX_STREAMS]; int64_t part_stride[NB_MAX_STREAMS];
    int fold;                      // covariate (0-based) folded into the intercept stream by stream, or -1 (see nb_regress)
    double fold_val[NB_MAX_STREAMS];   // its value on every stream
    int nattr, nattr_pad;
    double *beta, *se;             // [KMAX][nattr_pad]
    double *dev_old;               // [nattr_pad]
    // exact stopping decision (see exact_dev_kernel): coefficients of the previous iteration, attributes whose deviance test is
    // too close to glm.fit's epsilon to be decided from the streamed deviance, their exactly re-evaluated deviances
    double *beta_prev;             // [KMAX][nattr_pad]
    unsigned char *ambig;          // [nattr_pad]
    unsigned int *n_ambig;
    const double *dev_exact;       // [2][nattr_pad]: deviance at beta, at beta_prev
    int resolve;                   // 1: second visit of the ambiguous attributes, with dev_exact
    double *hit_d, *miss_d;        // [nattr_pad] sum of the attribute diff over the hit / miss records (from pass 0)
    const double *dmax;            // [nattr] upper bound of the attribute diff over any pair
    double cmax[NB_FAST_COVARS]; // upper bounds of the covariate diffs
    int speculate;                 // 1: finish without the confirming pass when the predicted deviance change is far below epsilon
    unsigned char *active;         // [nattr_pad]
    unsigned char *tile_active;    // [n attr tiles]
    int *iters; int *flags;        // [nattr_pad]
    double *stats;                 // [nattr][8]
    double *dev_final;
    unsigned int *n_active;
    int pass;                      // 0 = after INIT, t >= 1 = after FULL pass t
    double M;                      // number of regression rows (weights included)
    double shared[NSHARED_MAX];    // lm shared moments
};

// Binary-covariate folding.  A match-mismatch covariate diff is 0 or 1, so the pair list is split by its value into streams
// on which that covariate is CONSTANT (c): its design column is c times the intercept column, the streams are evaluated by the
// kernels of the next smaller design (Q - 1 covariates, intercept shifted by c * beta_f: 24 instead of 31 FP64 instructions
// per evaluation at q = 2) and the full statistics are rebuilt here:
//   G[r][s] = cr cs G'[e(r)][e(s)],  b[r] = cr b'[e(r)],  with e(f) = 0 (intercept), cr = c for r = f and 1 otherwise.
// Exact: the same sums in a different grouping.
__device__ __forceinline__ int eff_col(int i, int f) {            // full design column (0 intercept, 1 d, 2+u covariates) -> column of the reduced design
    return (f < 0 || i < 2 + f) ? i : (i == 2 + f ? 0 : i - 1);
}

// moment statistics (MODE_LM / MODE_SIGN): [S d, S d^2, S d y | S d s, S d c_u ...]
template <int Q>
__device__ __forceinline__ void reduce_moments(const SolveArgs &S, int a, double (&out)[3 + Q]) {
    constexpr int NA = 3 + Q;
    for (int n = 0; n < NA; ++n) out[n] = 0.0;
    const int f = S.fold, qe = Q - (f >= 0 ? 1 : 0), nae = 3 + qe;
    const double *base = S.partials;
    for (int st = 0; st < NB_MAX_STREAMS; ++st) {
        if (S.n_chunks[st] == 0) continue;
        double tmp[NA];
        for (int n = 0; n < NA; ++n) tmp[n] = 0.0;
        for (int c = 0; c < S.n_chunks[st]; ++c)
            for (int n = 0; n < nae; ++n) tmp[n] += base[((int64_t)c * nae + n) * S.nattr_pad + a];
        const double w = S.weight[st];
        for (int n = 0; n < 3; ++n) out[n] += w * tmp[n];
        for (int u = 0; u < Q; ++u)
            out[3 + u] += w * ((u == f) ? S.fold_val[st] * tmp[0] : tmp[3 + ((f >= 0 && u > f) ? u - 1 : u)]);
        base += S.part_stride[st];
    }
}

// IRLS statistics: X'WX upper triangle (row-major), the K score sums, log-product and clamp correction
template <int Q>
__device__ __forceinline__ void reduce_irls(const SolveArgs &S, int a, double (&out)[(2 + Q) * (3 + Q) / 2 + (2 + Q) + 2]) {
    constexpr int K = 2 + Q, NA = K * (K + 1) / 2 + K + 2;
    for (int n = 0; n < NA; ++n) out[n] = 0.0;
    const int f = S.fold, ke = K - (f >= 0 ? 1 : 0), nae = ke * (ke + 1) / 2 + ke + 2;
    const int fc = f >= 0 ? 2 + f : -1;                    // design column of the folded covariate
    const double *base = S.partials;
    for (int st = 0; st < NB_MAX_STREAMS; ++st) {
        if (S.n_chunks[st] == 0) continue;
        double tmp[NA];
        for (int n = 0; n < NA; ++n) tmp[n] = 0.0;
        for (int c = 0; c < S.n_chunks[st]; ++c)
            for (int n = 0; n < nae; ++n) tmp[n] += base[((int64_t)c * nae + n) * S.nattr_pad + a];
        const double w = S.weight[st], cv = S.fold_val[st];
        int n = 0;
        for (int r = 0; r < K; ++r)
            for (int c = r; c < K; ++c) {
                int er = eff_col(r, f), ec = eff_col(c, f);
                if (er > ec) { const int t = er; er = ec; ec = t; }
                const int idx = er * ke - er * (er - 1) / 2 + (ec - er);          // upper-triangle position in the reduced design
                const double sc = ((r == fc) ? cv : 1.0) * ((c == fc) ? cv : 1.0);
                out[n] += w * sc * tmp[idx];
                ++n;
            }
        const int nge = ke * (ke + 1) / 2;
        for (int r = 0; r < K; ++r) out[n + r] += w * ((r == fc) ? cv : 1.0) * tmp[nge + eff_col(r, f)];
        out[NA - 2] += w * tmp[nae - 2];
        out[NA - 1] += w * tmp[nae - 1];
        base += S.part_stride[st];
    }
}

// coefficients a stream's pass evaluates with: the reduced design's [b0 + c b_f, b1, the other covariates]
__global__ void fold_beta_kernel(const double *__restrict__ beta, int nattr_pad, int K, int fold, double cval, double *__restrict__ out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nattr_pad) return;
    int e = 0;
    for (int r = 0; r < K; ++r) {
        if (r == 2 + fold) continue;
        double v = beta[(int64_t)r * nattr_pad + a];
        if (r == 0) v += cval * beta[(int64_t)(2 + fold) * nattr_pad + a];
        out[(int64_t)e * nattr_pad + a] = v;
        ++e;
    }
}

__device__ __forceinline__ void write_stats(double *st, const double *beta, const double *se, const bool *aliased, int K,
                                            double df, double extra, int iters, int flags) {
    int row2 = -1;
    for (int j = 1; j < K; ++j) if (!aliased[j]) { row2 = j; break; }   // summary() drops aliased rows (quirk 3)
    const double nanv = __longlong_as_double(0x7ff8000000000000ll);
    st[0] = row2 >= 0 ? beta[row2] : nanv;
    st[1] = row2 >= 0 ? beta[row2] / se[row2] : nanv;
    st[2] = aliased[0] ? nanv : beta[0];
    st[3] = aliased[0] ? nanv : beta[0] / se[0];
    st[4] = df;
    st[5] = extra;
    st[6] = (double)iters;
    st[7] = (double)(flags | (aliased[1] ? NPDRB_FLAG_ALIASED : 0));
}

// lm: closed-form OLS from the moments (lm + summary.lm, R/npdr.R:34,41-67)
template <int Q>
__global__ void lm_finalize_kernel(SolveArgs S) {
    constexpr int K = 2 + Q, NA = 3 + Q;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= S.nattr) return;
    double m[NA];
    reduce_moments<Q>(S, a, m);
    // shared: [M, Sy, Syy, Sc_r, Scy_r, Scc_rs]
    const double M = S.shared[0], Sy = S.shared[1], Syy = S.shared[2];
    const double *Sc = S.shared + 3, *Scy = S.shared + 3 + Q, *Scc = S.shared + 3 + 2 * Q;
    double A[K][K], b[K];
    // uncentred normal equations; dqrdc2's alias test compares against uncentred column norms
    A[0][0] = M; A[0][1] = A[1][0] = m[0]; A[1][1] = m[1]; b[0] = Sy; b[1] = m[2];
    int cc = 0;
    for (int r = 0; r < Q; ++r) {
        A[0][2 + r] = A[2 + r][0] = Sc[r];
        A[1][2 + r] = A[2 + r][1] = m[3 + r];
        b[2 + r] = Scy[r];
        for (int u = r; u < Q; ++u) { A[2 + r][2 + u] = A[2 + u][2 + r] = Scc[cc]; ++cc; }
    }
    // centre everything but the intercept (Schur complement on column 0) for conditioning
    double G[K][K], g[K], mean[K];
    for (int i = 0; i < K; ++i) mean[i] = A[0][i] / M;
    const double ybar = Sy / M;
    for (int i = 1; i < K; ++i) {
        g[i] = b[i] - A[0][i] * ybar;
        for (int j = 1; j < K; ++j) G[i][j] = A[i][j] - A[0][i] * mean[j];
    }
    // solve the centred (K-1) system through chol_solve<K> with a unit intercept block
    G[0][0] = 1.0; g[0] = 0.0;
    for (int j = 1; j < K; ++j) { G[0][j] = 0.0; G[j][0] = 0.0; }
    // alias test relative to the UNCENTRED norm: scale tolerance per column below
    double x[K], dinv[K];
    bool al[K];
    // pre-mark columns whose centred norm is negligible against the uncentred one (dqrdc2 tol 1e-7 on norms)
    double Gs[K][K];
    for (int i = 0; i < K; ++i) for (int j = 0; j < K; ++j) Gs[i][j] = G[i][j];
    for (int j = 1; j < K; ++j) {
        if (!(A[j][j] > 0.0) || !(G[j][j] > 1e-14 * A[j][j])) {   // constant column (incl. all zero)
            for (int i = 0; i < K; ++i) { Gs[i][j] = 0.0; Gs[j][i] = 0.0; }
        }
    }
    chol_solve<K>(Gs, g, 1e-14, x, dinv, al);
    double beta[K], se[K];
    double b0 = ybar;
    for (int j = 1; j < K; ++j) { beta[j] = al[j] ? 0.0 : x[j]; b0 -= beta[j] * mean[j]; }
    beta[0] = b0;
    int rank = 1;
    for (int j = 1; j < K; ++j) rank += al[j] ? 0 : 1;
    const double Syy_c = Syy - Sy * ybar;
    double rss = Syy_c;
    for (int j = 1; j < K; ++j) rss -= beta[j] * g[j];
    const double rdf = M - (double)rank;
    const double resvar = rss / rdf;
    // var(beta_0) = resvar * (1/M + mean' G^-1 mean): needs the full inverse -> solve G z = mean
    double gm[K], z[K], d2[K];
    bool al2[K];
    gm[0] = 0.0;
    for (int j = 1; j < K; ++j) gm[j] = al[j] ? 0.0 : mean[j];
    chol_solve<K>(Gs, gm, 1e-14, z, d2, al2);
    double quad = 0.0;
    for (int j = 1; j < K; ++j) quad += gm[j] * z[j];
    se[0] = sqrt(resvar * (1.0 / M + quad));
    for (int j = 1; j < K; ++j) se[j] = sqrt(resvar * dinv[j]);
    al[0] = false;
    const double rsq = 1.0 - rss / Syy_c;                      // mss/(mss+rss) with an intercept
    write_stats(S.stats + (int64_t)a * 8, beta, se, al, K, rdf, rsq, 0, 0);
}

// binomial, first solve: glm.fit starts from mustart = (y + 0.5)/2, i.e. mu = 0.75 / 0.25, eta = +-log 3, constant
// weight w0 = mu(1-mu) = 3/16 and working response z = s * z0 (s = +1 miss, -1 hit).  X'WX and X'Wz are therefore
// w0 times the plain moments the MODE_SIGN pass accumulated, and the deviance at mustart is M * 2 log(4/3).
template <int Q>
__global__ void irls_init_kernel(SolveArgs S) {
    constexpr int K = 2 + Q, NA = 3 + Q;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= S.nattr) return;
    double m[NA];
    reduce_moments<Q>(S, a, m);
    const double M = S.shared[0], Ss = S.shared[1];
    const double *Sc = S.shared + 3, *Scs = S.shared + 3 + Q, *Scc = S.shared + 3 + 2 * Q;
    const double w0 = 0.1875, z0 = 1.0986122886681098 + 0.25 / 0.1875;
    double A[K][K], b[K];
    A[0][0] = w0 * M; A[0][1] = A[1][0] = w0 * m[0]; A[1][1] = w0 * m[1];
    b[0] = w0 * z0 * Ss; b[1] = w0 * z0 * m[2];
    int cc = 0;
    for (int r = 0; r < Q; ++r) {
        A[0][2 + r] = A[2 + r][0] = w0 * Sc[r];
        A[1][2 + r] = A[2 + r][1] = w0 * m[3 + r];
        b[2 + r] = w0 * z0 * Scs[r];
        for (int u = r; u < Q; ++u) { A[2 + r][2 + u] = A[2 + u][2 + r] = w0 * Scc[cc]; ++cc; }
    }
    S.dev_old[a] = M * (2.0 * 0.2876820724517809);         // sum of dev.resids at mustart: 2*log(4/3) each
    S.hit_d[a] = 0.5 * (m[0] - m[2]);                       // S(d) = hits + misses, S(d*s) = misses - hits
    S.miss_d[a] = 0.5 * (m[0] + m[2]);
    double x[K], dinv[K];
    bool al[K];
    chol_solve<K>(A, b, 1e-13, x, dinv, al);
    const double nanv = __longlong_as_double(0x7ff8000000000000ll);
    bool ok = true;
    for (int r = 0; r < K; ++r) ok = ok && isfinite(x[r]);
    if (!ok) {
        double beta[K], se[K];
        for (int r = 0; r < K; ++r) { beta[r] = nanv; se[r] = nanv; al[r] = false; }
        write_stats(S.stats + (int64_t)a * 8, beta, se, al, K, M - (double)K, nanv, 0, NPDRB_FLAG_NONFINITE);
        S.active[a] = 0;
        return;
    }
    for (int r = 0; r < K; ++r) {
        S.beta[(int64_t)r * S.nattr_pad + a] = x[r];
        S.se[(int64_t)r * S.nattr_pad + a] = al[r] ? nanv : sqrt(dinv[r]);
    }
    S.iters[a] = 1;
    S.tile_active[a / TA] = 1;
    atomicAdd(S.n_active, 1u);
}

// binomial: one IRLS bookkeeping step per attribute (glm.fit's loop body)
template <int Q>
__global__ void irls_step_kernel(SolveArgs S) {
    constexpr int K = 2 + Q, NA = K * (K + 1) / 2 + K + 2;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= S.nattr) return;
    if (!S.active[a]) return;
    double m[NA];
    reduce_irls<Q>(S, a, m);
    double A[K][K], b[K];                                   // X'WX and the score X'(y - mu)
    int n = 0;
    for (int r = 0; r < K; ++r) for (int c = r; c < K; ++c) { A[r][c] = A[c][r] = m[n]; ++n; }
    // hit / miss moments of the features v = [1, d, c_1..c_q]: attribute-independent ones from the shared moments
    // (M, S(s), ., S(c), S(c s), ...), the attribute diff from pass 0
    const double Mtot = S.shared[0], Ss = S.shared[1];
    const double *Sc = S.shared + 3, *Scs = S.shared + 3 + Q;
    double hitv[K], missv[K];
    hitv[0] = 0.5 * (Mtot - Ss); missv[0] = 0.5 * (Mtot + Ss);
    hitv[1] = S.hit_d[a]; missv[1] = S.miss_d[a];
    for (int r = 0; r < Q; ++r) { hitv[2 + r] = 0.5 * (Sc[r] - Scs[r]); missv[2 + r] = 0.5 * (Sc[r] + Scs[r]); }
    // score X'(y - mu) = sum over all of (1 - mu) v  -  sum over hits of v
    for (int r = 0; r < K; ++r) b[r] = m[n + r] - hitv[r];
    // residual deviance = 2 * sum[ log(1 + t) - y * log(t) ], log t = eta = beta . v except for clamped records
    double lin = 0.0;
    for (int r = 0; r < K; ++r) lin += S.beta[(int64_t)r * S.nattr_pad + a] * missv[r];
    double dev = 2.0 * (m[NA - 2] - lin - m[NA - 1]);
    bool finish = false;
    int flags = 0;
    {
        double devold = S.dev_old[a];
        // glm.fit stops when |dev - devold| / (|dev| + 0.1) < 1e-8 and reports the standard errors of THAT iteration's
        // weights, so stopping one iteration early or late moves Z by ~1e-5: the decision has to be R's.  The streamed
        // deviance carries the exp polynomial's error (<= 1.7e-12 per record, systematic when the diffs take few distinct
        // values, e.g. SNP data); when the test lands inside that error band the attribute is parked, its two deviances
        // are re-evaluated record by record with full-accuracy exp / log (exact_dev_kernel) and this kernel visits it again.
        if (S.resolve) {
            if (!S.ambig[a]) return;
            S.ambig[a] = 0;
            dev = S.dev_exact[a];
            if (S.pass >= 2) devold = S.dev_exact[S.nattr_pad + a];
        } else if (isfinite(dev) && S.ambig) {
            const double thr = 1e-8 * (fabs(dev) + 0.1);
            if (fabs(fabs(dev - devold) - thr) <= 8e-12 * S.M + 1e-3 * thr * 1e-3) {
                S.ambig[a] = 1;
                atomicAdd(S.n_ambig, 1u);
                return;
            }
        }
        if (!isfinite(dev)) { finish = true; flags |= NPDRB_FLAG_NONFINITE; }
        else if (fabs(dev - devold) / (fabs(dev) + 0.1) < 1e-8) finish = true;      // glm.control(epsilon = 1e-8)
        else if (S.pass >= 25) { finish = true; flags |= NPDRB_FLAG_NOT_CONVERGED; }   // maxit = 25
        else S.dev_old[a] = dev;
    }
    if (finish) {
        double beta[K], se[K];
        bool al[K];
        for (int r = 0; r < K; ++r) {
            beta[r] = S.beta[(int64_t)r * S.nattr_pad + a];
            se[r] = S.se[(int64_t)r * S.nattr_pad + a];
            al[r] = !(se[r] == se[r]);                         // NaN se marks an aliased coefficient
        }
#ifdef NB_DEBUG_SE
        if (a == 0 && Q == 4) printf("finish pass %d: se %g %g %g %g %g %g\n", S.pass, se[0], se[1], se[2], se[3], se[4], se[5 < K ? 5 : 0]);
#endif
        int rank = 0;
        for (int r = 0; r < K; ++r) rank += al[r] ? 0 : 1;
        write_stats(S.stats + (int64_t)a * 8, beta, se, al, K, S.M - (double)rank, dev, S.pass, flags);
        S.active[a] = 0;
        return;
    }
    double x[K], dinv[K];
    bool al[K];
    chol_solve<K>(A, b, 1e-13, x, dinv, al);
    bool ok = true;
    for (int r = 0; r < K; ++r) ok = ok && isfinite(x[r]);
    const double nanv = __longlong_as_double(0x7ff8000000000000ll);
    if (!ok) {
        // non-finite coefficients: glm.fit stops with a warning; report what we have
        double beta[K], se[K];
        for (int r = 0; r < K; ++r) { beta[r] = S.beta[(int64_t)r * S.nattr_pad + a]; se[r] = S.se[(int64_t)r * S.nattr_pad + a]; al[r] = false; }
        write_stats(S.stats + (int64_t)a * 8, beta, se, al, K, S.M - (double)K, dev, S.pass, NPDRB_FLAG_NONFINITE);
        S.active[a] = 0;
        return;
    }
#ifdef NB_DEBUG_SE
    if (a == 0 && Q == 4) {
        printf("step pass %d: dinv %g %g %g %g %g %g  al %d%d%d%d%d%d\n", S.pass, dinv[0], dinv[1], dinv[2], dinv[3], dinv[4], dinv[5 < K ? 5 : 0],
               (int)al[0], (int)al[1], (int)al[2], (int)al[3], (int)al[4], (int)al[5 < K ? 5 : 0]);
        for (int r = 0; r < K; ++r) printf("   A[%d]: %g %g %g %g %g %g | b %g\n", r, A[r][0], A[r][1], A[r][2], A[r][3], A[r][4], A[r][5 < K ? 5 : 0], b[r]);
    }
#endif
    double bnew[K], senew[K];
    for (int r = 0; r < K; ++r) {                           // Newton / IRLS update: beta + (X'WX)^-1 X'(y - mu)
        const double old = S.beta[(int64_t)r * S.nattr_pad + a];
        bnew[r] = al[r] ? 0.0 : old + x[r];
        senew[r] = al[r] ? nanv : sqrt(dinv[r]);
    }
    if (S.speculate && S.pass < 25) {
        // glm.fit would now run iteration pass+1: weights at the CURRENT beta (the system just solved), new coefficients
        // bnew, then test |dev_new - dev| / (|dev_new| + 0.1) < 1e-8.  With A delta = s the deviance along the step is
        //   dev_new - dev = -delta.s + R3,   |R3| <= (max|g3|/3) * sum |x.delta|^3 <= 0.0321 * M * B^3,
        // g = log(1 + e^eta), g3 its third derivative (max 1/(6 sqrt 3)), B >= max |x.delta| from the per-attribute range
        // bounds.  If even the bound is 10x below epsilon (and no eta can reach the +-30 clamps), iteration pass+1 is known
        // to converge: its coefficients are bnew and its QR (standard errors) is the system just solved, so the confirming
        // pass over the pairs is skipped.  Same beta, se and iteration count as running it; the reported deviance is
        // dev - delta.s (error <= R3).
        double vmax[K];
        vmax[0] = 1.0; vmax[1] = S.dmax[a];
        for (int r = 0; r < Q; ++r) vmax[2 + r] = S.cmax[r];
        double pred = 0.0, B = 0.0, E = 0.0;
        for (int r = 0; r < K; ++r) {
            if (al[r]) continue;
            pred += x[r] * b[r];
            B += fabs(x[r]) * vmax[r];
            E += fmax(fabs(bnew[r]), fabs(bnew[r] - x[r])) * vmax[r];
        }
        const double r3 = 0.0321 * S.M * B * B * B;
        const double dev_new_lo = dev - fabs(pred) - r3;
        if (E < 30.0 && pred >= 0.0 && isfinite(pred) && (fabs(pred) + r3) < 1e-9 * (fabs(dev_new_lo) + 0.1)) {
            int rank = 0;
            for (int r = 0; r < K; ++r) rank += al[r] ? 0 : 1;
            write_stats(S.stats + (int64_t)a * 8, bnew, senew, al, K, S.M - (double)rank, dev - pred, S.pass + 1, 0);
            S.active[a] = 0;
            return;
        }
    }
    for (int r = 0; r < K; ++r) {
        if (S.beta_prev) S.beta_prev[(int64_t)r * S.nattr_pad + a] = S.beta[(int64_t)r * S.nattr_pad + a];
        S.beta[(int64_t)r * S.nattr_pad + a] = bnew[r];
        S.se[(int64_t)r * S.nattr_pad + a] = senew[r];
    }
    S.iters[a] = S.pass + 1;
    S.tile_active[a / TA] = 1;
    atomicAdd(S.n_active, 1u);
}

// Exact residual deviance of the parked attributes at two coefficient vectors, record by record over the ORIGINAL pair list,
// with libdevice exp / log and binomial()'s eta clamps -- dev.resids as glm.fit forms them (stats family.c, R/npdr.R:36).
// grid = (attributes, splits of the pair list); CTAs of attributes that are not parked exit at once; fixed-order reductions
// (inside a CTA, then over the splits in exact_dev_reduce_kernel), so the result does not depend on scheduling.
struct ExactArgs {
    const double *X; int64_t ldx; int64_t attr0; int nattr, nattr_pad; int64_t N;
    const int32_t *Ri, *NN; int64_t M;
    const double *y, *C; int q, diff_type, y_raw;
    int cov_type[NB_FAST_COVARS];
    const double *beta, *beta_prev; const unsigned char *ambig;
    double *dev_exact;             // [2][nattr_pad]
    double *part;                  // [nattr][n_split][2]
    int n_split;
};
__device__ __forceinline__ double exact_dev_term(double eta, bool miss) {
    const double t = (eta < -30.0) ? 2.220446049250313e-16 : ((eta > 30.0) ? 4503599627370496.0 : exp(eta));
    const double mu = t / (1.0 + t);
    return miss ? 2.0 * log(1.0 / mu) : 2.0 * log(1.0 / (1.0 - mu));         // 2 (y log(y/mu) + (1-y) log((1-y)/(1-mu)))
}
__global__ void __launch_bounds__(256)
exact_dev_kernel(ExactArgs E) {
    const int a = blockIdx.x;
    if (a >= E.nattr || !E.ambig[a]) return;
    __shared__ double red0[256], red1[256];
    const int K = 2 + E.q;
    double b0[KMAX], b1[KMAX];
    for (int r = 0; r < K; ++r) { b0[r] = E.beta[(int64_t)r * E.nattr_pad + a]; b1[r] = E.beta_prev[(int64_t)r * E.nattr_pad + a]; }
    const double *xcol = E.X + (E.attr0 + a) * E.ldx;
    double s0 = 0.0, s1 = 0.0;
    const int64_t per = (E.M + E.n_split - 1) / E.n_split;
    const int64_t m_lo = (int64_t)blockIdx.y * per, m_hi = (m_lo + per < E.M) ? m_lo + per : E.M;
    for (int64_t m = m_lo + threadIdx.x; m < m_hi; m += 256) {
        const int64_t i = E.Ri[m] - 1, j = E.NN[m] - 1;
        const double d = nb_diff1(xcol[i], xcol[j], E.diff_type);
        const bool miss = E.y_raw ? (E.y[i] != 0.0) : (E.y[i] != E.y[j]);
        double e0 = fma(b0[1], d, b0[0]), e1 = fma(b1[1], d, b1[0]);
        for (int u = 0; u < E.q; ++u) {
            const double *col = E.C + (int64_t)u * E.N;
            const double c = E.y_raw ? col[i] : nb_diff1(col[i], col[j], E.cov_type[u]);
            e0 = fma(b0[2 + u], c, e0); e1 = fma(b1[2 + u], c, e1);
        }
        s0 += exact_dev_term(e0, miss); s1 += exact_dev_term(e1, miss);
    }
    red0[threadIdx.x] = s0; red1[threadIdx.x] = s1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { red0[threadIdx.x] += red0[threadIdx.x + o]; red1[threadIdx.x] += red1[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double *o = E.part + ((int64_t)a * E.n_split + blockIdx.y) * 2;
        o[0] = red0[0]; o[1] = red1[0];
    }
}
__global__ void exact_dev_reduce_kernel(ExactArgs E) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= E.nattr || !E.ambig[a]) return;
    double s0 = 0.0, s1 = 0.0;
    for (int y = 0; y < E.n_split; ++y) { const double *o = E.part + ((int64_t)a * E.n_split + y) * 2; s0 += o[0]; s1 += o[1]; }
    E.dev_exact[a] = s0; E.dev_exact[E.nattr_pad + a] = s1;
}

// ------------------------------------------------------------------ host side
template <typename T>
int dmalloc(T **p, size_t n) {
    cudaError_t e = cudaMalloc((void **)p, sizeof(T) * (n ? n : 1));
    if (e != cudaSuccess) { npdrb_set_error("cudaMalloc of %zu bytes failed: %s", sizeof(T) * n, cudaGetErrorString(e)); return NPDRB_ENOMEM; }
    return NPDRB_OK;
}

nb_tables *g_tables_dev[64] = {nullptr};
nb_tables1024 *g_tables1024_dev[64] = {nullptr};

std::mutex g_tables_mu;
int get_tables1024(npdrb_ctx *ctx, const nb_tables1024 **out) {
    std::lock_guard<std::mutex> lk(g_tables_mu);
    if (!g_tables1024_dev[ctx->device & 63]) {
        static const nb_tables1024 host_tables = NB_TABLES1024_STATIC_INIT;
        nb_tables1024 *d = nullptr;
        cudaError_t e = (cudaMalloc)((void **)&d, sizeof(nb_tables1024));       // lives as long as the process: not from the cache
        if (e != cudaSuccess) { npdrb_set_error("cudaMalloc(tables1024): %s", cudaGetErrorString(e)); return NPDRB_ENOMEM; }
        NB_CUDA(cudaMemcpyAsync(d, &host_tables, sizeof(nb_tables1024), cudaMemcpyHostToDevice, ctx->stream));
        NB_CUDA(cudaStreamSynchronize(ctx->stream));
        g_tables1024_dev[ctx->device & 63] = d;
    }
    *out = g_tables1024_dev[ctx->device & 63];
    return NPDRB_OK;
}

int get_tables(npdrb_ctx *ctx, const nb_tables **out) {
    std::lock_guard<std::mutex> lk(g_tables_mu);
    if (!g_tables_dev[ctx->device & 63]) {
        static const nb_tables host_tables = NB_TABLES_STATIC_INIT;
        nb_tables *d = nullptr;
        cudaError_t e = (cudaMalloc)((void **)&d, sizeof(nb_tables));           // lives as long as the process: not from the cache
        if (e != cudaSuccess) { npdrb_set_error("cudaMalloc(tables): %s", cudaGetErrorString(e)); return NPDRB_ENOMEM; }
        NB_CUDA(cudaMemcpyAsync(d, &host_tables, sizeof(nb_tables), cudaMemcpyHostToDevice, ctx->stream));
        NB_CUDA(cudaStreamSynchronize(ctx->stream));
        g_tables_dev[ctx->device & 63] = d;
    }
    *out = g_tables_dev[ctx->device & 63];
    return NPDRB_OK;
}

void free_stream(nb_stream &st) {
    cudaFree(st.rec); cudaFree(st.yv);
    for (int c = 0; c < NB_FAST_COVARS; ++c) cudaFree(st.cd[c]);
    cudaFree(st.tile_ptr); cudaFree(st.tile_mid); cudaFree(st.tile_ib); cudaFree(st.tile_jb); cudaFree(st.chunk_tile);
    st = nb_stream();
}

// Build one bucketed stream from a pair list (device).  y_raw: uniReg mode.
struct CovPlan {                    // which covariates a stream stores per record (all but the folded one)
    int q = 0;                      // stored covariates
    int src[NB_FAST_COVARS] = {0, 0, 0, 0};    // column of C behind stored covariate u
    int type[NB_FAST_COVARS] = {0, 0, 0, 0};   // its diff type
    int fold = -1;                  // folded covariate (column of C), or -1
};
int build_stream(npdrb_session *s, const int32_t *dRi, const int32_t *dNN, int64_t M, int reg_type,
                 const CovPlan &cp, int y_raw, double weight, double fold_val, double work_share, nb_stream &st, int64_t *n_miss_out) {
    npdrb_ctx *ctx = s->ctx;
    const int64_t N = s->N;
    const int q = cp.q;
    const int n_ib = (int)nb_cdiv(N, BI), n_jb = (int)nb_cdiv(N, BJ);
    const int64_t n_tiles_all = (int64_t)n_ib * n_jb;
    st = nb_stream();
    st.n_rec = M; st.weight = weight; st.fold_val = fold_val;
    uint32_t *keys = nullptr, *idx = nullptr, *keys2 = nullptr, *perm = nullptr;
    unsigned *hist = nullptr;
    unsigned long long *d_miss = nullptr;
    void *tmp = nullptr;
    nb_scope scope;                                     // temporaries: released on every return path
    NB_CHECK(scope.alloc(&keys, (size_t)M)); NB_CHECK(scope.alloc(&idx, (size_t)M));
    NB_CHECK(scope.alloc(&keys2, (size_t)M)); NB_CHECK(scope.alloc(&perm, (size_t)M));
    NB_CHECK(scope.alloc(&hist, (size_t)2 * n_tiles_all)); NB_CHECK(scope.alloc(&d_miss, 1));
    NB_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned) * (size_t)2 * n_tiles_all, ctx->stream));
    NB_CUDA(cudaMemsetAsync(d_miss, 0, sizeof(unsigned long long), ctx->stream));
    const unsigned grid = (unsigned)nb_cdiv(M, 256);
    const int ybits = (reg_type == NPDRB_REG_BINOMIAL) ? (y_raw ? 2 : 1) : 0;
    tile_keys_kernel<<<grid, 256, 0, ctx->stream>>>(dRi, dNN, M, n_jb, s->dy, ybits, keys, idx);
    NB_LAUNCH_CHECK(ctx);
    int end_bit = 1;
    while ((1ll << end_bit) < 2 * n_tiles_all && end_bit < 32) ++end_bit;
    size_t tmp_bytes = 0;
    NB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, idx, perm, (int)M, 0, end_bit, ctx->stream));
    NB_CHECK(scope.alloc((unsigned char **)&tmp, tmp_bytes ? tmp_bytes : 1));
    NB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, idx, perm, (int)M, 0, end_bit, ctx->stream));
    ctx->launches += 3;
    tile_hist_kernel<<<grid, 256, 0, ctx->stream>>>(keys2, M, hist);
    NB_LAUNCH_CHECK(ctx);

    NB_CHECK(dmalloc(&st.rec, (size_t)M));
    if (reg_type == NPDRB_REG_LM) NB_CHECK(dmalloc(&st.yv, (size_t)M));
    for (int c = 0; c < q; ++c) NB_CHECK(dmalloc(&st.cd[c], (size_t)M));
    GatherArgs g;
    g.Ri = dRi; g.NN = dNN; g.perm = perm; g.y = s->dy; g.C = s->dC; g.M = M; g.N = N; g.q = q; g.reg_type = reg_type;
    for (int c = 0; c < NB_FAST_COVARS; ++c) { g.cov_type[c] = cp.type[c]; g.cov_src[c] = cp.src[c]; g.cd[c] = st.cd[c]; }
    g.y_raw = y_raw; g.rec = st.rec; g.yv = st.yv; g.n_miss = d_miss;
    gather_records_kernel<<<grid, 256, 0, ctx->stream>>>(g);
    NB_LAUNCH_CHECK(ctx);

    std::vector<unsigned> h_hist((size_t)2 * n_tiles_all);
    unsigned long long h_miss = 0;
    NB_CUDA(cudaMemcpyAsync(h_hist.data(), hist, sizeof(unsigned) * h_hist.size(), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaMemcpyAsync(&h_miss, d_miss, sizeof(h_miss), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n_miss_out) *n_miss_out = (int64_t)h_miss;

    std::vector<int64_t> tptr, tmid; std::vector<int32_t> tib, tjb;
    int64_t run = 0;
    for (int64_t t = 0; t < n_tiles_all; ++t) {
        const int64_t n0 = h_hist[2 * t], n1 = h_hist[2 * t + 1];
        if (!(n0 + n1)) continue;
        tptr.push_back(run); tmid.push_back(run + n0);
        tib.push_back((int32_t)(t / n_jb)); tjb.push_back((int32_t)(t % n_jb));
        run += n0 + n1;
    }
    tptr.push_back(run);
    st.n_tiles = (int32_t)tib.size();
    // work chunks: contiguous tile ranges balanced by record count
    const int64_t n_at = nb_cdiv(s->xt_nattr > 0 ? s->xt_nattr : s->p, TA);
    int64_t target = (int64_t)(16.0 * ctx->sm_count * work_share);
    int64_t C = nb_cdiv(target > 1 ? target : 1, n_at);
    {   // L2 residency: sm_count / (total chunks) attribute-tile slabs are live at once; keep them within ~48 MB
        const double slab = (double)nb_round_up(N, BI > BJ ? BI : BJ) * TA * sizeof(double);
        const int64_t c_l2 = (int64_t)ceil(ctx->sm_count * slab / (48.0 * 1024 * 1024) * work_share);
        if (C < c_l2) C = c_l2;
    }
    if (C < 1) C = 1;
    if (C > st.n_tiles) C = st.n_tiles;
    if (C < 1) C = 1;
    std::vector<int32_t> ctile;
    ctile.push_back(0);
    for (int64_t c = 1; c < C; ++c) {
        const int64_t want = run * c / C;
        int32_t t = ctile.back();
        while (t < st.n_tiles && tptr[t] < want) ++t;
        if (t <= ctile.back()) t = ctile.back() + 1;
        if (t > st.n_tiles) t = st.n_tiles;
        ctile.push_back(t);
    }
    ctile.push_back(st.n_tiles);
    st.n_chunks = (int32_t)ctile.size() - 1;
    NB_CHECK(dmalloc(&st.tile_ptr, tptr.size())); NB_CHECK(dmalloc(&st.tile_mid, tmid.size() + 1));
    NB_CHECK(dmalloc(&st.tile_ib, tib.size() + 1));
    NB_CHECK(dmalloc(&st.tile_jb, tjb.size() + 1)); NB_CHECK(dmalloc(&st.chunk_tile, ctile.size()));
    NB_CUDA(cudaMemcpyAsync(st.tile_ptr, tptr.data(), sizeof(int64_t) * tptr.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (!tib.empty()) {
        NB_CUDA(cudaMemcpyAsync(st.tile_mid, tmid.data(), sizeof(int64_t) * tmid.size(), cudaMemcpyHostToDevice, ctx->stream));
        NB_CUDA(cudaMemcpyAsync(st.tile_ib, tib.data(), sizeof(int32_t) * tib.size(), cudaMemcpyHostToDevice, ctx->stream));
        NB_CUDA(cudaMemcpyAsync(st.tile_jb, tjb.data(), sizeof(int32_t) * tjb.size(), cudaMemcpyHostToDevice, ctx->stream));
    }
    NB_CUDA(cudaMemcpyAsync(st.chunk_tile, ctile.data(), sizeof(int32_t) * ctile.size(), cudaMemcpyHostToDevice, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    return NPDRB_OK;
}

size_t pass_smem(int mode, int q) {
    size_t b = sizeof(double) * (size_t)(BI * TA + BJ * TA) + sizeof(double) * (size_t)(q + 1) * RCHUNK +
               sizeof(uint32_t) * RCHUNK + sizeof(nb_tables);
    size_t red = sizeof(double) * (size_t)NSLICE * nacc_for(mode, q) * TA;
    return b > red ? b : red;
}

template <int MODE, int Q, int DIFF>
int launch_pass_t(npdrb_ctx *ctx, const PassArgs &P, int n_at, int n_chunks) {
    const size_t smem = pass_smem(MODE, Q);
    NB_CUDA(cudaFuncSetAttribute(pass_kernel<MODE, Q, DIFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (n_at > 65535) { npdrb_set_error("regress: more than 65535 attribute tiles in one call; split the attribute range"); return NPDRB_EINVAL; }
    dim3 grid((unsigned)n_chunks, (unsigned)n_at);
    pass_kernel<MODE, Q, DIFF><<<grid, RTHREADS, smem, ctx->stream>>>(P);
    NB_LAUNCH_CHECK(ctx);
    return NPDRB_OK;
}
template <int MODE, int Q>
int launch_pass_q(npdrb_ctx *ctx, const PassArgs &P, int n_at, int n_chunks) {
    return P.diff_type == NPDRB_DIFF_NUMERIC_ABS ? launch_pass_t<MODE, Q, 0>(ctx, P, n_at, n_chunks)
                                                 : launch_pass_t<MODE, Q, 1>(ctx, P, n_at, n_chunks);
}
template <int MODE, int Q, int DIFF>
int launch_pass2_t(npdrb_ctx *ctx, const PassArgs &P, int n_at, int n_chunks) {
    const size_t smem = V2Layout<MODE, Q>::SMEM;
    NB_CUDA(cudaFuncSetAttribute(pass2_kernel<MODE, Q, DIFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (n_at > 65535) { npdrb_set_error("regress: more than 65535 attribute tiles in one call; split the attribute range"); return NPDRB_EINVAL; }
    dim3 grid((unsigned)n_chunks, (unsigned)n_at);
    pass2_kernel<MODE, Q, DIFF><<<grid, V2Layout<MODE, Q>::THREADS, smem, ctx->stream>>>(P);
    NB_LAUNCH_CHECK(ctx);
    return NPDRB_OK;
}
template <int MODE, int Q, int DIFF>
int launch_pass3_t(npdrb_ctx *ctx, const PassArgs &P, int n_at, int n_chunks) {
    const size_t smem = V3Layout<MODE, Q>::SMEM;
    NB_CUDA(cudaFuncSetAttribute(pass3_kernel<MODE, Q, DIFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (n_at > 65535) { npdrb_set_error("regress: more than 65535 attribute tiles in one call; split the attribute range"); return NPDRB_EINVAL; }
    dim3 grid((unsigned)n_chunks, (unsigned)n_at);
    pass3_kernel<MODE, Q, DIFF><<<grid, V3Layout<MODE, Q>::THREADS, smem, ctx->stream>>>(P);
    NB_LAUNCH_CHECK(ctx);
    return NPDRB_OK;
}
bool use_v2_kernel() {
    const char *e = getenv("NPDRB_PASS_KERNEL");
    return e && strcmp(e, "v2") == 0;
}
// v2 / v3 cover q <= 2 (shared memory: I tile + two J tiles + record buffers <= 227 KB); larger q uses the v1 kernel
template <int MODE>
int launch_pass2(npdrb_ctx *ctx, int q, const PassArgs &P, int n_at, int n_chunks) {
    const bool ad = P.diff_type == NPDRB_DIFF_NUMERIC_ABS;
    if (MODE != MODE_IRLS_FULL && !use_v2_kernel()) {        // moment passes: four attributes per thread (v3); IRLS pass: v2
        switch (q) {
        case 0: return ad ? launch_pass3_t<MODE, 0, 0>(ctx, P, n_at, n_chunks) : launch_pass3_t<MODE, 0, 1>(ctx, P, n_at, n_chunks);
        case 1: return ad ? launch_pass3_t<MODE, 1, 0>(ctx, P, n_at, n_chunks) : launch_pass3_t<MODE, 1, 1>(ctx, P, n_at, n_chunks);
        case 2: return ad ? launch_pass3_t<MODE, 2, 0>(ctx, P, n_at, n_chunks) : launch_pass3_t<MODE, 2, 1>(ctx, P, n_at, n_chunks);
        }
    }
    switch (q) {
    case 0: return ad ? launch_pass2_t<MODE, 0, 0>(ctx, P, n_at, n_chunks) : launch_pass2_t<MODE, 0, 1>(ctx, P, n_at, n_chunks);
    case 1: return ad ? launch_pass2_t<MODE, 1, 0>(ctx, P, n_at, n_chunks) : launch_pass2_t<MODE, 1, 1>(ctx, P, n_at, n_chunks);
    case 2: return ad ? launch_pass2_t<MODE, 2, 0>(ctx, P, n_at, n_chunks) : launch_pass2_t<MODE, 2, 1>(ctx, P, n_at, n_chunks);
    }
    npdrb_set_error("pass2: q > 2");
    return NPDRB_EINVAL;
}
bool use_v1_only() {
    const char *e = getenv("NPDRB_PASS_KERNEL");
    return e && strcmp(e, "v1") == 0;
}
bool use_v2(int q) { return q <= 2 && !use_v1_only(); }

template <int MODE>
int launch_pass(npdrb_ctx *ctx, int q, const PassArgs &P, int n_at, int n_chunks) {
    switch (q) {
    case 0: return launch_pass_q<MODE, 0>(ctx, P, n_at, n_chunks);
    case 1: return launch_pass_q<MODE, 1>(ctx, P, n_at, n_chunks);
    case 2: return launch_pass_q<MODE, 2>(ctx, P, n_at, n_chunks);
    case 3: return launch_pass_q<MODE, 3>(ctx, P, n_at, n_chunks);
    case 4: return launch_pass_q<MODE, 4>(ctx, P, n_at, n_chunks);
    }
    npdrb_set_error("regress: at most %d covariates are supported", NB_FAST_COVARS);
    return NPDRB_EINVAL;
}

enum { SOLVE_LM = 0, SOLVE_IRLS_INIT = 1, SOLVE_IRLS_STEP = 2 };
int launch_solve(npdrb_ctx *ctx, int q, int kind, const SolveArgs &S) {
    const unsigned grid = (unsigned)nb_cdiv(S.nattr, 128);
#define NB_SOLVE_CASE(QQ)                                                                  \
    case QQ:                                                                               \
        if (kind == SOLVE_LM) lm_finalize_kernel<QQ><<<grid, 128, 0, ctx->stream>>>(S);    \
        else if (kind == SOLVE_IRLS_INIT) irls_init_kernel<QQ><<<grid, 128, 0, ctx->stream>>>(S); \
        else irls_step_kernel<QQ><<<grid, 128, 0, ctx->stream>>>(S);                       \
        break;
    switch (q) {
        NB_SOLVE_CASE(0) NB_SOLVE_CASE(1) NB_SOLVE_CASE(2) NB_SOLVE_CASE(3) NB_SOLVE_CASE(4)
    default: npdrb_set_error("regress: bad q"); return NPDRB_EINVAL;
    }
#undef NB_SOLVE_CASE
    NB_LAUNCH_CHECK(ctx);
    return NPDRB_OK;
}

}  // namespace

void nb_free_streams(npdrb_session *s) {
    for (int i = 0; i < NB_MAX_STREAMS; ++i) free_stream(s->streams[i]);
    s->n_streams = 0;
    s->streams_valid = false;
}

int nb_regress(npdrb_session *s, int64_t attr0, int64_t nattr, int32_t reg_type, int32_t diff_type,
               const int32_t *cov_type, double *stats_out_host, int32_t *irls_passes) {
    npdrb_ctx *ctx = s->ctx;
    const int q = s->q;
    const int64_t N = s->N;
    if (!s->dX || !s->dy) { npdrb_set_error("regress: inputs not set"); return NPDRB_EINVAL; }
    NB_CHECK(nb_wait_upload(s));
    if (q > 0 && !s->dC) { npdrb_set_error("regress: covariates missing"); return NPDRB_EINVAL; }
    if (!s->dRi || s->M <= 0) { npdrb_set_error("regress: empty pair list"); return NPDRB_EINVAL; }
    if (attr0 < 0 || nattr < 1 || attr0 + nattr > s->p) { npdrb_set_error("regress: bad attribute range"); return NPDRB_EINVAL; }
    if (reg_type != NPDRB_REG_LM && reg_type != NPDRB_REG_BINOMIAL) { npdrb_set_error("regress: bad regression type"); return NPDRB_EINVAL; }
    if (diff_type < 0 || diff_type > NPDRB_DIFF_RAW) { npdrb_set_error("regress: bad attr diff type"); return NPDRB_EINVAL; }
    if (s->M >= (1ll << 31)) { npdrb_set_error("regress: more than 2^31 pairs per device"); return NPDRB_EINVAL; }
    if (q > NB_FAST_COVARS) return nb_regress_wide(s, attr0, nattr, reg_type, diff_type, cov_type, stats_out_host, irls_passes);
    const int y_raw = (diff_type == NPDRB_DIFF_RAW) ? 1 : 0;

    cudaEvent_t e0 = s->ev[0], e1 = s->ev[1];
    NB_CUDA(cudaEventRecord(e0, ctx->stream));
    // instance-major attribute slab
    const int64_t n_at = nb_cdiv(nattr, TA);
    const int64_t nattr_pad = n_at * TA;
    if (s->xt_attr0 != attr0 || s->xt_nattr != nattr) {
        if (s->dXT) { cudaFree(s->dXT); s->dXT = nullptr; }
        const int64_t rows = nb_round_up(N, BI > BJ ? BI : BJ);
        s->ldt = nattr_pad;
        NB_CHECK(dmalloc(&s->dXT, (size_t)rows * (size_t)s->ldt));
        NB_CUDA(cudaMemsetAsync(s->dXT, 0, sizeof(double) * (size_t)rows * (size_t)s->ldt, ctx->stream));
        dim3 tg((unsigned)nb_cdiv(N, 32), (unsigned)nb_cdiv(nattr, 32));
        transpose_kernel<<<tg, dim3(32, 8), 0, ctx->stream>>>(s->dX, s->ldx, N, attr0, nattr, s->dXT, s->ldt);
        NB_LAUNCH_CHECK(ctx);
        s->xt_attr0 = attr0; s->xt_nattr = nattr;
        s->streams_valid = false;                       // chunking depends on the attribute tile count
    }
    // Which covariates the records carry.  The first match-mismatch covariate (diff in {0, 1}) is folded: the pair list is
    // split by its value and every part runs the kernels of the next smaller design (reduce_irls rebuilds the statistics).
    CovPlan cp;
    {
        const char *ncf = getenv("NPDRB_NO_COVFOLD");
        if (!y_raw && !(ncf && ncf[0] == '1'))
            for (int c = 0; c < q && cp.fold < 0; ++c)
                if ((cov_type ? cov_type[c] : NPDRB_DIFF_MATCH_MISMATCH) == NPDRB_DIFF_MATCH_MISMATCH) cp.fold = c;
        for (int c = 0; c < q; ++c) {
            if (c == cp.fold) continue;
            cp.src[cp.q] = c; cp.type[cp.q] = cov_type ? cov_type[c] : NPDRB_DIFF_MATCH_MISMATCH; ++cp.q;
        }
    }
    const int qe = cp.q;                                   // covariates of the design the pass kernels evaluate
    s->q_eff = qe; s->n_exact_stops = 0;
    // bucketed pair stream(s)
    int sig = y_raw * 1000003 + (cp.fold + 1) * 7919;
    for (int c = 0; c < q; ++c) sig = sig * 31 + (cov_type ? cov_type[c] : NPDRB_DIFF_MATCH_MISMATCH) + 7;
    if (!s->streams_valid || s->streams_reg_type != reg_type || s->streams_cov_sig != sig) {
        nb_free_streams(s);
        struct Seg { const int32_t *Ri, *NN; int64_t n; double w, cval; };
        std::vector<Seg> segs;
        nb_scope seg_scope;                                // the split pair lists live until the streams are built
        const int64_t M = s->M;
        const char *nf = getenv("NPDRB_NO_FOLD");
        // (RAW = uniReg: the row is x[Ri], y[Ri] -- not symmetric in (i, j), so mutual pairs must not be folded)
        if (!y_raw && !(nf && nf[0] == '1') && M >= 2) {
            // classify: one-directional (weight 1) / mutual with i<j (weight 2) / mutual reverse (dropped)
            const size_t words = ((size_t)N * (size_t)N + 31) / 32;
            unsigned *d_bits = nullptr; unsigned char *d_f1 = nullptr, *d_f2 = nullptr; unsigned long long *d_cnt = nullptr;
            int32_t *d_sel = nullptr; int *d_nsel = nullptr; void *d_tmp = nullptr;
            nb_scope fold_scope;
            NB_CHECK(fold_scope.alloc(&d_bits, words)); NB_CHECK(fold_scope.alloc(&d_f1, (size_t)M)); NB_CHECK(fold_scope.alloc(&d_f2, (size_t)M));
            NB_CHECK(fold_scope.alloc(&d_cnt, 1)); NB_CHECK(seg_scope.alloc(&d_sel, (size_t)4 * M)); NB_CHECK(fold_scope.alloc(&d_nsel, 4));
            NB_CUDA(cudaMemsetAsync(d_bits, 0, sizeof(unsigned) * words, ctx->stream));
            NB_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), ctx->stream));
            const unsigned grid = (unsigned)nb_cdiv(M, 256);
            mark_bits_kernel<<<grid, 256, 0, ctx->stream>>>(s->dRi, s->dNN, M, N, d_bits);
            NB_LAUNCH_CHECK(ctx);
            popcount_kernel<<<1024, 256, 0, ctx->stream>>>(d_bits, words, d_cnt);
            NB_LAUNCH_CHECK(ctx);
            classify_pairs_kernel<<<grid, 256, 0, ctx->stream>>>(s->dRi, s->dNN, M, N, d_bits, d_f1, d_f2);
            NB_LAUNCH_CHECK(ctx);
            size_t tmp_bytes = 0;
            NB_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, s->dRi, d_f1, d_sel, d_nsel, (int)M, ctx->stream));
            NB_CHECK(fold_scope.alloc((unsigned char **)&d_tmp, tmp_bytes ? tmp_bytes : 1));
            const int32_t *src[4] = {s->dRi, s->dNN, s->dRi, s->dNN};
            const unsigned char *flg[4] = {d_f1, d_f1, d_f2, d_f2};
            for (int k4 = 0; k4 < 4; ++k4)
                NB_CUDA(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, src[k4], flg[k4], d_sel + (size_t)k4 * M, d_nsel + k4, (int)M, ctx->stream));
            ctx->launches += 8;
            unsigned long long h_cnt = 0; int h_nsel[4] = {0, 0, 0, 0};
            NB_CUDA(cudaMemcpyAsync(&h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, ctx->stream));
            NB_CUDA(cudaMemcpyAsync(h_nsel, d_nsel, sizeof(h_nsel), cudaMemcpyDeviceToHost, ctx->stream));
            NB_CUDA(cudaStreamSynchronize(ctx->stream));
            const int64_t n1 = h_nsel[0], n2 = h_nsel[2];
            if ((int64_t)h_cnt == M && n2 > 0 && n1 + 2 * n2 == M) {       // no repeated ordered pair: folding is exact
                if (n1 > 0) segs.push_back({d_sel, d_sel + M, n1, 1.0, 0.0});
                segs.push_back({d_sel + 2 * M, d_sel + 3 * M, n2, 2.0, 0.0});
            }
        }
        if (segs.empty()) segs.push_back({s->dRi, s->dNN, M, 1.0, 0.0});
        if (cp.fold >= 0) {
            // split every part by the value of the folded covariate's diff (match-mismatch of column cp.fold over the pair)
            std::vector<Seg> parts;
            for (const Seg &g : segs) {
                unsigned char *d_c1 = nullptr, *d_c0 = nullptr; int32_t *d_out = nullptr; int *d_n = nullptr; void *d_tmp = nullptr;
                nb_scope tmp_scope;
                NB_CHECK(tmp_scope.alloc(&d_c1, (size_t)g.n)); NB_CHECK(tmp_scope.alloc(&d_c0, (size_t)g.n));
                NB_CHECK(seg_scope.alloc(&d_out, (size_t)4 * g.n)); NB_CHECK(tmp_scope.alloc(&d_n, 4));
                cov_flags_kernel<<<(unsigned)nb_cdiv(g.n, 256), 256, 0, ctx->stream>>>(g.Ri, g.NN, g.n, s->dC + (int64_t)cp.fold * N, d_c0, d_c1);
                NB_LAUNCH_CHECK(ctx);
                size_t tmp_bytes = 0;
                NB_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, g.Ri, d_c0, d_out, d_n, (int)g.n, ctx->stream));
                NB_CHECK(tmp_scope.alloc((unsigned char **)&d_tmp, tmp_bytes ? tmp_bytes : 1));
                const int32_t *src[4] = {g.Ri, g.NN, g.Ri, g.NN};
                const unsigned char *flg[4] = {d_c0, d_c0, d_c1, d_c1};
                for (int k4 = 0; k4 < 4; ++k4)
                    NB_CUDA(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, src[k4], flg[k4], d_out + (size_t)k4 * g.n, d_n + k4, (int)g.n, ctx->stream));
                ctx->launches += 4;
                int h_n[4] = {0, 0, 0, 0};
                NB_CUDA(cudaMemcpyAsync(h_n, d_n, sizeof(h_n), cudaMemcpyDeviceToHost, ctx->stream));
                NB_CUDA(cudaStreamSynchronize(ctx->stream));
                if (h_n[0] > 0) parts.push_back({d_out, d_out + g.n, h_n[0], g.w, 0.0});
                if (h_n[2] > 0) parts.push_back({d_out + 2 * g.n, d_out + 3 * g.n, h_n[2], g.w, 1.0});
            }
            segs.swap(parts);
        }
        double tot = 0;
        for (const Seg &g : segs) tot += (double)g.n;
        int64_t n_miss = 0;
        int ns = 0;
        for (const Seg &g : segs) {
            int64_t miss = 0;
            int rc = build_stream(s, g.Ri, g.NN, g.n, reg_type, cp, y_raw, g.w, g.cval, (double)g.n / tot, s->streams[ns], &miss);
            ++ns;
            s->n_streams = ns;
            if (rc) { nb_free_streams(s); return rc; }
            n_miss += (int64_t)(g.w * (double)miss + 0.5);
        }
        s->streams_reg_type = reg_type; s->streams_cov_sig = sig; s->streams_valid = true;
        if (reg_type == NPDRB_REG_BINOMIAL && (n_miss == 0 || n_miss == s->M)) {
            npdrb_set_error("binomial regression: the pair outcome has a single level (%lld of %lld pairs are misses)",
                            (long long)n_miss, (long long)s->M);
            s->streams_valid = false;
            return NPDRB_EDATA;
        }
    }
    NB_CUDA(cudaEventRecord(e1, ctx->stream));
    NB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    NB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    s->ms_prepare = ms;
    NB_CUDA(cudaEventRecord(e0, ctx->stream));

    const bool lm = reg_type == NPDRB_REG_LM;
    const int mode_na = nacc_for(lm ? MODE_LM : MODE_IRLS_FULL, qe);
    double *d_part = nullptr, *d_beta = nullptr, *d_se = nullptr, *d_devold = nullptr, *d_stats = nullptr;
    unsigned char *d_active = nullptr, *d_tile_active = nullptr;
    int *d_iters = nullptr, *d_flags = nullptr;
    unsigned int *d_nactive = nullptr;
    int64_t part_stride[NB_MAX_STREAMS] = {0, 0, 0, 0}, part_total = 0;
    for (int i = 0; i < s->n_streams; ++i) {
        part_stride[i] = (int64_t)s->streams[i].n_chunks * mode_na * nattr_pad;
        part_total += part_stride[i];
    }
    nb_scope scope;                                     // per-call buffers: released on every return path
    NB_CHECK(scope.alloc(&d_part, (size_t)part_total));
    NB_CHECK(scope.alloc(&d_stats, (size_t)nattr * 8));
    NB_CHECK(scope.alloc(&d_beta, (size_t)KMAX * nattr_pad)); NB_CHECK(scope.alloc(&d_se, (size_t)KMAX * nattr_pad));
    double *d_beta_prev = nullptr, *d_dev_exact = nullptr; unsigned char *d_ambig = nullptr; unsigned int *d_nambig = nullptr;
    if (!lm) {
        NB_CHECK(scope.alloc(&d_beta_prev, (size_t)KMAX * nattr_pad)); NB_CHECK(scope.alloc(&d_dev_exact, (size_t)2 * nattr_pad));
        NB_CHECK(scope.alloc(&d_ambig, (size_t)nattr_pad)); NB_CHECK(scope.alloc(&d_nambig, 1));
        NB_CUDA(cudaMemsetAsync(d_beta_prev, 0, sizeof(double) * KMAX * nattr_pad, ctx->stream));
        NB_CUDA(cudaMemsetAsync(d_ambig, 0, (size_t)nattr_pad, ctx->stream));
        NB_CUDA(cudaMemsetAsync(d_nambig, 0, sizeof(unsigned int), ctx->stream));
    }
    double *d_beta_eff = nullptr;                       // per stream: the reduced design's coefficients (covariate folding)
    if (cp.fold >= 0 && !lm) NB_CHECK(scope.alloc(&d_beta_eff, (size_t)NB_MAX_STREAMS * KMAX * nattr_pad));
    double *d_hitd = nullptr, *d_missd = nullptr;
    NB_CHECK(scope.alloc(&d_hitd, (size_t)nattr_pad)); NB_CHECK(scope.alloc(&d_missd, (size_t)nattr_pad));
    double *d_dmax = nullptr;
    NB_CHECK(scope.alloc(&d_dmax, (size_t)nattr_pad));
    NB_CUDA(cudaMemsetAsync(d_dmax, 0, sizeof(double) * nattr_pad, ctx->stream));
    if (!lm) {
        column_dmax_kernel<<<(unsigned)nattr, 256, 0, ctx->stream>>>(s->dX, s->ldx, N, attr0, nattr, diff_type, d_dmax);
        NB_LAUNCH_CHECK(ctx);
    }
    NB_CHECK(scope.alloc(&d_devold, (size_t)nattr_pad)); NB_CHECK(scope.alloc(&d_active, (size_t)nattr_pad));
    NB_CHECK(scope.alloc(&d_tile_active, (size_t)n_at)); NB_CHECK(scope.alloc(&d_iters, (size_t)nattr_pad));
    NB_CHECK(scope.alloc(&d_flags, (size_t)nattr_pad)); NB_CHECK(scope.alloc(&d_nactive, 1));
    NB_CUDA(cudaMemsetAsync(d_beta, 0, sizeof(double) * KMAX * nattr_pad, ctx->stream));
    NB_CUDA(cudaMemsetAsync(d_se, 0, sizeof(double) * KMAX * nattr_pad, ctx->stream));
    NB_CUDA(cudaMemsetAsync(d_active, 1, (size_t)nattr_pad, ctx->stream));
    NB_CUDA(cudaMemsetAsync(d_iters, 0, sizeof(int) * nattr_pad, ctx->stream));

    const nb_tables *d_tab = nullptr;
    const nb_tables1024 *d_tab1024 = nullptr;
    NB_CHECK(get_tables(ctx, &d_tab));
    NB_CHECK(get_tables1024(ctx, &d_tab1024));

    SolveArgs S;
    memset(&S, 0, sizeof(S));
    S.partials = d_part; S.nattr = (int)nattr; S.nattr_pad = (int)nattr_pad;
    S.hit_d = d_hitd; S.miss_d = d_missd; S.dmax = d_dmax;
    {
        const char *ns = getenv("NPDRB_NO_SPECULATE");
        S.speculate = (ns && ns[0] == '1') ? 0 : 1;
        if (!lm && q > 0) {                                  // covariate diff bounds from the column ranges (N x q is tiny)
            std::vector<double> hc((size_t)N * q);
            NB_CUDA(cudaMemcpyAsync(hc.data(), s->dC, sizeof(double) * hc.size(), cudaMemcpyDeviceToHost, ctx->stream));
            NB_CUDA(cudaStreamSynchronize(ctx->stream));
            for (int c = 0; c < q; ++c) {
                double mn = hc[(size_t)c * N], mx = mn;
                for (int64_t i = 1; i < N; ++i) { const double v = hc[(size_t)c * N + i]; if (v < mn) mn = v; if (v > mx) mx = v; }
                const int ct = cov_type ? cov_type[c] : NPDRB_DIFF_MATCH_MISMATCH;
                const double range = mx - mn;
                S.cmax[c] = y_raw ? fmax(fabs(mx), fabs(mn))
                          : ct == NPDRB_DIFF_NUMERIC_SQR ? range * range
                          : ct == NPDRB_DIFF_ALLELE_SHARING ? 0.5 * range
                          : ct == NPDRB_DIFF_MATCH_MISMATCH ? 1.0 : range;
            }
        }
    }
    S.beta = d_beta; S.se = d_se; S.dev_old = d_devold; S.active = d_active; S.tile_active = d_tile_active;
    S.iters = d_iters; S.flags = d_flags; S.stats = d_stats; S.n_active = d_nactive;
    {
        const char *ne = getenv("NPDRB_NO_EXACT_STOP");        // tests: decide from the streamed deviance alone
        const bool exact_stop = !(ne && ne[0] == '1');
        S.beta_prev = d_beta_prev; S.ambig = exact_stop ? d_ambig : nullptr; S.n_ambig = d_nambig; S.dev_exact = d_dev_exact; S.resolve = 0;
    }
    double Mw = 0;
    S.fold = cp.fold;
    for (int i = 0; i < NB_MAX_STREAMS; ++i) {
        S.n_chunks[i] = i < s->n_streams ? s->streams[i].n_chunks : 0;
        S.weight[i] = i < s->n_streams ? s->streams[i].weight : 0.0;
        S.fold_val[i] = i < s->n_streams ? s->streams[i].fold_val : 0.0;
        S.part_stride[i] = part_stride[i];
        if (i < s->n_streams) Mw += s->streams[i].weight * (double)s->streams[i].n_rec;
    }
    S.M = Mw;

    auto run_pass = [&](int mode, const unsigned char *tile_active) -> int {
        PassArgs P;
        memset(&P, 0, sizeof(P));
        P.XT = s->dXT; P.ldt = s->ldt; P.N = N; P.nattr = (int)nattr; P.nattr_pad = (int)nattr_pad;
        P.tile_active = tile_active; P.tables = d_tab; P.tables1024 = d_tab1024; P.diff_type = diff_type;
        P.dmax = d_dmax;
        { const char *dbg = getenv("NPDRB_V3_DBG"); P.dbg = dbg ? atoi(dbg) : 0; }
        for (int c = 0; c < qe; ++c) P.cmax[c] = S.cmax[cp.src[c]];
        double *pp = d_part;
        int total_chunks = 0;
        for (int i = 0; i < s->n_streams; ++i) {
            const nb_stream &st = s->streams[i];
            StreamArgs &A = P.st[i];
            A.rec = st.rec; A.yv = st.yv;
            for (int c = 0; c < NB_FAST_COVARS; ++c) A.cd[c] = st.cd[c];
            A.tile_ptr = st.tile_ptr; A.tile_mid = st.tile_mid; A.tile_ib = st.tile_ib; A.tile_jb = st.tile_jb;
            A.chunk_tile = st.chunk_tile; A.partials = pp;
            A.beta = d_beta;
            if (d_beta_eff && mode == MODE_IRLS_FULL) {       // this stream's view of the coefficients: [b0 + c b_f, b1, the rest]
                double *be = d_beta_eff + (size_t)i * KMAX * nattr_pad;
                fold_beta_kernel<<<(unsigned)nb_cdiv(nattr_pad, 256), 256, 0, ctx->stream>>>(d_beta, (int)nattr_pad, 2 + q, cp.fold,
                                                                                            st.fold_val, be);
                NB_LAUNCH_CHECK(ctx);
                A.beta = be;
            }
            pp += part_stride[i];
            P.chunk_start[i] = total_chunks;
            total_chunks += st.n_chunks;
        }
        for (int i = s->n_streams; i <= NB_MAX_STREAMS; ++i) P.chunk_start[i] = total_chunks;
        if (use_v2(qe)) {
            return mode == MODE_LM ? launch_pass2<MODE_LM>(ctx, qe, P, (int)n_at, total_chunks)
                 : mode == MODE_SIGN ? launch_pass2<MODE_SIGN>(ctx, qe, P, (int)n_at, total_chunks)
                                     : launch_pass2<MODE_IRLS_FULL>(ctx, qe, P, (int)n_at, total_chunks);
        }
        return mode == MODE_LM ? launch_pass<MODE_LM>(ctx, qe, P, (int)n_at, total_chunks)
             : mode == MODE_SIGN ? launch_pass<MODE_SIGN>(ctx, qe, P, (int)n_at, total_chunks)
                                 : launch_pass<MODE_IRLS_FULL>(ctx, qe, P, (int)n_at, total_chunks);
    };

    int passes = 0;
    struct EventList : std::vector<cudaEvent_t> { ~EventList() { for (cudaEvent_t e : *this) cudaEventDestroy(e); } } pev;
    auto mark = [&]() -> int {
        cudaEvent_t e;
        NB_CUDA(cudaEventCreate(&e));
        NB_CUDA(cudaEventRecord(e, ctx->stream));
        pev.push_back(e);
        return NPDRB_OK;
    };
    std::vector<unsigned char> h_tile_active((size_t)n_at, 1);
    std::vector<double> pass_evals;
    double rec_total = 0;
    for (int i = 0; i < s->n_streams; ++i) rec_total += (double)s->streams[i].n_rec;
    {
        // attribute-independent moments of the record stream(s): M, S(y), S(y^2), S(c), S(c y), S(c c') with y the
        // outcome diff (lm) or the sign s = +-1 of the hit/miss bit (binomial first pass); weights applied on the host
        const int nblk = 128;
        double *d_sh = nullptr;
        NB_CHECK(scope.alloc(&d_sh, (size_t)nblk * NSHARED_MAX));
        long double tot[NSHARED_MAX] = {0};
        std::vector<double> h((size_t)nblk * NSHARED_MAX);
        const int ns = 3 + 2 * q + q * (q + 1) / 2;
        for (int i = 0; i < s->n_streams; ++i) {
            SharedArgs a;
            a.yv = lm ? s->streams[i].yv : nullptr; a.rec = s->streams[i].rec;
            a.n = s->streams[i].n_rec; a.q = q; a.out = d_sh; a.fold_val = s->streams[i].fold_val;
            for (int c = 0; c < NB_FAST_COVARS; ++c) a.cd[c] = nullptr;      // full covariate order; NULL = the folded (constant) one
            for (int u = 0; u < qe; ++u) a.cd[cp.src[u]] = s->streams[i].cd[u];
            NB_CUDA(cudaMemsetAsync(d_sh, 0, sizeof(double) * h.size(), ctx->stream));
            shared_moments_kernel<<<nblk, 256, 0, ctx->stream>>>(a);
            NB_LAUNCH_CHECK(ctx);
            NB_CUDA(cudaMemcpyAsync(h.data(), d_sh, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
            NB_CUDA(cudaStreamSynchronize(ctx->stream));
            for (int b = 0; b < nblk; ++b)
                for (int k = 0; k < ns; ++k) tot[k] += (long double)s->streams[i].weight * h[(size_t)b * NSHARED_MAX + k];
        }
        for (int k = 0; k < NSHARED_MAX; ++k) S.shared[k] = (double)tot[k];
    }
    if (lm) {
        NB_CHECK(mark());
        NB_CHECK(run_pass(MODE_LM, nullptr));
        NB_CHECK(mark());
        NB_CHECK(launch_solve(ctx, q, SOLVE_LM, S));
        passes = 1;
    } else {
        // pass 0: closed-form weights at mustart -> plain moments; then one full pass per IRLS iteration
        NB_CUDA(cudaMemsetAsync(d_tile_active, 0, (size_t)n_at, ctx->stream));
        NB_CUDA(cudaMemsetAsync(d_nactive, 0, sizeof(unsigned int), ctx->stream));
        NB_CHECK(mark());
        NB_CHECK(run_pass(MODE_SIGN, nullptr));
        NB_CHECK(mark());
        S.pass = 0;
        NB_CHECK(launch_solve(ctx, q, SOLVE_IRLS_INIT, S));
        for (int pass = 1; pass <= 25; ++pass) {
            unsigned int h_active = 0;
            NB_CUDA(cudaMemcpyAsync(&h_active, d_nactive, sizeof(h_active), cudaMemcpyDeviceToHost, ctx->stream));
            NB_CUDA(cudaMemcpyAsync(h_tile_active.data(), d_tile_active, (size_t)n_at, cudaMemcpyDeviceToHost, ctx->stream));
            NB_CUDA(cudaStreamSynchronize(ctx->stream));
            if (h_active == 0) break;
            int act = 0;
            for (int64_t t = 0; t < n_at; ++t) act += h_tile_active[(size_t)t] ? 1 : 0;
            pass_evals.push_back((double)act * TA * rec_total);
            NB_CHECK(mark());
            NB_CHECK(run_pass(MODE_IRLS_FULL, d_tile_active));
            NB_CHECK(mark());
            passes++;
            // the step kernel rebuilds the active flags; the pass above has consumed the old ones (stream order)
            NB_CUDA(cudaMemsetAsync(d_tile_active, 0, (size_t)n_at, ctx->stream));
            NB_CUDA(cudaMemsetAsync(d_nactive, 0, sizeof(unsigned int), ctx->stream));
            S.pass = pass;
            NB_CHECK(launch_solve(ctx, q, SOLVE_IRLS_STEP, S));
            if (S.ambig) {
                unsigned int h_ambig = 0;
                NB_CUDA(cudaMemcpyAsync(&h_ambig, d_nambig, sizeof(h_ambig), cudaMemcpyDeviceToHost, ctx->stream));
                NB_CUDA(cudaStreamSynchronize(ctx->stream));
                if (h_ambig > 0) {
                    // stopping test inside the streamed deviance's error band: exact deviances, then a second visit
                    ExactArgs E;
                    memset(&E, 0, sizeof(E));
                    E.X = s->dX; E.ldx = s->ldx; E.attr0 = attr0; E.nattr = (int)nattr; E.nattr_pad = (int)nattr_pad; E.N = N;
                    E.Ri = s->dRi; E.NN = s->dNN; E.M = s->M; E.y = s->dy; E.C = s->dC; E.q = q; E.diff_type = diff_type; E.y_raw = y_raw;
                    for (int c = 0; c < q; ++c) E.cov_type[c] = cov_type ? cov_type[c] : NPDRB_DIFF_MATCH_MISMATCH;
                    E.beta = d_beta; E.beta_prev = d_beta_prev; E.ambig = d_ambig; E.dev_exact = d_dev_exact;
                    int n_split = (int)(s->M >> 20);             // ~1 M pairs per CTA
                    n_split = n_split < 1 ? 1 : (n_split > 64 ? 64 : n_split);
                    { const char *xs = getenv("NPDRB_EXACT_SPLIT"); if (xs && atoi(xs) >= 1 && atoi(xs) <= 1024) n_split = atoi(xs); }   // tests
                    double *d_xpart = nullptr;
                    NB_CHECK(scope.alloc(&d_xpart, (size_t)nattr * n_split * 2));
                    E.part = d_xpart; E.n_split = n_split;
                    exact_dev_kernel<<<dim3((unsigned)nattr, (unsigned)n_split), 256, 0, ctx->stream>>>(E);
                    NB_LAUNCH_CHECK(ctx);
                    exact_dev_reduce_kernel<<<(unsigned)nb_cdiv(nattr, 256), 256, 0, ctx->stream>>>(E);
                    NB_LAUNCH_CHECK(ctx);
                    S.resolve = 1;
                    NB_CHECK(launch_solve(ctx, q, SOLVE_IRLS_STEP, S));
                    S.resolve = 0;
                    NB_CUDA(cudaMemsetAsync(d_nambig, 0, sizeof(unsigned int), ctx->stream));
                    s->n_exact_stops += (int)h_ambig;
                }
            }
        }
    }
    NB_CUDA(cudaMemcpyAsync(stats_out_host, d_stats, sizeof(double) * (size_t)nattr * 8, cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaEventRecord(e1, ctx->stream));
    NB_CUDA(cudaEventSynchronize(e1));
    NB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    s->ms_regress = ms;
    s->ms_pass_first = 0; s->ms_pass_full = 0; s->n_pass_full = 0; s->evals_pass_full = 0;
    for (size_t e = 0; e + 1 < pev.size(); e += 2) {
        float pm = 0;
        cudaEventElapsedTime(&pm, pev[e], pev[e + 1]);
        if (e == 0) s->ms_pass_first = pm;
        else { s->ms_pass_full += pm; s->n_pass_full++; }
    }
    for (double v : pass_evals) s->evals_pass_full += v;
    if (irls_passes) *irls_passes = passes;
    return NPDRB_OK;
}
