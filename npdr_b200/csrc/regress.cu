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
#include <math.h>
#include "common.cuh"
#include "dmath.cuh"

namespace {

constexpr int BI = 64, BJ = 64;        // instance block sizes (tile = BI x BJ)
constexpr int TA = 128;                // attributes per CTA
constexpr int NSLICE = 4;              // pair slices per attribute
constexpr int RTHREADS = TA * NSLICE;  // 512
constexpr int RCHUNK = 1024;           // records staged per step
constexpr int KMAX = 2 + NPDRB_MAX_COVARS;

enum { MODE_LM = 0, MODE_IRLS_INIT = 1, MODE_IRLS_FULL = 2 };

__host__ __device__ constexpr int nacc_for(int mode, int q) {
    return mode == MODE_LM ? (3 + q) : ((2 + q) * (3 + q) / 2 + (2 + q) + 1);
}

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

__global__ void tile_keys_kernel(const int32_t *__restrict__ Ri, const int32_t *__restrict__ NN, int64_t M, int n_jb,
                                 uint32_t *__restrict__ keys, uint32_t *__restrict__ idx) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    keys[m] = (uint32_t)((Ri[m] - 1) / BI) * (uint32_t)n_jb + (uint32_t)((NN[m] - 1) / BJ);
    idx[m] = (uint32_t)m;
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
    int q, reg_type;
    int cov_type[NPDRB_MAX_COVARS];
    int y_raw;                 // uniReg: outcome of instance Ri itself, no diff
    uint32_t *rec;
    double *yv;
    double *cd[NPDRB_MAX_COVARS];
    unsigned long long *n_miss;
};
__global__ void gather_records_kernel(GatherArgs g) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int miss = 0;
    if (t < g.M) {
        const int64_t m = g.perm[t];
        const int64_t i = g.Ri[m] - 1, j = g.NN[m] - 1;
        uint32_t r = (uint32_t)(i % BI) | ((uint32_t)(j % BJ) << 8);
        const double yi = g.y[i], yj = g.y[j];
        if (g.reg_type == NPDRB_REG_BINOMIAL) {
            miss = g.y_raw ? (yi != 0.0) : (yi != yj);     // match-mismatch: hit = 0, miss = 1
            r |= (uint32_t)miss << 16;
        } else {
            g.yv[t] = g.y_raw ? yi : fabs(yi - yj);         // always numeric-abs (R/npdr.R:303)
        }
        g.rec[t] = r;
        for (int c = 0; c < g.q; ++c) {
            const double *col = g.C + (int64_t)c * g.N;
            g.cd[c][t] = g.y_raw ? col[i] : nb_diff1(col[i], col[j], g.cov_type[c]);
        }
    }
    unsigned mask = __ballot_sync(0xffffffffu, miss);
    if ((threadIdx.x & 31) == 0 && mask) atomicAdd(g.n_miss, (unsigned long long)__popc(mask));
}

// shared (attribute-independent) lm moments of the record stream: M, Sy, Syy, Sc_r, Scy_r, Scc_rs
constexpr int NSHARED_MAX = 3 + 2 * NPDRB_MAX_COVARS + NPDRB_MAX_COVARS * (NPDRB_MAX_COVARS + 1) / 2;
struct SharedArgs { const double *yv; const double *cd[NPDRB_MAX_COVARS]; int64_t n; int q; double *out; };
__global__ void __launch_bounds__(256)
shared_moments_kernel(SharedArgs a) {
    __shared__ double red[256];
    double acc[NSHARED_MAX];
    const int ns = 3 + 2 * a.q + a.q * (a.q + 1) / 2;
    for (int s = 0; s < NSHARED_MAX; ++s) acc[s] = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < a.n; t += (int64_t)gridDim.x * 256) {
        const double y = a.yv[t];
        double c[NPDRB_MAX_COVARS];
        for (int r = 0; r < a.q; ++r) c[r] = a.cd[r][t];
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
struct PassArgs {
    const double *XT; int64_t ldt; int64_t N; int nattr; int nattr_pad;
    const uint32_t *rec; const double *yv; const double *cd[NPDRB_MAX_COVARS];
    const int64_t *tile_ptr; const int32_t *tile_ib; const int32_t *tile_jb; const int32_t *chunk_tile;
    const double *beta;            // [KMAX][nattr_pad] current coefficients (IRLS_FULL)
    const unsigned char *tile_active;  // per attribute tile, or NULL
    double *partials;              // [n_chunks][NACC][nattr_pad]
    const nb_tables *tables;       // device copy of the exp/log tables
    int diff_type;
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
    constexpr int NA = (MODE == MODE_LM) ? (3 + Q) : (K * (K + 1) / 2 + K + 1);
    extern __shared__ unsigned char smem_raw[];
    double *sXI = (double *)smem_raw;                   // [BI][TA]
    double *sXJ = sXI + BI * TA;                        // [BJ][TA]
    double *sCD = sXJ + BJ * TA;                        // [max(Q,1)][RCHUNK]  (lm: slot Q holds yv)
    uint32_t *sREC = (uint32_t *)(sCD + (Q + 1) * RCHUNK);
    nb_tables *sT = (nb_tables *)(sREC + RCHUNK);

    const int tid = threadIdx.x;
    const int al = tid % TA, slice = tid / TA;
    const int at = blockIdx.x, chunk = blockIdx.y;
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
        for (int r = 0; r < K; ++r) beta[r] = P.beta[(int64_t)r * P.nattr_pad + a0 + al];
    }
    double acc[NA];
#pragma unroll
    for (int n = 0; n < NA; ++n) acc[n] = 0.0;

    const int t_begin = P.chunk_tile[chunk], t_end = P.chunk_tile[chunk + 1];
    int cur_ib = -1;
    for (int t = t_begin; t < t_end; ++t) {
        const int ib = P.tile_ib[t], jb = P.tile_jb[t];
        const int64_t r0 = P.tile_ptr[t], r1 = P.tile_ptr[t + 1];
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
                sREC[e] = P.rec[rc + e];
#pragma unroll
                for (int c = 0; c < Q; ++c) sCD[c * RCHUNK + e] = P.cd[c][rc + e];
                if (MODE == MODE_LM) sCD[Q * RCHUNK + e] = P.yv[rc + e];
            }
            __syncthreads();
#pragma unroll 2
            for (int m = slice; m < n; m += NSLICE) {
                const uint32_t r = sREC[m];
                const double xi = sXI[(r & 0xFFu) * TA + al];
                const double xj = sXJ[((r >> 8) & 0xFFu) * TA + al];
                const double d = proj_diff<DIFF>(xi, xj, P.diff_type);
                double c[Q > 0 ? Q : 1];
#pragma unroll
                for (int u = 0; u < Q; ++u) c[u] = sCD[u * RCHUNK + m];
                if (MODE == MODE_LM) {
                    const double yv = sCD[Q * RCHUNK + m];
                    acc[0] += d;
                    acc[1] = fma(d, d, acc[1]);
                    acc[2] = fma(d, yv, acc[2]);
#pragma unroll
                    for (int u = 0; u < Q; ++u) acc[3 + u] = fma(d, c[u], acc[3 + u]);
                } else {
                    const bool yb = (r >> 16) & 1u;
                    double w, wz;
                    if (MODE == MODE_IRLS_INIT) {
                        // mustart = (y + 0.5)/2 -> mu = 0.75 / 0.25, eta = +-log 3, w = mu(1-mu) = 0.1875,
                        // z = eta + (y - mu)/w  (base-R glm.fit + binomial()$initialize)
                        const double z0 = 1.0986122886681098 + 0.25 / 0.1875;
                        w = 0.1875;
                        wz = yb ? (0.1875 * z0) : -(0.1875 * z0);
                    } else {
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
                        const double yd = yb ? 1.0 : 0.0;
                        acc[NA - 1] += fma(-yd, eta_dev, lg);
                        wz = fma(w, eta, yd - mu);          // w*z with z = eta + (y-mu)/mu.eta
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
        P.partials[((int64_t)chunk * NA + n) * P.nattr_pad + a0 + a] = s;
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
    // diag of inverse: columns of L^-1
    for (int c = 0; c < K; ++c) dinv[c] = 0.0;
    for (int c = 0; c < K; ++c) {
        if (aliased[c]) continue;
        double col[K];                                     // L^-1 e_c
        for (int i = 0; i < K; ++i) col[i] = 0.0;
        for (int i = c; i < K; ++i) {
            if (aliased[i]) continue;
            double u = (i == c) ? 1.0 : 0.0;
            for (int t = c; t < i; ++t) if (!aliased[t]) u -= L[i][t] * col[t];
            col[i] = u / L[i][i];
        }
        // (A^-1)_{rr} = sum_i (L^-1)_{i r}^2 ; accumulate column c's contribution to row index c only needs
        // all rows i >= c of column c: (A^-1)_{cc} = sum_{i>=c} col[i]^2
        double d = 0.0;
        for (int i = c; i < K; ++i) d += col[i] * col[i];
        dinv[c] = d * s[c] * s[c];
    }
}

struct SolveArgs {
    const double *partials; int n_chunks[2]; double weight[2]; int64_t part_stride[2];  // up to two streams
    int nattr, nattr_pad;
    double *beta, *se;             // [KMAX][nattr_pad]
    double *dev_old;               // [nattr_pad]
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

// sum the per-chunk partials of statistic n for attribute a, fixed order, weights applied per stream
template <int NA>
__device__ __forceinline__ void reduce_partials(const SolveArgs &S, int a, double (&out)[NA]) {
    for (int n = 0; n < NA; ++n) out[n] = 0.0;
    const double *base = S.partials;
    for (int st = 0; st < 2; ++st) {
        if (S.n_chunks[st] == 0) continue;
        double tmp[NA];
        for (int n = 0; n < NA; ++n) tmp[n] = 0.0;
        for (int c = 0; c < S.n_chunks[st]; ++c)
            for (int n = 0; n < NA; ++n) tmp[n] += base[((int64_t)c * NA + n) * S.nattr_pad + a];
        for (int n = 0; n < NA; ++n) out[n] += S.weight[st] * tmp[n];
        base += S.part_stride[st];
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
    reduce_partials<NA>(S, a, m);
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

// binomial: one IRLS bookkeeping step per attribute (glm.fit's loop body)
template <int Q>
__global__ void irls_step_kernel(SolveArgs S) {
    constexpr int K = 2 + Q, NA = K * (K + 1) / 2 + K + 1;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= S.nattr) return;
    if (!S.active[a]) return;
    double m[NA];
    reduce_partials<NA>(S, a, m);
    double A[K][K], b[K];
    int n = 0;
    for (int r = 0; r < K; ++r) for (int c = r; c < K; ++c) { A[r][c] = A[c][r] = m[n]; ++n; }
    for (int r = 0; r < K; ++r) b[r] = m[n + r];
    const double dev = 2.0 * m[NA - 1];
    bool finish = false;
    int flags = 0;
    if (S.pass == 0) {
        S.dev_old[a] = S.M * (2.0 * 0.2876820724517809);      // sum of dev.resids at mustart: 2*log(4/3) each
    } else {
        const double devold = S.dev_old[a];
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
    for (int r = 0; r < K; ++r) {
        S.beta[(int64_t)r * S.nattr_pad + a] = x[r];
        S.se[(int64_t)r * S.nattr_pad + a] = al[r] ? nanv : sqrt(dinv[r]);
    }
    S.iters[a] = S.pass + 1;
    S.tile_active[a / TA] = 1;
    atomicAdd(S.n_active, 1u);
}

// ------------------------------------------------------------------ host side
template <typename T>
int dmalloc(T **p, size_t n) {
    cudaError_t e = cudaMalloc((void **)p, sizeof(T) * (n ? n : 1));
    if (e != cudaSuccess) { npdrb_set_error("cudaMalloc of %zu bytes failed: %s", sizeof(T) * n, cudaGetErrorString(e)); return NPDRB_ENOMEM; }
    return NPDRB_OK;
}

nb_tables *g_tables_dev[64] = {nullptr};

int get_tables(npdrb_ctx *ctx, const nb_tables **out) {
    if (!g_tables_dev[ctx->device & 63]) {
        static const nb_tables host_tables = NB_TABLES_STATIC_INIT;
        nb_tables *d = nullptr;
        NB_CHECK(dmalloc(&d, 1));
        NB_CUDA(cudaMemcpy(d, &host_tables, sizeof(nb_tables), cudaMemcpyHostToDevice));
        g_tables_dev[ctx->device & 63] = d;
    }
    *out = g_tables_dev[ctx->device & 63];
    return NPDRB_OK;
}

void free_stream(nb_stream &st) {
    cudaFree(st.rec); cudaFree(st.yv);
    for (int c = 0; c < NPDRB_MAX_COVARS; ++c) cudaFree(st.cd[c]);
    cudaFree(st.tile_ptr); cudaFree(st.tile_ib); cudaFree(st.tile_jb); cudaFree(st.chunk_tile);
    st = nb_stream();
}

// Build one bucketed stream from a pair list (device).  y_raw: uniReg mode.
int build_stream(npdrb_session *s, const int32_t *dRi, const int32_t *dNN, int64_t M, int reg_type,
                 const int32_t *cov_type, int y_raw, double weight, nb_stream &st, int64_t *n_miss_out) {
    npdrb_ctx *ctx = s->ctx;
    const int64_t N = s->N;
    const int q = s->q;
    const int n_ib = (int)nb_cdiv(N, BI), n_jb = (int)nb_cdiv(N, BJ);
    const int64_t n_tiles_all = (int64_t)n_ib * n_jb;
    st = nb_stream();
    st.n_rec = M; st.weight = weight;
    uint32_t *keys = nullptr, *idx = nullptr, *keys2 = nullptr, *perm = nullptr;
    unsigned *hist = nullptr;
    unsigned long long *d_miss = nullptr;
    void *tmp = nullptr;
    NB_CHECK(dmalloc(&keys, (size_t)M)); NB_CHECK(dmalloc(&idx, (size_t)M));
    NB_CHECK(dmalloc(&keys2, (size_t)M)); NB_CHECK(dmalloc(&perm, (size_t)M));
    NB_CHECK(dmalloc(&hist, (size_t)n_tiles_all)); NB_CHECK(dmalloc(&d_miss, 1));
    NB_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned) * (size_t)n_tiles_all, ctx->stream));
    NB_CUDA(cudaMemsetAsync(d_miss, 0, sizeof(unsigned long long), ctx->stream));
    const unsigned grid = (unsigned)nb_cdiv(M, 256);
    tile_keys_kernel<<<grid, 256, 0, ctx->stream>>>(dRi, dNN, M, n_jb, keys, idx);
    NB_LAUNCH_CHECK(ctx);
    int end_bit = 1;
    while ((1ll << end_bit) < n_tiles_all && end_bit < 32) ++end_bit;
    size_t tmp_bytes = 0;
    NB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, idx, perm, (int)M, 0, end_bit, ctx->stream));
    NB_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    NB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, idx, perm, (int)M, 0, end_bit, ctx->stream));
    ctx->launches += 3;
    tile_hist_kernel<<<grid, 256, 0, ctx->stream>>>(keys2, M, hist);
    NB_LAUNCH_CHECK(ctx);

    NB_CHECK(dmalloc(&st.rec, (size_t)M));
    if (reg_type == NPDRB_REG_LM) NB_CHECK(dmalloc(&st.yv, (size_t)M));
    for (int c = 0; c < q; ++c) NB_CHECK(dmalloc(&st.cd[c], (size_t)M));
    GatherArgs g;
    g.Ri = dRi; g.NN = dNN; g.perm = perm; g.y = s->dy; g.C = s->dC; g.M = M; g.N = N; g.q = q; g.reg_type = reg_type;
    for (int c = 0; c < NPDRB_MAX_COVARS; ++c) { g.cov_type[c] = (c < q && cov_type) ? cov_type[c] : NPDRB_DIFF_MATCH_MISMATCH; g.cd[c] = st.cd[c]; }
    g.y_raw = y_raw; g.rec = st.rec; g.yv = st.yv; g.n_miss = d_miss;
    gather_records_kernel<<<grid, 256, 0, ctx->stream>>>(g);
    NB_LAUNCH_CHECK(ctx);

    std::vector<unsigned> h_hist((size_t)n_tiles_all);
    unsigned long long h_miss = 0;
    NB_CUDA(cudaMemcpyAsync(h_hist.data(), hist, sizeof(unsigned) * h_hist.size(), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaMemcpyAsync(&h_miss, d_miss, sizeof(h_miss), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n_miss_out) *n_miss_out = (int64_t)h_miss;

    std::vector<int64_t> tptr; std::vector<int32_t> tib, tjb;
    int64_t run = 0;
    for (int64_t t = 0; t < n_tiles_all; ++t) {
        if (!h_hist[t]) continue;
        tptr.push_back(run); tib.push_back((int32_t)(t / n_jb)); tjb.push_back((int32_t)(t % n_jb));
        run += h_hist[t];
    }
    tptr.push_back(run);
    st.n_tiles = (int32_t)tib.size();
    // work chunks: contiguous tile ranges balanced by record count
    const int64_t n_at = nb_cdiv(s->xt_nattr > 0 ? s->xt_nattr : s->p, TA);
    int64_t target = (int64_t)16 * ctx->sm_count;
    int64_t C = nb_cdiv(target, n_at);
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
    NB_CHECK(dmalloc(&st.tile_ptr, tptr.size())); NB_CHECK(dmalloc(&st.tile_ib, tib.size() + 1));
    NB_CHECK(dmalloc(&st.tile_jb, tjb.size() + 1)); NB_CHECK(dmalloc(&st.chunk_tile, ctile.size()));
    NB_CUDA(cudaMemcpyAsync(st.tile_ptr, tptr.data(), sizeof(int64_t) * tptr.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (!tib.empty()) {
        NB_CUDA(cudaMemcpyAsync(st.tile_ib, tib.data(), sizeof(int32_t) * tib.size(), cudaMemcpyHostToDevice, ctx->stream));
        NB_CUDA(cudaMemcpyAsync(st.tile_jb, tjb.data(), sizeof(int32_t) * tjb.size(), cudaMemcpyHostToDevice, ctx->stream));
    }
    NB_CUDA(cudaMemcpyAsync(st.chunk_tile, ctile.data(), sizeof(int32_t) * ctile.size(), cudaMemcpyHostToDevice, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(keys); cudaFree(idx); cudaFree(keys2); cudaFree(perm); cudaFree(hist); cudaFree(d_miss); cudaFree(tmp);
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
    dim3 grid((unsigned)n_at, (unsigned)n_chunks);
    pass_kernel<MODE, Q, DIFF><<<grid, RTHREADS, smem, ctx->stream>>>(P);
    NB_LAUNCH_CHECK(ctx);
    return NPDRB_OK;
}
template <int MODE, int Q>
int launch_pass_q(npdrb_ctx *ctx, const PassArgs &P, int n_at, int n_chunks) {
    return P.diff_type == NPDRB_DIFF_NUMERIC_ABS ? launch_pass_t<MODE, Q, 0>(ctx, P, n_at, n_chunks)
                                                 : launch_pass_t<MODE, Q, 1>(ctx, P, n_at, n_chunks);
}
template <int MODE>
int launch_pass(npdrb_ctx *ctx, int q, const PassArgs &P, int n_at, int n_chunks) {
    switch (q) {
    case 0: return launch_pass_q<MODE, 0>(ctx, P, n_at, n_chunks);
    case 1: return launch_pass_q<MODE, 1>(ctx, P, n_at, n_chunks);
    case 2: return launch_pass_q<MODE, 2>(ctx, P, n_at, n_chunks);
    case 3: return launch_pass_q<MODE, 3>(ctx, P, n_at, n_chunks);
    case 4: return launch_pass_q<MODE, 4>(ctx, P, n_at, n_chunks);
    }
    npdrb_set_error("regress: at most %d covariates are supported", NPDRB_MAX_COVARS);
    return NPDRB_EINVAL;
}

int launch_solve(npdrb_ctx *ctx, int q, bool lm, const SolveArgs &S) {
    const unsigned grid = (unsigned)nb_cdiv(S.nattr, 128);
#define NB_SOLVE_CASE(QQ)                                                                  \
    case QQ:                                                                               \
        if (lm) lm_finalize_kernel<QQ><<<grid, 128, 0, ctx->stream>>>(S);                  \
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
    for (int i = 0; i < 2; ++i) free_stream(s->streams[i]);
    s->n_streams = 0;
    s->streams_valid = false;
}

int nb_regress(npdrb_session *s, int64_t attr0, int64_t nattr, int32_t reg_type, int32_t diff_type,
               const int32_t *cov_type, double *stats_out_host, int32_t *irls_passes) {
    npdrb_ctx *ctx = s->ctx;
    const int q = s->q;
    const int64_t N = s->N;
    if (!s->dX || !s->dy) { npdrb_set_error("regress: inputs not set"); return NPDRB_EINVAL; }
    if (q > 0 && !s->dC) { npdrb_set_error("regress: covariates missing"); return NPDRB_EINVAL; }
    if (!s->dRi || s->M <= 0) { npdrb_set_error("regress: empty pair list"); return NPDRB_EINVAL; }
    if (attr0 < 0 || nattr < 1 || attr0 + nattr > s->p) { npdrb_set_error("regress: bad attribute range"); return NPDRB_EINVAL; }
    if (reg_type != NPDRB_REG_LM && reg_type != NPDRB_REG_BINOMIAL) { npdrb_set_error("regress: bad regression type"); return NPDRB_EINVAL; }
    if (diff_type < 0 || diff_type > NPDRB_DIFF_RAW) { npdrb_set_error("regress: bad attr diff type"); return NPDRB_EINVAL; }
    if (s->M >= (1ll << 31)) { npdrb_set_error("regress: more than 2^31 pairs per device"); return NPDRB_EINVAL; }
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
    // bucketed pair stream(s)
    int sig = y_raw * 1000003;
    for (int c = 0; c < q; ++c) sig = sig * 31 + (cov_type ? cov_type[c] : NPDRB_DIFF_MATCH_MISMATCH) + 7;
    if (!s->streams_valid || s->streams_reg_type != reg_type || s->streams_cov_sig != sig) {
        nb_free_streams(s);
        int64_t n_miss = 0;
        NB_CHECK(build_stream(s, s->dRi, s->dNN, s->M, reg_type, cov_type, y_raw, 1.0, s->streams[0], &n_miss));
        s->n_streams = 1;
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
    const int mode_na = nacc_for(lm ? MODE_LM : MODE_IRLS_FULL, q);
    double *d_part = nullptr, *d_beta = nullptr, *d_se = nullptr, *d_devold = nullptr, *d_stats = nullptr;
    unsigned char *d_active = nullptr, *d_tile_active = nullptr;
    int *d_iters = nullptr, *d_flags = nullptr;
    unsigned int *d_nactive = nullptr;
    int64_t part_stride[2] = {0, 0}, part_total = 0;
    for (int i = 0; i < s->n_streams; ++i) {
        part_stride[i] = (int64_t)s->streams[i].n_chunks * mode_na * nattr_pad;
        part_total += part_stride[i];
    }
    NB_CHECK(dmalloc(&d_part, (size_t)part_total));
    NB_CHECK(dmalloc(&d_stats, (size_t)nattr * 8));
    NB_CHECK(dmalloc(&d_beta, (size_t)KMAX * nattr_pad)); NB_CHECK(dmalloc(&d_se, (size_t)KMAX * nattr_pad));
    NB_CHECK(dmalloc(&d_devold, (size_t)nattr_pad)); NB_CHECK(dmalloc(&d_active, (size_t)nattr_pad));
    NB_CHECK(dmalloc(&d_tile_active, (size_t)n_at)); NB_CHECK(dmalloc(&d_iters, (size_t)nattr_pad));
    NB_CHECK(dmalloc(&d_flags, (size_t)nattr_pad)); NB_CHECK(dmalloc(&d_nactive, 1));
    NB_CUDA(cudaMemsetAsync(d_beta, 0, sizeof(double) * KMAX * nattr_pad, ctx->stream));
    NB_CUDA(cudaMemsetAsync(d_se, 0, sizeof(double) * KMAX * nattr_pad, ctx->stream));
    NB_CUDA(cudaMemsetAsync(d_active, 1, (size_t)nattr_pad, ctx->stream));
    NB_CUDA(cudaMemsetAsync(d_iters, 0, sizeof(int) * nattr_pad, ctx->stream));

    const nb_tables *d_tab = nullptr;
    NB_CHECK(get_tables(ctx, &d_tab));

    SolveArgs S;
    memset(&S, 0, sizeof(S));
    S.partials = d_part; S.nattr = (int)nattr; S.nattr_pad = (int)nattr_pad;
    S.beta = d_beta; S.se = d_se; S.dev_old = d_devold; S.active = d_active; S.tile_active = d_tile_active;
    S.iters = d_iters; S.flags = d_flags; S.stats = d_stats; S.n_active = d_nactive;
    double Mw = 0;
    for (int i = 0; i < 2; ++i) {
        S.n_chunks[i] = i < s->n_streams ? s->streams[i].n_chunks : 0;
        S.weight[i] = i < s->n_streams ? s->streams[i].weight : 0.0;
        S.part_stride[i] = part_stride[i];
        if (i < s->n_streams) Mw += s->streams[i].weight * (double)s->streams[i].n_rec;
    }
    S.M = Mw;

    auto run_pass = [&](int mode, const unsigned char *tile_active) -> int {
        double *pp = d_part;
        for (int i = 0; i < s->n_streams; ++i) {
            const nb_stream &st = s->streams[i];
            PassArgs P;
            memset(&P, 0, sizeof(P));
            P.XT = s->dXT; P.ldt = s->ldt; P.N = N; P.nattr = (int)nattr; P.nattr_pad = (int)nattr_pad;
            P.rec = st.rec; P.yv = st.yv;
            for (int c = 0; c < NPDRB_MAX_COVARS; ++c) P.cd[c] = st.cd[c];
            P.tile_ptr = st.tile_ptr; P.tile_ib = st.tile_ib; P.tile_jb = st.tile_jb; P.chunk_tile = st.chunk_tile;
            P.beta = d_beta; P.tile_active = tile_active; P.partials = pp; P.tables = d_tab; P.diff_type = diff_type;
            int rc = mode == MODE_LM ? launch_pass<MODE_LM>(ctx, q, P, (int)n_at, st.n_chunks)
                   : mode == MODE_IRLS_INIT ? launch_pass<MODE_IRLS_INIT>(ctx, q, P, (int)n_at, st.n_chunks)
                                            : launch_pass<MODE_IRLS_FULL>(ctx, q, P, (int)n_at, st.n_chunks);
            if (rc) return rc;
            pp += part_stride[i];
        }
        return NPDRB_OK;
    };

    int passes = 0;
    std::vector<cudaEvent_t> pev;
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
    if (lm) {
        // shared moments over all streams (weights applied on the host)
        const int nblk = 128;
        double *d_sh = nullptr;
        NB_CHECK(dmalloc(&d_sh, (size_t)nblk * NSHARED_MAX));
        long double tot[NSHARED_MAX] = {0};
        std::vector<double> h((size_t)nblk * NSHARED_MAX);
        const int ns = 3 + 2 * q + q * (q + 1) / 2;
        for (int i = 0; i < s->n_streams; ++i) {
            SharedArgs a;
            a.yv = s->streams[i].yv; a.n = s->streams[i].n_rec; a.q = q; a.out = d_sh;
            for (int c = 0; c < NPDRB_MAX_COVARS; ++c) a.cd[c] = s->streams[i].cd[c];
            NB_CUDA(cudaMemsetAsync(d_sh, 0, sizeof(double) * h.size(), ctx->stream));
            shared_moments_kernel<<<nblk, 256, 0, ctx->stream>>>(a);
            NB_LAUNCH_CHECK(ctx);
            NB_CUDA(cudaMemcpyAsync(h.data(), d_sh, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
            NB_CUDA(cudaStreamSynchronize(ctx->stream));
            for (int b = 0; b < nblk; ++b)
                for (int k = 0; k < ns; ++k) tot[k] += (long double)s->streams[i].weight * h[(size_t)b * NSHARED_MAX + k];
        }
        cudaFree(d_sh);
        for (int k = 0; k < NSHARED_MAX; ++k) S.shared[k] = (double)tot[k];
        NB_CHECK(mark());
        NB_CHECK(run_pass(MODE_LM, nullptr));
        NB_CHECK(mark());
        NB_CHECK(launch_solve(ctx, q, true, S));
        passes = 1;
    } else {
        NB_CUDA(cudaMemsetAsync(d_tile_active, 1, (size_t)n_at, ctx->stream));
        NB_CHECK(mark());
        NB_CHECK(run_pass(MODE_IRLS_INIT, nullptr));
        NB_CHECK(mark());
        for (int pass = 0; pass <= 25; ++pass) {
            if (pass > 0) {
                int act = 0;
                for (int64_t t = 0; t < n_at; ++t) act += h_tile_active[(size_t)t] ? 1 : 0;
                pass_evals.push_back((double)act * TA * rec_total);
                NB_CHECK(mark());
                NB_CHECK(run_pass(MODE_IRLS_FULL, d_tile_active));
                NB_CHECK(mark());
                passes++;
            }
            NB_CUDA(cudaMemsetAsync(d_tile_active, 0, (size_t)n_at, ctx->stream));
            NB_CUDA(cudaMemsetAsync(d_nactive, 0, sizeof(unsigned int), ctx->stream));
            S.pass = pass;
            NB_CHECK(launch_solve(ctx, q, false, S));
            unsigned int h_active = 0;
            NB_CUDA(cudaMemcpyAsync(&h_active, d_nactive, sizeof(h_active), cudaMemcpyDeviceToHost, ctx->stream));
            NB_CUDA(cudaMemcpyAsync(h_tile_active.data(), d_tile_active, (size_t)n_at, cudaMemcpyDeviceToHost, ctx->stream));
            NB_CUDA(cudaStreamSynchronize(ctx->stream));
            if (h_active == 0) break;
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
    for (cudaEvent_t e : pev) cudaEventDestroy(e);
    if (irls_passes) *irls_passes = passes;
    cudaFree(d_part); cudaFree(d_beta); cudaFree(d_se); cudaFree(d_devold); cudaFree(d_stats);
    cudaFree(d_active); cudaFree(d_tile_active); cudaFree(d_iters); cudaFree(d_flags); cudaFree(d_nactive);
    return NPDRB_OK;
}
