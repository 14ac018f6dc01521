// regress_wide.cu -- the regression loop for designs with MANY covariates (q > NB_FAST_COVARS).
//
// R/npdr.R:315-345 puts no limit on the number of covariate columns (mdd.RNAseq.small$covs.short has 39); the streaming
// kernels of regress.cu keep the (2+q)(3+q)/2 statistics of an attribute in registers and stop at q = 4.  Here the design
// row v = [1, d, c_1 .. c_q, z] is wide, so the per-attribute statistic is a weighted Gram matrix G = sum_m W_m v_m v_m'
// of order K + 1 = q + 3 <= 64, i.e. a small rank-M update per attribute and IRLS iteration:
//   * one CTA per attribute runs the WHOLE fit of that attribute -- lm: one pass; binomial: glm.fit's loop (R/npdr.R:36;
//     mustart, eta clamps, deviance stopping rule, maxit 25, standard errors from the last weighted system) with no host
//     round trip between iterations;
//   * the pair list is walked in blocks of 128 pairs: (1) the covariate diffs of the block (pair-major, formed once per
//     call) are copied to shared memory with coalesced loads, (2) one thread per pair gathers x_i, x_j, forms d, eta, mu,
//     W, z and the deviance term, (3) every thread owns a 4 x 4 register tile of G (columns strided by KT/4, so operand
//     reads are conflict-free 8-byte loads) and adds the block's pairs; for narrow designs (KT = 8, 16, 32) the spare
//     threads form 64 / 16 / 4 groups that split the pairs of a block and are summed in a fixed order at the end of a pass;
//   * the solve is a Jacobi-scaled Cholesky of the K x K system in shared memory (same aliasing rule as chol_solve in
//     regress.cu), with L^-1 formed column-parallel for diag(G^-1) and the solution.
// FP64 work per attribute-iteration: M * KT^2 / 2 FMAs (cfg 3 with all 39 covariates: 7,536 * 64^2 / 2 = 1.5e7).
#include <cuda_runtime.h>
#include <math.h>

#include "common.cuh"

namespace {

constexpr int WB = 128;             // pairs per staged block
constexpr int WT = 256;             // threads per CTA

struct WideArgs {
    const double *X; int64_t ldx; int64_t attr0; int nattr; int64_t N;
    const int32_t *Ri, *NN; int64_t M;
    const double *yd;               // per pair: lm outcome diff; binomial 0 (hit) / 1 (miss)
    const double *CD;               // M x q covariate diffs, pair-major
    int q, diff_type, reg_type;
    double *stats;                  // nattr x 8
    int *max_passes;                // passes over the pair list of the slowest attribute
};

__device__ __forceinline__ double wide_linkinv(double eta) {          // stats family.c logit_linkinv
    const double t = (eta < -30.0) ? 2.220446049250313e-16 : ((eta > 30.0) ? 4503599627370496.0 : exp(eta));
    return t / (1.0 + t);
}
__device__ __forceinline__ double wide_mu_eta(double eta) {           // logit_mu_eta
    const double op = 1.0 + exp(eta);
    return (eta > 30.0 || eta < -30.0) ? 2.220446049250313e-16 : exp(eta) / (op * op);
}
__device__ __forceinline__ double wide_ylogy(double y, double mu) { return (y != 0.0) ? y * log(y / mu) : 0.0; }

// Shared-memory image of one CTA.  LD = KT + 1 doubles per row: odd, so a column walk (one row per lane) is conflict free.
template <int KT>
struct WideSmem {
    static constexpr int LD = KT + 1;
    static constexpr int NT = KT / 4;                    // register tiles per dimension
    static constexpr int TPG = NT * NT;                  // threads per pair group
    static constexpr int NG = WT / TPG;                  // pair groups
    static constexpr size_t V_DOUBLES = (size_t)WB * LD;
    static constexpr size_t RED_DOUBLES = (size_t)WT * 16;
    static constexpr size_t VR = V_DOUBLES > RED_DOUBLES ? V_DOUBLES : RED_DOUBLES;
    static constexpr size_t DOUBLES = VR + 3 * (size_t)KT * LD + WB /* w */ + 12 * (size_t)KT + WT /* dev */ + 16;
    static constexpr size_t BYTES = DOUBLES * sizeof(double) + 2 * KT /* alias flags */ + 16;
};

// Jacobi-scaled Cholesky of the leading K x K block of A (symmetric, full storage, leading dimension LD), all threads.
// On return: L (lower), Linv = L^-1 (lower), s = 1/sqrt(diag), al = aliased flags (pivot <= tol on the scaled system or a
// non-positive diagonal), exactly the rule of chol_solve<K> in regress.cu.
template <int LD>
__device__ void wide_chol(int K, const double *A, double *L, double *Linv, double *s, unsigned char *al, double *tmp, double tol) {
    const int tid = threadIdx.x;
    if (tid < K) {
        const double a = A[tid * LD + tid];
        const bool bad = !(a > 0.0);
        al[tid] = bad;
        s[tid] = bad ? 0.0 : 1.0 / sqrt(a);
    }
    __syncthreads();
    for (int j = 0; j < K; ++j) {
        if (tid >= j && tid < K) {
            double dot = 0.0;
            for (int t = 0; t < j; ++t) dot += L[tid * LD + t] * L[j * LD + t];      // aliased columns hold zeros
            tmp[tid] = dot;
        }
        __syncthreads();
        const double v = 1.0 - tmp[j];
        const bool bad = al[j] || !(v > tol);
        const double ljj = bad ? 0.0 : sqrt(v);
        double mine = 0.0;
        if (tid > j && tid < K && !bad && !al[tid]) mine = (A[tid * LD + j] * s[tid] * s[j] - tmp[tid]) / ljj;
        __syncthreads();
        if (tid < K) {
            if (tid == j) { L[j * LD + j] = ljj; if (bad) al[j] = 1; }
            else if (tid > j) L[tid * LD + j] = mine;
            else L[tid * LD + j] = 0.0;
        }
        __syncthreads();
    }
    if (tid < K) {                                       // column tid of L^-1 by forward substitution
        const int c = tid;
        for (int i = 0; i < K; ++i) {
            double u = 0.0;
            if (i >= c && !al[c] && !al[i]) {
                u = (i == c) ? 1.0 : 0.0;
                for (int t = c; t < i; ++t) u -= L[i * LD + t] * Linv[t * LD + c];
                u /= L[i * LD + i];
            }
            Linv[i * LD + c] = u;
        }
    }
    __syncthreads();
}

// x = A^-1 b and dinv = diag(A^-1) from the factors of wide_chol (zeros for aliased columns); threads < K write.
template <int LD>
__device__ void wide_solve(int K, const double *Linv, const double *s, const double *b, double *yv, double *x, double *dinv) {
    const int tid = threadIdx.x;
    if (tid < K) {
        double u = 0.0;
        for (int c = 0; c <= tid; ++c) u += Linv[tid * LD + c] * (b[c] * s[c]);
        yv[tid] = u;
    }
    __syncthreads();
    if (tid < K) {
        double u = 0.0, d = 0.0;
        for (int t = tid; t < K; ++t) { const double l = Linv[t * LD + tid]; u += l * yv[t]; d += l * l; }
        x[tid] = u * s[tid];
        dinv[tid] = d * s[tid] * s[tid];
    }
    __syncthreads();
}

// summary()$coefficients rows (R/npdr.R:41-68), as write_stats in regress.cu
__device__ void wide_write_stats(double *st, const double *beta, const double *se, const unsigned char *al, int K, double df,
                                 double extra, int iters, int flags) {
    int row2 = -1;
    for (int j = 1; j < K; ++j) if (!al[j]) { row2 = j; break; }
    const double nanv = __longlong_as_double(0x7ff8000000000000ll);
    st[0] = row2 >= 0 ? beta[row2] : nanv;
    st[1] = row2 >= 0 ? beta[row2] / se[row2] : nanv;
    st[2] = al[0] ? nanv : beta[0];
    st[3] = al[0] ? nanv : beta[0] / se[0];
    st[4] = df;
    st[5] = extra;
    st[6] = (double)iters;
    st[7] = (double)(flags | (al[1] ? NPDRB_FLAG_ALIASED : 0));
}

template <int KT>
__global__ void __launch_bounds__(WT, KT == 64 ? 1 : 2)      // narrow designs: two CTAs per SM hide the block loads' latency
wide_fit_kernel(WideArgs P) {
    using S = WideSmem<KT>;
    constexpr int LD = S::LD, NT = S::NT, TPG = S::TPG, NG = S::NG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *V = (double *)smem_raw;                      // [WB][LD] design rows of the block; reused for the group reduction
    double *G = V + S::VR;                               // [KT][LD] Gram matrix
    double *L = G + (size_t)KT * LD;
    double *Linv = L + (size_t)KT * LD;
    double *wv = Linv + (size_t)KT * LD;                 // [WB] weights
    double *beta = wv + WB, *se = beta + KT, *sc = se + KT, *tmp = sc + KT, *xs = tmp + KT, *dinv = xs + KT,
           *bvec = dinv + KT, *mean = bvec + KT, *gvec = mean + KT, *diag0 = gvec + KT, *zq = diag0 + KT, *dq = zq + KT;
    double *devs = dq + KT;                              // [WT]
    double *scal = devs + WT;                            // [16] scalars shared by the CTA
    unsigned char *al = (unsigned char *)(scal + 16);    // [KT] aliased flags of the current solve
    unsigned char *al_prev = al + KT;

    const int tid = threadIdx.x;
    const int a = blockIdx.x;
    if (a >= P.nattr) return;
    const int q = P.q, K = 2 + q;                        // column K of the design row holds the response
    const bool lm = P.reg_type == NPDRB_REG_LM;
    const double *xcol = P.X + (P.attr0 + a) * P.ldx;
    const int g = tid / TPG, within = tid % TPG, ti = within / NT, tj = within % NT;
    const bool tile_on = tj >= ti;                       // upper tiles; results are mirrored when G is assembled
    const int64_t nblk = (P.M + WB - 1) / WB;
    double *st = P.stats + (int64_t)a * 8;
    const double nanv = __longlong_as_double(0x7ff8000000000000ll);

    if (tid < KT) { beta[tid] = 0.0; se[tid] = nanv; al_prev[tid] = 0; }
    double dev_old = (double)P.M * (2.0 * 0.2876820724517809);       // deviance at mustart: 2 log(4/3) per pair
    int passes = 0;
    __syncthreads();

    // pass t accumulates the weighted system of IRLS iteration t at the coefficients of iteration t - 1 (t = 1: mustart)
    // together with the deviance AT those coefficients, which decides whether iteration t - 1 had converged
    for (int pass = 1; pass <= 26; ++pass) {
        double acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
        double dev_mine = 0.0;
        for (int64_t blk = 0; blk < nblk; ++blk) {
            const int64_t m0 = blk * WB;
            const int nb = (int)((P.M - m0 < WB) ? (P.M - m0) : WB);
            // (1) covariate diffs of the block: contiguous nb * q doubles
            for (int e = tid; e < nb * q; e += WT) V[(e / q) * LD + 2 + (e % q)] = P.CD[m0 * q + e];
            for (int e = tid; e < WB * (KT - K); e += WT) V[(e / (KT - K)) * LD + K + (e % (KT - K))] = 0.0;   // response + padding
            __syncthreads();
            // (2) one thread per pair: design row, weight, response, deviance term
            if (tid < WB) {
                double *row = V + tid * LD;
                if (tid < nb) {
                    const int64_t m = m0 + tid;
                    const double xi = xcol[P.Ri[m] - 1], xj = xcol[P.NN[m] - 1];
                    const double d = nb_diff1(xi, xj, P.diff_type);
                    const double y = P.yd[m];
                    row[0] = 1.0; row[1] = d;
                    if (lm) { wv[tid] = 1.0; row[K] = y; }
                    else {
                        double eta, mu;
                        if (pass == 1) { const double ms = (y + 0.5) / 2.0; eta = log(ms / (1.0 - ms)); mu = wide_linkinv(eta); }
                        else {
                            eta = 0.0;
                            for (int r = 0; r < K; ++r) eta += row[r] * beta[r];
                            mu = wide_linkinv(eta);
                            dev_mine += 2.0 * (wide_ylogy(y, mu) + wide_ylogy(1.0 - y, 1.0 - mu));
                        }
                        const double me = wide_mu_eta(eta), varmu = mu * (1.0 - mu);
                        wv[tid] = me * me / varmu;
                        row[K] = eta + (y - mu) / me;
                    }
                } else {
                    for (int r = 0; r < K; ++r) row[r] = 0.0;
                    wv[tid] = 0.0;
                }
            }
            __syncthreads();
            // (3) register tiles: acc[r][c] += W * v[ti + r NT] * v[tj + c NT]
            if (tile_on) {
                for (int m = g; m < WB; m += NG) {
                    const double *row = V + m * LD;
                    const double w = wv[m];
                    double av[4], bv[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) { av[r] = w * row[ti + r * NT]; bv[r] = row[tj + r * NT]; }
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[r][c] = fma(av[r], bv[c], acc[r][c]);
                }
            }
            __syncthreads();
        }
        ++passes;
        // group partials -> G, fixed order; deviance: fixed-order sum of the per-thread partials
        double *red = V;
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) red[(size_t)tid * 16 + r * 4 + c] = acc[r][c];
        devs[tid] = dev_mine;
        __syncthreads();
        if (tid < TPG && tile_on) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    double sum = 0.0;
                    for (int gg = 0; gg < NG; ++gg) sum += red[((size_t)gg * TPG + tid) * 16 + r * 4 + c];
                    const int rr = ti + r * NT, cc = tj + c * NT;
                    if (ti != tj || r <= c) { G[rr * LD + cc] = sum; G[cc * LD + rr] = sum; }
                }
        }
        if (tid == 0) {
            double dsum = 0.0;
            for (int t = 0; t < WB; ++t) dsum += devs[t];
            scal[0] = dsum;
        }
        __syncthreads();

        if (lm) {
            // lm + summary.lm (R/npdr.R:34,41-67): centred normal equations (Schur complement on the intercept), the same
            // algebra as lm_finalize_kernel in regress.cu.  Row / column K of G is the response.
            const double Mn = G[0], Sy = G[K], Syy = G[K * LD + K], ybar = Sy / Mn;
            if (tid < K) {
                mean[tid] = G[tid] / Mn;
                gvec[tid] = (tid == 0) ? 0.0 : G[tid * LD + K] - G[tid] * ybar;
                diag0[tid] = G[tid * LD + tid];
            }
            __syncthreads();
            for (int e = tid; e < K * K; e += WT) {              // centred matrix, staged in Linv's storage
                const int i = e / K, j = e % K;
                Linv[i * LD + j] = (i == 0 || j == 0) ? ((i == j) ? 1.0 : 0.0) : G[i * LD + j] - G[i] * mean[j];
            }
            __syncthreads();
            for (int e = tid; e < K * K; e += WT) {              // constant columns (alias test relative to the UNCENTRED norm) -> zero
                const int i = e / K, j = e % K;
                const bool ci = i > 0 && (!(diag0[i] > 0.0) || !(Linv[i * LD + i] > 1e-14 * diag0[i]));
                const bool cj = j > 0 && (!(diag0[j] > 0.0) || !(Linv[j * LD + j] > 1e-14 * diag0[j]));
                G[i * LD + j] = (ci || cj) ? 0.0 : Linv[i * LD + j];
            }
            __syncthreads();
            wide_chol<LD>(K, G, L, Linv, sc, al, tmp, 1e-14);
            wide_solve<LD>(K, Linv, sc, gvec, tmp, xs, dinv);
            if (tid < K) bvec[tid] = (tid == 0 || al[tid]) ? 0.0 : mean[tid];
            __syncthreads();
            wide_solve<LD>(K, Linv, sc, bvec, tmp, zq, dq);      // var(beta_0) needs mean' Gc^-1 mean
            if (tid == 0) {
                double b0 = ybar, rss = Syy - Sy * ybar, quad = 0.0;
                const double syy_c = rss;
                int rank = 1;
                for (int j = 1; j < K; ++j) {
                    const double bj = al[j] ? 0.0 : xs[j];
                    beta[j] = bj;
                    b0 -= bj * mean[j];
                    rss -= bj * gvec[j];
                    quad += bvec[j] * zq[j];
                    rank += al[j] ? 0 : 1;
                }
                beta[0] = b0;
                const double rdf = Mn - (double)rank, resvar = rss / rdf;
                se[0] = sqrt(resvar * (1.0 / Mn + quad));
                for (int j = 1; j < K; ++j) se[j] = sqrt(resvar * dinv[j]);
                al[0] = 0;
                wide_write_stats(st, beta, se, al, K, rdf, 1.0 - rss / syy_c, 0, 0);
            }
            break;
        }

        // ---- binomial: glm.fit bookkeeping
        const double dev = (pass == 1) ? dev_old : scal[0];
        int finish = 0, flags = 0;
        if (pass > 1) {
            if (!isfinite(dev)) { finish = 1; flags = NPDRB_FLAG_NONFINITE; }
            else if (fabs(dev - dev_old) / (fabs(dev) + 0.1) < 1e-8) finish = 1;         // glm.control(epsilon = 1e-8)
            else if (pass - 1 >= 25) { finish = 1; flags = NPDRB_FLAG_NOT_CONVERGED; }  // maxit = 25
            else dev_old = dev;
        }
        if (finish) {
            if (tid == 0) {
                int rank = 0;
                for (int r = 0; r < K; ++r) rank += al_prev[r] ? 0 : 1;
                wide_write_stats(st, beta, se, al_prev, K, (double)P.M - (double)rank, dev, pass - 1, flags);
            }
            break;
        }
        if (tid < K) bvec[tid] = G[tid * LD + K];        // X'Wz
        __syncthreads();
        wide_chol<LD>(K, G, L, Linv, sc, al, tmp, 1e-13);
        wide_solve<LD>(K, Linv, sc, bvec, tmp, xs, dinv);
        if (tid == 0) {
            bool ok = true;
            for (int r = 0; r < K; ++r) ok = ok && isfinite(xs[r]);
            scal[4] = ok ? 1.0 : 0.0;
        }
        __syncthreads();
        if (scal[4] == 0.0) {
            // non-finite coefficients: glm.fit stops with a warning; report the previous iterate
            if (tid == 0) {
                unsigned char none[64];
                for (int r = 0; r < K; ++r) none[r] = 0;
                wide_write_stats(st, beta, se, none, K, (double)P.M - (double)K, dev, pass - 1, NPDRB_FLAG_NONFINITE);
            }
            break;
        }
        if (tid < K) {
            beta[tid] = al[tid] ? 0.0 : xs[tid];
            se[tid] = al[tid] ? nanv : sqrt(dinv[tid]);
            al_prev[tid] = al[tid];
        }
        __syncthreads();
    }
    if (tid == 0 && P.max_passes) atomicMax(P.max_passes, passes);
}

// outcome and covariate diffs of every pair (R/npdr.R:299-345), pair-major; raw = uniReg (the instance's own values)
struct WidePairArgs {
    const int32_t *Ri, *NN; int64_t M, N;
    const double *y, *C; int q, reg_type, y_raw;
    int cov_type[NPDRB_MAX_COVARS];
    double *yd, *CD; unsigned long long *n_miss;
};
__global__ void wide_pairs_kernel(WidePairArgs g) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int miss = 0;
    if (m < g.M) {
        const int64_t i = g.Ri[m] - 1, j = g.NN[m] - 1;
        const double yi = g.y[i], yj = g.y[j];
        if (g.reg_type == NPDRB_REG_BINOMIAL) { miss = g.y_raw ? (yi != 0.0) : (yi != yj); g.yd[m] = (double)miss; }
        else g.yd[m] = g.y_raw ? yi : fabs(yi - yj);
        for (int c = 0; c < g.q; ++c) {
            const double *col = g.C + (int64_t)c * g.N;
            g.CD[m * g.q + c] = g.y_raw ? col[i] : nb_diff1(col[i], col[j], g.cov_type[c]);
        }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, miss);
    if ((threadIdx.x & 31) == 0 && mask) atomicAdd(g.n_miss, (unsigned long long)__popc(mask));
}

template <int KT>
int wide_launch(npdrb_ctx *ctx, const WideArgs &P) {
    const size_t smem = WideSmem<KT>::BYTES;
    NB_CUDA(cudaFuncSetAttribute(wide_fit_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wide_fit_kernel<KT><<<(unsigned)P.nattr, WT, smem, ctx->stream>>>(P);
    NB_LAUNCH_CHECK(ctx);
    return NPDRB_OK;
}

}  // namespace

int nb_regress_wide(npdrb_session *s, int64_t attr0, int64_t nattr, int32_t reg_type, int32_t diff_type,
                    const int32_t *cov_type, double *stats_out_host, int32_t *irls_passes) {
    npdrb_ctx *ctx = s->ctx;
    const int q = s->q;
    if (q + 3 > 64) { npdrb_set_error("regress: at most %d covariates are supported, got %d", NPDRB_MAX_COVARS, q); return NPDRB_EINVAL; }
    const int y_raw = (diff_type == NPDRB_DIFF_RAW) ? 1 : 0;
    nb_scope scope;
    double *d_yd = nullptr, *d_cd = nullptr, *d_stats = nullptr;
    unsigned long long *d_miss = nullptr;
    int *d_passes = nullptr;
    cudaEvent_t e0 = s->ev[0], e1 = s->ev[1];
    NB_CUDA(cudaEventRecord(e0, ctx->stream));
    NB_CHECK(scope.alloc(&d_yd, (size_t)s->M)); NB_CHECK(scope.alloc(&d_cd, (size_t)s->M * (size_t)q));
    NB_CHECK(scope.alloc(&d_stats, (size_t)nattr * 8)); NB_CHECK(scope.alloc(&d_miss, 1)); NB_CHECK(scope.alloc(&d_passes, 1));
    NB_CUDA(cudaMemsetAsync(d_miss, 0, sizeof(unsigned long long), ctx->stream));
    NB_CUDA(cudaMemsetAsync(d_passes, 0, sizeof(int), ctx->stream));
    WidePairArgs g;
    memset(&g, 0, sizeof(g));
    g.Ri = s->dRi; g.NN = s->dNN; g.M = s->M; g.N = s->N; g.y = s->dy; g.C = s->dC; g.q = q; g.reg_type = reg_type; g.y_raw = y_raw;
    for (int c = 0; c < q; ++c) g.cov_type[c] = cov_type ? cov_type[c] : NPDRB_DIFF_MATCH_MISMATCH;
    g.yd = d_yd; g.CD = d_cd; g.n_miss = d_miss;
    wide_pairs_kernel<<<(unsigned)nb_cdiv(s->M, 256), 256, 0, ctx->stream>>>(g);
    NB_LAUNCH_CHECK(ctx);
    unsigned long long h_miss = 0;
    NB_CUDA(cudaMemcpyAsync(&h_miss, d_miss, sizeof(h_miss), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaEventRecord(e1, ctx->stream));
    NB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    NB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    s->ms_prepare = ms;
    if (reg_type == NPDRB_REG_BINOMIAL && (h_miss == 0 || (int64_t)h_miss == s->M)) {
        npdrb_set_error("binomial regression: the pair outcome has a single level (%lld of %lld pairs are misses)",
                        (long long)h_miss, (long long)s->M);
        return NPDRB_EDATA;
    }
    NB_CUDA(cudaEventRecord(e0, ctx->stream));
    WideArgs P;
    memset(&P, 0, sizeof(P));
    P.X = s->dX; P.ldx = s->ldx; P.attr0 = attr0; P.nattr = (int)nattr; P.N = s->N;
    P.Ri = s->dRi; P.NN = s->dNN; P.M = s->M; P.yd = d_yd; P.CD = d_cd; P.q = q; P.diff_type = diff_type; P.reg_type = reg_type;
    P.stats = d_stats; P.max_passes = d_passes;
    const int kk = q + 3;                                // design columns + response
    NB_CHECK(kk <= 8 ? wide_launch<8>(ctx, P) : kk <= 16 ? wide_launch<16>(ctx, P) : kk <= 32 ? wide_launch<32>(ctx, P)
                                                                                              : wide_launch<64>(ctx, P));
    int h_passes = 0;
    NB_CUDA(cudaMemcpyAsync(stats_out_host, d_stats, sizeof(double) * (size_t)nattr * 8, cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaMemcpyAsync(&h_passes, d_passes, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaEventRecord(e1, ctx->stream));
    NB_CUDA(cudaEventSynchronize(e1));
    NB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    s->ms_regress = ms;
    s->ms_pass_first = ms; s->ms_pass_full = 0; s->n_pass_full = 0; s->evals_pass_full = 0; s->q_eff = q;
    if (irls_passes) *irls_passes = h_passes;
    return NPDRB_OK;
}
