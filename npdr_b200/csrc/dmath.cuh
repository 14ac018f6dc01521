// dmath.cuh -- table-driven FP64 exp / log for the binomial IRLS kernel (host + device).
//
// binomial(logit) in base-R glm.fit (called at R/npdr.R:36) needs, per pair and attribute,
// t = exp(eta) (clamped like C logit_linkinv: eta < -30 -> DBL_EPSILON, eta > 30 -> 1/DBL_EPSILON)
// and the deviance term log(1 + t).  libm-grade exp/log cost ~25-30 FP64 instructions each on
// the FP64 pipe, which is the binding unit of that kernel; these versions use small lookup
// tables (shared memory on the device) and short polynomials: ~11 + ~9 FP64 instructions,
// relative error < 2e-16 (exp) and absolute error < 2e-16 (log on [1, 2^53]).
#pragma once
#include <stdint.h>
#include "dmath_tables.h"

#if defined(__CUDACC__)
#define NB_HD __host__ __device__ __forceinline__
#else
#define NB_HD inline
#endif

struct nb_tables {
    double exp_t[64];      // 2^(j/64)
    double log_rc[128];    // ~1/centre of mantissa bin i (24 significant bits)
    double log_lrc[128];   // -log(log_rc[i])
    double log_eln2[64];   // e * ln2
};

#define NB_TABLES_STATIC_INIT { NB_EXP_TABLE_INIT, NB_LOG_RC_TABLE_INIT, NB_LOG_LRC_TABLE_INIT, NB_LOG_ELN2_TABLE_INIT }

NB_HD uint64_t nb_d2u(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    union { double d; uint64_t u; } c; c.d = d; return c.u;
#endif
}
NB_HD double nb_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    union { double d; uint64_t u; } c; c.u = u; return c.d;
#endif
}
NB_HD double nb_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

// exp(eta) for |eta| <= 30 (callers clamp first).
NB_HD double nb_exp(double eta, const double* __restrict__ exp_t) {
    const double magic = 6755399441055744.0;                 // 1.5 * 2^52
    double t = nb_fma(eta, NB_64_OVER_LN2, magic);
    int32_t ki = (int32_t)(uint32_t)nb_d2u(t);               // low word = round(eta * 64/ln2)
    double kf = t - magic;
    double r = nb_fma(kf, -NB_LN2_64_HI, eta);
    r = nb_fma(kf, -NB_LN2_64_LO, r);                        // |r| <= ln2/128
    double r2 = r * r;
    double q = nb_fma(r, 1.0 / 120.0, 1.0 / 24.0);
    q = nb_fma(q, r, 1.0 / 6.0);
    q = nb_fma(q, r, 0.5);
    q = nb_fma(q, r2, r);                                    // exp(r) - 1
    double tj = exp_t[ki & 63];
    double v = nb_fma(tj, q, tj);
    int32_t e = ki >> 6;                                     // arithmetic shift: floor
    return nb_u2d(nb_d2u(v) + ((uint64_t)(int64_t)e << 52)); // scale by 2^e (|e| <= 44: stays normal)
}

// log(u) for 1 <= u < 2^63.
NB_HD double nb_log_ge1(double u, const double* __restrict__ log_rc, const double* __restrict__ log_lrc,
                        const double* __restrict__ log_eln2) {
    uint64_t b = nb_d2u(u);
    int32_t e = (int32_t)(b >> 52) - 1023;
    int32_t i = (int32_t)((b >> 45) & 127);
    double m = nb_u2d((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
    double f = nb_fma(m, log_rc[i], -1.0);                   // |f| <= 2^-8 + 2^-24
    double p = nb_fma(f, -1.0 / 6.0, 0.2);
    p = nb_fma(p, f, -0.25);
    p = nb_fma(p, f, 1.0 / 3.0);
    p = nb_fma(p, f, -0.5);
    double f2 = f * f;
    p = nb_fma(p, f2, f);                                    // log1p(f)
    return (log_eln2[e & 63] + log_lrc[i]) + p;   // & 63: a NaN/Inf argument must not index out of the table
}

// ---- exp of the v2 IRLS pass kernel (pass2_kernel<IRLS_FULL>): 1024 table entries, degree-2 polynomial --------------
// (the v1 kernel, used for more than two covariates, evaluates nb_exp / nb_log_ge1 above)
// eta in units of ln2/1024: k = rint(s), frac = s - k (exact, |frac| <= 1/2), x = frac * ln2/1024, |x| <= h = ln2/2048.
// exp(x) ~ 1 + (1 + h^2/8) x + x^2/2: the linear coefficient absorbs the Chebyshev-economised cubic term, leaving
// max |error| = h^3/24 + h^4/192 = 1.6e-12 relative -- one FMA less per evaluation than a degree-3 / 256-entry form.
// The kernel needs no per-record log: the deviance is accumulated as a running product (regress.cu: v2_tail / v2_renorm).
struct nb_tables1024 { double exp_t[1024]; };
#define NB_TABLES1024_STATIC_INIT { NB_EXP1024_TABLE_INIT }
#define NB_EXP2_C1 (NB_LN2_1024 * (1.0 + (NB_LN2_1024 * 0.5) * (NB_LN2_1024 * 0.5) / 8.0))
#define NB_EXP2_C2 (NB_LN2_1024 * NB_LN2_1024 / 2.0)

NB_HD double nb_one_plus_exp1024(double sv, const double* __restrict__ exp_t) {     // sv = eta * 1024/ln2, |eta| < 30
    const double magic = 6755399441055744.0;
    double tt = sv + magic;
    int32_t ki = (int32_t)(uint32_t)nb_d2u(tt);
    double kf = tt - magic;
    double fr = sv - kf;
    double qq = nb_fma(fr, NB_EXP2_C2, NB_EXP2_C1);
    qq = nb_fma(qq, fr, 1.0);
    uint32_t j = (uint32_t)ki & 1023u;
    uint64_t b = nb_d2u(exp_t[j]);
    uint32_t hi = ((uint32_t)(b >> 32) - j * 1024u) + (uint32_t)ki * 1024u;      // the kernel's table image holds the first term
    return nb_fma(nb_u2d(((uint64_t)hi << 32) | (b & 0xFFFFFFFFull)), qq, 1.0);
}
