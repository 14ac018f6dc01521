// Host accuracy check of npdr_b200/csrc/dmath.cuh against libm (long double).
#include <cstdio>
#include <cmath>
#include <random>
#include "dmath.cuh"
static const nb_tables T = NB_TABLES_STATIC_INIT;
static const nb_tables256 T2 = NB_TABLES256_STATIC_INIT;
static const nb_tables512 T3 = NB_TABLES512_STATIC_INIT;
static const nb_tables1024 T4 = NB_TABLES1024_STATIC_INIT;
int main() {
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(-30.0, 30.0);
    double max_rel_exp = 0, max_abs_log = 0, max_rel_log = 0;
    for (int it = 0; it < 4000000; ++it) {
        double eta = (it < 2000000) ? U(rng) : U(rng) / 30.0;
        double t = nb_exp(eta, T.exp_t);
        long double te = expl((long double)eta);
        double rel = (double)fabsl(((long double)t - te) / te);
        if (rel > max_rel_exp) max_rel_exp = rel;
        double u = 1.0 + t;
        double l = nb_log_ge1(u, T.log_rc, T.log_lrc, T.log_eln2);
        long double le = logl((long double)u);
        double ab = (double)fabsl((long double)l - le);
        if (ab > max_abs_log) max_abs_log = ab;
        if (le > 1e-3L) { double r = (double)(ab / le); if (r > max_rel_log) max_rel_log = r; }
    }
    // edge values
    double edges[] = {1.0, 1.0 + 2.2e-16, 2.0, 1.9999999999999998, 4503599627370497.0, 1.0078125, 1.00781249999};
    for (double u : edges) {
        double ab = (double)fabsl((long double)nb_log_ge1(u, T.log_rc, T.log_lrc, T.log_eln2) - logl((long double)u));
        if (ab > max_abs_log) max_abs_log = ab;
    }
    printf("max_rel_exp=%.3e max_abs_log=%.3e max_rel_log=%.3e\n", max_rel_exp, max_abs_log, max_rel_log);
    // trimmed variants used by the IRLS kernel
    double fr_exp = 0, fa_log = 0;
    for (int it = 0; it < 4000000; ++it) {
        double eta = (it < 2000000) ? U(rng) : U(rng) / 30.0;
        double t = nb_exp_fast(eta, T.exp_t);
        long double te = expl((long double)eta);
        double rel = (double)fabsl(((long double)t - te) / te);
        if (rel > fr_exp) fr_exp = rel;
        double u = 1.0 + (double)te;
        double ab = (double)fabsl((long double)nb_log_ge1_fast(u, T.log_rc, T.log_lrc, T.log_eln2) - logl((long double)u));
        double sc = (double)logl((long double)u); if (sc < 1.0) sc = 1.0;
        if (ab / sc > fa_log) fa_log = ab / sc;
    }
    printf("fast: max_rel_exp=%.3e max_log_err(rel to max(1,log))=%.3e\n", fr_exp, fa_log);
    // 256-entry variants used by pass2_kernel<IRLS_FULL>
    double g_exp = 0, g_log = 0;
    for (int it = 0; it < 4000000; ++it) {
        double eta = (it < 2000000) ? U(rng) : U(rng) / 30.0;
        double t = nb_exp_fast256(eta * NB_256_OVER_LN2, T2.exp_t);
        long double te = expl((long double)eta);
        double rel = (double)fabsl(((long double)t - te) / te);
        if (rel > g_exp) g_exp = rel;
        double u = 1.0 + (double)te;
        double ab = (double)fabsl((long double)nb_log_ge1_fast256(u, T2.log_rc, T2.log_lrc, T2.log_eln2) - logl((long double)u));
        double sc = (double)logl((long double)u); if (sc < 1.0) sc = 1.0;
        if (ab / sc > g_log) g_log = ab / sc;
    }
    printf("fast256: max_rel_exp=%.3e max_log_err(rel to max(1,log))=%.3e\n", g_exp, g_log);
    // v3 variants used by pass2_kernel<IRLS_FULL>: u = 1 + exp(eta) in one go, 512-bin degree-3 log with integer exponent
    double h_u = 0, h_log = 0, h_bias = 0; long n_bias = 0;
    for (int it = 0; it < 4000000; ++it) {
        double eta = (it < 2000000) ? U(rng) : U(rng) / 30.0;
        double u = nb_one_plus_exp256(eta * NB_256_OVER_LN2, T3.exp_t);
        long double ue = 1.0L + expl((long double)eta);
        double rel = (double)fabsl(((long double)u - ue) / ue);                   // the kernel only ever uses u = 1 + t
        if (rel > h_u) h_u = rel;
        int32_t e;
        double lm = nb_log_mant512(u, T3.log_rc, T3.log_lrc, &e);
        long double le = logl((long double)u);
        long double err = ((long double)lm + (long double)e * (long double)NB_LN2) - le;
        double sc = (double)le; if (sc < 1.0) sc = 1.0;
        if ((double)fabsl(err) / sc > h_log) h_log = (double)fabsl(err) / sc;
        h_bias += (double)err; ++n_bias;
    }
    h_bias /= (double)n_bias;
    printf("v3: max_rel_err(1+exp)=%.3e max_log_err(rel to max(1,log))=%.3e mean_log_err=%.3e\n", h_u, h_log, h_bias);
    if (!(h_u < 2e-13 && h_log < 3e-13 && fabs(h_bias) < 6e-14)) return 1;
    // v4 exp used by pass2_kernel<IRLS_FULL>: 1024-entry table, economised degree-2 polynomial
    double k_u = 0;
    for (int it = 0; it < 4000000; ++it) {
        double eta = (it < 2000000) ? U(rng) : U(rng) / 30.0;
        double u = nb_one_plus_exp1024(eta * NB_1024_OVER_LN2, T4.exp_t);
        long double ue = 1.0L + expl((long double)eta);
        double rel = (double)fabsl(((long double)u - ue) / ue);
        if (rel > k_u) k_u = rel;
    }
    printf("v4: max_rel_err(1+exp)=%.3e\n", k_u);
    if (!(k_u < 1.8e-12)) return 1;
    return !(max_rel_exp < 4e-16 && max_abs_log < 8e-15 && max_rel_log < 1e-15 && fr_exp < 6e-14 && fa_log < 2e-15 &&
             g_exp < 2e-13 && g_log < 8e-15);
}
