// Host accuracy check of npdr_b200/csrc/dmath.cuh against libm (long double).
#include <cstdio>
#include <cmath>
#include <random>
#include "dmath.cuh"
static const nb_tables T = NB_TABLES_STATIC_INIT;
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
    // exp of the v2 IRLS kernel (pass2_kernel<IRLS_FULL>): 1024-entry table, economised degree-2 polynomial
    double k_u = 0;
    for (int it = 0; it < 4000000; ++it) {
        double eta = (it < 2000000) ? U(rng) : U(rng) / 30.0;
        double u = nb_one_plus_exp1024(eta * NB_1024_OVER_LN2, T4.exp_t);
        long double ue = 1.0L + expl((long double)eta);
        double rel = (double)fabsl(((long double)u - ue) / ue);
        if (rel > k_u) k_u = rel;
    }
    printf("v2 kernel exp: max_rel_err(1+exp)=%.3e\n", k_u);
    if (!(k_u < 1.8e-12)) return 1;
    return !(max_rel_exp < 4e-16 && max_abs_log < 8e-15 && max_rel_log < 1e-15);
}
