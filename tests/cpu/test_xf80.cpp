// Host check of npdr_b200/csrc/xf80.cuh against the compiler's real x87 `long double`.
// Build: g++ -O1 -ffp-contract=off -I npdr_b200/csrc tests/cpu/test_xf80.cpp -o /tmp/test_xf80
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <random>
#include "xf80.cuh"

static xf80 from_ld(long double v) {           // exact import of an x87 value
    xf80 r; uint64_t m; uint16_t se;
    memcpy(&m, &v, 8); memcpy(&se, (char*)&v + 8, 2);
    r.m = m; r.neg = se >> 15; r.e = (int)(se & 0x7FFF) - 16383;
    if (m == 0) r.e = 0;
    return r;
}
static bool same(xf80 a, long double v) {
    xf80 b = from_ld(v);
    if (a.m == 0 && b.m == 0) return true;
    return a.m == b.m && a.e == b.e && a.neg == b.neg;
}
int main() {
    std::mt19937_64 rng(1618);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    std::uniform_int_distribution<int> E(-40, 40);
    long fails = 0, n = 0;
    auto rnd = [&]() { return std::ldexp(U(rng), E(rng)); };
    for (int it = 0; it < 2000000; ++it) {
        double x = rnd(), y = rnd(), z = rnd();
        volatile long double a = (long double)x * (long double)y;     // a full 64-bit significand
        volatile long double b = (long double)z + a;
        xf80 A = xf_mul(xf_from_double(x), xf_from_double(y));
        xf80 B = xf_add(xf_from_double(z), A);
        volatile long double s = a + b, d = a - b, m = a * b, q = a / b;
        fails += !same(A, a); fails += !same(B, b);
        fails += !same(xf_add(A, B), s); fails += !same(xf_sub(A, B), d);
        fails += !same(xf_mul(A, B), m);
        if (b != 0) fails += !same(xf_div(A, B), q);
        fails += (xf_to_double(A) != (double)a) ; fails += (xf_to_double(xf_add(A,B)) != (double)s);
        // near-cancellation and tie cases
        volatile long double c = a + (long double)std::ldexp(1.0, E(rng)) ;
        fails += !same(xf_add(A, xf_from_double(std::ldexp(1.0, 0))), a + 1.0L) ; (void)c;
        n += 9;
    }
    // integer divisions as used for means
    for (int it = 0; it < 200000; ++it) {
        double x = std::fabs(rnd()) * 1e4; long k = 2 + (long)(rng() % 100000);
        volatile long double a = (long double)x * 3.0L + (long double)rnd();
        volatile long double q = a / k;
        fails += !same(xf_div(from_ld(a), xf_from_int(k)), q); n++;
    }
    // the exact radius recipe (oracle r_var + colSums) on random "distance" rows
    for (int it = 0; it < 300; ++it) {
        int N = 3 + (int)(rng() % 3000), skip = (int)(rng() % N);
        double *x = (double*)malloc(sizeof(double) * N);
        double base = 50.0 + 100.0 * std::fabs(U(rng));
        for (int i = 0; i < N; ++i) x[i] = base + 5.0 * U(rng);
        x[skip] = 0.0;
        long nobs = N - 1;
        long double sum = 0, tmp; for (int k = 0; k < N; ++k) if (k != skip) sum += x[k];
        double colsum = (double)sum;
        tmp = sum / nobs; long double s2 = 0; for (int k = 0; k < N; ++k) if (k != skip) s2 += (x[k] - tmp);
        tmp = tmp + s2 / nobs; long double xm = (double)tmp; long double s3 = 0;
        for (int k = 0; k < N; ++k) if (k != skip) s3 += (x[k] - xm) * (x[k] - xm);
        double var = (double)(s3 / (nobs - 1));
        xf80 S = xf_zero(); for (int k = 0; k < N; ++k) if (k != skip) S = xf_add(S, xf_from_double(x[k]));
        xf80 T = xf_div(S, xf_from_int(nobs)); xf80 S2 = xf_zero();
        for (int k = 0; k < N; ++k) if (k != skip) S2 = xf_add(S2, xf_sub(xf_from_double(x[k]), T));
        T = xf_add(T, xf_div(S2, xf_from_int(nobs))); xf80 XM = xf_from_double(xf_to_double(T)); xf80 S3 = xf_zero();
        for (int k = 0; k < N; ++k) if (k != skip) { xf80 dd = xf_sub(xf_from_double(x[k]), XM); S3 = xf_add(S3, xf_mul(dd, dd)); }
        double var2 = xf_to_double(xf_div(S3, xf_from_int(nobs - 1)));
        fails += (xf_to_double(S) != colsum); fails += (var != var2); n += 2;
        free(x);
    }
    printf("checks=%ld fails=%ld\n", n, fails);
    return fails != 0;
}
