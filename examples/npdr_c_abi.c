/* Minimal C caller of libnpdr_b200.so: the whole path on a small synthetic case through npdrb_npdr.
 *   gcc -Iinclude examples/npdr_c_abi.c -Lnpdr_b200 -lnpdr_b200 -Wl,-rpath,$PWD/npdr_b200 -lm -o /tmp/npdr_c_abi
 * Exit status: 0 = ran on a B200; 2 = no usable device (the library never computes on the CPU); 1 = any other error. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "npdr_b200.h"

int main(void) {
    const int64_t N = 200, p = 50;
    double *X = malloc(sizeof(double) * N * p), *y = malloc(sizeof(double) * N);
    unsigned s = 12345u;
    for (int64_t i = 0; i < N; ++i) y[i] = (double)(i % 2);
    for (int64_t k = 0; k < p; ++k)
        for (int64_t i = 0; i < N; ++i) {                 /* column-major, as R holds a matrix */
            s = s * 1664525u + 1013904223u;
            X[k * N + i] = (double)(s >> 8) / 16777216.0 + (k < 3 ? 0.8 * y[i] : 0.0);
        }
    npdrb_ctx *ctx = NULL;
    int rc = npdrb_create(0, &ctx);
    if (rc) { fprintf(stderr, "npdrb_create: %s\n", npdrb_last_error()); return rc == NPDRB_ENODEVICE ? 2 : 1; }
    npdrb_opts o;
    npdrb_opts_default(&o);                               /* binomial, numeric-abs, multisurf, manhattan, sd.frac 0.5 */
    npdrb_result r;
    rc = npdrb_npdr(ctx, X, N, p, y, NULL, &o, &r);
    if (rc) { fprintf(stderr, "npdrb_npdr: %s\n", npdrb_last_error()); npdrb_destroy(ctx); return 1; }
    printf("%lld pairs, %lld unique; attribute 1: beta = %.6f, z = %.3f (%d IRLS passes)\n", (long long)r.n_pairs,
           (long long)r.n_unique_pairs, r.stats[0], r.stats[1], (int)r.irls_passes);
    npdrb_result_release(&r);
    npdrb_destroy(ctx);
    free(X); free(y);
    return 0;
}
