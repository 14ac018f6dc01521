// ubench_fp64.cu -- issue-slot micro-benchmarks behind DESIGN.md's model of the IRLS pass kernel:
// does an integer / LDS / MUFU instruction issue "under" an FP64 instruction on sm_100a, or does every FP64 warp
// instruction hold the dispatch port for its 2 cycles?  Cycles come from clock64() inside the kernel (clock independent).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp64 ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

// per loop iteration: 16 DFMA (8 independent chains) + NI integer (mad.lo / lop3 alternating, 8 chains) + NL LDS.64
// (independent destinations, address from an integer chain) + NM MUFU.RCP64H (rcp.approx.ftz.f64 of a chain value)
template <int NI, int NL, int NM>
__global__ void __launch_bounds__(1024) k(double *out, long long *cyc, double a, double b, int iters, int ia) {
    __shared__ double tab[1024];
    for (int e = threadIdx.x; e < 1024; e += blockDim.x) tab[e] = 1.0 + e * 1e-9;
    __syncthreads();
    double x[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = threadIdx.x + c;
    int v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = ia + c + threadIdx.x;
    double ld[4] = {b, b, b, b}, mr[4] = {a, a, a, a};
    unsigned la[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) la[c] = (threadIdx.x * 8 + c * 64) & 0x1ff8u;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(tab);
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int c = 0; c < 8; ++c)                       // loaded / MUFU values feed the chains (as in the real kernel), so nothing can be sunk out of the loop
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(NM ? mr[c & 3] : a), "d"(NL ? ld[c & 3] : b));
#pragma unroll
            for (int n = r * NI / 2; n < (r + 1) * NI / 2; ++n) {
                if (n & 1) asm volatile("lop3.b32 %0, %0, %1, 0x1ff8, 0x78;" : "+r"(v[n & 7]) : "r"(ia));
                else asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v[n & 7]) : "r"(ia), "r"(n + 1));
            }
#pragma unroll
            for (int n = r * NL / 2; n < (r + 1) * NL / 2; ++n)
            {
                la[n & 7] = (la[n & 7] + 264u) & 0x1ff8u;          // 2 integer instructions per LDS (counted in the model)
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(ld[n & 3]) : "r"(sbase + la[n & 7]) : "memory");
            }
#pragma unroll
            for (int n = r * NM / 2; n < (r + 1) * NM / 2; ++n)
                asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(mr[n & 3]) : "d"(x[n & 7]) : "memory");
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s += x[c] + v[c];
    s += ld[0] + ld[1] + ld[2] + ld[3] + mr[0] + mr[1] + mr[2] + mr[3] + la[0] + la[1] + la[2] + la[3] + la[4] + la[5] + la[6] + la[7];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[(size_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int NI, int NL, int NM>
void run(int threads, int sm) {
    double *d; long long *dc;
    const int blocks = sm, warps = threads / 32;
    cudaMalloc(&d, sizeof(double) * blocks * threads); cudaMalloc(&dc, sizeof(long long) * blocks * warps);
    const int iters = 20000;
    k<NI, NL, NM><<<blocks, threads>>>(d, dc, 1.0000001, 1e-9, iters, 3);
    k<NI, NL, NM><<<blocks, threads>>>(d, dc, 1.0000001, 1e-9, iters, 3);
    cudaDeviceSynchronize();
    static long long h[148 * 32 * 2];
    cudaMemcpy(h, dc, sizeof(long long) * blocks * warps, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < blocks * warps; ++i) mean += (double)h[i];
    mean /= (double)blocks * warps * iters;                 // cycles per iteration as one warp sees them
    const double wps = threads / 128.0;                     // warps per SM sub-partition
    printf("16 DFMA + %2d INT + %d LDS + %d MUFU | %4d threads (%2.0f warps/SMSP): %6.2f cycles/iter/warp -> %5.2f cycles per warp-iteration of the SMSP"
           " (FP64 floor 32.00, serial-issue model %d)\n", NI, NL, NM, threads, wps, mean, mean / wps, 32 + NI + NL * 3 + NM * 2 + 3);
    cudaFree(d); cudaFree(dc);
}

int main() {
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    for (int threads : {256, 512, 1024}) {
        run<0, 0, 0>(threads, sm);
        run<8, 0, 0>(threads, sm);
        run<16, 0, 0>(threads, sm);
        run<32, 0, 0>(threads, sm);
        run<0, 4, 0>(threads, sm);
        run<0, 8, 0>(threads, sm);
        run<0, 0, 4>(threads, sm);
        run<0, 0, 8>(threads, sm);
        run<16, 8, 8>(threads, sm);
    }
    return 0;
}
