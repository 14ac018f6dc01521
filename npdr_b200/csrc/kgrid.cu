// kgrid.cu -- batched-k ReliefF neighbourhoods: vwok's loop over knn (R/adaptive.R:85-111) on ONE resident problem.
//
// get_Ri_nearest (R/nearestNeighbors.R:622-634) keeps, for instance Ri, the rows with min_rank(d) <= k + 1 in (d, row)
// order and drops d == 0.  For a row sorted once by (d, row) that is a PREFIX: with cut = the (k+1)-th smallest distance,
// the list is the sorted entries [#zeros, #{d <= cut}).  So the k grid needs one stable segmented sort of the distance block
// (cub::DeviceSegmentedRadixSort, stable => ties in d stay in ascending row order, which is dplyr's order) and, per k, two
// binary searches per row + one copy -- instead of a radix select and a bitonic sort per row and per k.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "common.cuh"

namespace {

struct SegBegin { int ld; __host__ __device__ int operator()(int r) const { return r * ld; } };
struct SegEnd { int ld, n; __host__ __device__ int operator()(int r) const { return r * ld + n; } };

__global__ void iota_rows_kernel(int32_t *__restrict__ idx, int64_t ld, int64_t N, int64_t nrows) {
    const int64_t total = nrows * N;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
        idx[(e / N) * ld + (e % N)] = (int32_t)(e % N) + 1;
}

// first position in the sorted row whose value is > v
__device__ __forceinline__ int64_t upper_bound_row(const double *__restrict__ row, int64_t n, double v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (row[mid] <= v) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void kgrid_count_kernel(const double *__restrict__ SD, int64_t ld, int64_t N, int64_t nrows, int64_t kplus1,
                                   int64_t *__restrict__ counts, int64_t *__restrict__ zeros) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const double *row = SD + r * ld;
    const int64_t want = kplus1 < N ? kplus1 : N;
    const int64_t hi = upper_bound_row(row, N, row[want - 1]);
    const int64_t z = upper_bound_row(row, N, 0.0);
    zeros[r] = z;
    counts[r] = hi > z ? hi - z : 0;
}

// one warp per row: copy the prefix [zeros, zeros + n) of the sorted indices
__global__ void kgrid_write_kernel(const int32_t *__restrict__ SJ, int64_t ld, int64_t row0, int64_t nrows,
                                   const int64_t *__restrict__ zeros, const int64_t *__restrict__ ptr,
                                   int32_t *__restrict__ Ri, int32_t *__restrict__ NN) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= nrows) return;
    const int64_t off = ptr[r], n = ptr[r + 1] - off;
    const int32_t *src = SJ + r * ld + zeros[r];
    const int32_t ri = (int32_t)(row0 + r + 1);
    for (int64_t t = lane; t < n; t += 32) { NN[off + t] = src[t]; Ri[off + t] = ri; }
}

__global__ void __launch_bounds__(1024)
kgrid_scan_kernel(const int64_t *__restrict__ counts, int64_t n, int64_t *__restrict__ ptr, int64_t *__restrict__ n_empty) {
    // single-block exclusive scan (n <= a few 10^4 rows): chunked Hillis-Steele with a running carry
    __shared__ int64_t buf[1024];
    __shared__ int64_t carry_s, empty_s;
    if (threadIdx.x == 0) { carry_s = 0; empty_s = 0; }
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < n ? counts[i] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int64_t t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
            __syncthreads();
            buf[threadIdx.x] += t;
            __syncthreads();
        }
        const int64_t carry = carry_s;
        if (i < n) { ptr[i] = carry + buf[threadIdx.x] - v; if (v == 0) atomicAdd((unsigned long long *)&empty_s, 1ull); }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + buf[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) { ptr[n] = carry_s; *n_empty = empty_s; }
}

}  // namespace

// Sort every owned row of the distance block by (d, row) once; the sorted copy stays with the session.
int nb_sort_rows(npdrb_session *s) {
    npdrb_ctx *ctx = s->ctx;
    if (!s->dD) { npdrb_set_error("sort rows: compute or set distances first"); return NPDRB_EINVAL; }
    const int64_t N = s->N, nrows = s->nrows, ld = s->ldd;
    if (ld >= (1ll << 30)) { npdrb_set_error("sort rows: N too large"); return NPDRB_EINVAL; }
    NB_CUDA(cudaEventRecord(s->ev[0], ctx->stream));
    if (s->d_sortD) { cudaFree(s->d_sortD); s->d_sortD = nullptr; }
    if (s->d_sortJ) { cudaFree(s->d_sortJ); s->d_sortJ = nullptr; }
    NB_CUDA(cudaMalloc(&s->d_sortD, sizeof(double) * (size_t)nrows * (size_t)ld));
    NB_CUDA(cudaMalloc(&s->d_sortJ, sizeof(int32_t) * (size_t)nrows * (size_t)ld));
    nb_scope scope;
    int32_t *d_iota = nullptr;
    NB_CHECK(scope.alloc(&d_iota, (size_t)nrows * (size_t)ld));
    iota_rows_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_iota, ld, N, nrows);
    NB_LAUNCH_CHECK(ctx);
    // row chunks so that the item count of one call fits an int
    const int64_t rows_per = ((1ll << 30) / ld) > 0 ? ((1ll << 30) / ld) : 1;
    thrust::counting_iterator<int> cnt(0);
    auto seg_b = thrust::make_transform_iterator(cnt, SegBegin{(int)ld});
    auto seg_e = thrust::make_transform_iterator(cnt, SegEnd{(int)ld, (int)N});
    size_t tmp_bytes = 0;
    {
        const int nr = (int)(nrows < rows_per ? nrows : rows_per);
        NB_CUDA(cub::DeviceSegmentedRadixSort::SortPairs(nullptr, tmp_bytes, s->dD, s->d_sortD, d_iota, s->d_sortJ, (int)(nr * ld), nr,
                                                         seg_b, seg_e, 0, 64, ctx->stream));
    }
    void *d_tmp = nullptr;
    NB_CHECK(scope.alloc((unsigned char **)&d_tmp, tmp_bytes ? tmp_bytes : 1));
    for (int64_t r0 = 0; r0 < nrows; r0 += rows_per) {
        const int nr = (int)((nrows - r0) < rows_per ? (nrows - r0) : rows_per);
        NB_CUDA(cub::DeviceSegmentedRadixSort::SortPairs(d_tmp, tmp_bytes, s->dD + r0 * ld, s->d_sortD + r0 * ld, d_iota + r0 * ld,
                                                         s->d_sortJ + r0 * ld, (int)(nr * ld), nr, seg_b, seg_e, 0, 64, ctx->stream));
        ctx->launches += 2;
    }
    NB_CUDA(cudaEventRecord(s->ev[1], ctx->stream));
    NB_CUDA(cudaEventSynchronize(s->ev[1]));
    float ms = 0;
    NB_CUDA(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
    s->ms_sort = ms;
    s->sorted_valid = true;
    return NPDRB_OK;
}

// ReliefF neighbour list of the owned rows for one k, from the sorted rows (same list, bit for bit, as nb_neighbors)
int nb_neighbors_sorted(npdrb_session *s, int64_t k, int64_t *M_local, int64_t *n_empty) {
    npdrb_ctx *ctx = s->ctx;
    if (!s->sorted_valid || !s->d_sortD) { npdrb_set_error("neighbors (k grid): call npdrb_session_sort_rows first"); return NPDRB_EINVAL; }
    const int64_t N = s->N, nrows = s->nrows, ld = s->ldd;
    if (N < 3) { npdrb_set_error("neighbors: need at least 3 instances"); return NPDRB_EINVAL; }
    if (k <= 0) { npdrb_set_error("neighbors (k grid): k must be positive"); return NPDRB_EINVAL; }
    NB_CUDA(cudaEventRecord(s->ev[0], ctx->stream));
    if (s->dRiL) { cudaFree(s->dRiL); s->dRiL = nullptr; }
    if (s->dNNL) { cudaFree(s->dNNL); s->dNNL = nullptr; }
    const int64_t nmax = N > nrows ? N : nrows;
    if (!s->d_rowptr) NB_CUDA(cudaMalloc(&s->d_rowptr, sizeof(int64_t) * (size_t)(nmax + 2)));
    nb_scope scope;
    int64_t *d_counts = nullptr, *d_zeros = nullptr, *d_nempty = nullptr;
    NB_CHECK(scope.alloc(&d_counts, (size_t)nrows)); NB_CHECK(scope.alloc(&d_zeros, (size_t)nrows)); NB_CHECK(scope.alloc(&d_nempty, 1));
    kgrid_count_kernel<<<(unsigned)nb_cdiv(nrows, 128), 128, 0, ctx->stream>>>(s->d_sortD, ld, N, nrows, k + 1, d_counts, d_zeros);
    NB_LAUNCH_CHECK(ctx);
    kgrid_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_counts, nrows, s->d_rowptr, d_nempty);
    NB_LAUNCH_CHECK(ctx);
    int64_t tot[2] = {0, 0};
    NB_CUDA(cudaMemcpyAsync(&tot[0], s->d_rowptr + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaMemcpyAsync(&tot[1], d_nempty, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    const int64_t M = tot[0];
    NB_CUDA(cudaMalloc(&s->dRiL, sizeof(int32_t) * (size_t)(M > 0 ? M : 1)));
    NB_CUDA(cudaMalloc(&s->dNNL, sizeof(int32_t) * (size_t)(M > 0 ? M : 1)));
    kgrid_write_kernel<<<(unsigned)nb_cdiv(nrows * 32, 256), 256, 0, ctx->stream>>>(s->d_sortJ, ld, s->row0, nrows, d_zeros, s->d_rowptr,
                                                                                   s->dRiL, s->dNNL);
    NB_LAUNCH_CHECK(ctx);
    s->M_local = M; s->n_empty_local = tot[1];
    if (M_local) *M_local = M;
    if (n_empty) *n_empty = tot[1];
    NB_CUDA(cudaEventRecord(s->ev[1], ctx->stream));
    NB_CUDA(cudaEventSynchronize(s->ev[1]));
    float ms = 0;
    NB_CUDA(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
    s->ms_neighbors = ms;
    return NPDRB_OK;
}
