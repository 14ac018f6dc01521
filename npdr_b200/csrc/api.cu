// api.cu -- the C ABI of libnpdr_b200.so (see include/npdr_b200.h): contexts, sessions and the
// host-buffer operator entry points that replace the reference's exported R functions.
#include <math.h>
#include <stdarg.h>
#include <condition_variable>
#include <map>
#include <chrono>
#include <mutex>
#include <thread>
#include <unordered_map>
#include "common.cuh"

static thread_local char g_err[1024] = "";

void npdrb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------ device-memory cache (see common.cuh)
namespace {
struct DevCache {
    std::multimap<size_t, void *> free_blocks;          // size -> block
    std::unordered_map<void *, size_t> live;            // blocks handed out
    size_t cached_bytes = 0;
};
std::mutex g_cache_mu;
DevCache g_cache[64];

void cache_release_device(int dev) {
    DevCache &c = g_cache[dev & 63];
    for (auto &kv : c.free_blocks) (cudaFree)(kv.second);
    c.free_blocks.clear();
    c.cached_bytes = 0;
}
}  // namespace

cudaError_t nb_cached_malloc(void **p, size_t bytes) {
    *p = nullptr;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const size_t n = (bytes + 511) & ~(size_t)511;
    std::lock_guard<std::mutex> lk(g_cache_mu);
    DevCache &c = g_cache[dev & 63];
    auto it = c.free_blocks.lower_bound(n);
    if (it != c.free_blocks.end() && it->first <= n + n / 8 + 4096) {       // close fit only: no silent bloat
        *p = it->second;
        c.live[*p] = it->first;
        c.cached_bytes -= it->first;
        c.free_blocks.erase(it);
        return cudaSuccess;
    }
    e = (cudaMalloc)(p, n ? n : 512);
    if (e != cudaSuccess) {                                                   // out of memory: drop the cache, retry once
        cudaGetLastError();
        cache_release_device(dev);
        e = (cudaMalloc)(p, n ? n : 512);
        if (e != cudaSuccess) return e;
    }
    c.live[*p] = n ? n : 512;
    return cudaSuccess;
}

cudaError_t nb_cached_free(void *p) {
    if (!p) return cudaSuccess;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (int d = 0; d < 64; ++d) {                                            // usually found on the current device
        DevCache &c = g_cache[(dev + d) & 63];
        auto it = c.live.find(p);
        if (it == c.live.end()) continue;
        c.free_blocks.emplace(it->second, p);
        c.cached_bytes += it->second;
        c.live.erase(it);
        return cudaSuccess;
    }
    return (cudaFree)(p);                                                     // not ours (should not happen)
}

// ------------------------------------------------------------------ pinned staging ring (pageable host sources)
// R's heap is pageable: cudaMemcpyAsync from it is staged by the driver on the calling thread and does not overlap with
// anything.  The ring gives a pageable source the behaviour of a pinned one: helper threads memcpy a chunk into one of
// three pinned buffers, the chunk goes to the device asynchronously on the copy stream, and the buffer is reused once
// its copy has completed.
struct nb_stage_ring {
    static constexpr int NBUF = 3;
    size_t buf_bytes = (size_t)32 << 20;
    void *buf[NBUF] = {nullptr, nullptr, nullptr};
    cudaEvent_t done[NBUF] = {nullptr, nullptr, nullptr};
    bool used[NBUF] = {false, false, false};
    int next = 0;
    // helper pool: piece t of the current job is copied by helper t (the caller copies piece 0)
    int n_helpers = 0;
    std::vector<std::thread> helpers;
    std::mutex mu; std::condition_variable cv_go, cv_done;
    unsigned gen = 0; int remaining = 0; bool stop = false;
    char *job_dst = nullptr; const char *job_src = nullptr; size_t job_bytes = 0; int job_pieces = 1;
    std::mutex api_mu;                // one staged copy at a time per ring (workers that share a context in the emulated multi-device mode)

    static void copy_piece(char *dst, const char *src, size_t bytes, int piece, int pieces) {
        const size_t per = ((bytes / (size_t)pieces) + 4095) & ~(size_t)4095;
        const size_t lo = per * (size_t)piece;
        if (lo >= bytes) return;
        const size_t n = (lo + per < bytes && piece + 1 < pieces) ? per : bytes - lo;
        memcpy(dst + lo, src + lo, n);
    }
    void helper_main(int t) {
        unsigned seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            cv_go.wait(lk, [&] { return stop || gen != seen; });
            if (stop) return;
            seen = gen;
            char *d = job_dst; const char *sr = job_src; const size_t n = job_bytes; const int pcs = job_pieces;
            lk.unlock();
            if (t < pcs) copy_piece(d, sr, n, t, pcs);
            lk.lock();
            if (--remaining == 0) cv_done.notify_one();
        }
    }
    void start(int helpers_wanted) {
        n_helpers = helpers_wanted;
        for (int t = 1; t <= n_helpers; ++t) helpers.emplace_back(&nb_stage_ring::helper_main, this, t);
    }
    void parallel_copy(void *dst, const void *src, size_t bytes) {
        const int pieces = (bytes >= ((size_t)1 << 20)) ? n_helpers + 1 : 1;
        if (pieces == 1) { memcpy(dst, src, bytes); return; }
        {
            std::lock_guard<std::mutex> lk(mu);
            job_dst = (char *)dst; job_src = (const char *)src; job_bytes = bytes; job_pieces = pieces;
            remaining = n_helpers; ++gen;
        }
        cv_go.notify_all();
        copy_piece((char *)dst, (const char *)src, bytes, 0, pieces);
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return remaining == 0; });
    }
    ~nb_stage_ring() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv_go.notify_all();
        for (auto &t : helpers) t.join();
        for (int b = 0; b < NBUF; ++b) { if (buf[b]) cudaFreeHost(buf[b]); if (done[b]) cudaEventDestroy(done[b]); }
    }
};

static std::mutex g_ring_mu;          // creation of a context's ring / copy stream
static int get_ring(npdrb_ctx *ctx, nb_stage_ring **out) {
    std::lock_guard<std::mutex> lk(g_ring_mu);
    if (!ctx->copy_stream) NB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->ring) {
        nb_stage_ring *r = new nb_stage_ring();
        for (int b = 0; b < nb_stage_ring::NBUF; ++b) {
            cudaError_t e = cudaHostAlloc(&r->buf[b], r->buf_bytes, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->done[b], cudaEventDisableTiming);
            if (e != cudaSuccess) { delete r; npdrb_set_error("pinned staging ring: %s", cudaGetErrorString(e)); return NPDRB_ENOMEM; }
        }
        // copiers (helpers + the calling thread): the host's cores shared out among the box's GPUs (every GPU may be staging at
        // the same time), between 4 and 8 -- one core moves ~10 GB/s, a PCIe 5 x16 link ~55 GB/s; NPDRB_STAGE_THREADS overrides
        const unsigned hw = std::thread::hardware_concurrency();
        int ndev = 1;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); ndev = 1; }
        int copiers = hw ? (int)hw / ndev : 4;
        copiers = copiers < 4 ? 4 : (copiers > 8 ? 8 : copiers);
        int want = copiers - 1;
        const char *e = getenv("NPDRB_STAGE_THREADS");
        if (e && atoi(e) >= 1) want = atoi(e) - 1;
        if (hw && want > (int)hw - 1) want = (int)hw - 1;
        if (want > 15) want = 15;
        r->start(want < 0 ? 0 : want);
        ctx->ring = r;
    }
    *out = ctx->ring;
    return NPDRB_OK;
}

int nb_staged_h2d(npdrb_ctx *ctx, double *dst, int64_t ld_dst, const double *src, int64_t rows, int64_t ncols) {
    const size_t col_bytes = sizeof(double) * (size_t)rows;
    nb_stage_ring *r = nullptr;
    NB_CHECK(get_ring(ctx, &r));
    std::lock_guard<std::mutex> api_lk(r->api_mu);
    if (col_bytes > r->buf_bytes) {                       // a single column does not fit a ring buffer: let the driver stage it
        NB_CUDA(cudaMemcpy2DAsync(dst, sizeof(double) * (size_t)ld_dst, src, col_bytes, col_bytes, (size_t)ncols,
                                  cudaMemcpyHostToDevice, ctx->copy_stream));
        return NPDRB_OK;
    }
    const int64_t cols_per_chunk = (int64_t)(r->buf_bytes / col_bytes);
    for (int64_t c = 0; c < ncols; c += cols_per_chunk) {
        const int64_t nc = (c + cols_per_chunk < ncols) ? cols_per_chunk : ncols - c;
        const int b = r->next;
        r->next = (b + 1) % nb_stage_ring::NBUF;
        if (r->used[b]) NB_CUDA(cudaEventSynchronize(r->done[b]));
        r->parallel_copy(r->buf[b], src + c * rows, col_bytes * (size_t)nc);
        NB_CUDA(cudaMemcpy2DAsync(dst + c * ld_dst, sizeof(double) * (size_t)ld_dst, r->buf[b], col_bytes, col_bytes, (size_t)nc,
                                  cudaMemcpyHostToDevice, ctx->copy_stream));
        NB_CUDA(cudaEventRecord(r->done[b], ctx->copy_stream));
        r->used[b] = true;
    }
    return NPDRB_OK;
}

bool nb_host_pointer_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

namespace {

__global__ void check_finite_kernel(const double *__restrict__ v, int64_t n, int *__restrict__ bad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int b = 0;
    for (; i < n; i += (int64_t)gridDim.x * blockDim.x) b |= !isfinite(v[i]);
    if (b) atomicOr(bad, 1);
}

// FP64 issue-rate micro-kernels: 8 independent chains per thread, 4096 x 8 ops each
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < 4096; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void __launch_bounds__(256) dadd_peak_kernel(double *out, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < 4096; ++i) {
        x0 = __dadd_rn(x0, b); x1 = __dadd_rn(x1, b); x2 = __dadd_rn(x2, b); x3 = __dadd_rn(x3, b);
        x4 = __dadd_rn(x4, b); x5 = __dadd_rn(x5, b); x6 = __dadd_rn(x6, b); x7 = __dadd_rn(x7, b);
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int check_finite(npdrb_ctx *ctx, const double *d, int64_t n, const char *what) {
    int *d_bad = nullptr, h_bad = 0;
    NB_CUDA(cudaMalloc(&d_bad, sizeof(int)));
    NB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    int grid = (int)((n + 255) / 256);
    if (grid > 4096) grid = 4096;
    if (grid < 1) grid = 1;
    check_finite_kernel<<<grid, 256, 0, ctx->stream>>>(d, n, d_bad);
    NB_LAUNCH_CHECK(ctx);
    NB_CUDA(cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_bad);
    if (h_bad) { npdrb_set_error("%s contains NA/NaN/Inf (rejected: SURVEY 8a quirk 8)", what); return NPDRB_EDATA; }
    return NPDRB_OK;
}

}  // namespace

int nb_check_finite(npdrb_ctx *ctx, const double *d, int64_t n, const char *what) { return check_finite(ctx, d, n, what); }

int nb_issue_slab(npdrb_session *s, int slab) {
    npdrb_ctx *ctx = s->ctx;
    if (s->slab_ev_external) {
        // the caller's transfers are in flight (or, with ready flags, being issued by another host thread right now: an event
        // must have been recorded before a stream may wait for it, so wait for the producer's flag first)
        for (int i = s->n_slabs_issued; i <= slab && i < s->n_slabs_pending; ++i) {
            if (s->slab_ready) {
                const auto t0 = std::chrono::steady_clock::now();
                while (__atomic_load_n(&s->slab_ready[i], __ATOMIC_ACQUIRE) == 0) {
                    std::this_thread::yield();
                    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) {
                        npdrb_set_error("column slab %d of the adopted matrix never became ready (producer stalled?)", i);
                        return NPDRB_ECUDA;
                    }
                }
            }
            s->n_slabs_issued = i + 1;
        }
        return NPDRB_OK;
    }
    for (int i = s->n_slabs_issued; i <= slab && i < s->n_slabs_pending; ++i) {
        const int64_t c0 = (int64_t)i * s->slab_cols, c1 = (c0 + s->slab_cols < s->p) ? c0 + s->slab_cols : s->p;
        if (!s->hX_lazy) { npdrb_set_error("upload: slab %d has no host source", i); return NPDRB_EINVAL; }
        NB_CHECK(nb_staged_h2d(ctx, s->dX + c0 * s->ldx, s->ldx, s->hX_lazy + c0 * s->N, s->N, c1 - c0));
        NB_CUDA(cudaEventRecord(s->slab_ev[(size_t)i], ctx->copy_stream));
        s->n_slabs_issued = i + 1;
    }
    return NPDRB_OK;
}

int nb_wait_upload(npdrb_session *s) {
    if (s->n_slabs_pending == 0) return NPDRB_OK;
    npdrb_ctx *ctx = s->ctx;
    NB_CHECK(nb_issue_slab(s, s->n_slabs_pending - 1));
    for (int i = 0; i < s->n_slabs_pending; ++i) NB_CUDA(cudaStreamWaitEvent(ctx->stream, s->slab_ev[(size_t)i], 0));
    s->n_slabs_pending = 0; s->hX_lazy = nullptr;
    return check_finite(ctx, s->dX, s->ldx * s->p, "the attribute matrix");
}

extern "C" {

const char *npdrb_version(void) { return "npdr_b200 0.1.0 (sm_100a)"; }
const char *npdrb_last_error(void) { return g_err; }
void npdrb_free(void *p) { free(p); }

void npdrb_opts_default(npdrb_opts *o) {
    memset(o, 0, sizeof(*o));
    o->regression_type = NPDRB_REG_BINOMIAL;          // npdr() defaults, R/npdr.R:151-164
    o->attr_diff_type = NPDRB_DIFF_NUMERIC_ABS;
    o->nbd_method = NPDRB_NBD_MULTISURF;
    o->nbd_metric = NPDRB_METRIC_MANHATTAN;
    o->knn = 0;
    o->msurf_sd_frac = 0.5;
    o->n_covars = 0;
    for (int c = 0; c < NPDRB_MAX_COVARS; ++c) o->covar_diff_type[c] = NPDRB_DIFF_MATCH_MISMATCH;
}

void npdrb_result_release(npdrb_result *r) {
    if (!r) return;
    free(r->stats); free(r->Ri); free(r->NN);
    r->stats = nullptr; r->Ri = nullptr; r->NN = nullptr;
}

int64_t npdrb_knn_surf(int64_t m_samples, double sd_frac) {
    // knnSURF, R/utils.R:12-18: floor((m - 1) * (1 - erf(sd.frac/sqrt(2))) / 2), erf(x) = 2*pnorm(x*sqrt(2)) - 1
    double x = sd_frac / sqrt(2.0);
    double z = x * sqrt(2.0);
    double pn = 0.5 * erfc(-z / sqrt(2.0));
    double erfv = 2.0 * pn - 1.0;
    return (int64_t)floor((double)(m_samples - 1) * (1.0 - erfv) / 2.0);
}

int npdrb_create(int device, npdrb_ctx **out) {
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        npdrb_set_error("no CUDA device available (%s); libnpdr_b200 has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        cudaGetLastError();
        return NPDRB_ENODEVICE;
    }
    if (device < 0 || device >= n) { npdrb_set_error("device %d out of range (0..%d)", device, n - 1); return NPDRB_ENODEVICE; }
    cudaDeviceProp prop;
    NB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        npdrb_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return NPDRB_ENODEVICE;
    }
    NB_CUDA(cudaSetDevice(device));
    npdrb_ctx *c = new npdrb_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    NB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    *out = c;
    return NPDRB_OK;
}

int npdrb_release_cache(npdrb_ctx *ctx) {
    NB_CUDA(cudaSetDevice(ctx->device));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    std::lock_guard<std::mutex> lk(g_cache_mu);
    cache_release_device(ctx->device);
    return NPDRB_OK;
}

void npdrb_destroy(npdrb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        cache_release_device(ctx->device);
    }
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); }
    delete ctx->ring;
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int npdrb_set_stream(npdrb_ctx *ctx, void *cuda_stream) {
    NB_CUDA(cudaSetDevice(ctx->device));
    if (ctx->stream) NB_CUDA(cudaStreamSynchronize(ctx->stream));      // the memory cache assumes one stream at a time
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return NPDRB_OK;
}

int npdrb_synchronize(npdrb_ctx *ctx) {
    NB_CUDA(cudaSetDevice(ctx->device));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    return NPDRB_OK;
}

int64_t npdrb_launch_count(npdrb_ctx *ctx) { return ctx->launches; }

int npdrb_upload_columns(npdrb_ctx *ctx, void *dst_device, int64_t ld_dst, const double *src_host, int64_t rows, int64_t ncols) {
    if (!ctx) { npdrb_set_error("npdrb_upload_columns: null context (no device)"); return NPDRB_ENODEVICE; }
    if (!dst_device || !src_host || rows < 1 || ncols < 0 || ld_dst < rows) { npdrb_set_error("npdrb_upload_columns: bad arguments"); return NPDRB_EINVAL; }
    NB_CUDA(cudaSetDevice(ctx->device));
    if (ncols == 0) return NPDRB_OK;
    if (nb_host_pointer_is_pinned(src_host)) {
        // on the copy stream, never the context's main stream: that one may already hold distance launches waiting for this
        // very data (block-cyclic upload pipeline, where another host thread calls this function)
        {
            std::lock_guard<std::mutex> lk(g_ring_mu);
            if (!ctx->copy_stream) NB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        }
        NB_CUDA(cudaMemcpy2DAsync(dst_device, sizeof(double) * (size_t)ld_dst, src_host, sizeof(double) * (size_t)rows,
                                  sizeof(double) * (size_t)rows, (size_t)ncols, cudaMemcpyHostToDevice, ctx->copy_stream));
        NB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        return NPDRB_OK;
    }
    NB_CHECK(nb_staged_h2d(ctx, (double *)dst_device, ld_dst, src_host, rows, ncols));
    NB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    return NPDRB_OK;
}

int npdrb_measure_fp64_peak(npdrb_ctx *ctx, double *dfma_tflops, double *dadd_tflops) {
    NB_CUDA(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256;
    double *d_out = nullptr;
    NB_CUDA(cudaMalloc(&d_out, sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t e0, e1;
    NB_CUDA(cudaEventCreate(&e0)); NB_CUDA(cudaEventCreate(&e1));
    double best[2] = {0, 0};
    for (int which = 0; which < 2; ++which) {
        for (int rep = 0; rep < 5; ++rep) {
            NB_CUDA(cudaEventRecord(e0, ctx->stream));
            if (which == 0) dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, 1.0000001, 1e-9);
            else dadd_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, 1e-9);
            NB_LAUNCH_CHECK(ctx);
            NB_CUDA(cudaEventRecord(e1, ctx->stream));
            NB_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            NB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            const double ops = (double)blocks * threads * 4096.0 * 8.0;
            const double rate = ops / (ms * 1e-3) / 1e12 * (which == 0 ? 2.0 : 1.0);
            if (rep > 0 && rate > best[which]) best[which] = rate;
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    if (dfma_tflops) *dfma_tflops = best[0];
    if (dadd_tflops) *dadd_tflops = best[1];
    return NPDRB_OK;
}

// ------------------------------------------------------------------ sessions
int npdrb_session_create(npdrb_ctx *ctx, int64_t N, int64_t p, int32_t q, npdrb_session **out) {
    *out = nullptr;
    if (!ctx) { npdrb_set_error("session: null context"); return NPDRB_EINVAL; }
    if (N < 1 || p < 1) { npdrb_set_error("session: need N >= 1 and p >= 1"); return NPDRB_EINVAL; }
    if (N >= (1ll << 31) - 256) { npdrb_set_error("session: N too large"); return NPDRB_EINVAL; }
    if (q < 0 || q > NPDRB_MAX_COVARS) { npdrb_set_error("session: 0..%d covariates supported, got %d", NPDRB_MAX_COVARS, q); return NPDRB_EINVAL; }
    NB_CUDA(cudaSetDevice(ctx->device));
    npdrb_session *s = new npdrb_session();
    s->ctx = ctx; s->N = N; s->p = p; s->q = q;
    NB_CUDA(cudaEventCreate(&s->ev[0]));
    NB_CUDA(cudaEventCreate(&s->ev[1]));
    *out = s;
    return NPDRB_OK;
}

void npdrb_session_destroy(npdrb_session *s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    if (s->ctx->copy_stream) cudaStreamSynchronize(s->ctx->copy_stream);
    if (!s->slab_ev_external) for (cudaEvent_t e : s->slab_ev) if (e) cudaEventDestroy(e);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->own_X) cudaFree(s->dX);
    cudaFree(s->dXs);
    if (s->own_y) cudaFree(s->dy);
    if (s->own_C) cudaFree(s->dC);
    if (s->own_D) cudaFree(s->dD);
    cudaFree(s->d_sd_vec); cudaFree(s->d_radius); cudaFree(s->d_radius2);
    cudaFree(s->dRiL); cudaFree(s->dNNL); cudaFree(s->d_rowptr);
    cudaFree(s->d_sortD); cudaFree(s->d_sortJ);
    if (s->own_pairs) { cudaFree(s->dRi); cudaFree(s->dNN); }
    cudaFree(s->dXT);
    nb_free_streams(s);
    if (s->ev[0]) cudaEventDestroy(s->ev[0]);
    if (s->ev[1]) cudaEventDestroy(s->ev[1]);
    delete s;
}

int npdrb_session_upload(npdrb_session *s, const double *X, const double *y, const double *C) {
    npdrb_ctx *ctx = s->ctx;
    NB_CUDA(cudaSetDevice(ctx->device));
    const int64_t N = s->N, p = s->p;
    if (X) {
        if (!s->dX || !s->own_X) {
            s->ldx = nb_round_up(N, 16);
            NB_CUDA(cudaMalloc(&s->dX, sizeof(double) * (size_t)s->ldx * (size_t)p));
            s->own_X = true;
            if (s->ldx != N) NB_CUDA(cudaMemsetAsync(s->dX, 0, sizeof(double) * (size_t)s->ldx * (size_t)p, ctx->stream));
        }
        // Large matrices travel as column slabs on a second stream: the distance kernel starts on slab 0 while the rest
        // is still crossing PCIe (it accumulates attribute ranges in order, so the result is bit-identical).  The host
        // buffer must stay valid until the first call that consumes X returns.
        const int64_t bytes = (int64_t)sizeof(double) * N * p;
        const char *np_ = getenv("NPDRB_NO_PIPELINED_UPLOAD");
        const int n_slabs = (bytes >= (64ll << 20) && !(np_ && np_[0] == '1')) ? 8 : 1;
        if (n_slabs > 1) {
            {
                std::lock_guard<std::mutex> lk(g_ring_mu);
                if (!ctx->copy_stream) NB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
            }
            if (!s->slab_ev_external) for (cudaEvent_t e : s->slab_ev) cudaEventDestroy(e);
            s->slab_ev_external = false;
            s->slab_ev.assign((size_t)n_slabs, nullptr);
            s->slab_cols = nb_round_up(nb_cdiv(p, n_slabs), 16);
            cudaEvent_t ready;                              // the memset / previous users of dX on the main stream come first
            NB_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
            NB_CUDA(cudaEventRecord(ready, ctx->stream));
            NB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ready, 0));
            cudaEventDestroy(ready);
            // pinned source: every slab is enqueued now.  Pageable source: nothing is copied here -- nb_issue_slab()
            // stages slab i through the pinned ring when its consumer is about to be launched (distance.cu), so the
            // host-side memcpy of slab i+1 overlaps the kernel of slab i.
            const bool pinned = nb_host_pointer_is_pinned(X);
            int used = 0;
            for (int i = 0; i < n_slabs; ++i) {
                const int64_t c0 = (int64_t)i * s->slab_cols, c1 = (c0 + s->slab_cols < p) ? c0 + s->slab_cols : p;
                if (c0 >= p) break;
                NB_CUDA(cudaEventCreateWithFlags(&s->slab_ev[(size_t)i], cudaEventDisableTiming));
                if (pinned) {
                    NB_CUDA(cudaMemcpy2DAsync(s->dX + c0 * s->ldx, sizeof(double) * (size_t)s->ldx, X + c0 * N, sizeof(double) * (size_t)N,
                                              sizeof(double) * (size_t)N, (size_t)(c1 - c0), cudaMemcpyHostToDevice, ctx->copy_stream));
                    NB_CUDA(cudaEventRecord(s->slab_ev[(size_t)i], ctx->copy_stream));
                }
                used = i + 1;
            }
            s->n_slabs_pending = used;
            s->slab_start.assign((size_t)used + 1, p);
            for (int i = 0; i < used; ++i) s->slab_start[(size_t)i] = (int64_t)i * s->slab_cols;
            s->n_slabs_issued = pinned ? used : 0;
            s->hX_lazy = pinned ? nullptr : X;
        } else {
            NB_CUDA(cudaMemcpy2DAsync(s->dX, sizeof(double) * (size_t)s->ldx, X, sizeof(double) * (size_t)N,
                                      sizeof(double) * (size_t)N, (size_t)p, cudaMemcpyHostToDevice, ctx->stream));
        }
        s->xt_attr0 = -1; s->xt_nattr = 0;
    }
    if (y) {
        if (!s->dy || !s->own_y) { NB_CUDA(cudaMalloc(&s->dy, sizeof(double) * (size_t)N)); s->own_y = true; }
        NB_CUDA(cudaMemcpyAsync(s->dy, y, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
        s->streams_valid = false;
    }
    if (C && s->q > 0) {
        if (!s->dC || !s->own_C) { NB_CUDA(cudaMalloc(&s->dC, sizeof(double) * (size_t)N * s->q)); s->own_C = true; }
        NB_CUDA(cudaMemcpyAsync(s->dC, C, sizeof(double) * (size_t)N * s->q, cudaMemcpyHostToDevice, ctx->stream));
        s->streams_valid = false;
    }
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (X && s->n_slabs_pending == 0) NB_CHECK(check_finite(ctx, s->dX, s->ldx * p, "the attribute matrix"));
    if (y) NB_CHECK(check_finite(ctx, s->dy, N, "the outcome"));
    if (C && s->q > 0) NB_CHECK(check_finite(ctx, s->dC, N * s->q, "the covariates"));
    return NPDRB_OK;
}

int npdrb_session_set_device_inputs(npdrb_session *s, const double *dX, int64_t ldx, const double *dy, const double *dC) {
    if (dX) {
        if (s->n_slabs_pending) {
            if (!s->slab_ev_external && s->ctx->copy_stream) cudaStreamSynchronize(s->ctx->copy_stream);
            s->n_slabs_pending = 0; s->n_slabs_issued = 0; s->hX_lazy = nullptr;
        }
        if (s->slab_ev_external) { s->slab_ev.clear(); s->slab_ev_external = false; s->slab_ready = nullptr; }
        if (ldx < s->N || (ldx & 1)) { npdrb_set_error("set_device_inputs: ldx must be even and >= N"); return NPDRB_EINVAL; }
        if (((uintptr_t)dX) & 15) { npdrb_set_error("set_device_inputs: dX must be 16-byte aligned"); return NPDRB_EINVAL; }
        if (s->own_X) cudaFree(s->dX);
        s->dX = const_cast<double *>(dX); s->ldx = ldx; s->own_X = false;
        s->xt_attr0 = -1; s->xt_nattr = 0;
    }
    if (dy) { if (s->own_y) cudaFree(s->dy); s->dy = const_cast<double *>(dy); s->own_y = false; s->streams_valid = false; }
    if (dC) { if (s->own_C) cudaFree(s->dC); s->dC = const_cast<double *>(dC); s->own_C = false; s->streams_valid = false; }
    return NPDRB_OK;
}

int npdrb_session_set_device_slab_events_v(npdrb_session *s, int32_t n_slabs, const int64_t *slab_starts, void *const *events,
                                           const int32_t *ready) {
    // The adopted device matrix is still ARRIVING (e.g. an NCCL broadcast per rank's attribute block on the caller's
    // communication stream): column slab i = attributes [slab_starts[i], slab_starts[i + 1]) is complete once events[i] fires.
    // The distance stage then runs one launch per slab, each waiting for its own event only (same per-element order
    // k = 0..p-1, so the result is bit-identical), and the transfer of slab i+1 hides behind the kernel of slab i.
    if (!s->dX) { npdrb_set_error("set_device_slab_events: the session has no device matrix yet (npdrb_session_set_device_inputs)"); return NPDRB_EINVAL; }
    bool ok = n_slabs >= 1 && events && slab_starts && slab_starts[0] == 0 && slab_starts[n_slabs] == s->p;
    for (int i = 0; ok && i < n_slabs; ++i)
        ok = slab_starts[i] < slab_starts[i + 1] && (slab_starts[i] % 16) == 0 && events[i] != nullptr;
    if (!ok) {
        npdrb_set_error("set_device_slab_events: need increasing slab bounds 0 = s_0 < ... < s_n = p (multiples of 16 except p) and one event per slab");
        return NPDRB_EINVAL;
    }
    if (!s->slab_ev_external) for (cudaEvent_t e : s->slab_ev) if (e) cudaEventDestroy(e);
    s->slab_ev.assign((size_t)n_slabs, nullptr);
    for (int i = 0; i < n_slabs; ++i) s->slab_ev[(size_t)i] = (cudaEvent_t)events[i];
    s->slab_start.assign(slab_starts, slab_starts + n_slabs + 1);
    s->slab_ev_external = true;
    s->slab_ready = ready;
    s->slab_cols = 0; s->n_slabs_pending = n_slabs; s->n_slabs_issued = ready ? 0 : n_slabs; s->hX_lazy = nullptr;
    return NPDRB_OK;
}

int npdrb_session_set_device_slab_events(npdrb_session *s, int32_t n_slabs, int64_t slab_cols, void *const *events, const int32_t *ready) {
    if (n_slabs < 1 || slab_cols < 1 || (int64_t)(n_slabs - 1) * slab_cols >= s->p || (int64_t)n_slabs * slab_cols < s->p) {
        npdrb_set_error("set_device_slab_events: need n_slabs slabs of slab_cols (a multiple of 16) columns covering the p attributes");
        return NPDRB_EINVAL;
    }
    std::vector<int64_t> st((size_t)n_slabs + 1, s->p);
    for (int i = 0; i < n_slabs; ++i) st[(size_t)i] = (int64_t)i * slab_cols;
    return npdrb_session_set_device_slab_events_v(s, n_slabs, st.data(), events, ready);
}

int npdrb_session_distances(npdrb_session *s, int32_t metric, int32_t fast_euclidean, int64_t row0, int64_t nrows) {
    NB_CUDA(cudaSetDevice(s->ctx->device));
    s->sorted_valid = false;
    return nb_distances(s, metric, fast_euclidean, row0, nrows);
}

int npdrb_session_distances_balanced(npdrb_session *s, int32_t metric, int32_t fast_euclidean, const int64_t *row_starts,
                                     int32_t rank, int32_t world) {
    NB_CUDA(cudaSetDevice(s->ctx->device));
    s->sorted_valid = false;
    return nb_distances_balanced(s, metric, fast_euclidean, row_starts, rank, world);
}
int npdrb_session_balanced_counts(npdrb_session *s, int64_t *send_doubles, int64_t *recv_doubles) {
    return nb_balanced_counts(s, send_doubles, recv_doubles);
}
int npdrb_session_balanced_pack(npdrb_session *s, double *d_send) {
    NB_CUDA(cudaSetDevice(s->ctx->device));
    return nb_balanced_pack(s, d_send);
}
int npdrb_session_balanced_unpack(npdrb_session *s, const double *d_recv) {
    NB_CUDA(cudaSetDevice(s->ctx->device));
    s->sorted_valid = false;
    return nb_balanced_unpack(s, d_recv);
}

int npdrb_session_set_device_distances(npdrb_session *s, const double *dD, int64_t ldd, int64_t row0, int64_t nrows) {
    if (ldd < s->N || row0 < 0 || nrows < 1 || row0 + nrows > s->N) { npdrb_set_error("set_device_distances: bad block"); return NPDRB_EINVAL; }
    if (s->own_D) cudaFree(s->dD);
    s->dD = const_cast<double *>(dD); s->ldd = ldd; s->row0 = row0; s->nrows = nrows; s->own_D = false;
    s->sorted_valid = false;
    return NPDRB_OK;
}

int npdrb_session_set_sd_vec(npdrb_session *s, const double *sd_vec_host) {
    NB_CUDA(cudaSetDevice(s->ctx->device));
    if (!sd_vec_host) { cudaFree(s->d_sd_vec); s->d_sd_vec = nullptr; return NPDRB_OK; }
    if (!s->d_sd_vec) NB_CUDA(cudaMalloc(&s->d_sd_vec, sizeof(double) * (size_t)s->N));
    NB_CUDA(cudaMemcpyAsync(s->d_sd_vec, sd_vec_host, sizeof(double) * (size_t)s->N, cudaMemcpyHostToDevice, s->ctx->stream));
    NB_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return NPDRB_OK;
}

int npdrb_session_neighbors(npdrb_session *s, int32_t nbd_method, double sd_frac, int64_t k, int64_t *M_local,
                            int64_t *n_empty_local) {
    NB_CUDA(cudaSetDevice(s->ctx->device));
    return nb_neighbors(s, nbd_method, sd_frac, k, M_local, n_empty_local);
}

int npdrb_session_neighbors_hitmiss(npdrb_session *s, int32_t nbd_method, double sd_frac, int64_t k, int64_t *M_local,
                                    int64_t *n_empty_local) {
    NB_CUDA(cudaSetDevice(s->ctx->device));
    return nb_neighbors_hitmiss(s, nbd_method, sd_frac, k, M_local, n_empty_local);
}

int npdrb_session_sort_rows(npdrb_session *s) {
    NB_CUDA(cudaSetDevice(s->ctx->device));
    return nb_sort_rows(s);
}

int npdrb_session_neighbors_sorted(npdrb_session *s, int64_t k, int64_t *M_local, int64_t *n_empty_local) {
    NB_CUDA(cudaSetDevice(s->ctx->device));
    return nb_neighbors_sorted(s, k, M_local, n_empty_local);
}

int npdrb_session_copy_radius2(npdrb_session *s, double *radius_host) {
    if (!s->d_radius2) { npdrb_set_error("copy_radius2: no hit/miss radii"); return NPDRB_EINVAL; }
    NB_CUDA(cudaSetDevice(s->ctx->device));
    NB_CUDA(cudaMemcpyAsync(radius_host, s->d_radius2, sizeof(double) * (size_t)s->nrows, cudaMemcpyDeviceToHost, s->ctx->stream));
    NB_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return NPDRB_OK;
}

int npdrb_session_set_surf_radius(npdrb_session *s, double radius) {
    s->surf_radius = radius; s->have_surf_radius = true;
    return NPDRB_OK;
}

int npdrb_session_local_pairs(npdrb_session *s, const int32_t **dRi, const int32_t **dNN, int64_t *M_local) {
    if (!s->dRiL) { npdrb_set_error("local_pairs: run neighbors first"); return NPDRB_EINVAL; }
    *dRi = s->dRiL; *dNN = s->dNNL; *M_local = s->M_local;
    return NPDRB_OK;
}

int npdrb_session_copy_local_pairs_device(npdrb_session *s, int32_t *dRi_dst, int32_t *dNN_dst) {
    if (!s->dRiL) { npdrb_set_error("copy_local_pairs_device: run neighbors first"); return NPDRB_EINVAL; }
    NB_CUDA(cudaSetDevice(s->ctx->device));
    if (s->M_local > 0) {
        NB_CUDA(cudaMemcpyAsync(dRi_dst, s->dRiL, sizeof(int32_t) * (size_t)s->M_local, cudaMemcpyDeviceToDevice, s->ctx->stream));
        NB_CUDA(cudaMemcpyAsync(dNN_dst, s->dNNL, sizeof(int32_t) * (size_t)s->M_local, cudaMemcpyDeviceToDevice, s->ctx->stream));
    }
    NB_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return NPDRB_OK;
}

int npdrb_session_copy_radius(npdrb_session *s, double *radius_host) {
    if (!s->d_radius) { npdrb_set_error("copy_radius: no radii"); return NPDRB_EINVAL; }
    NB_CUDA(cudaSetDevice(s->ctx->device));
    NB_CUDA(cudaMemcpyAsync(radius_host, s->d_radius, sizeof(double) * (size_t)s->nrows, cudaMemcpyDeviceToHost, s->ctx->stream));
    NB_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return NPDRB_OK;
}

int npdrb_session_copy_distances(npdrb_session *s, double *D_host) {
    if (!s->dD) { npdrb_set_error("copy_distances: no distances"); return NPDRB_EINVAL; }
    NB_CUDA(cudaSetDevice(s->ctx->device));
    NB_CUDA(cudaMemcpy2DAsync(D_host, sizeof(double) * (size_t)s->N, s->dD, sizeof(double) * (size_t)s->ldd,
                              sizeof(double) * (size_t)s->N, (size_t)s->nrows, cudaMemcpyDeviceToHost, s->ctx->stream));
    NB_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return NPDRB_OK;
}

int npdrb_session_import_pairs(npdrb_session *s, const int32_t *dRi, const int32_t *dNN, int64_t M) {
    if (M < 0 || (M > 0 && (!dRi || !dNN))) { npdrb_set_error("import_pairs: bad arguments"); return NPDRB_EINVAL; }
    if (s->own_pairs) { cudaFree(s->dRi); cudaFree(s->dNN); }
    s->dRi = const_cast<int32_t *>(dRi); s->dNN = const_cast<int32_t *>(dNN); s->M = M; s->own_pairs = false;
    s->streams_valid = false;
    return NPDRB_OK;
}

int npdrb_session_import_pairs_host(npdrb_session *s, const int32_t *Ri, const int32_t *NN, int64_t M) {
    if (M < 0 || (M > 0 && (!Ri || !NN))) { npdrb_set_error("import_pairs_host: bad arguments"); return NPDRB_EINVAL; }
    for (int64_t m = 0; m < M; ++m) {
        if (Ri[m] < 1 || Ri[m] > s->N || NN[m] < 1 || NN[m] > s->N) {
            npdrb_set_error("import_pairs_host: pair %lld = (%d, %d) out of range 1..%lld", (long long)m, Ri[m], NN[m], (long long)s->N);
            return NPDRB_EINVAL;
        }
    }
    NB_CUDA(cudaSetDevice(s->ctx->device));
    if (s->own_pairs) { cudaFree(s->dRi); cudaFree(s->dNN); }
    s->dRi = nullptr; s->dNN = nullptr;
    NB_CUDA(cudaMalloc(&s->dRi, sizeof(int32_t) * (size_t)(M > 0 ? M : 1)));
    NB_CUDA(cudaMalloc(&s->dNN, sizeof(int32_t) * (size_t)(M > 0 ? M : 1)));
    s->own_pairs = true; s->M = M; s->streams_valid = false;
    NB_CUDA(cudaMemcpyAsync(s->dRi, Ri, sizeof(int32_t) * (size_t)M, cudaMemcpyHostToDevice, s->ctx->stream));
    NB_CUDA(cudaMemcpyAsync(s->dNN, NN, sizeof(int32_t) * (size_t)M, cudaMemcpyHostToDevice, s->ctx->stream));
    NB_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return NPDRB_OK;
}

int npdrb_session_unique_pairs(npdrb_session *s, int32_t keep_unique_only, int64_t *n_unique) {
    NB_CUDA(cudaSetDevice(s->ctx->device));
    return nb_unique_pairs(s, keep_unique_only, n_unique);
}

int npdrb_session_regress(npdrb_session *s, int64_t attr0, int64_t nattr, int32_t regression_type, int32_t attr_diff_type,
                          const int32_t *covar_diff_type, double *stats_out, int32_t *irls_passes) {
    NB_CUDA(cudaSetDevice(s->ctx->device));
    return nb_regress(s, attr0, nattr, regression_type, attr_diff_type, covar_diff_type, stats_out, irls_passes);
}

int npdrb_session_copy_pairs(npdrb_session *s, int32_t *Ri_host, int32_t *NN_host) {
    if (!s->dRi && s->M > 0) { npdrb_set_error("copy_pairs: no pair list"); return NPDRB_EINVAL; }
    NB_CUDA(cudaSetDevice(s->ctx->device));
    NB_CUDA(cudaMemcpyAsync(Ri_host, s->dRi, sizeof(int32_t) * (size_t)s->M, cudaMemcpyDeviceToHost, s->ctx->stream));
    NB_CUDA(cudaMemcpyAsync(NN_host, s->dNN, sizeof(int32_t) * (size_t)s->M, cudaMemcpyDeviceToHost, s->ctx->stream));
    NB_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return NPDRB_OK;
}

int npdrb_session_stage_ms(npdrb_session *s, double *ms_distance, double *ms_neighbors, double *ms_prepare, double *ms_regress) {
    if (ms_distance) *ms_distance = s->ms_distance;
    if (ms_neighbors) *ms_neighbors = s->ms_neighbors;
    if (ms_prepare) *ms_prepare = s->ms_prepare;
    if (ms_regress) *ms_regress = s->ms_regress;
    return NPDRB_OK;
}

int npdrb_session_kernel_ms(npdrb_session *s, double *out5) {
    out5[0] = s->ms_pass_first; out5[1] = s->ms_pass_full; out5[2] = (double)s->n_pass_full; out5[3] = s->evals_pass_full;
    double rec = 0;
    for (int i = 0; i < s->n_streams; ++i) rec += (double)s->streams[i].n_rec;
    out5[4] = rec;
    return NPDRB_OK;
}

// ------------------------------------------------------------------ host-buffer operator API
int npdrb_session_regress_design(npdrb_session *s, int32_t *q_eff, int32_t *n_streams, int32_t *n_exact_stops) {
    if (q_eff) *q_eff = s->q_eff;
    if (n_streams) *n_streams = s->n_streams;
    if (n_exact_stops) *n_exact_stops = s->n_exact_stops;
    return NPDRB_OK;
}

int npdrb_diff(npdrb_ctx *ctx, const double *a, const double *b, int64_t n, int32_t diff_type, double norm_fac, double *out) {
    if (!ctx) { npdrb_set_error("npdrb_diff: null context (no device)"); return NPDRB_ENODEVICE; }
    if (diff_type < 0 || diff_type > NPDRB_DIFF_RAW) { npdrb_set_error("npdrb_diff: unknown diff type %d", diff_type); return NPDRB_EINVAL; }
    if (n <= 0) return NPDRB_OK;
    NB_CUDA(cudaSetDevice(ctx->device));
    double *d = nullptr;
    NB_CUDA(cudaMalloc(&d, sizeof(double) * 3 * (size_t)n));
    NB_CUDA(cudaMemcpyAsync(d, a, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    NB_CUDA(cudaMemcpyAsync(d + n, b, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = nb_diff_vectors(ctx, d, d + n, n, diff_type, norm_fac, d + 2 * n);
    if (rc == NPDRB_OK) {
        cudaError_t e = cudaMemcpyAsync(out, d + 2 * n, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { npdrb_set_error("npdrb_diff: %s", cudaGetErrorString(e)); rc = NPDRB_ECUDA; }
    }
    cudaFree(d);
    return rc;
}

int npdrb_distances(npdrb_ctx *ctx, const double *X, int64_t N, int64_t p, int32_t metric, int32_t fast_euclidean, double *D_out) {
    if (!ctx) { npdrb_set_error("npdrb_distances: null context (no device)"); return NPDRB_ENODEVICE; }
    npdrb_session *s = nullptr;
    NB_CHECK(npdrb_session_create(ctx, N, p, 0, &s));
    int rc = npdrb_session_upload(s, X, nullptr, nullptr);
    if (!rc) rc = nb_distances(s, metric, fast_euclidean, 0, N);
    if (!rc) rc = npdrb_session_copy_distances(s, D_out);     // symmetric: row-major block == column-major matrix
    npdrb_session_destroy(s);
    return rc;
}

// X2 (m2 x p, host, column-major) -> device copy with the leading dimension padded like the session's X
static int upload_padded(npdrb_ctx *ctx, const double *X, int64_t m, int64_t p, double **dX_out, int64_t *ld_out) {
    const int64_t ld = nb_round_up(m, 16);
    double *d = nullptr;
    NB_CUDA(cudaMalloc(&d, sizeof(double) * (size_t)ld * (size_t)p));
    if (ld != m) NB_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * (size_t)ld * (size_t)p, ctx->stream));
    NB_CUDA(cudaMemcpy2DAsync(d, sizeof(double) * (size_t)ld, X, sizeof(double) * (size_t)m, sizeof(double) * (size_t)m, (size_t)p,
                              cudaMemcpyHostToDevice, ctx->stream));
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    int rc = check_finite(ctx, d, ld * p, "the attribute matrix");
    if (rc) { cudaFree(d); return rc; }
    *dX_out = d; *ld_out = ld;
    return NPDRB_OK;
}

// distances block of a rectangular problem: rows = X2 instances, columns = X1 instances (session N = m1)
static int rect_block(npdrb_session *s, const double *X1, const double *X2, int64_t m2, int32_t metric) {
    npdrb_ctx *ctx = s->ctx;
    const int64_t m1 = s->N, p = s->p;
    int32_t base_metric = metric;
    std::vector<double> h1, h2;
    if (metric == NPDRB_METRIC_ALLELE_SHARING_MANHATTAN) {          // attr.mat / 2 on both sides (R/npdrLearner.R:330-332)
        h1.resize((size_t)m1 * p); h2.resize((size_t)m2 * p);
        for (size_t i = 0; i < h1.size(); ++i) h1[i] = X1[i] / 2.0;
        for (size_t i = 0; i < h2.size(); ++i) h2[i] = X2[i] / 2.0;
        X1 = h1.data(); X2 = h2.data();
        base_metric = NPDRB_METRIC_MANHATTAN;
    } else if (metric != NPDRB_METRIC_MANHATTAN && metric != NPDRB_METRIC_EUCLIDEAN) {
        npdrb_set_error("distances2: metric must be manhattan, euclidean or allele-sharing-manhattan");
        return NPDRB_EINVAL;
    }
    NB_CHECK(npdrb_session_upload(s, X1, nullptr, nullptr));
    NB_CHECK(nb_wait_upload(s));
    double *dX2 = nullptr; int64_t ld2 = 0;
    NB_CHECK(upload_padded(ctx, X2, m2, p, &dX2, &ld2));
    if (s->own_D) cudaFree(s->dD);
    s->ldd = nb_round_up(m1, 2);
    s->dD = nullptr;
    cudaError_t e = cudaMalloc(&s->dD, sizeof(double) * (size_t)m2 * (size_t)s->ldd);
    if (e != cudaSuccess) { cudaFree(dX2); npdrb_set_error("distances2: out of device memory"); return NPDRB_ENOMEM; }
    s->own_D = true; s->row0 = 0; s->nrows = m2; s->rect = true;
    int rc = nb_distances2(ctx, s->dX, s->ldx, m1, dX2, ld2, m2, p, base_metric, 0, s->dD, s->ldd);
    cudaFree(dX2);
    return rc;
}

int npdrb_distances2(npdrb_ctx *ctx, const double *X1, int64_t m1, const double *X2, int64_t m2, int64_t p, int32_t metric,
                     double *D_out) {
    if (!ctx) { npdrb_set_error("npdrb_distances2: null context (no device)"); return NPDRB_ENODEVICE; }
    if (m2 < 1) { npdrb_set_error("npdrb_distances2: empty second matrix"); return NPDRB_EINVAL; }
    npdrb_session *s = nullptr;
    NB_CHECK(npdrb_session_create(ctx, m1, p, 0, &s));
    int rc = rect_block(s, X1, X2, m2, metric);
    // block row i (X2 instance) holds the m1 distances to X1's instances = column i of the m1 x m2 column-major result
    if (!rc) {
        cudaError_t e = cudaMemcpy2DAsync(D_out, sizeof(double) * (size_t)m1, s->dD, sizeof(double) * (size_t)s->ldd,
                                          sizeof(double) * (size_t)m1, (size_t)m2, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { npdrb_set_error("npdrb_distances2: %s", cudaGetErrorString(e)); rc = NPDRB_ECUDA; }
    }
    npdrb_session_destroy(s);
    return rc;
}

int npdrb_nearest_neighbors2(npdrb_ctx *ctx, const double *X1, int64_t m1, const double *X2, int64_t m2, int64_t p,
                             int32_t nbd_method, int32_t nbd_metric, const double *sd_vec, double sd_frac, int64_t k,
                             int64_t **offsets, int32_t **idx, int64_t *M, double *radius_out) {
    if (!ctx) { npdrb_set_error("npdrb_nearest_neighbors2: null context (no device)"); return NPDRB_ENODEVICE; }
    if (m2 < 1) { npdrb_set_error("npdrb_nearest_neighbors2: empty second matrix"); return NPDRB_EINVAL; }
    if (offsets) *offsets = nullptr;
    if (idx) *idx = nullptr;
    if (!offsets || !idx || !M) { npdrb_set_error("npdrb_nearest_neighbors2: null output pointer"); return NPDRB_EINVAL; }
    npdrb_session *s = nullptr;
    NB_CHECK(npdrb_session_create(ctx, m1, p, 0, &s));
    int rc = rect_block(s, X1, X2, m2, nbd_metric);
    if (!rc && sd_vec) {                                             // one sd per X2 instance
        cudaError_t e = cudaMalloc(&s->d_sd_vec, sizeof(double) * (size_t)m2);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_sd_vec, sd_vec, sizeof(double) * (size_t)m2, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { npdrb_set_error("npdrb_nearest_neighbors2: %s", cudaGetErrorString(e)); rc = NPDRB_ECUDA; }
    }
    int64_t Ml = 0, ne = 0;
    if (!rc) rc = nb_neighbors(s, nbd_method, sd_frac, k, &Ml, &ne);
    if (!rc) {
        *offsets = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m2 + 1));
        *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(Ml > 0 ? Ml : 1));
        if (!*offsets || !*idx) { npdrb_set_error("out of host memory"); rc = NPDRB_ENOMEM; }
    }
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(*offsets, s->d_rowptr, sizeof(int64_t) * (size_t)(m2 + 1), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && Ml > 0) e = cudaMemcpyAsync(*idx, s->dNNL, sizeof(int32_t) * (size_t)Ml, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && radius_out && nbd_method != NPDRB_NBD_RELIEFF)
            e = cudaMemcpyAsync(radius_out, s->d_radius, sizeof(double) * (size_t)m2, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { npdrb_set_error("npdrb_nearest_neighbors2: %s", cudaGetErrorString(e)); rc = NPDRB_ECUDA; }
        *M = Ml;
    }
    if (rc) { free(*offsets); free(*idx); *offsets = nullptr; *idx = nullptr; }   // nothing is handed out on failure
    npdrb_session_destroy(s);
    return rc;
}

static int pairs_to_host(npdrb_session *s, int32_t **Ri, int32_t **NN) {
    const int64_t M = s->M;
    *Ri = (int32_t *)malloc(sizeof(int32_t) * (size_t)(M > 0 ? M : 1));
    *NN = (int32_t *)malloc(sizeof(int32_t) * (size_t)(M > 0 ? M : 1));
    int rc = NPDRB_OK;
    if (!*Ri || !*NN) { npdrb_set_error("out of host memory for %lld pairs", (long long)M); rc = NPDRB_ENOMEM; }
    if (!rc && M > 0) rc = npdrb_session_copy_pairs(s, *Ri, *NN);
    if (rc) { free(*Ri); free(*NN); *Ri = nullptr; *NN = nullptr; }              // nothing is handed out on failure
    return rc;
}

// distances (or adoption of a precomputed matrix) + neighbours on one device; leaves the list imported in s
static int neighbors_stage(npdrb_session *s, const double *X_or_D_host, int32_t nbd_method, int32_t nbd_metric,
                           const double *sd_vec, double sd_frac, int64_t k, int32_t fast_euclidean,
                           int64_t *M, int64_t *n_empty, int32_t hitmiss = 0) {
    npdrb_ctx *ctx = s->ctx;
    const int64_t N = s->N;
    if (nbd_metric == NPDRB_METRIC_PRECOMPUTED) {
        // attr.mat is the distance matrix (R/nearestNeighbors.R:171-176); column Ri of D is row Ri of the block
        if (!X_or_D_host) { npdrb_set_error("precomputed metric needs the distance matrix"); return NPDRB_EINVAL; }
        double *dD = nullptr;
        NB_CUDA(cudaMalloc(&dD, sizeof(double) * (size_t)N * (size_t)N));
        NB_CUDA(cudaMemcpyAsync(dD, X_or_D_host, sizeof(double) * (size_t)N * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
        NB_CUDA(cudaStreamSynchronize(ctx->stream));
        int rc = check_finite(ctx, dD, N * N, "the distance matrix");
        if (rc) { cudaFree(dD); return rc; }
        if (s->own_D) cudaFree(s->dD);
        s->dD = dD; s->ldd = N; s->row0 = 0; s->nrows = N; s->own_D = true;
    } else {
        NB_CHECK(nb_distances(s, nbd_metric, fast_euclidean, 0, N));
    }
    NB_CHECK(npdrb_session_set_sd_vec(s, sd_vec));
    if (hitmiss) NB_CHECK(nb_neighbors_hitmiss(s, nbd_method, sd_frac, k, M, n_empty));
    else NB_CHECK(nb_neighbors(s, nbd_method, sd_frac, k, M, n_empty));
    NB_CHECK(npdrb_session_import_pairs(s, s->dRiL, s->dNNL, s->M_local));
    return NPDRB_OK;
}

int npdrb_nearest_neighbors(npdrb_ctx *ctx, const double *X, int64_t N, int64_t p, int32_t nbd_method, int32_t nbd_metric,
                            const double *sd_vec, double sd_frac, int64_t k, int32_t unique_sampling, int32_t **Ri,
                            int32_t **NN, int64_t *M, int64_t *n_unique, int64_t *n_empty, double *radius_out) {
    if (!ctx) { npdrb_set_error("npdrb_nearest_neighbors: null context (no device)"); return NPDRB_ENODEVICE; }
    npdrb_session *s = nullptr;
    const bool pre = nbd_metric == NPDRB_METRIC_PRECOMPUTED;
    NB_CHECK(npdrb_session_create(ctx, N, pre ? N : p, 0, &s));
    int rc = pre ? NPDRB_OK : npdrb_session_upload(s, X, nullptr, nullptr);
    int64_t Ml = 0, ne = 0, nu = 0;
    if (!rc) rc = neighbors_stage(s, X, nbd_method, nbd_metric, sd_vec, sd_frac, k, 0, &Ml, &ne);
    if (!rc && radius_out && nbd_method != NPDRB_NBD_RELIEFF) rc = npdrb_session_copy_radius(s, radius_out);
    if (!rc) rc = nb_unique_pairs(s, unique_sampling, &nu);
    if (!rc) rc = pairs_to_host(s, Ri, NN);
    if (!rc) { *M = s->M; if (n_unique) *n_unique = nu; if (n_empty) *n_empty = ne; }
    npdrb_session_destroy(s);
    return rc;
}

int npdrb_nearest_neighbors_hitmiss(npdrb_ctx *ctx, const double *X, int64_t N, int64_t p, const double *pheno,
                                    int32_t nbd_method, int32_t nbd_metric, double sd_frac, int64_t k,
                                    int32_t unique_sampling, int32_t **Ri, int32_t **NN, int64_t *M, int64_t *n_unique,
                                    double *r_hit_out, double *r_miss_out) {
    if (!ctx) { npdrb_set_error("npdrb_nearest_neighbors_hitmiss: null context (no device)"); return NPDRB_ENODEVICE; }
    if (!pheno) { npdrb_set_error("npdrb_nearest_neighbors_hitmiss: class labels missing"); return NPDRB_EINVAL; }
    npdrb_session *s = nullptr;
    const bool pre = nbd_metric == NPDRB_METRIC_PRECOMPUTED;
    NB_CHECK(npdrb_session_create(ctx, N, pre ? N : p, 0, &s));
    int rc = npdrb_session_upload(s, pre ? nullptr : X, pheno, nullptr);
    int64_t Ml = 0, ne = 0, nu = 0;
    if (!rc) rc = neighbors_stage(s, X, nbd_method, nbd_metric, nullptr, sd_frac, k, 0, &Ml, &ne, 1);
    if (!rc && r_hit_out && nbd_method != NPDRB_NBD_RELIEFF) rc = npdrb_session_copy_radius(s, r_hit_out);
    if (!rc && r_miss_out && nbd_method != NPDRB_NBD_RELIEFF) rc = npdrb_session_copy_radius2(s, r_miss_out);
    if (!rc) rc = nb_unique_pairs(s, unique_sampling, &nu);
    if (!rc) rc = pairs_to_host(s, Ri, NN);
    if (!rc) { *M = s->M; if (n_unique) *n_unique = nu; }
    npdrb_session_destroy(s);
    return rc;
}

int npdrb_regress(npdrb_ctx *ctx, const double *X, int64_t N, int64_t p, const int32_t *Ri, const int32_t *NN, int64_t M,
                  const double *y, const double *C, int32_t q, const int32_t *covar_diff_type, int32_t attr_diff_type,
                  int32_t regression_type, double *stats_out) {
    if (!ctx) { npdrb_set_error("npdrb_regress: null context (no device)"); return NPDRB_ENODEVICE; }
    npdrb_session *s = nullptr;
    NB_CHECK(npdrb_session_create(ctx, N, p, q, &s));
    int rc = npdrb_session_upload(s, X, y, C);
    if (!rc) rc = npdrb_session_import_pairs_host(s, Ri, NN, M);
    if (!rc) rc = nb_regress(s, 0, p, regression_type, attr_diff_type, covar_diff_type, stats_out, nullptr);
    npdrb_session_destroy(s);
    return rc;
}

int npdrb_npdr(npdrb_ctx *ctx, const double *X, int64_t N, int64_t p, const double *y, const double *C,
               const npdrb_opts *o, npdrb_result *res) {
    if (!ctx) { npdrb_set_error("npdrb_npdr: null context (no device)"); return NPDRB_ENODEVICE; }
    if (!o || !res) { npdrb_set_error("npdrb_npdr: null opts/result"); return NPDRB_EINVAL; }
    memset(res, 0, sizeof(*res));
    if (o->nbd_metric == NPDRB_METRIC_PRECOMPUTED) {
        npdrb_set_error("npdrb_npdr: precomputed distances go through npdrb_nearest_neighbors + npdrb_regress");
        return NPDRB_EINVAL;
    }
    cudaEvent_t t0, t1, t2;
    NB_CUDA(cudaSetDevice(ctx->device));
    NB_CUDA(cudaEventCreate(&t0)); NB_CUDA(cudaEventCreate(&t1)); NB_CUDA(cudaEventCreate(&t2));
    NB_CUDA(cudaEventRecord(t0, ctx->stream));
    npdrb_session *s = nullptr;
    int rc = npdrb_session_create(ctx, N, p, o->n_covars, &s);
    if (rc) return rc;
    rc = npdrb_session_upload(s, X, y, C);
    if (!rc) { cudaEventRecord(t1, ctx->stream); }
    int64_t M = 0, ne = 0, nu = 0;
    int64_t k = o->knn;
    if (o->nbd_method == NPDRB_NBD_RELIEFF && k <= 0) k = npdrb_knn_surf(N, o->msurf_sd_frac);
    if (!rc) rc = neighbors_stage(s, nullptr, o->nbd_method, o->nbd_metric, nullptr, o->msurf_sd_frac, k, o->fast_euclidean, &M, &ne,
                                  o->separate_hitmiss_nbds);
    if (!rc) rc = nb_unique_pairs(s, o->neighbor_sampling_unique, &nu);
    if (!rc && s->M <= 0) { npdrb_set_error("npdr: no neighbour pairs found"); rc = NPDRB_EDATA; }
    int32_t passes = 0;
    if (!rc) {
        res->stats = (double *)malloc(sizeof(double) * (size_t)p * NPDRB_STATS_PER_ATTR);
        if (!res->stats) { npdrb_set_error("out of host memory"); rc = NPDRB_ENOMEM; }
    }
    if (!rc) rc = nb_regress(s, 0, p, o->regression_type, o->attr_diff_type, o->covar_diff_type, res->stats, &passes);
    if (!rc && o->return_pairs) rc = pairs_to_host(s, &res->Ri, &res->NN);
    if (!rc) {
        cudaEventRecord(t2, ctx->stream);
        cudaEventSynchronize(t2);
        float ms = 0;
        cudaEventElapsedTime(&ms, t0, t1); res->ms_h2d = ms;
        cudaEventElapsedTime(&ms, t0, t2); res->ms_total = ms;
        res->n_attr = p; res->n_pairs = M; res->n_unique_pairs = nu; res->n_pairs_used = s->M;
        res->n_empty = ne; res->knn_used = (o->nbd_method == NPDRB_NBD_RELIEFF) ? k : 0;
        res->k_ave_empirical = (N - ne) > 0 ? (double)M / (double)(N - ne) : 0.0;   // mean(knnVec()): only Ri that appear
        res->ms_distance = s->ms_distance; res->ms_neighbors = s->ms_neighbors;
        res->ms_prepare = s->ms_prepare; res->ms_regress = s->ms_regress; res->irls_passes = passes;
    } else {
        npdrb_result_release(res);
    }
    npdrb_session_destroy(s);
    cudaEventDestroy(t0); cudaEventDestroy(t1); cudaEventDestroy(t2);
    return rc;
}

int npdrb_npdr_k_grid(npdrb_ctx *ctx, const double *X, int64_t N, int64_t p, const double *y, const double *C,
                      const npdrb_opts *o, const int64_t *k_grid, int32_t n_k, double *stats_out, int64_t *n_pairs_out) {
    // vwok's foreach over k (R/adaptive.R:85-111): one upload, one distance matrix, one (d, row) sort of every row and one
    // instance-major attribute copy serve the whole grid; per k only the list prefix, the pair stream and the fit are redone.
    if (!ctx) { npdrb_set_error("npdrb_npdr_k_grid: null context (no device)"); return NPDRB_ENODEVICE; }
    if (!o || !k_grid || n_k < 1 || !stats_out) { npdrb_set_error("npdrb_npdr_k_grid: bad arguments"); return NPDRB_EINVAL; }
    if (o->nbd_method != NPDRB_NBD_RELIEFF) { npdrb_set_error("npdrb_npdr_k_grid: the k grid is defined for nbd.method = relieff"); return NPDRB_EINVAL; }
    if (o->nbd_metric == NPDRB_METRIC_PRECOMPUTED) { npdrb_set_error("npdrb_npdr_k_grid: precomputed distances are not supported here"); return NPDRB_EINVAL; }
    for (int32_t c = 0; c < n_k; ++c)
        if (k_grid[c] < 1 || k_grid[c] > N - 1) { npdrb_set_error("npdrb_npdr_k_grid: k = %lld outside 1..N-1", (long long)k_grid[c]); return NPDRB_EINVAL; }
    NB_CUDA(cudaSetDevice(ctx->device));
    npdrb_session *s = nullptr;
    int rc = npdrb_session_create(ctx, N, p, o->n_covars, &s);
    if (rc) return rc;
    rc = npdrb_session_upload(s, X, y, C);
    if (!rc) rc = nb_distances(s, o->nbd_metric, o->fast_euclidean, 0, N);
    if (!rc && !o->separate_hitmiss_nbds) rc = nb_sort_rows(s);
    for (int32_t c = 0; c < n_k && !rc; ++c) {
        int64_t M = 0, ne = 0, nu = 0;
        rc = o->separate_hitmiss_nbds ? nb_neighbors_hitmiss(s, NPDRB_NBD_RELIEFF, o->msurf_sd_frac, k_grid[c], &M, &ne)
                                      : nb_neighbors_sorted(s, k_grid[c], &M, &ne);
        if (!rc) rc = npdrb_session_import_pairs(s, s->dRiL, s->dNNL, s->M_local);
        if (!rc) rc = nb_unique_pairs(s, o->neighbor_sampling_unique, &nu);
        if (!rc && s->M <= 0) { npdrb_set_error("npdr (k = %lld): no neighbour pairs found", (long long)k_grid[c]); rc = NPDRB_EDATA; }
        if (!rc) rc = nb_regress(s, 0, p, o->regression_type, o->attr_diff_type, o->covar_diff_type,
                                 stats_out + (size_t)c * (size_t)p * NPDRB_STATS_PER_ATTR, nullptr);
        if (!rc && n_pairs_out) n_pairs_out[c] = M;
    }
    npdrb_session_destroy(s);
    return rc;
}

int npdrb_unireg(npdrb_ctx *ctx, const double *X, int64_t N, int64_t p, const double *y, const double *C, int32_t q,
                 int32_t regression_type, double *stats_out) {
    // uniReg (R/utils.R:66-157): the same per-column GLM on the N raw rows -- the identity pair list (i, i)
    // with the RAW diff (the attribute value itself), outcome and covariates taken as they are.
    if (!ctx) { npdrb_set_error("npdrb_unireg: null context (no device)"); return NPDRB_ENODEVICE; }
    npdrb_session *s = nullptr;
    NB_CHECK(npdrb_session_create(ctx, N, p, q, &s));
    int rc = npdrb_session_upload(s, X, y, C);
    std::vector<int32_t> id((size_t)N);
    for (int64_t i = 0; i < N; ++i) id[(size_t)i] = (int32_t)(i + 1);
    if (!rc) rc = npdrb_session_import_pairs_host(s, id.data(), id.data(), N);
    if (!rc) rc = nb_regress(s, 0, p, regression_type, NPDRB_DIFF_RAW, nullptr, stats_out, nullptr);
    npdrb_session_destroy(s);
    return rc;
}

}  // extern "C"
