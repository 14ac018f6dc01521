// multi.cu -- the whole path on several GPUs of one node, driven from ONE host process (npdrb_npdr_multi).
//
// This is what an R session uses: R is single threaded and cannot be launched once per GPU, so the `.Call` spawns one host
// thread per device for the duration of the call.  The partitioning is the one of npdr_b200/distributed.py:
//   rows/G      distances + neighbour lists; D is symmetric, so each device computes its diagonal square and a
//               checkerboard half of every off-diagonal square and PULLS the other half from its peers' packed send
//               buffers over NVLink (cudaMemcpyPeerAsync), then unpacks (distance.cu: nb_distances_balanced / pack / unpack);
//   exchange    every device pulls the peers' (Ri, NN) segments into its copy of the full list (rank order = reference order);
//   attrs/G     regression of the device's attribute columns; statistics land in the caller's p x 8 array.
// No data-path reduction: distances, radii and pair lists are bit-identical to the one-device run.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace {

constexpr int TILE_ROWS = 128;      // row blocks start on distance-tile boundaries

// Host barrier of the device threads.  wait() returns the value of `failed` as the LAST arriver saw it, so every thread
// leaves the barrier with the same verdict: a thread that races ahead and fails in the next phase cannot make a slower
// peer quit while the others still expect it at the following barrier (the exit decision is collective).
struct Barrier {
    std::mutex mu; std::condition_variable cv; int n, waiting = 0; unsigned gen = 0; int verdict = 0;
    explicit Barrier(int n_) : n(n_) {}
    int wait(const std::atomic<int> &failed) {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned g = gen;
        if (++waiting == n) { verdict = failed.load(); waiting = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
        return verdict;                              // stable until all n threads have arrived at the NEXT barrier
    }
};

struct Worker {
    int dev = 0;
    npdrb_ctx *ctx = nullptr;
    npdrb_session *s = nullptr;
    int rc = NPDRB_OK;
    char err[600] = "";
    int64_t row0 = 0, nrows = 0, attr0 = 0, nattr = 0;
    int live = -1;                                   // index among the devices that own rows, or -1
    double *d_send = nullptr;
    std::vector<int64_t> send_cnt, recv_cnt;         // doubles per live peer
    int64_t M_local = 0, n_empty = 0, n_unique = 0;
    double surf0[4] = {0, 0, 0, 0}, surf1[4] = {0, 0, 0, 0};
    int64_t surf_band = 0;                           // distances of my block inside the SURF guard band
    int32_t irls_passes = 0;
    double ms_distance = 0, ms_neighbors = 0, ms_prepare = 0, ms_regress = 0;
    // block-cyclic upload + peer-pull pipeline (run_producer): this device's upload events (one per sub-slab), the events and
    // host flags the distance call waits for (one per global block), the stream the peer pulls are issued on
    std::vector<cudaEvent_t> ev_up;
    std::vector<void *> ev_slab;
    std::vector<cudaEvent_t> ev_pull;                // the events of ev_slab this worker owns (blocks pulled from peers)
    std::vector<int32_t> slab_ready;
    cudaStream_t pull_stream = nullptr;
    ~Worker() {
        for (cudaEvent_t e : ev_up) if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : ev_pull) if (e) cudaEventDestroy(e);
        if (pull_stream) cudaStreamDestroy(pull_stream);
    }
};

struct Shared {
    int G = 0;
    std::vector<Worker> w;
    std::vector<int> live_devs;                      // worker indices that own rows
    std::vector<int64_t> row_starts;                 // live + 1 bounds
    std::atomic<int> failed{0};
    Barrier *bar = nullptr;
    const double *X = nullptr, *y = nullptr, *C = nullptr;
    int64_t N = 0, p = 0;
    const npdrb_opts *o = nullptr;
    double *stats = nullptr;                         // p x 8, host
    int64_t k = 0;
    int64_t M_total = 0;
    // Pipelined replication of X (pipe_nb > 0).  Ownership of the attribute columns is block-cyclic: sub-slab c of device g is
    // global block c * G + g (pipe_nb columns).  Every device stages its sub-slabs one after the other over its own PCIe
    // link; as soon as sub-slab c of a peer is on that peer (up_ready), the block is pulled over NVLink; the distance kernel
    // of every device consumes block after block in attribute order while later sub-slabs are still being staged.
    int64_t pipe_nb = 0; int pipe_nsub = 0;          // pipe_nb > 0: pipeline on (value = width of block 0)
    std::vector<int64_t> pipe_start;                 // G * pipe_nsub + 1 column bounds of the blocks (multiples of 16 except the last)
    std::unique_ptr<std::atomic<int>[]> up_ready;    // [G][pipe_nsub]
    std::vector<double> hostD;                       // SURF exact replay: the gathered distance matrix
    double surf_radius_exact = 0.0;
};

// One context per device for the life of the process: the device-memory cache (api.cu) then survives between calls, which
// saves ~200 ms of cudaMalloc / cudaFree per call and device.  npdrb_multi_shutdown() releases them.
std::mutex g_ctx_mu;
npdrb_ctx *g_ctx[64] = {nullptr};
int get_ctx(int dev, npdrb_ctx **out) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (dev < 0 || dev >= 64) { npdrb_set_error("device %d out of range", dev); return NPDRB_ENODEVICE; }
    if (!g_ctx[dev]) NB_CHECK(npdrb_create(dev, &g_ctx[dev]));
    *out = g_ctx[dev];
    return NPDRB_OK;
}

void fail(Shared &S, Worker &w, int rc) {
    w.rc = rc;
    snprintf(w.err, sizeof(w.err), "device %d: %s", w.dev, npdrb_last_error());
    S.failed.store(1);
}
#define STEP(call) do { if (!S.failed.load() && w.rc == NPDRB_OK) { int rc__ = (call); if (rc__ != NPDRB_OK) fail(S, w, rc__); } } while (0)
#define CU(call) ([&]() -> int { cudaError_t e__ = (call); if (e__ != cudaSuccess) { npdrb_set_error("%s: %s", #call, cudaGetErrorString(e__)); return NPDRB_ECUDA; } return NPDRB_OK; }())

// The upload / pull side of the pipeline for device g (its own host thread; the device's worker thread is inside the
// distance call meanwhile and waits, block by block, for slab_ready + the block's event).
void run_producer(Shared &S, int g) {
    Worker &w = S.w[(size_t)g];
    const int G = S.G, nsub = S.pipe_nsub;
    const int64_t ldx = w.s->ldx;
    const std::vector<int64_t> &st = S.pipe_start;
    int rc = NPDRB_OK;
    if (cudaSetDevice(w.dev) != cudaSuccess) rc = NPDRB_ECUDA;
    for (int c = 0; c < nsub && rc == NPDRB_OK && !S.failed.load(); ++c) {
        const int64_t b_own = (int64_t)c * G + g;
        rc = npdrb_upload_columns(w.ctx, w.s->dX + st[(size_t)b_own] * ldx, ldx, S.X + st[(size_t)b_own] * S.N, S.N,
                                  st[(size_t)b_own + 1] - st[(size_t)b_own]);                      // pageable: pinned ring
        if (rc == NPDRB_OK) rc = CU(cudaEventRecord(w.ev_up[(size_t)c], w.ctx->copy_stream));
        S.up_ready[(size_t)g * nsub + c].store(1, std::memory_order_release);
        for (int r = 0; r < G && rc == NPDRB_OK; ++r) {
            const int64_t b = (int64_t)c * G + r;
            if (r != g) {
                const Worker &peer = S.w[(size_t)r];
                const auto t0 = std::chrono::steady_clock::now();
                while (!S.up_ready[(size_t)r * nsub + c].load(std::memory_order_acquire) && !S.failed.load()) {
                    std::this_thread::yield();
                    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) { npdrb_set_error("peer upload stalled"); rc = NPDRB_ECUDA; break; }
                }
                if (rc != NPDRB_OK || S.failed.load()) break;
                rc = CU(cudaStreamWaitEvent(w.pull_stream, peer.ev_up[(size_t)c], 0));
                if (rc == NPDRB_OK)
                    rc = CU(cudaMemcpyPeerAsync(w.s->dX + st[(size_t)b] * ldx, w.dev, peer.s->dX + st[(size_t)b] * ldx, peer.dev,
                                                sizeof(double) * (size_t)ldx * (size_t)(st[(size_t)b + 1] - st[(size_t)b]), w.pull_stream));
                if (rc == NPDRB_OK) rc = CU(cudaEventRecord((cudaEvent_t)w.ev_slab[(size_t)b], w.pull_stream));
            }
            if (rc == NPDRB_OK) __atomic_store_n(&w.slab_ready[(size_t)b], 1, __ATOMIC_RELEASE);
        }
    }
    if (rc != NPDRB_OK || S.failed.load()) {             // unblock everybody who waits for this device; the call fails as a whole
        if (rc != NPDRB_OK) fail(S, w, rc);
        for (int c = 0; c < nsub; ++c) S.up_ready[(size_t)g * nsub + c].store(1, std::memory_order_release);
        for (size_t b = 0; b < w.slab_ready.size(); ++b) __atomic_store_n(&w.slab_ready[b], 1, __ATOMIC_RELEASE);
    }
}

void run_worker(Shared &S, int g) {
    Worker &w = S.w[(size_t)g];
    const npdrb_opts &o = *S.o;
    const int W = (int)S.live_devs.size();
    STEP(CU(cudaSetDevice(w.dev)));
    STEP(get_ctx(w.dev, &w.ctx));
    STEP(CU(cudaSetDevice(w.dev)));
    if (w.ctx) {
        for (int h = 0; h < S.G; ++h)                // NVLink peer access (already-enabled / unsupported are not errors)
            if (S.w[(size_t)h].dev != w.dev) { cudaDeviceEnablePeerAccess(S.w[(size_t)h].dev, 0); cudaGetLastError(); }
    }
    STEP(npdrb_session_create(w.ctx, S.N, S.p, o.n_covars, &w.s));
    if (S.G == 1) {
        STEP(npdrb_session_upload(w.s, S.X, S.y, S.C));
    } else {
        // every device uploads ITS attribute columns over its own PCIe link (1/G of the matrix), then pulls the other
        // columns from its peers over NVLink: G x less host traffic than G full uploads
        STEP(npdrb_session_upload(w.s, nullptr, S.y, S.C));
        const int64_t ldx = nb_round_up(S.N, 16);
        double *dX = nullptr;
        STEP(CU(cudaMalloc(&dX, sizeof(double) * (size_t)ldx * (size_t)S.p)));
        if (dX && ldx != S.N) STEP(CU(cudaMemsetAsync(dX, 0, sizeof(double) * (size_t)ldx * (size_t)S.p, w.ctx->stream)));
        if (S.pipe_nb > 0) {
            // pipelined replication: events and flags first, the transfers themselves run beside the distance call (phase 1)
            STEP(CU(cudaStreamSynchronize(w.ctx->stream)));
            if (w.s && dX) { w.s->dX = dX; w.s->ldx = ldx; w.s->own_X = true; }
            else if (dX) cudaFree(dX);
            const size_t n_blocks = (size_t)S.pipe_nsub * (size_t)S.G;
            w.ev_up.assign((size_t)S.pipe_nsub, nullptr);
            w.ev_slab.assign(n_blocks, nullptr);
            w.slab_ready.assign(n_blocks, 0);
            STEP(CU(cudaStreamCreateWithFlags(&w.pull_stream, cudaStreamNonBlocking)));
            for (int c = 0; c < S.pipe_nsub; ++c) STEP(CU(cudaEventCreateWithFlags(&w.ev_up[(size_t)c], cudaEventDisableTiming)));
            for (size_t b = 0; b < n_blocks; ++b) {
                if ((int)(b % (size_t)S.G) == g) { w.ev_slab[b] = (void *)w.ev_up[b / (size_t)S.G]; continue; }
                cudaEvent_t e = nullptr;
                STEP(CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)));
                w.ev_pull.push_back(e);
                w.ev_slab[b] = (void *)e;
            }
            if (S.bar->wait(S.failed)) return;           // every device's buffer and events exist
        } else {
        if (dX && w.nattr > 0) {
            if (nb_host_pointer_is_pinned(S.X)) {
                STEP(CU(cudaMemcpy2DAsync(dX + w.attr0 * ldx, sizeof(double) * (size_t)ldx, S.X + w.attr0 * S.N, sizeof(double) * (size_t)S.N,
                                          sizeof(double) * (size_t)S.N, (size_t)w.nattr, cudaMemcpyHostToDevice, w.ctx->stream)));
            } else {                                 // pageable (R's heap): through this device's pinned ring, after the memset
                STEP(CU(cudaStreamSynchronize(w.ctx->stream)));
                STEP(nb_staged_h2d(w.ctx, dX + w.attr0 * ldx, ldx, S.X + w.attr0 * S.N, S.N, w.nattr));
                if (w.ctx && w.ctx->copy_stream) STEP(CU(cudaStreamSynchronize(w.ctx->copy_stream)));
            }
        }
        STEP(CU(cudaStreamSynchronize(w.ctx->stream)));
        if (w.s && dX) { w.s->dX = dX; w.s->ldx = ldx; w.s->own_X = true; }
        else if (dX) cudaFree(dX);
        if (S.bar->wait(S.failed)) return;
        for (int h = 0; h < S.G; ++h) {
            const Worker &peer = S.w[(size_t)h];
            if (h == g || peer.nattr == 0) continue;
            STEP(CU(cudaMemcpyPeerAsync(w.s->dX + peer.attr0 * ldx, w.dev, peer.s->dX + peer.attr0 * ldx, peer.dev,
                                        sizeof(double) * (size_t)ldx * (size_t)peer.nattr, w.ctx->stream)));
        }
        STEP(CU(cudaStreamSynchronize(w.ctx->stream)));
        STEP(nb_check_finite(w.ctx, w.s ? w.s->dX : nullptr, ldx * S.p, "the attribute matrix"));
        }
    }
    // ---- phase 1: this device's share of the distance matrix, packed halves for the peers
    std::thread producer;
    const bool piped = S.pipe_nb > 0 && S.G > 1 && !S.failed.load() && w.rc == NPDRB_OK && w.s;
    if (piped) {
        STEP(npdrb_session_set_device_slab_events_v(w.s, (int32_t)w.ev_slab.size(), S.pipe_start.data(), w.ev_slab.data(), w.slab_ready.data()));
        producer = std::thread(run_producer, std::ref(S), g);
    }
    if (w.live >= 0) {
        if (W == 1) STEP(nb_distances(w.s, o.nbd_metric, o.fast_euclidean, 0, S.N));
        else {
            STEP(nb_distances_balanced(w.s, o.nbd_metric, o.fast_euclidean, S.row_starts.data(), w.live, W));
            w.send_cnt.assign((size_t)W, 0); w.recv_cnt.assign((size_t)W, 0);
            STEP(nb_balanced_counts(w.s, w.send_cnt.data(), w.recv_cnt.data()));
            int64_t ns = 0;
            for (int64_t c : w.send_cnt) ns += c;
            STEP(CU(cudaMalloc(&w.d_send, sizeof(double) * (size_t)(ns > 0 ? ns : 1))));
            if (ns > 0) STEP(nb_balanced_pack(w.s, w.d_send));
        }
        if (w.s) w.ms_distance = w.s->ms_distance;
    }
    if (producer.joinable()) producer.join();
    if (S.bar->wait(S.failed)) return;
    // ---- phase 2: pull the halves the peers computed for my rows, unpack
    double *d_recv = nullptr;
    if (w.live >= 0 && W > 1) {
        int64_t nr = 0;
        for (int64_t c : w.recv_cnt) nr += c;
        STEP(CU(cudaMalloc(&d_recv, sizeof(double) * (size_t)(nr > 0 ? nr : 1))));
        int64_t roff = 0;
        for (int h = 0; h < W; ++h) {
            const int64_t n = w.recv_cnt[(size_t)h];
            if (n == 0) continue;
            const Worker &peer = S.w[(size_t)S.live_devs[(size_t)h]];
            int64_t soff = 0;                        // my segment inside the peer's send buffer (peers in rank order)
            for (int t = 0; t < w.live; ++t) soff += peer.send_cnt[(size_t)t];
            STEP(CU(cudaMemcpyPeerAsync(d_recv + roff, w.dev, peer.d_send + soff, peer.dev, sizeof(double) * (size_t)n, w.ctx->stream)));
            roff += n;
        }
        STEP(CU(cudaStreamSynchronize(w.ctx->stream)));
    }
    const int bad2 = S.bar->wait(S.failed);          // every pull has completed: send buffers can go
    if (w.d_send) { cudaFree(w.d_send); w.d_send = nullptr; }
    if (bad2) { if (d_recv) cudaFree(d_recv); return; }
    if (w.live >= 0 && W > 1) {
        int64_t nr = 0;
        for (int64_t c : w.recv_cnt) nr += c;
        if (nr > 0) STEP(nb_balanced_unpack(w.s, d_recv));
        cudaFree(d_recv);
    }
    // ---- SURF on row blocks: one global mean / sd from per-block moments (R/nearestNeighbors.R:234-241)
    const bool surf_multi = (o.nbd_method == NPDRB_NBD_SURF) && W > 1;
    if (surf_multi) {
        if (w.live >= 0) STEP(npdrb_session_surf_partial(w.s, 0.0, w.surf0));
        if (S.bar->wait(S.failed)) return;
        double m0[4] = {0, 0, 0, 0};
        for (int h : S.live_devs) for (int t = 0; t < 4; ++t) m0[t] += S.w[(size_t)h].surf0[t];
        const double mean_u = m0[1] / m0[2];
        if (w.live >= 0) STEP(npdrb_session_surf_partial(w.s, mean_u, w.surf1));
        if (S.bar->wait(S.failed)) return;
        double m1[4] = {0, 0, 0, 0};
        for (int h : S.live_devs) for (int t = 0; t < 4; ++t) m1[t] += S.w[(size_t)h].surf1[t];
        // radius from the combined moments + guard band; if a distance of any block falls inside the band the blocks are
        // gathered on the host and R's sequential long-double chains are replayed exactly (neighbors.cu)
        double radius = 0.0, eps = 0.0;
        STEP(npdrb_surf_radius_from_moments(m0, m1, S.N, o.msurf_sd_frac, &radius, &eps));
        if (w.live >= 0) STEP(npdrb_session_surf_band_count(w.s, radius - eps, radius + eps, &w.surf_band));
        if (S.bar->wait(S.failed)) return;
        int64_t in_band = 0;
        for (int h : S.live_devs) in_band += S.w[(size_t)h].surf_band;
        const char *force = getenv("NPDRB_SURF_FORCE_EXACT");
        if (in_band > 0 || (force && force[0] == '1')) {
            if (g == S.live_devs[0]) S.hostD.assign((size_t)S.N * (size_t)S.N, 0.0);
            if (S.bar->wait(S.failed)) return;
            if (w.live >= 0) STEP(npdrb_session_copy_distances(w.s, S.hostD.data() + (size_t)w.row0 * (size_t)S.N));
            if (S.bar->wait(S.failed)) return;
            if (g == S.live_devs[0]) {
                STEP(npdrb_surf_radius_exact_host(S.hostD.data(), S.N, S.N, o.msurf_sd_frac, &S.surf_radius_exact));
                std::vector<double>().swap(S.hostD);
            }
            if (S.bar->wait(S.failed)) return;
            radius = S.surf_radius_exact;
        }
        if (w.live >= 0) STEP(npdrb_session_set_surf_radius(w.s, radius));
    }
    // ---- phase 3: neighbour lists of my rows
    if (w.live >= 0) {
        if (o.separate_hitmiss_nbds) STEP(nb_neighbors_hitmiss(w.s, o.nbd_method, o.msurf_sd_frac, S.k, &w.M_local, &w.n_empty));
        else STEP(nb_neighbors(w.s, o.nbd_method, o.msurf_sd_frac, S.k, &w.M_local, &w.n_empty));
        if (w.s) w.ms_neighbors = w.s->ms_neighbors;
    }
    if (S.bar->wait(S.failed)) return;
    // ---- phase 4: the full pair list on every device (segments in rank order = reference order)
    int64_t M = 0;
    for (int h : S.live_devs) M += S.w[(size_t)h].M_local;
    int32_t *dRi = nullptr, *dNN = nullptr;
    if (M > 0 && w.nattr > 0) {
        STEP(CU(cudaMalloc(&dRi, sizeof(int32_t) * (size_t)M)));
        STEP(CU(cudaMalloc(&dNN, sizeof(int32_t) * (size_t)M)));
        int64_t off = 0;
        for (int h : S.live_devs) {
            const Worker &peer = S.w[(size_t)h];
            if (peer.M_local > 0) {
                STEP(CU(cudaMemcpyPeerAsync(dRi + off, w.dev, peer.s->dRiL, peer.dev, sizeof(int32_t) * (size_t)peer.M_local, w.ctx->stream)));
                STEP(CU(cudaMemcpyPeerAsync(dNN + off, w.dev, peer.s->dNNL, peer.dev, sizeof(int32_t) * (size_t)peer.M_local, w.ctx->stream)));
            }
            off += peer.M_local;
        }
        STEP(CU(cudaStreamSynchronize(w.ctx->stream)));
    }
    if (S.bar->wait(S.failed)) {                     // every device has its copy: the local segments may be reused
        if (dRi) cudaFree(dRi);
        if (dNN) cudaFree(dNN);
        return;
    }
    if (g == 0) S.M_total = M;
    // ---- phase 5: regression of my attribute columns
    if (M > 0 && w.nattr > 0) {
        STEP(npdrb_session_import_pairs(w.s, dRi, dNN, M));
        if (w.s) w.s->own_pairs = true;              // the session frees the list
        STEP(nb_unique_pairs(w.s, o.neighbor_sampling_unique, &w.n_unique));
        STEP(nb_regress(w.s, w.attr0, w.nattr, o.regression_type, o.attr_diff_type, o.covar_diff_type,
                        S.stats + (size_t)w.attr0 * NPDRB_STATS_PER_ATTR, &w.irls_passes));
        if (w.s) { w.ms_prepare = w.s->ms_prepare; w.ms_regress = w.s->ms_regress; }
    }
}

}  // namespace

extern "C" int npdrb_npdr_multi(int32_t n_devices, const int32_t *devices, const double *X, int64_t N, int64_t p, const double *y,
                                const double *C, const npdrb_opts *o, npdrb_result *res) {
    if (!o || !res) { npdrb_set_error("npdrb_npdr_multi: null opts/result"); return NPDRB_EINVAL; }
    memset(res, 0, sizeof(*res));
    if (n_devices < 1 || n_devices > 64) { npdrb_set_error("npdrb_npdr_multi: 1..64 devices"); return NPDRB_EINVAL; }
    if (o->nbd_metric == NPDRB_METRIC_PRECOMPUTED) { npdrb_set_error("npdrb_npdr_multi: precomputed distances are a one-device path"); return NPDRB_EINVAL; }
    if (o->nbd_method == NPDRB_NBD_SURF && o->separate_hitmiss_nbds && n_devices > 1) {
        npdrb_set_error("npdrb_npdr_multi: surf hit/miss radii need the whole distance matrix on one device");
        return NPDRB_EINVAL;
    }
    // NPDRB_MULTI_ALLOW_DUPLICATES=1 (tests): the same ordinal may be listed several times; the workers then share that
    // device's context and stream (stream order keeps the memory cache's reuse rule valid) and "peer" copies are local.
    const char *dup_env = getenv("NPDRB_MULTI_ALLOW_DUPLICATES");
    const bool allow_dup = dup_env && dup_env[0] == '1';
    for (int a = 0; a < n_devices && !allow_dup; ++a)
        for (int b = a + 1; b < n_devices; ++b)
            if (devices && devices[a] == devices[b]) { npdrb_set_error("npdrb_npdr_multi: device %d listed twice", devices[a]); return NPDRB_EINVAL; }
    Shared S;
    S.G = n_devices; S.X = X; S.y = y; S.C = C; S.N = N; S.p = p; S.o = o;
    S.w.resize((size_t)n_devices);
    S.k = o->knn;
    if (o->nbd_method == NPDRB_NBD_RELIEFF && S.k <= 0) S.k = npdrb_knn_surf(N, o->msurf_sd_frac);
    // row blocks (tile aligned, trailing devices may be empty) and attribute blocks, as distributed.py partitions them
    const int64_t n_tiles = nb_cdiv(N, TILE_ROWS);
    int64_t t0 = 0, a0 = 0;
    for (int g = 0; g < n_devices; ++g) {
        Worker &w = S.w[(size_t)g];
        w.dev = devices ? devices[g] : g;
        const int64_t nt = n_tiles / n_devices + (g < n_tiles % n_devices ? 1 : 0);
        w.row0 = (t0 * TILE_ROWS < N) ? t0 * TILE_ROWS : N;
        const int64_t row1 = ((t0 + nt) * TILE_ROWS < N) ? (t0 + nt) * TILE_ROWS : N;
        w.nrows = row1 - w.row0;
        t0 += nt;
        if (w.nrows > 0) { w.live = (int)S.live_devs.size(); S.live_devs.push_back(g); S.row_starts.push_back(w.row0); }
        w.attr0 = a0;
        w.nattr = p / n_devices + (g < p % n_devices ? 1 : 0);
        a0 += w.nattr;
    }
    S.row_starts.push_back(N);
    {
        // pipelined replication of X: every device owns rows (it consumes the blocks itself) and the columns split into
        // G x n_sub blocks of a multiple of 16 columns
        const char *np_ = getenv("NPDRB_NO_PIPELINED_GATHER");
        if (n_devices > 1 && (int)S.live_devs.size() == n_devices && !(np_ && np_[0] == '1')) {
            // G * n_sub blocks of (nearly) equal width, bounds on multiples of 16 columns (the distance kernel's chunk)
            const int64_t chunks = p / 16;
            int n_sub = (int)(chunks / n_devices < 5 ? chunks / n_devices : 5);
            if (n_sub >= 2) {
                const int64_t n_blocks = (int64_t)n_devices * n_sub;
                S.pipe_start.assign((size_t)n_blocks + 1, p);
                for (int64_t b = 0; b < n_blocks; ++b) S.pipe_start[(size_t)b] = 16 * (b * chunks / n_blocks);
                S.pipe_nb = S.pipe_start[1]; S.pipe_nsub = n_sub;
            }
            if (S.pipe_nb > 0) {
                S.up_ready.reset(new std::atomic<int>[(size_t)n_devices * (size_t)S.pipe_nsub]);
                for (size_t i = 0; i < (size_t)n_devices * (size_t)S.pipe_nsub; ++i) S.up_ready[i].store(0);
            }
        }
    }
    res->stats = (double *)malloc(sizeof(double) * (size_t)p * NPDRB_STATS_PER_ATTR);
    if (!res->stats) { npdrb_set_error("out of host memory"); return NPDRB_ENOMEM; }
    S.stats = res->stats;
    Barrier bar(n_devices);
    S.bar = &bar;
    const auto wall0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int g = 1; g < n_devices; ++g) th.emplace_back(run_worker, std::ref(S), g);
    run_worker(S, 0);
    for (auto &t : th) t.join();
    const double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();

    int rc = NPDRB_OK;
    for (const Worker &w : S.w)
        if (w.rc != NPDRB_OK && rc == NPDRB_OK) { rc = w.rc; npdrb_set_error("%s", w.err); }
    if (rc == NPDRB_OK && S.M_total <= 0) { npdrb_set_error("npdr: no neighbour pairs found"); rc = NPDRB_EDATA; }
    if (rc == NPDRB_OK && o->return_pairs) {
        const Worker *src = nullptr;
        for (const Worker &w : S.w) if (w.s && w.s->dRi && w.s->M > 0) { src = &w; break; }
        if (src) {
            cudaSetDevice(src->dev);
            res->Ri = (int32_t *)malloc(sizeof(int32_t) * (size_t)src->s->M);
            res->NN = (int32_t *)malloc(sizeof(int32_t) * (size_t)src->s->M);
            if (!res->Ri || !res->NN) { npdrb_set_error("out of host memory"); rc = NPDRB_ENOMEM; }
            else rc = npdrb_session_copy_pairs(src->s, res->Ri, res->NN);
        }
    }
    if (rc == NPDRB_OK) {
        int64_t ne = 0;
        for (const Worker &w : S.w) {
            ne += w.n_empty;
            if (w.ms_distance > res->ms_distance) res->ms_distance = w.ms_distance;
            if (w.ms_neighbors > res->ms_neighbors) res->ms_neighbors = w.ms_neighbors;
            if (w.ms_prepare > res->ms_prepare) res->ms_prepare = w.ms_prepare;
            if (w.ms_regress > res->ms_regress) res->ms_regress = w.ms_regress;
            if (w.irls_passes > res->irls_passes) res->irls_passes = w.irls_passes;
        }
        const Worker &w0 = S.w[0];
        res->n_attr = p; res->n_pairs = S.M_total; res->n_unique_pairs = w0.n_unique;
        res->n_pairs_used = (w0.s ? w0.s->M : S.M_total);
        res->n_empty = ne; res->knn_used = (o->nbd_method == NPDRB_NBD_RELIEFF) ? S.k : 0;
        res->k_ave_empirical = (N - ne) > 0 ? (double)S.M_total / (double)(N - ne) : 0.0;
        res->ms_total = wall_ms;                     // host wall clock around the G device threads
    }
    for (Worker &w : S.w) {
        if (w.ctx) cudaSetDevice(w.dev);
        if (w.s) npdrb_session_destroy(w.s);         // the contexts (and their memory caches) stay for the next call
    }
    if (rc != NPDRB_OK) npdrb_result_release(res);
    return rc;
}

extern "C" void npdrb_multi_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    for (int d = 0; d < 64; ++d)
        if (g_ctx[d]) { npdrb_destroy(g_ctx[d]); g_ctx[d] = nullptr; }
}
