// cgs_b200.cu -- C ABI (include/cgs_b200.h) over the sm_100a kernels.
#include "../../include/cgs_b200.h"

#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <mutex>
#include <numeric>
#include <thread>
#include <type_traits>
#include <vector>

#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>

#include "common.cuh"
#include "corpus_kernels.cuh"
#include "exchange_kernels.cuh"
#include "fragment_kernels.cuh"
#include "impute_kernels.cuh"
#include "metrics_kernels.cuh"
#include "normalize_kernels.cuh"
#include "ranksum_kernels.cuh"
#include "rankings_kernels.cuh"
#include "sampler_kernels.cuh"
#include "stats_kernels.cuh"

namespace cgs {
thread_local std::string g_last_error;

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// NVTX range for the lifetime of a scope (sweeps, exchanges, exports, log-likelihood show up by name in a timeline)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// Device buffers come from the device's stream-ordered pool (cudaMallocAsync): creating and
// destroying models inside a topic sweep then costs microseconds instead of the device-wide
// synchronisation of cudaMalloc / cudaFree.  `stream` is the stream that first uses the buffer
// (dmalloc) or last used it (dfree).  Buffers that are exported over CUDA IPC (AD-LDA deltas)
// use dmalloc_ipc: legacy IPC handles only exist for cudaMalloc memory.
template <typename T>
static int dmalloc(T **p, size_t n, cudaStream_t stream) {
    *p = nullptr;
    if (n == 0) n = 1;
    CGS_CUDA(cudaMallocAsync((void **)p, n * sizeof(T), stream));
    return 0;
}

template <typename T>
static void dfree(T *p, cudaStream_t stream) {
    if (p) cudaFreeAsync((void *)p, stream);
}

template <typename T>
static int dmalloc_ipc(T **p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    CGS_CUDA(cudaMalloc((void **)p, n * sizeof(T)));
    return 0;
}

static inline cudaError_t cgs_memcpy_sync(cudaStream_t stream, void *dst, const void *src, size_t n, cudaMemcpyKind kind) {
    cudaError_t e = cudaMemcpyAsync(dst, src, n, kind, stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(stream);
}

// CGS_B200_TIMING=1 prints host wall-clock of the construction stages to stderr (diagnostics only).
struct StageTimer {
    bool on;
    std::chrono::steady_clock::time_point t;
    StageTimer() : on(getenv("CGS_B200_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
    void lap(const char *what) {
        if (!on) return;
        cudaDeviceSynchronize();
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[cgs timing] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

// Keep freed blocks in the pool (the default threshold of 0 returns them to the driver at every
// synchronisation, which is what made cudaFree expensive in the first place).
static int configure_pool(int device) {
    cudaMemPool_t pool;
    CGS_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t keep = UINT64_MAX;
    CGS_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    return 0;
}

// Device -> pageable host memory at more than the few GB/s a plain cudaMemcpy gives there (its staging copy and the
// first-touch page faults of a fresh result array run on ONE host thread): the data is DMA'd into two pinned staging
// buffers in turn and a handful of host threads copy each buffer to its destination while the other one fills.
// copy() is ordered after everything already enqueued on `after` (an event is taken from it); finish() drains.
struct HostSink {
    static constexpr size_t kStage = (size_t)64 << 20;
    static constexpr int kThreads = 8;
    int device = 0;
    cudaStream_t stream = nullptr;
    uint8_t *stage[2] = {nullptr, nullptr};
    cudaEvent_t landed[2] = {nullptr, nullptr};
    cudaEvent_t src_read[2] = {nullptr, nullptr};
    cudaEvent_t ready = nullptr;
    std::thread mover[2];
    int next = 0;
    size_t stage_bytes = kStage;

    // total_bytes: what will travel through the sink (small transfers get small staging buffers: pinning costs time)
    int open(int device_, size_t total_bytes = (size_t)-1) {
        device = device_;
        stage_bytes = std::max<size_t>((size_t)1 << 20, std::min<size_t>(kStage, (total_bytes + 4095) & ~(size_t)4095));
        CGS_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CGS_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
        for (int b = 0; b < 2; ++b) {
            CGS_CUDA(cudaHostAlloc((void **)&stage[b], stage_bytes, cudaHostAllocDefault));
            CGS_CUDA(cudaEventCreateWithFlags(&landed[b], cudaEventDisableTiming));
            CGS_CUDA(cudaEventCreateWithFlags(&src_read[b], cudaEventDisableTiming));
        }
        return 0;
    }
    // dst[0, bytes) = d_src[0, bytes); d_src must stay untouched until finish() or wait_sources()
    int copy(void *dst, const void *d_src, size_t bytes, cudaStream_t after) {
        CGS_CUDA(cudaEventRecord(ready, after));
        CGS_CUDA(cudaStreamWaitEvent(stream, ready, 0));
        for (size_t off = 0; off < bytes; off += stage_bytes) {
            const size_t n = std::min(stage_bytes, bytes - off);
            const int b = next;
            next ^= 1;
            if (mover[b].joinable()) mover[b].join();  // the buffer's previous contents have reached the host array
            CGS_CUDA(cudaMemcpyAsync(stage[b], (const uint8_t *)d_src + off, n, cudaMemcpyDeviceToHost, stream));
            CGS_CUDA(cudaEventRecord(landed[b], stream));
            uint8_t *to = (uint8_t *)dst + off;
            const uint8_t *from = stage[b];
            cudaEvent_t ev = landed[b];
            const int dev = device;
            mover[b] = std::thread([=]() {
                cudaSetDevice(dev);
                cudaEventSynchronize(ev);
                const size_t per = ((n + kThreads - 1) / kThreads + 4095) & ~(size_t)4095;
                std::thread workers[kThreads];
                int started = 0;
                for (size_t o = 0; o < n; o += per) workers[started++] = std::thread([=]() { memcpy(to + o, from + o, std::min(per, n - o)); });
                for (int t = 0; t < started; ++t) workers[t].join();
            });
        }
        return 0;
    }
    // height rows of `width` bytes, contiguous on the device, to rows `dpitch` bytes apart on the host
    int copy2d(void *dst, size_t dpitch, const void *d_src, size_t width, size_t height, cudaStream_t after) {
        if (dpitch == width) return copy(dst, d_src, width * height, after);
        CGS_CUDA(cudaEventRecord(ready, after));
        CGS_CUDA(cudaStreamWaitEvent(stream, ready, 0));
        const size_t rows_per = std::max<size_t>(1, stage_bytes / width);
        if (width > stage_bytes) CGS_FAIL("row wider than the staging buffer");
        for (size_t r0 = 0; r0 < height; r0 += rows_per) {
            const size_t nr = std::min(rows_per, height - r0);
            const int b = next;
            next ^= 1;
            if (mover[b].joinable()) mover[b].join();
            CGS_CUDA(cudaMemcpyAsync(stage[b], (const uint8_t *)d_src + r0 * width, nr * width, cudaMemcpyDeviceToHost, stream));
            CGS_CUDA(cudaEventRecord(landed[b], stream));
            uint8_t *to = (uint8_t *)dst + r0 * dpitch;
            const uint8_t *from = stage[b];
            cudaEvent_t ev = landed[b];
            const int dev = device;
            mover[b] = std::thread([=]() {
                cudaSetDevice(dev);
                cudaEventSynchronize(ev);
                const size_t per = (nr + kThreads - 1) / kThreads;
                std::thread workers[kThreads];
                int started = 0;
                for (size_t o = 0; o < nr; o += per)
                    workers[started++] = std::thread([=]() {
                        for (size_t r = o; r < std::min(nr, o + per); ++r) memcpy(to + r * dpitch, from + r * width, width);
                    });
                for (int t = 0; t < started; ++t) workers[t].join();
            });
        }
        return 0;
    }
    // Source buffers are recycled by slot (0 / 1): mark_sources(slot) after the copies out of a buffer were issued,
    // wait_sources(slot, s) before stream `s` overwrites that buffer (no-op until the slot was marked).
    int mark_sources(int slot) {
        CGS_CUDA(cudaEventRecord(src_read[slot], stream));
        return 0;
    }
    int wait_sources(int slot, cudaStream_t s) {
        CGS_CUDA(cudaStreamWaitEvent(s, src_read[slot], 0));
        return 0;
    }
    void finish() {
        for (int b = 0; b < 2; ++b)
            if (mover[b].joinable()) mover[b].join();
    }
    void close() {
        finish();
        for (int b = 0; b < 2; ++b) {
            if (stage[b]) cudaFreeHost(stage[b]);
            if (landed[b]) cudaEventDestroy(landed[b]);
            if (src_read[b]) cudaEventDestroy(src_read[b]);
            stage[b] = nullptr; landed[b] = nullptr; src_read[b] = nullptr;
        }
        if (ready) cudaEventDestroy(ready);
        if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
        ready = nullptr; stream = nullptr;
    }
};

// Pageable host memory -> device, the mirror image: host threads fill one pinned staging buffer while the other one
// is DMA'd.  copy() returns when the last piece has been ENQUEUED on `stream` (the caller's stream).
struct HostSource {
    static constexpr size_t kStage = (size_t)64 << 20;
    static constexpr int kThreads = 8;
    uint8_t *stage[2] = {nullptr, nullptr};
    cudaEvent_t sent[2] = {nullptr, nullptr};
    int next = 0;
    size_t stage_bytes = kStage;

    int open(size_t total_bytes = (size_t)-1) {
        stage_bytes = std::max<size_t>((size_t)1 << 20, std::min<size_t>(kStage, (total_bytes + 4095) & ~(size_t)4095));
        for (int b = 0; b < 2; ++b) {
            CGS_CUDA(cudaHostAlloc((void **)&stage[b], stage_bytes, cudaHostAllocDefault));
            CGS_CUDA(cudaEventCreateWithFlags(&sent[b], cudaEventDisableTiming));
        }
        return 0;
    }
    int copy(void *d_dst, const void *src, size_t bytes, cudaStream_t stream) {
        for (size_t off = 0; off < bytes; off += stage_bytes) {
            const size_t n = std::min(stage_bytes, bytes - off);
            const int b = next;
            next ^= 1;
            CGS_CUDA(cudaEventSynchronize(sent[b]));  // the buffer's previous DMA has left (no-op the first time)
            const uint8_t *from = (const uint8_t *)src + off;
            uint8_t *to = stage[b];
            const size_t per = ((n + kThreads - 1) / kThreads + 4095) & ~(size_t)4095;
            std::thread workers[kThreads];
            int started = 0;
            for (size_t o = 0; o < n; o += per) workers[started++] = std::thread([=]() { memcpy(to + o, from + o, std::min(per, n - o)); });
            for (int t = 0; t < started; ++t) workers[t].join();
            CGS_CUDA(cudaMemcpyAsync((uint8_t *)d_dst + off, stage[b], n, cudaMemcpyHostToDevice, stream));
            CGS_CUDA(cudaEventRecord(sent[b], stream));
        }
        return 0;
    }
    // height rows of `width` bytes, `spitch` bytes apart on the host, to contiguous rows on the device
    int copy2d(void *d_dst, const void *src, size_t spitch, size_t width, size_t height, cudaStream_t stream) {
        if (spitch == width) return copy(d_dst, src, width * height, stream);
        if (width > stage_bytes) CGS_FAIL("row wider than the staging buffer");
        const size_t rows_per = std::max<size_t>(1, stage_bytes / width);
        for (size_t r0 = 0; r0 < height; r0 += rows_per) {
            const size_t nr = std::min(rows_per, height - r0);
            const int b = next;
            next ^= 1;
            CGS_CUDA(cudaEventSynchronize(sent[b]));
            const uint8_t *from = (const uint8_t *)src + r0 * spitch;
            uint8_t *to = stage[b];
            const size_t per = (nr + kThreads - 1) / kThreads;
            std::thread workers[kThreads];
            int started = 0;
            for (size_t o = 0; o < nr; o += per)
                workers[started++] = std::thread([=]() {
                    for (size_t r = o; r < std::min(nr, o + per); ++r) memcpy(to + r * width, from + r * spitch, width);
                });
            for (int t = 0; t < started; ++t) workers[t].join();
            CGS_CUDA(cudaMemcpyAsync((uint8_t *)d_dst + r0 * width, stage[b], nr * width, cudaMemcpyHostToDevice, stream));
            CGS_CUDA(cudaEventRecord(sent[b], stream));
        }
        return 0;
    }
    void close() {
        for (int b = 0; b < 2; ++b) {
            if (sent[b]) { cudaEventSynchronize(sent[b]); cudaEventDestroy(sent[b]); }
            if (stage[b]) cudaFreeHost(stage[b]);
            stage[b] = nullptr; sent[b] = nullptr;
        }
    }
};

}  // namespace cgs

using namespace cgs;

struct cgs_corpus {
    int device = 0;
    int num_sms = kNumSMsFallback;
    int32_t n_cells = 0, n_regions = 0;
    int64_t n_tokens = 0;
    int64_t token_offset = 0;
    // size of the WHOLE matrix when this corpus is one shard of it (cgs_corpus_set_global_size; 0 = not a shard)
    int64_t global_cells = 0, global_tokens = 0;
    int64_t *d_cell_ptr = nullptr;    // D+1
    int32_t *d_tok_region = nullptr;  // N
    int32_t *d_cell_order = nullptr;  // D, longest cell first
    // region blocks (cgs_corpus_set_region_blocks): block b = regions [block_bounds[b], block_bounds[b + 1]);
    // d_cell_split[(b - 1) * D + d] = first token of cell d at or above block_bounds[b], b = 1 .. n_blocks - 1
    int64_t *d_cell_split = nullptr;
    std::vector<int32_t> block_bounds;
    bool owns_tokens = true;
    int n_models = 0;
    // Construction runs on a private non-blocking stream: pool blocks freed on one non-blocking stream are
    // reused by another without a trip to the driver, which the legacy default stream does not get
    // (measured: 65 ms per 260 MB corpus build on stream 0, 7 ms here).
    cudaStream_t stream = nullptr;
};

struct cgs_model {
    cgs_corpus *corpus = nullptr;
    int32_t K = 0, Kp = 0;
    double alpha = 0, eta = 0;
    uint64_t seed = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    uint16_t *d_z = nullptr;
    int32_t *d_nwk = nullptr, *d_ndk = nullptr, *d_nk = nullptr;
    int32_t *d_snap = nullptr, *d_delta = nullptr;  // AD-LDA, allocated on first use
    int32_t *d_work = nullptr;                      // work-queue counters, one per sweep slot
    double *d_partials = nullptr;                   // reduction scratch
    int *d_flag = nullptr;
    unsigned long long *d_visits = nullptr;         // set only inside cgs_model_debug_sweep_visits
    int64_t launches = 0;
    int64_t sweeps = 0;
    int64_t work_slot = 0;
    int tpt = 1, ch = 1, threads = 128;
    double nk_divisor = 16.0;
    // Distinct uniform draws per sweep.  The reference reuses 131 072 numbers cyclically inside a sweep
    // (lda's rands[i % n_rand]), which correlates far-apart tokens and measurably speeds up the
    // transient; the same reuse density keeps the trace on the reference's (0 = every token its own draw).
    int64_t rng_period = 131072;
    int max_blocks = 0;   // user override, 0 = automatic
    int auto_blocks = 1 << 30;  // concurrency cap chosen by pick_sweep_shape
    int sweep_blocks = 0;
    // deterministic mode (cgs_model_set_deterministic): sweeps read copies of n_wk / n_k taken at the start of
    // each of det_parts sub-sweeps, one warp per cell
    int sweep_block = -1, sweep_x_ctas = 0;  // set around the launches of cgs_model_sweep_overlapped / _blocked
    bool fp64 = false;  // conditional in double (A/B build, cgs_model_set_conditional_precision)
    int det_parts = 0;
    int32_t *d_nwk_ro = nullptr, *d_nk_ro = nullptr;
    // asynchronous peer-memory exchange (cgs_model_exchange_*): arena = [delta | total | flags]
    char *d_arena = nullptr;
    size_t arena_bytes = 0;
    ExchangePeers xpeers;
    int x_ranks = 0, x_rank = 0;
    uint32_t x_round = 0;
    int x_sync_blocks = 0;
    int sweep_blocks_fused = 0;
    int sweep_blocks_det = 0;
    ExchangeParams x_fused;      // the exchange of the current fused launch
    int x_pending_block = -1;    // overlapped sweeps: the region block that was sampled and not yet exchanged
    cudaStream_t x_stream = nullptr;
    cudaEvent_t x_event = nullptr, x_done = nullptr;
    bool x_pending = false;
    int64_t x_launches = 0;
};

namespace cgs {

static constexpr int kWorkSlots = 1024;
static constexpr int kPartialBlocks = 1184;  // 148 SMs x 8

static int query_sms(int device, int *sms) {
    CGS_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, device));
    if (*sms <= 0) *sms = kNumSMsFallback;
    return 0;
}

// Orders cells longest-first on the host (D counts only) and uploads the schedule.
static int build_cell_order(cgs_corpus *c) {
    std::vector<int64_t> ptr((size_t)c->n_cells + 1);
    CGS_CUDA(cgs_memcpy_sync(c->stream, ptr.data(), c->d_cell_ptr, sizeof(int64_t) * ptr.size(), cudaMemcpyDeviceToHost));
    std::vector<int32_t> order((size_t)c->n_cells);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
        return (ptr[a + 1] - ptr[a]) > (ptr[b + 1] - ptr[b]);
    });
    CGS_TRY(dmalloc(&c->d_cell_order, order.size(), c->stream));
    CGS_CUDA(cgs_memcpy_sync(c->stream, c->d_cell_order, order.data(), sizeof(int32_t) * order.size(), cudaMemcpyHostToDevice));
    // models run on non-blocking streams: everything the corpus owns must have landed before they start
    CGS_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// Steps (4)-(5) of the construction: canonicalise every cell's segment and, if duplicates were
// dropped, compact.  d_tmp holds the unsorted region ids, d_raw_ptr the segment offsets.
static int canonicalise(cgs_corpus *c, int32_t *d_tmp, int64_t *d_raw_ptr, int64_t raw_total) {
    const int32_t D = c->n_cells;
    int32_t *d_sorted = nullptr;
    unsigned long long *d_uniq = nullptr;
    int *d_flag = nullptr;
    CGS_TRY(dmalloc(&d_sorted, (size_t)raw_total, c->stream));
    CGS_TRY(dmalloc(&d_uniq, (size_t)D, c->stream));
    CGS_TRY(dmalloc(&d_flag, 1, c->stream));
    CGS_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), c->stream));

    int max_smem = 0;
    CGS_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    int64_t words_needed = ((int64_t)c->n_regions + 31) / 32;
    int64_t words_cap = (max_smem - 1024) / 4;
    int32_t words = (int32_t)std::max<int64_t>(1, std::min(words_needed, words_cap));
    size_t smem = (size_t)words * 4;
    CGS_CUDA(cudaFuncSetAttribute(bitmap_sort_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = std::max(1, std::min(D, c->num_sms * (smem > 100 * 1024 ? 1 : 2)));
    StageTimer tm;
    bitmap_sort_cells_kernel<<<blocks, 512, smem, c->stream>>>(d_tmp, d_raw_ptr, D, c->n_regions, words, d_sorted, d_uniq, d_flag);
    CGS_CUDA(cudaGetLastError());
    tm.lap("  bitmap_sort_cells");
    int flag = 0;
    CGS_CUDA(cgs_memcpy_sync(c->stream, &flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) {
        dfree(d_sorted, c->stream); dfree(d_uniq, c->stream); dfree(d_flag, c->stream);
        CGS_FAIL("matrix index out of range while building the token list");
    }
    CGS_TRY(dmalloc(&c->d_cell_ptr, (size_t)D + 1, c->stream));
    exclusive_scan_i64_kernel<<<1, 1024, 0, c->stream>>>(d_uniq, c->d_cell_ptr, D);
    CGS_CUDA(cudaGetLastError());
    int64_t total = 0;
    CGS_CUDA(cgs_memcpy_sync(c->stream, &total, c->d_cell_ptr + D, sizeof(int64_t), cudaMemcpyDeviceToHost));
    c->n_tokens = total;
    if (total == raw_total) {
        c->d_tok_region = d_sorted;
    } else {
        CGS_TRY(dmalloc(&c->d_tok_region, (size_t)total, c->stream));
        compact_cells_kernel<<<std::max(1, std::min(D, c->num_sms * 8)), 256, 0, c->stream>>>(d_sorted, d_raw_ptr, c->d_cell_ptr, D,
                                                                               c->d_tok_region);
        CGS_CUDA(cudaGetLastError());
        CGS_CUDA(cudaStreamSynchronize(c->stream));
        dfree(d_sorted, c->stream);
    }
    dfree(d_uniq, c->stream);
    dfree(d_flag, c->stream);
    tm.lap("  scan + compact");
    int rc = build_cell_order(c);
    tm.lap("  build_cell_order");
    return rc;
}

static int corpus_common_init(cgs_corpus *c, int device) {
    c->device = device;
    CGS_TRY(query_sms(device, &c->num_sms));
    CGS_TRY(configure_pool(device));
    CGS_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    return 0;
}

}  // namespace cgs

// =============================================================================================
extern "C" {

const char *cgs_last_error(void) { return g_last_error.c_str(); }
int cgs_version(void) { return 100; }
int cgs_max_topics(void) { return kMaxTopics; }

int cgs_device_count(int *count) {
    if (!count) CGS_FAIL("count is NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        CGS_FAIL(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    return 0;
}

int cgs_corpus_from_regions_by_cells(cgs_corpus **out, int device, const int64_t *indptr, const int32_t *indices,
                                     int32_t n_regions, int32_t n_cells, int64_t nnz, int32_t cell_begin,
                                     int32_t cell_end) {
    if (!out || !indptr || (!indices && nnz > 0)) CGS_FAIL("NULL argument");
    if (n_regions <= 0 || n_cells <= 0) CGS_FAIL("empty matrix");
    if (cell_begin < 0 || cell_end > n_cells || cell_begin >= cell_end) CGS_FAIL("bad cell range");
    if (indptr[n_regions] != nnz) CGS_FAIL("indptr[n_regions] != nnz");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    cgs_corpus *c = new cgs_corpus();
    *out = nullptr;
    int rc = [&]() -> int {
        CGS_TRY(corpus_common_init(c, device));
        c->n_cells = cell_end - cell_begin;
        c->n_regions = n_regions;
        const int32_t D = c->n_cells;
        int64_t *d_row_ptr = nullptr;
        int32_t *d_idx = nullptr;
        unsigned long long *d_count = nullptr, *d_fill = nullptr;
        int64_t *d_raw_ptr = nullptr;
        CGS_TRY(dmalloc(&d_row_ptr, (size_t)n_regions + 1, c->stream));
        CGS_TRY(dmalloc(&d_idx, (size_t)nnz, c->stream));
        CGS_TRY(dmalloc(&d_count, (size_t)D, c->stream));
        CGS_TRY(dmalloc(&d_fill, (size_t)D, c->stream));
        CGS_TRY(dmalloc(&d_raw_ptr, (size_t)D + 1, c->stream));
        CGS_CUDA(cgs_memcpy_sync(c->stream, d_row_ptr, indptr, sizeof(int64_t) * ((size_t)n_regions + 1), cudaMemcpyHostToDevice));
        CGS_CUDA(cgs_memcpy_sync(c->stream, d_idx, indices, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice));
        CGS_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long) * (size_t)D, c->stream));
        CGS_CUDA(cudaMemsetAsync(d_fill, 0, sizeof(unsigned long long) * (size_t)D, c->stream));
        const int blocks = c->num_sms * 8;
        count_cells_kernel<<<blocks, 256, 0, c->stream>>>(d_idx, nnz, cell_begin, cell_end, d_count);
        CGS_CUDA(cudaGetLastError());
        exclusive_scan_i64_kernel<<<1, 1024, 0, c->stream>>>(d_count, d_raw_ptr, D);
        CGS_CUDA(cudaGetLastError());
        int64_t raw_total = 0;
        CGS_CUDA(cgs_memcpy_sync(c->stream, &raw_total, d_raw_ptr + D, sizeof(int64_t), cudaMemcpyDeviceToHost));
        int32_t *d_tmp = nullptr;
        CGS_TRY(dmalloc(&d_tmp, (size_t)raw_total, c->stream));
        scatter_rows_kernel<<<blocks, 256, 0, c->stream>>>(d_row_ptr, d_idx, n_regions, cell_begin, cell_end, d_raw_ptr, d_fill, d_tmp);
        CGS_CUDA(cudaGetLastError());
        CGS_CUDA(cudaStreamSynchronize(c->stream));
        dfree(d_row_ptr, c->stream); dfree(d_idx, c->stream); dfree(d_count, c->stream); dfree(d_fill, c->stream);
        int r = canonicalise(c, d_tmp, d_raw_ptr, raw_total);
        dfree(d_tmp, c->stream); dfree(d_raw_ptr, c->stream);
        return r;
    }();
    if (rc != 0) { cgs_corpus_destroy(c); return rc; }
    *out = c;
    return 0;
}

int cgs_corpus_from_cells_by_regions(cgs_corpus **out, int device, const int64_t *indptr, const int32_t *indices,
                                     int32_t n_cells, int32_t n_regions, int64_t nnz, int32_t cell_begin,
                                     int32_t cell_end) {
    if (!out || !indptr || (!indices && nnz > 0)) CGS_FAIL("NULL argument");
    if (n_regions <= 0 || n_cells <= 0) CGS_FAIL("empty matrix");
    if (cell_begin < 0 || cell_end > n_cells || cell_begin >= cell_end) CGS_FAIL("bad cell range");
    if (indptr[n_cells] != nnz) CGS_FAIL("indptr[n_cells] != nnz");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    cgs_corpus *c = new cgs_corpus();
    *out = nullptr;
    int rc = [&]() -> int {
        CGS_TRY(corpus_common_init(c, device));
        c->n_cells = cell_end - cell_begin;
        c->n_regions = n_regions;
        const int32_t D = c->n_cells;
        const int64_t t0 = indptr[cell_begin], t1 = indptr[cell_end];
        const int64_t raw_total = t1 - t0;
        std::vector<int64_t> local_ptr((size_t)D + 1);
        for (int32_t d = 0; d <= D; ++d) local_ptr[d] = indptr[cell_begin + d] - t0;
        int64_t *d_raw_ptr = nullptr;
        int32_t *d_tmp = nullptr;
        CGS_TRY(dmalloc(&d_raw_ptr, (size_t)D + 1, c->stream));
        CGS_TRY(dmalloc(&d_tmp, (size_t)raw_total, c->stream));
        StageTimer tm;
        CGS_CUDA(cgs_memcpy_sync(c->stream, d_raw_ptr, local_ptr.data(), sizeof(int64_t) * local_ptr.size(), cudaMemcpyHostToDevice));
        CGS_CUDA(cgs_memcpy_sync(c->stream, d_tmp, indices + t0, sizeof(int32_t) * (size_t)raw_total, cudaMemcpyHostToDevice));
        tm.lap("H2D indices");
        int r = canonicalise(c, d_tmp, d_raw_ptr, raw_total);
        tm.lap("canonicalise + cell order");
        dfree(d_tmp, c->stream); dfree(d_raw_ptr, c->stream);
        return r;
    }();
    if (rc != 0) { cgs_corpus_destroy(c); return rc; }
    *out = c;
    return 0;
}

int cgs_corpus_from_device_tokens(cgs_corpus **out, int device, const int64_t *d_cell_ptr, const int32_t *d_tok_region,
                                  int32_t n_cells, int32_t n_regions, int64_t n_tokens, int64_t global_token_offset) {
    if (!out || !d_cell_ptr || !d_tok_region) CGS_FAIL("NULL argument");
    if (n_regions <= 0 || n_cells <= 0) CGS_FAIL("empty matrix");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    cgs_corpus *c = new cgs_corpus();
    *out = nullptr;
    int rc = [&]() -> int {
        CGS_TRY(corpus_common_init(c, device));
        c->n_cells = n_cells;
        c->n_regions = n_regions;
        c->n_tokens = n_tokens;
        c->token_offset = global_token_offset;
        c->owns_tokens = false;
        c->d_cell_ptr = const_cast<int64_t *>(d_cell_ptr);
        c->d_tok_region = const_cast<int32_t *>(d_tok_region);
        return build_cell_order(c);
    }();
    if (rc != 0) { cgs_corpus_destroy(c); return rc; }
    *out = c;
    return 0;
}

int cgs_corpus_destroy(cgs_corpus *c) {
    if (!c) return 0;
    if (c->n_models > 0) CGS_FAIL("corpus still has live models");
    DeviceGuard g(c->device);
    if (c->owns_tokens) {
        if (c->d_cell_ptr) dfree(c->d_cell_ptr, c->stream);
        if (c->d_tok_region) dfree(c->d_tok_region, c->stream);
    }
    if (c->d_cell_order) dfree(c->d_cell_order, c->stream);
    if (c->d_cell_split) dfree(c->d_cell_split, c->stream);
    if (c->stream) cudaStreamDestroy(c->stream);  // pending frees still run
    delete c;
    return 0;
}

int cgs_corpus_dims(const cgs_corpus *c, int32_t *n_cells, int32_t *n_regions, int64_t *n_tokens) {
    if (!c) CGS_FAIL("NULL corpus");
    if (n_cells) *n_cells = c->n_cells;
    if (n_regions) *n_regions = c->n_regions;
    if (n_tokens) *n_tokens = c->n_tokens;
    return 0;
}

int cgs_corpus_set_token_offset(cgs_corpus *c, int64_t off) {
    if (!c) CGS_FAIL("NULL corpus");
    c->token_offset = off;
    return 0;
}

int cgs_corpus_set_region_blocks(cgs_corpus *c, int32_t n_blocks, const int32_t *bounds) {
    if (!c || !bounds) CGS_FAIL("NULL argument");
    if (n_blocks < 1 || n_blocks > 64) CGS_FAIL("n_blocks must be in [1, 64]");
    if (bounds[0] != 0 || bounds[n_blocks] != c->n_regions) CGS_FAIL("bounds must start at 0 and end at n_regions");
    for (int b = 0; b < n_blocks; ++b)
        if (bounds[b + 1] < bounds[b]) CGS_FAIL("bounds must not decrease");
    DeviceGuard g(c->device);
    if (c->d_cell_split) { dfree(c->d_cell_split, c->stream); c->d_cell_split = nullptr; }
    if (n_blocks > 1) {
        CGS_TRY(dmalloc(&c->d_cell_split, (size_t)(n_blocks - 1) * (size_t)c->n_cells, c->stream));
        for (int b = 1; b < n_blocks; ++b) {
            cell_split_kernel<<<(unsigned)ceil_div(c->n_cells, 256), 256, 0, c->stream>>>(
                c->d_cell_ptr, c->d_tok_region, c->n_cells, bounds[b], c->d_cell_split + (size_t)(b - 1) * c->n_cells);
            CGS_CUDA(cudaGetLastError());
        }
    }
    CGS_CUDA(cudaStreamSynchronize(c->stream));
    c->block_bounds.assign(bounds, bounds + n_blocks + 1);
    return 0;
}

int cgs_corpus_set_region_split(cgs_corpus *c, int32_t region_split) {
    if (!c) CGS_FAIL("NULL corpus");
    if (region_split < 0 || region_split > c->n_regions) CGS_FAIL("region_split outside [0, n_regions]");
    const int32_t bounds[3] = {0, region_split, c->n_regions};
    return cgs_corpus_set_region_blocks(c, 2, bounds);
}

int cgs_corpus_set_global_size(cgs_corpus *c, int64_t n_cells_global, int64_t n_tokens_global) {
    if (!c) CGS_FAIL("NULL corpus");
    if (n_cells_global < c->n_cells || n_tokens_global < c->n_tokens) CGS_FAIL("global size smaller than the shard");
    c->global_cells = n_cells_global;
    c->global_tokens = n_tokens_global;
    return 0;
}

int cgs_corpus_export_tokens(const cgs_corpus *c, int32_t *WS, int32_t *DS) {
    if (!c) CGS_FAIL("NULL corpus");
    DeviceGuard g(c->device);
    CGS_CUDA(cudaStreamSynchronize(c->stream));
    if (WS)
        CGS_CUDA(cgs_memcpy_sync(c->stream, WS, c->d_tok_region, sizeof(int32_t) * (size_t)c->n_tokens, cudaMemcpyDeviceToHost));
    if (DS) {
        int32_t *d_ds = nullptr;
        CGS_TRY(dmalloc(&d_ds, (size_t)c->n_tokens, c->stream));
        expand_cells_kernel<<<std::max(1, std::min(c->n_cells, c->num_sms * 8)), 256, 0, c->stream>>>(c->d_cell_ptr, c->n_cells, d_ds);
        CGS_CUDA(cudaGetLastError());
        CGS_CUDA(cgs_memcpy_sync(c->stream, DS, d_ds, sizeof(int32_t) * (size_t)c->n_tokens, cudaMemcpyDeviceToHost));
        dfree(d_ds, c->stream);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Shape of the sweep kernel for a model: lanes per token (tpt), 16-byte chunks per lane (ch),
// warps per CTA (= per cell) and the number of CTAs (= cells open at the same time).
// Throughput wants the fewest lanes per token and the GPU full of warps.  Closeness to the
// sequential reference wants little concurrency; measured on config C1 (8 seeds):
//     lag of the log-likelihood trace in the first sweeps
//         ~ 0.08 % per 1 % of all tokens in flight  +  0.28 % x (fraction of cells open at once),
// while tokens of ONE cell sampled together are harmless up to ~8 % of the cell at K = 10 but not at
// K = 40 on short cells (see cap_cell below).  So: wide CTAs, in-cell concurrency <=
// mean_cell_tokens / 16 (K <= 10) ... / 64 (K >= 40), open cells <= D / 8, tokens in flight <= N / 128.
// At benchmark scale (D >= 10^4) none of the caps binds before the GPU is full.
// Lanes per token the throughput alone asks for (a function of K only).
static int lanes_for_topics(int32_t K) {
    const int chunks = (K + 3) / 4;
    // fewest lanes per token with at most 4 chunks per lane, but two lanes as soon as there are two
    // chunks: one lane per token makes every row load and every count update an uncoalesced 32-address
    // access (measured on C2: K = 10 takes 0.93 ms per sweep with 2 lanes, 1.13 ms with 1)
    int t = 1;
    while (t * 4 < chunks) t <<= 1;
    if (chunks >= 2 && t < 2) t = 2;
    // ... and half the lanes with 5 chunks each when that covers the row (K = 33..40: 1.78-1.84 ms per sweep
    // on C2 with 2 lanes x 5 chunks, 2.04 ms with 4 lanes x 3; 6 and 7 chunks per lane do not pay)
    if (t >= 4 && chunks <= 5 * (t / 2)) t /= 2;
    return t;
}

// Row stride of n_wk / n_dk in counts: K rounded up so that the 16-byte-per-lane piece a token group loads
// (lanes x 16 B, at most 128 B) never straddles a 32-byte sector or a 128-byte line.  Depends on K only, so
// every rank of a sharded model uses the same layout.
static int32_t padded_topics(int32_t K) {
    const int piece = 4 * std::min(8, lanes_for_topics(K));  // counts per group load
    return (K + piece - 1) / piece * piece;
}

static void pick_sweep_shape(int32_t K, int32_t Kp, double n_tokens, double n_cells, int *tpt, int *ch, int *threads,
                             int *auto_blocks) {
    const int chunks = (K + 3) / 4;
    const double mean_cell = n_tokens / std::max(1.0, n_cells);
    int t = lanes_for_topics(K);
    // In-cell concurrency: at K = 10 up to 8 % of a cell sampled together is harmless (C1), at
    // K = 40 on 210-token cells 8 tokens per step (3.8 %) lag the trace by 0.4 % even with ONE cell open and
    // 1-2 tokens per step by 0.1 % (profiles/r1_trace_small40_concurrency.txt): 1/16 of a cell up to
    // K = 10, shrinking to 1/64 from K = 40.
    const double cell_div = std::min(64.0, std::max(16.0, 1.6 * (double)Kp));
    const int cap_cell = (int)std::max(1.0, mean_cell / cell_div);
    while (t < 32 && 32 / t > cap_cell) t <<= 1;
    // 4 warps per cell measure 1-4 % faster than 8 for every multi-lane shape (more, smaller CTAs per SM);
    // the one-lane shape (K <= 4) wants 8
    int warps = t == 1 ? 8 : 4;
    while (warps > 1 && warps * (32 / t) > cap_cell) warps >>= 1;
    const double in_flight_per_cta = warps * (32.0 / t);
    const double by_cells = std::max(1.0, n_cells / 8.0);
    const double by_tokens = std::max(1.0, n_tokens / (128.0 * in_flight_per_cta));
    *auto_blocks = (int)std::min(1.0e9, std::min(by_cells, by_tokens));
    *tpt = t;
    *ch = (chunks + t - 1) / t;
    *threads = warps * 32;
}

int cgs_model_create(cgs_model **out, cgs_corpus *corpus, int32_t n_topics, double alpha, double eta, uint64_t seed,
                     void *stream) {
    if (!out || !corpus) CGS_FAIL("NULL argument");
    if (n_topics < 1) CGS_FAIL("n_topics must be >= 1");
    if (n_topics > kMaxTopics) CGS_FAIL("n_topics exceeds cgs_max_topics()");
    if (!(alpha > 0) || !(eta > 0)) CGS_FAIL("alpha and eta must be greater than zero");
    DeviceGuard g(corpus->device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    cgs_model *m = new cgs_model();
    *out = nullptr;
    m->corpus = corpus;
    corpus->n_models++;
    int rc = [&]() -> int {
        m->K = n_topics;
        m->Kp = padded_topics(n_topics);
        m->alpha = alpha;
        m->eta = eta;
        m->seed = seed;
        // The concurrency caps compare what is in flight on THIS device with the statistics it is sampled against,
        // i.e. with the whole matrix: a shard of an AD-LDA run passes the global sizes.
        const double cap_tokens = (double)std::max<int64_t>(corpus->n_tokens, corpus->global_tokens);
        const double cap_cells = (double)std::max<int64_t>(corpus->n_cells, corpus->global_cells);
        pick_sweep_shape(m->K, m->Kp, cap_tokens, cap_cells, &m->tpt, &m->ch, &m->threads, &m->auto_blocks);
        if (stream) {
            m->stream = (cudaStream_t)stream;
        } else {
            CGS_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
            m->owns_stream = true;
        }
        const size_t W = (size_t)corpus->n_regions, D = (size_t)corpus->n_cells, Kp = (size_t)m->Kp;
        CGS_TRY(dmalloc(&m->d_z, (size_t)corpus->n_tokens, m->stream));
        CGS_TRY(dmalloc(&m->d_nwk, W * Kp, m->stream));
        CGS_TRY(dmalloc(&m->d_ndk, D * Kp, m->stream));
        CGS_TRY(dmalloc(&m->d_nk, Kp, m->stream));
        CGS_TRY(dmalloc(&m->d_work, (size_t)kWorkSlots, m->stream));
        CGS_TRY(dmalloc(&m->d_partials, (size_t)kPartialBlocks * 2 + 8, m->stream));
        CGS_TRY(dmalloc(&m->d_flag, 1, m->stream));
        CGS_CUDA(cudaMemsetAsync(m->d_nwk, 0, sizeof(int32_t) * W * Kp, m->stream));
        CGS_CUDA(cudaMemsetAsync(m->d_ndk, 0, sizeof(int32_t) * D * Kp, m->stream));
        CGS_CUDA(cudaMemsetAsync(m->d_nk, 0, sizeof(int32_t) * Kp, m->stream));
        CGS_CUDA(cudaMemsetAsync(m->d_z, 0, sizeof(uint16_t) * (size_t)corpus->n_tokens, m->stream));
        m->sweep_blocks = 0;
        return 0;
    }();
    if (rc != 0) { cgs_model_destroy(m); return rc; }
    *out = m;
    return 0;
}

int cgs_model_destroy(cgs_model *m) {
    if (!m) return 0;
    DeviceGuard g(m->corpus->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    dfree(m->d_z, m->stream); dfree(m->d_nwk, m->stream); dfree(m->d_ndk, m->stream); dfree(m->d_nk, m->stream);
    if (m->x_stream) { cudaStreamSynchronize(m->x_stream); cudaStreamDestroy(m->x_stream); }
    if (m->x_event) cudaEventDestroy(m->x_event);
    if (m->x_done) cudaEventDestroy(m->x_done);
    cudaFree(m->d_arena);
    dfree(m->d_nwk_ro, m->stream); dfree(m->d_nk_ro, m->stream);
    cudaFree(m->d_snap); cudaFree(m->d_delta); dfree(m->d_work, m->stream); dfree(m->d_partials, m->stream); dfree(m->d_flag, m->stream);
    if (m->owns_stream && m->stream) cudaStreamDestroy(m->stream);
    m->corpus->n_models--;
    delete m;
    return 0;
}

int cgs_model_set_stream(cgs_model *m, void *stream) {
    if (!m) CGS_FAIL("NULL model");
    DeviceGuard g(m->corpus->device);
    CGS_CUDA(cudaStreamSynchronize(m->stream));
    if (m->owns_stream) { cudaStreamDestroy(m->stream); m->owns_stream = false; }
    if (stream) {
        m->stream = (cudaStream_t)stream;
    } else {
        CGS_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
        m->owns_stream = true;
    }
    return 0;
}

int cgs_model_sync(cgs_model *m) {
    if (!m) CGS_FAIL("NULL model");
    DeviceGuard g(m->corpus->device);
    CGS_CUDA(cudaStreamSynchronize(m->stream));
    return 0;
}

static int rebuild_counts(cgs_model *m) {
    cgs_corpus *c = m->corpus;
    const size_t W = (size_t)c->n_regions, D = (size_t)c->n_cells, Kp = (size_t)m->Kp;
    CGS_CUDA(cudaMemsetAsync(m->d_nwk, 0, sizeof(int32_t) * W * Kp, m->stream));
    CGS_CUDA(cudaMemsetAsync(m->d_ndk, 0, sizeof(int32_t) * D * Kp, m->stream));
    CGS_CUDA(cudaMemsetAsync(m->d_nk, 0, sizeof(int32_t) * Kp, m->stream));
    if (m->d_arena) {
        // the tables are about to hold LOCAL counts again: the exchange state restarts from "nothing published"
        CGS_CUDA(cudaMemsetAsync(m->d_arena, 0, 2 * W * Kp * sizeof(int32_t), m->stream));
        if (m->d_snap) CGS_CUDA(cudaMemsetAsync(m->d_snap, 0, W * Kp * sizeof(int32_t), m->stream));
    }
    const int blocks = std::max(1, std::min(c->n_cells, c->num_sms * 8));
    count_kernel<<<blocks, 256, sizeof(int) * Kp, m->stream>>>(c->d_cell_ptr, c->d_tok_region, m->d_z, c->n_cells, m->K,
                                                              m->Kp, m->d_nwk, m->d_ndk, m->d_nk);
    CGS_CUDA(cudaGetLastError());
    m->launches += 1;
    return 0;
}

int cgs_model_init(cgs_model *m) {
    if (!m) CGS_FAIL("NULL model");
    cgs_corpus *c = m->corpus;
    DeviceGuard g(c->device);
    init_z_kernel<<<c->num_sms * 8, 256, 0, m->stream>>>(m->d_z, c->n_tokens, c->token_offset, m->K);
    CGS_CUDA(cudaGetLastError());
    m->launches += 1;
    m->sweeps = 0;
    m->x_pending_block = -1;
    return rebuild_counts(m);
}

int cgs_model_set_z(cgs_model *m, const int32_t *z) {
    if (!m || !z) CGS_FAIL("NULL argument");
    cgs_corpus *c = m->corpus;
    DeviceGuard g(c->device);
    int32_t *d_tmp = nullptr;
    CGS_TRY(dmalloc(&d_tmp, (size_t)c->n_tokens, m->stream));
    CGS_CUDA(cudaMemcpyAsync(d_tmp, z, sizeof(int32_t) * (size_t)c->n_tokens, cudaMemcpyHostToDevice, m->stream));
    CGS_CUDA(cudaMemsetAsync(m->d_flag, 0, sizeof(int), m->stream));
    z_from_i32_kernel<<<c->num_sms * 8, 256, 0, m->stream>>>(d_tmp, m->d_z, c->n_tokens, m->K, m->d_flag);
    CGS_CUDA(cudaGetLastError());
    m->launches += 1;
    int flag = 0;
    CGS_CUDA(cudaMemcpyAsync(&flag, m->d_flag, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    CGS_CUDA(cudaStreamSynchronize(m->stream));
    dfree(d_tmp, m->stream);
    if (flag) CGS_FAIL("z contains a topic outside [0, n_topics)");
    return rebuild_counts(m);
}

int cgs_model_get_z(cgs_model *m, int32_t *z) {
    if (!m || !z) CGS_FAIL("NULL argument");
    cgs_corpus *c = m->corpus;
    DeviceGuard g(c->device);
    int32_t *d_tmp = nullptr;
    CGS_TRY(dmalloc(&d_tmp, (size_t)c->n_tokens, m->stream));
    z_to_i32_kernel<<<c->num_sms * 8, 256, 0, m->stream>>>(m->d_z, d_tmp, c->n_tokens);
    CGS_CUDA(cudaGetLastError());
    m->launches += 1;
    CGS_CUDA(cudaMemcpyAsync(z, d_tmp, sizeof(int32_t) * (size_t)c->n_tokens, cudaMemcpyDeviceToHost, m->stream));
    CGS_CUDA(cudaStreamSynchronize(m->stream));
    dfree(d_tmp, m->stream);
    return 0;
}

}  // extern "C"

// ---- sweep dispatch --------------------------------------------------------------------------
static ExchangeParams g_no_exchange;  // zero-initialised: the argument of launches that do not exchange

template <int TPT, int CH, typename R>
static int launch_sweep(cgs_model *m, const SweepParams &p, int *blocks_cache) {
    const bool fused = p.x_ctas > 0;
    int *cache = fused ? &m->sweep_blocks_fused : blocks_cache;
    if (*cache == 0) {
        int per_sm = 0;
        if (fused) {
            if constexpr (std::is_same<R, float>::value)
                CGS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sweep_kernel<TPT, CH, float, true>, m->threads, 0));
        } else {
            CGS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sweep_kernel<TPT, CH, R, false>, m->threads, 0));
        }
        if (per_sm < 1) per_sm = 1;
        *cache = m->corpus->num_sms * per_sm;
    }
    int blocks = std::max(1, std::min(*cache, m->corpus->n_cells));
    blocks = std::min(blocks, m->max_blocks > 0 ? m->max_blocks : m->auto_blocks);
    if (fused) {
        if constexpr (std::is_same<R, float>::value) {
            // the exchange blocks come first in the grid and must all be resident: they wait for each other and
            // for the peers; the sampling blocks take whatever is left of the device and pull cells from the queue
            // (a request for more exchange blocks than fit beside one sampling block would leave some of them waiting
            // for an SM while the resident ones wait for them at the barrier: clamp it to what this kernel can hold)
            SweepParams q = p;
            q.x_ctas = std::min(p.x_ctas, *cache - 1);
            if (q.x_ctas < 1) CGS_FAIL("the device cannot hold an exchange block beside a sampling block");
            blocks = std::max(1, std::min(blocks, *cache - q.x_ctas));
            sweep_kernel<TPT, CH, float, true><<<blocks + q.x_ctas, m->threads, 0, m->stream>>>(q, m->x_fused);
        } else {
            CGS_FAIL("the fused sweep + exchange launch exists for the fp32 conditional only");
        }
    } else {
        sweep_kernel<TPT, CH, R, false><<<blocks, m->threads, 0, m->stream>>>(p, g_no_exchange);
    }
    CGS_CUDA(cudaGetLastError());
    return 0;
}

template <int TPT>
static int dispatch_ch(cgs_model *m, const SweepParams &p, int *bc) {
    if (m->fp64) {  // the A/B build exists for at most 4 chunks per lane (cgs_model_set_conditional_precision re-shapes)
        switch (m->ch) {
            case 1: return launch_sweep<TPT, 1, double>(m, p, bc);
            case 2: return launch_sweep<TPT, 2, double>(m, p, bc);
            case 3: return launch_sweep<TPT, 3, double>(m, p, bc);
            case 4: return launch_sweep<TPT, 4, double>(m, p, bc);
        }
        CGS_FAIL("unsupported chunk count for the fp64 conditional");
    }
    switch (m->ch) {
        case 1: return launch_sweep<TPT, 1, float>(m, p, bc);
        case 2: return launch_sweep<TPT, 2, float>(m, p, bc);
        case 3: return launch_sweep<TPT, 3, float>(m, p, bc);
        case 4: return launch_sweep<TPT, 4, float>(m, p, bc);
        case 5: if (TPT <= 16) return launch_sweep<(TPT <= 16 ? TPT : 16), 5, float>(m, p, bc); break;
        case 6: if (TPT <= 16) return launch_sweep<(TPT <= 16 ? TPT : 16), 6, float>(m, p, bc); break;
        case 7: if (TPT <= 16) return launch_sweep<(TPT <= 16 ? TPT : 16), 7, float>(m, p, bc); break;
        case 8: if (TPT <= 16) return launch_sweep<(TPT <= 16 ? TPT : 16), 8, float>(m, p, bc); break;
    }
    CGS_FAIL("unsupported chunk count");
}

template <int TPT, int CH, typename R>
static int launch_sweep_det(cgs_model *m, const SweepParams &p) {
    using L = CellSmem<TPT, CH, R>;
    const int nwarps = m->threads / 32;
    const size_t smem = (size_t)nwarps * L::WORDS * 4 + 2 * (size_t)L::KS * 4 + 16;
    if (m->sweep_blocks_det == 0) {
        CGS_CUDA(cudaFuncSetAttribute(sweep_det_kernel<TPT, CH, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        CGS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sweep_det_kernel<TPT, CH, R>, m->threads, smem));
        m->sweep_blocks_det = m->corpus->num_sms * std::max(1, per_sm);
    }
    const int blocks = std::max(1, std::min(m->sweep_blocks_det, m->corpus->n_cells));
    sweep_det_kernel<TPT, CH, R><<<blocks, m->threads, smem, m->stream>>>(p);
    CGS_CUDA(cudaGetLastError());
    return 0;
}

template <int TPT>
static int dispatch_det_ch(cgs_model *m, const SweepParams &p) {
    if (m->fp64) {
        switch (m->ch) {
            case 1: return launch_sweep_det<TPT, 1, double>(m, p);
            case 2: return launch_sweep_det<TPT, 2, double>(m, p);
            case 3: return launch_sweep_det<TPT, 3, double>(m, p);
            case 4: return launch_sweep_det<TPT, 4, double>(m, p);
        }
    } else {
        switch (m->ch) {
            case 1: return launch_sweep_det<TPT, 1, float>(m, p);
            case 2: return launch_sweep_det<TPT, 2, float>(m, p);
            case 3: return launch_sweep_det<TPT, 3, float>(m, p);
            case 4: return launch_sweep_det<TPT, 4, float>(m, p);
        }
    }
    CGS_FAIL("unsupported chunk count for the deterministic sweep");
}

static int dispatch_sweep_det(cgs_model *m, const SweepParams &p) {
    switch (m->tpt) {
        case 1: return dispatch_det_ch<1>(m, p);
        case 2: return dispatch_det_ch<2>(m, p);
        case 4: return dispatch_det_ch<4>(m, p);
        case 8: return dispatch_det_ch<8>(m, p);
        case 16: return dispatch_det_ch<16>(m, p);
        case 32: return dispatch_det_ch<32>(m, p);
    }
    CGS_FAIL("unsupported lanes-per-token");
}

static int dispatch_sweep(cgs_model *m, const SweepParams &p) {
    int *bc = &m->sweep_blocks;
    switch (m->tpt) {
        case 1: return dispatch_ch<1>(m, p, bc);
        case 2: return dispatch_ch<2>(m, p, bc);
        case 4: return dispatch_ch<4>(m, p, bc);
        case 8: return dispatch_ch<8>(m, p, bc);
        case 16: return dispatch_ch<16>(m, p, bc);
        case 32: return dispatch_ch<32>(m, p, bc);
    }
    CGS_FAIL("unsupported lanes-per-token");
}

extern "C" {

static int sweep_impl(cgs_model *m, int32_t n_sweeps, int32_t part, int32_t n_parts);

int cgs_model_sweep(cgs_model *m, int32_t n_sweeps) {
    if (!m) CGS_FAIL("NULL model");
    if (n_sweeps < 0) CGS_FAIL("n_sweeps < 0");
    return sweep_impl(m, n_sweeps, 0, 1);
}

int cgs_model_sweep_part(cgs_model *m, int32_t part, int32_t n_parts) {
    if (!m) CGS_FAIL("NULL model");
    if (n_parts < 1 || part < 0 || part >= n_parts) CGS_FAIL("bad part / n_parts");
    return sweep_impl(m, 1, part, n_parts);
}

static int sweep_impl(cgs_model *m, int32_t n_sweeps, int32_t part, int32_t n_parts) {
    cgs_corpus *c = m->corpus;
    DeviceGuard g(c->device);
    SweepParams p;
    p.cell_ptr = c->d_cell_ptr;
    p.tok_region = c->d_tok_region;
    p.z = m->d_z;
    p.n_wk = m->d_nwk;
    p.n_wk_read = m->d_nwk;
    p.nwk_write_delta = 0;
    p.n_dk = m->d_ndk;
    p.n_k = m->d_nk;
    p.n_k_read = m->d_nk;
    p.cell_order = c->d_cell_order;
    p.n_cells = c->n_cells;
    p.K = m->K;
    p.Kp = m->Kp;
    p.alpha = (float)m->alpha;
    p.eta = (float)m->eta;
    p.etaW = (float)(m->eta * (double)c->n_regions);
    p.alpha64 = m->alpha;
    p.eta64 = m->eta;
    p.etaW64 = m->eta * (double)c->n_regions;
    p.seed_lo = (uint32_t)m->seed;
    p.seed_hi = (uint32_t)(m->seed >> 32);
    p.rng_index_mask = m->rng_period == 0 ? ~0ull : (uint64_t)(m->rng_period / 4 - 1);
    p.token_offset = c->token_offset;
    {
        // n_k staleness budget: all warps together process at most 1/nk_divisor of a sweep
        // between two re-synchronisations of a warp's view of the topic totals
        const double warps_per_cell = m->threads / 32.0;
        const double warp_slots = (double)c->num_sms * 32.0;
        const double cta_cap = (double)(m->max_blocks > 0 ? m->max_blocks : m->auto_blocks);
        const double active_warps =
            std::max(1.0, std::min(warp_slots, std::min((double)c->n_cells, cta_cap) * warps_per_cell));
        const double tokens_per_batch = 32.0;  // a warp checks once per 32-token batch
        const double r = (double)c->n_tokens / (active_warps * tokens_per_batch * m->nk_divisor);
        p.refresh_every = (int32_t)std::max(1.0, std::min(1024.0, r));
    }
    p.part = part;
    p.n_parts = n_parts;
    p.visit_counter = m->d_visits;
    p.cell_lo = c->d_cell_ptr;
    p.cell_hi = c->d_cell_ptr + 1;
    if (m->sweep_block >= 0) {  // one region block of every cell
        const int nb = (int)c->block_bounds.size() - 1, b = m->sweep_block;
        if (b > 0) p.cell_lo = c->d_cell_split + (size_t)(b - 1) * c->n_cells;
        if (b < nb - 1) p.cell_hi = c->d_cell_split + (size_t)b * c->n_cells;
    }
    p.x_ctas = m->sweep_x_ctas;
    const bool det = m->det_parts > 0;
    const size_t WKp = (size_t)c->n_regions * (size_t)m->Kp;
    if (det) {
        if (n_parts != 1) CGS_FAIL("cgs_model_sweep_part is not available in deterministic mode");
        if (!m->d_nwk_ro) {
            CGS_TRY(dmalloc(&m->d_nwk_ro, WKp, m->stream));
            CGS_TRY(dmalloc(&m->d_nk_ro, (size_t)m->Kp, m->stream));
        }
        p.n_wk_read = m->d_nwk_ro;
        p.nwk_write_delta = (int64_t)((const char *)m->d_nwk - (const char *)m->d_nwk_ro);
        p.n_k_read = m->d_nk_ro;
        p.refresh_every = 0x7fffffff;  // the totals a cell works with are those of the copy
        p.n_parts = m->det_parts;
    }
    NvtxRange nvtx("cgs_model_sweep");
    for (int32_t s = 0; s < n_sweeps; ++s) {
        for (int32_t dp = 0; dp < (det ? m->det_parts : 1); ++dp) {
            if (det) {
                p.part = dp;
                CGS_CUDA(cudaMemcpyAsync(m->d_nwk_ro, m->d_nwk, WKp * 4, cudaMemcpyDeviceToDevice, m->stream));
                CGS_CUDA(cudaMemcpyAsync(m->d_nk_ro, m->d_nk, (size_t)m->Kp * 4, cudaMemcpyDeviceToDevice, m->stream));
            }
            const int slot = (int)(m->work_slot % kWorkSlots);
            if (slot == 0) {
                // a fresh bank of queue counters (stream-ordered after every launch that used the last one)
                CGS_CUDA(cudaMemsetAsync(m->d_work, 0, sizeof(int32_t) * kWorkSlots, m->stream));
            }
            m->work_slot += 1;
            p.work_counter = m->d_work + slot;
            p.sweep = (uint32_t)m->sweeps;  // all parts of one sweep share the Philox sweep index
            if (det) CGS_TRY(dispatch_sweep_det(m, p));
            else CGS_TRY(dispatch_sweep(m, p));
            m->launches += 1;
        }
        if (det || part == n_parts - 1) m->sweeps += 1;
    }
    return 0;
}

int cgs_model_debug_sweep_visits(cgs_model *m, int64_t *tokens_sampled) {
    if (!m || !tokens_sampled) CGS_FAIL("NULL argument");
    DeviceGuard g(m->corpus->device);
    unsigned long long *d = nullptr;
    CGS_TRY(dmalloc(&d, 1, m->stream));
    CGS_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), m->stream));
    m->d_visits = d;
    int rc = sweep_impl(m, 1, 0, 1);
    m->d_visits = nullptr;
    unsigned long long h = 0;
    if (rc == 0) {
        cudaError_t e = cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, m->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
        if (e != cudaSuccess) rc = fail(__FILE__, __LINE__, cudaGetErrorString(e));
    }
    dfree(d, m->stream);
    *tokens_sampled = (int64_t)h;
    return rc;
}

int cgs_model_set_sweep_shape(cgs_model *m, int32_t lanes_per_token, int32_t warps_per_cell) {
    if (!m) CGS_FAIL("NULL model");
    const int t = lanes_per_token;
    if (!(t == 1 || t == 2 || t == 4 || t == 8 || t == 16 || t == 32)) CGS_FAIL("lanes_per_token must be a power of two <= 32");
    if (!(warps_per_cell == 1 || warps_per_cell == 2 || warps_per_cell == 4 || warps_per_cell == 8 ||
          warps_per_cell == 12 || warps_per_cell == 16))
        CGS_FAIL("warps_per_cell must be 1, 2, 4, 8, 12 or 16");
    const int chunks = (m->K + 3) / 4;
    const int ch = (chunks + t - 1) / t;
    if (ch > kMaxChunksPerLane || t * ch * 4 > kMaxTopics) CGS_FAIL("lanes_per_token too small for this n_topics");
    if (warps_per_cell * 32 > (ch <= 1 ? 256 : ch <= 4 ? 512 : ch <= 6 ? 384 : 256))
        CGS_FAIL("too many warps per cell for this many chunks per lane");
    m->tpt = t;
    m->ch = ch;
    m->threads = warps_per_cell * 32;
    m->sweep_blocks = 0;
    m->auto_blocks = 1 << 30;  // an explicit shape lifts the automatic concurrency cap
    return 0;
}

int cgs_model_set_rng_period(cgs_model *m, int64_t period) {
    if (!m) CGS_FAIL("NULL model");
    if (period != 0 && (period < 128 || (period & (period - 1)) != 0)) CGS_FAIL("period must be 0 or a power of two >= 128");
    m->rng_period = period;
    return 0;
}

int cgs_model_set_max_blocks(cgs_model *m, int32_t max_blocks) {
    if (!m) CGS_FAIL("NULL model");
    if (max_blocks < 0) CGS_FAIL("max_blocks < 0");
    m->max_blocks = max_blocks;
    return 0;
}

int cgs_model_set_nk_staleness(cgs_model *m, double divisor) {
    if (!m) CGS_FAIL("NULL model");
    if (!(divisor >= 1.0)) CGS_FAIL("divisor must be >= 1");
    m->nk_divisor = divisor;
    return 0;
}

int cgs_model_get_sweep_shape(const cgs_model *m, int32_t *lanes_per_token, int32_t *chunks_per_lane,
                              int32_t *warps_per_cell) {
    if (!m) CGS_FAIL("NULL model");
    if (lanes_per_token) *lanes_per_token = m->tpt;
    if (chunks_per_lane) *chunks_per_lane = m->ch;
    if (warps_per_cell) *warps_per_cell = m->threads / 32;
    return 0;
}

int cgs_model_launch_count(const cgs_model *m, int64_t *count) {
    if (!m || !count) CGS_FAIL("NULL argument");
    *count = m->launches;
    return 0;
}

int cgs_model_sweeps_done(const cgs_model *m, int64_t *count) {
    if (!m || !count) CGS_FAIL("NULL argument");
    *count = m->sweeps;
    return 0;
}

// ---- log-likelihood ---------------------------------------------------------------------------
int cgs_model_loglik(cgs_model *m, int partial, double *ll) {
    if (!m || !ll) CGS_FAIL("NULL argument");
    if (partial < 0 || partial > 2) CGS_FAIL("partial must be 0, 1 or 2");
    NvtxRange nvtx("cgs_model_loglik");
    cgs_corpus *c = m->corpus;
    DeviceGuard g(c->device);
    const int64_t W = c->n_regions, D = c->n_cells;
    const int32_t Kp = m->Kp;
    double *part = m->d_partials;                    // [0, B): region side, [B, 2B): cell side
    double *res = m->d_partials + 2 * kPartialBlocks;  // 4 scalars
    CGS_CUDA(cudaMemsetAsync(res, 0, sizeof(double) * 4, m->stream));
    const int B = kPartialBlocks;
    if (partial == 0 || partial == 2) {
        lgamma_sum_kernel<<<B, kReduceThreads, 0, m->stream>>>(m->d_nwk, W * Kp, m->eta, part);
        final_sum_kernel<<<1, kReduceThreads, 0, m->stream>>>(part, B, res + 0);
        topic_norm_sum_kernel<<<1, kReduceThreads, 0, m->stream>>>(m->d_nk, m->K, m->eta * (double)W, res + 1);
        m->launches += 3;
    }
    if (partial == 0 || partial == 1) {
        lgamma_sum_kernel<<<B, kReduceThreads, 0, m->stream>>>(m->d_ndk, D * Kp, m->alpha, part + B);
        final_sum_kernel<<<1, kReduceThreads, 0, m->stream>>>(part + B, B, res + 2);
        const int cb = (int)std::max<int64_t>(1, std::min<int64_t>(B, ceil_div(D, kReduceThreads)));
        cell_norm_sum_kernel<<<cb, kReduceThreads, 0, m->stream>>>(c->d_cell_ptr, c->n_cells, m->alpha * m->K, part);
        final_sum_kernel<<<1, kReduceThreads, 0, m->stream>>>(part, cb, res + 3);
        m->launches += 4;
    }
    CGS_CUDA(cudaGetLastError());
    double h[4];
    CGS_CUDA(cudaMemcpyAsync(h, res, sizeof(double) * 4, cudaMemcpyDeviceToHost, m->stream));
    CGS_CUDA(cudaStreamSynchronize(m->stream));
    *ll = h[0] + h[1] + h[2] + h[3];
    return 0;
}

// ---- exports ----------------------------------------------------------------------------------
int cgs_model_export_counts(cgs_model *m, int32_t *nzw, int32_t *ndz, int32_t *nz) {
    if (!m) CGS_FAIL("NULL model");
    NvtxRange nvtx("cgs_model_export_counts");
    cgs_corpus *c = m->corpus;
    DeviceGuard g(c->device);
    const int64_t W = c->n_regions, D = c->n_cells;
    const int32_t K = m->K, Kp = m->Kp;
    if (nzw) {
        int32_t *d_t = nullptr;
        CGS_TRY(dmalloc(&d_t, (size_t)W * K, m->stream));
        dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(K, 32)), block(32, 8);
        transpose_counts_kernel<<<grid, block, 0, m->stream>>>(m->d_nwk, W, K, Kp, d_t);
        CGS_CUDA(cudaGetLastError());
        m->launches += 1;
        CGS_CUDA(cudaMemcpyAsync(nzw, d_t, sizeof(int32_t) * (size_t)W * K, cudaMemcpyDeviceToHost, m->stream));
        CGS_CUDA(cudaStreamSynchronize(m->stream));
        dfree(d_t, m->stream);
    }
    if (ndz) {
        int32_t *d_t = nullptr;
        CGS_TRY(dmalloc(&d_t, (size_t)D * K, m->stream));
        unpad_counts_kernel<<<c->num_sms * 8, 256, 0, m->stream>>>(m->d_ndk, D, K, Kp, d_t);
        CGS_CUDA(cudaGetLastError());
        m->launches += 1;
        CGS_CUDA(cudaMemcpyAsync(ndz, d_t, sizeof(int32_t) * (size_t)D * K, cudaMemcpyDeviceToHost, m->stream));
        CGS_CUDA(cudaStreamSynchronize(m->stream));
        dfree(d_t, m->stream);
    }
    if (nz) {
        CGS_CUDA(cudaMemcpyAsync(nz, m->d_nk, sizeof(int32_t) * (size_t)K, cudaMemcpyDeviceToHost, m->stream));
        CGS_CUDA(cudaStreamSynchronize(m->stream));
    }
    return 0;
}

static int export_dist(cgs_model *m, double *topic_side, int topic_transposed, double *cell_side, int cell_transposed) {
    NvtxRange nvtx("cgs_model_export_distributions");
    cgs_corpus *c = m->corpus;
    const int64_t W = c->n_regions, D = c->n_cells;
    const int32_t K = m->K, Kp = m->Kp;
    dim3 block(32, 8);
    if (topic_side) {
        double *d_t = nullptr;
        CGS_TRY(dmalloc(&d_t, (size_t)W * K, m->stream));
        dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(K, 32));
        topic_word_kernel<<<grid, block, 0, m->stream>>>(m->d_nwk, m->d_nk, W, K, Kp, m->eta, topic_transposed, d_t);
        CGS_CUDA(cudaGetLastError());
        m->launches += 1;
        CGS_CUDA(cudaMemcpyAsync(topic_side, d_t, sizeof(double) * (size_t)W * K, cudaMemcpyDeviceToHost, m->stream));
        CGS_CUDA(cudaStreamSynchronize(m->stream));
        dfree(d_t, m->stream);
    }
    if (cell_side) {
        double *d_t = nullptr;
        CGS_TRY(dmalloc(&d_t, (size_t)D * K, m->stream));
        dim3 grid((unsigned)ceil_div(D, 32), (unsigned)ceil_div(K, 32));
        doc_topic_kernel<<<grid, block, 0, m->stream>>>(m->d_ndk, c->d_cell_ptr, D, K, Kp, m->alpha, cell_transposed, d_t);
        CGS_CUDA(cudaGetLastError());
        m->launches += 1;
        CGS_CUDA(cudaMemcpyAsync(cell_side, d_t, sizeof(double) * (size_t)D * K, cudaMemcpyDeviceToHost, m->stream));
        CGS_CUDA(cudaStreamSynchronize(m->stream));
        dfree(d_t, m->stream);
    }
    return 0;
}

int cgs_model_export_distributions(cgs_model *m, double *topic_word, double *doc_topic) {
    if (!m) CGS_FAIL("NULL model");
    DeviceGuard g(m->corpus->device);
    return export_dist(m, topic_word, 1, doc_topic, 0);
}

int cgs_model_export_frames(cgs_model *m, double *topic_region, double *cell_topic) {
    if (!m) CGS_FAIL("NULL model");
    DeviceGuard g(m->corpus->device);
    return export_dist(m, topic_region, 0, cell_topic, 1);
}

// ---- AD-LDA exchange ---------------------------------------------------------------------------
static int ensure_exchange_buffers(cgs_model *m) {
    const size_t n = (size_t)m->corpus->n_regions * (size_t)m->Kp;
    if (!m->d_snap) {
        CGS_TRY(dmalloc_ipc(&m->d_snap, n));
        CGS_CUDA(cudaMemsetAsync(m->d_snap, 0, n * sizeof(int32_t), m->stream));
    }
    if (!m->d_delta) {
        CGS_TRY(dmalloc_ipc(&m->d_delta, n));
        CGS_CUDA(cudaMemsetAsync(m->d_delta, 0, n * sizeof(int32_t), m->stream));
    }
    return 0;
}

int cgs_model_device_buffer(cgs_model *m, int which, void **ptr, size_t *bytes) {
    if (!m || !ptr || !bytes) CGS_FAIL("NULL argument");
    cgs_corpus *c = m->corpus;
    DeviceGuard g(c->device);
    const size_t WK = (size_t)c->n_regions * (size_t)m->Kp;
    switch (which) {
        case 0: *ptr = m->d_nwk; *bytes = WK * 4; return 0;
        case 1: CGS_TRY(ensure_exchange_buffers(m)); *ptr = m->d_snap; *bytes = WK * 4; return 0;
        case 2: CGS_TRY(ensure_exchange_buffers(m)); *ptr = m->d_delta; *bytes = WK * 4; return 0;
        case 3: *ptr = m->d_nk; *bytes = (size_t)m->Kp * 4; return 0;
        case 4: *ptr = m->d_ndk; *bytes = (size_t)c->n_cells * (size_t)m->Kp * 4; return 0;
        case 5: *ptr = m->d_z; *bytes = (size_t)c->n_tokens * 2; return 0;
        case 6: *ptr = c->d_tok_region; *bytes = (size_t)c->n_tokens * 4; return 0;
        case 7: *ptr = c->d_cell_ptr; *bytes = ((size_t)c->n_cells + 1) * 8; return 0;
    }
    CGS_FAIL("unknown buffer id");
}

int cgs_model_padded_topics(const cgs_model *m, int32_t *kp) {
    if (!m || !kp) CGS_FAIL("NULL argument");
    *kp = m->Kp;
    return 0;
}

int cgs_model_snapshot(cgs_model *m) {
    if (!m) CGS_FAIL("NULL model");
    DeviceGuard g(m->corpus->device);
    CGS_TRY(ensure_exchange_buffers(m));
    const size_t n = (size_t)m->corpus->n_regions * (size_t)m->Kp;
    CGS_CUDA(cudaMemcpyAsync(m->d_snap, m->d_nwk, n * 4, cudaMemcpyDeviceToDevice, m->stream));
    return 0;
}

int cgs_model_delta_compute(cgs_model *m) {
    if (!m) CGS_FAIL("NULL model");
    DeviceGuard g(m->corpus->device);
    CGS_TRY(ensure_exchange_buffers(m));
    const int64_t n4 = (int64_t)m->corpus->n_regions * m->Kp / 4;
    delta_compute_kernel<<<m->corpus->num_sms * 8, 256, 0, m->stream>>>((const int4 *)m->d_nwk, (const int4 *)m->d_snap,
                                                                       (int4 *)m->d_delta, n4);
    CGS_CUDA(cudaGetLastError());
    m->launches += 1;
    return 0;
}

static int delta_apply_common(cgs_model *m, const PeerPtrs &peers, int n_ranks) {
    const int64_t n4 = (int64_t)m->corpus->n_regions * m->Kp / 4;
    CGS_CUDA(cudaMemsetAsync(m->d_nk, 0, sizeof(int32_t) * (size_t)m->Kp, m->stream));
    delta_apply_kernel<<<m->corpus->num_sms * 4, 256, sizeof(int) * (size_t)m->Kp, m->stream>>>(
        (int4 *)m->d_nwk, (int4 *)m->d_snap, peers, n_ranks, n4, m->Kp / 4, m->d_nk);
    CGS_CUDA(cudaGetLastError());
    m->launches += 1;
    return 0;
}

int cgs_model_delta_apply(cgs_model *m) {
    if (!m) CGS_FAIL("NULL model");
    DeviceGuard g(m->corpus->device);
    CGS_TRY(ensure_exchange_buffers(m));
    PeerPtrs peers;
    memset(&peers, 0, sizeof peers);
    peers.p[0] = (const int4 *)m->d_delta;
    return delta_apply_common(m, peers, 1);
}

int cgs_model_delta_apply_peers(cgs_model *m, void *const *peer_delta, int32_t n_ranks) {
    if (!m || !peer_delta) CGS_FAIL("NULL argument");
    if (n_ranks < 1 || n_ranks > kMaxPeers) CGS_FAIL("n_ranks out of range");
    DeviceGuard g(m->corpus->device);
    CGS_TRY(ensure_exchange_buffers(m));
    PeerPtrs peers;
    memset(&peers, 0, sizeof peers);
    for (int r = 0; r < n_ranks; ++r) {
        if (!peer_delta[r]) CGS_FAIL("NULL peer pointer");
        peers.p[r] = (const int4 *)peer_delta[r];
    }
    return delta_apply_common(m, peers, n_ranks);
}

// ---- CUDA IPC helpers for the peer-memory exchange (one process per GPU) -------------------------
int cgs_ipc_export(void *device_ptr, void *handle_out_64_bytes) {
    if (!device_ptr || !handle_out_64_bytes) CGS_FAIL("NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CGS_CUDA(cudaIpcGetMemHandle(&h, device_ptr));
    memcpy(handle_out_64_bytes, &h, sizeof h);
    return 0;
}

int cgs_ipc_open(int device, const void *handle_64_bytes, void **device_ptr) {
    if (!handle_64_bytes || !device_ptr) CGS_FAIL("NULL argument");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_64_bytes, sizeof h);
    CGS_CUDA(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int cgs_device_malloc(int device, size_t bytes, void **device_ptr) {
    if (!device_ptr || bytes == 0) CGS_FAIL("bad argument");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    CGS_CUDA(cudaMalloc(device_ptr, bytes));
    return 0;
}

int cgs_device_free(int device, void *device_ptr) {
    if (!device_ptr) return 0;
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    CGS_CUDA(cudaFree(device_ptr));
    return 0;
}

int cgs_ipc_close(int device, void *device_ptr) {
    if (!device_ptr) return 0;
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    CGS_CUDA(cudaIpcCloseMemHandle(device_ptr));
    return 0;
}


// ---- precision of the conditional (A/B) ------------------------------------------------------------
int cgs_model_set_conditional_precision(cgs_model *m, int32_t bits) {
    if (!m) CGS_FAIL("NULL model");
    if (bits != 32 && bits != 64) CGS_FAIL("bits must be 32 or 64");
    m->fp64 = bits == 64;
    m->sweep_blocks = 0;
    m->sweep_blocks_det = 0;
    if (m->fp64 && m->ch > 4) {  // the fp64 kernels are built for <= 4 chunks per lane: widen the token group
        const int chunks = (m->K + 3) / 4;
        while (m->tpt < 32 && (chunks + m->tpt - 1) / m->tpt > 4) m->tpt <<= 1;
        m->ch = (chunks + m->tpt - 1) / m->tpt;
        if (m->ch > 4) CGS_FAIL("n_topics too large for the fp64 conditional");
    }
    return 0;
}

// ---- deterministic mode --------------------------------------------------------------------------
int cgs_model_set_deterministic(cgs_model *m, int32_t parts) {
    if (!m) CGS_FAIL("NULL model");
    if (parts < 0 || parts > 1024) CGS_FAIL("parts must be in [0, 1024]");
    if (parts > 0 && m->ch > 4) {  // the deterministic kernels are built for <= 4 chunks per lane: widen the token group
        const int chunks = (m->K + 3) / 4;
        while (m->tpt < 32 && (chunks + m->tpt - 1) / m->tpt > 4) m->tpt <<= 1;
        m->ch = (chunks + m->tpt - 1) / m->tpt;
        if (m->ch > 4) CGS_FAIL("n_topics too large for the deterministic sweep");
        m->sweep_blocks = 0;
    }
    m->sweep_blocks_det = 0;
    m->det_parts = parts;
    return 0;
}

// ---- consistency checks (bench.py runs them after its timed regions) -------------------------------
int cgs_model_region_histogram(cgs_model *m, int32_t *d_hist, int64_t *ndk_mismatches) {
    if (!m || !d_hist) CGS_FAIL("NULL argument");
    cgs_corpus *c = m->corpus;
    DeviceGuard g(c->device);
    const size_t WKp = (size_t)c->n_regions * (size_t)m->Kp;
    unsigned long long *d_bad = nullptr;
    unsigned long long bad[2] = {0, 0};
    CGS_TRY(dmalloc(&d_bad, 2, m->stream));
    const int rc = [&]() -> int {
        CGS_CUDA(cudaMemsetAsync(d_bad, 0, 2 * sizeof(unsigned long long), m->stream));
        CGS_CUDA(cudaMemsetAsync(d_hist, 0, WKp * 4, m->stream));
        const int blocks = std::max(1, std::min(c->n_cells, c->num_sms * 8));
        recount_kernel<<<blocks, 256, sizeof(int) * (size_t)m->Kp, m->stream>>>(c->d_cell_ptr, c->d_tok_region, m->d_z, c->n_cells,
                                                                               m->K, m->Kp, m->d_ndk, d_hist, d_bad);
        CGS_CUDA(cudaGetLastError());
        m->launches += 1;
        CGS_CUDA(cgs_memcpy_sync(m->stream, bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost));
        return 0;
    }();
    dfree(d_bad, m->stream);
    if (rc != 0) return rc;
    if (ndk_mismatches) *ndk_mismatches = (int64_t)bad[0];
    return 0;
}

int cgs_model_verify_counts(cgs_model *m, const int32_t *d_hist, int64_t *out3) {
    if (!m || !out3) CGS_FAIL("NULL argument");
    cgs_corpus *c = m->corpus;
    DeviceGuard g(c->device);
    const size_t WKp = (size_t)c->n_regions * (size_t)m->Kp;
    int32_t *own_hist = nullptr;
    out3[0] = 0;
    if (!d_hist) {
        CGS_TRY(dmalloc(&own_hist, WKp, m->stream));
        int rc = cgs_model_region_histogram(m, own_hist, &out3[0]);
        if (rc != 0) { dfree(own_hist, m->stream); return rc; }
        d_hist = own_hist;
    }
    unsigned long long *d_bad = nullptr;
    long long *d_col = nullptr;
    unsigned long long bad[2] = {0, 0};
    std::vector<long long> col((size_t)m->Kp);
    std::vector<int32_t> nk((size_t)m->Kp);
    const int rc = [&]() -> int {
        CGS_TRY(dmalloc(&d_bad, 2, m->stream));
        CGS_TRY(dmalloc(&d_col, (size_t)m->Kp, m->stream));
        CGS_CUDA(cudaMemsetAsync(d_bad, 0, 2 * sizeof(unsigned long long), m->stream));
        CGS_CUDA(cudaMemsetAsync(d_col, 0, sizeof(long long) * (size_t)m->Kp, m->stream));
        compare_counts_kernel<<<c->num_sms * 4, 256, sizeof(long long) * (size_t)m->Kp, m->stream>>>(
            m->d_nwk, d_hist, (int64_t)WKp, m->Kp, d_bad, d_col);
        CGS_CUDA(cudaGetLastError());
        m->launches += 1;
        CGS_CUDA(cudaMemcpyAsync(bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, m->stream));
        CGS_CUDA(cudaMemcpyAsync(col.data(), d_col, sizeof(long long) * col.size(), cudaMemcpyDeviceToHost, m->stream));
        CGS_CUDA(cudaMemcpyAsync(nk.data(), m->d_nk, sizeof(int32_t) * nk.size(), cudaMemcpyDeviceToHost, m->stream));
        CGS_CUDA(cudaStreamSynchronize(m->stream));
        return 0;
    }();
    dfree(d_bad, m->stream); dfree(d_col, m->stream); dfree(own_hist, m->stream);
    if (rc != 0) return rc;
    out3[1] = (int64_t)bad[1];
    out3[2] = 0;
    for (int k = 0; k < m->Kp; ++k)
        if (col[(size_t)k] != (long long)nk[(size_t)k]) out3[2] += 1;
    return 0;
}

int cgs_model_state_hash(cgs_model *m, uint64_t *out4) {
    if (!m || !out4) CGS_FAIL("NULL argument");
    cgs_corpus *c = m->corpus;
    DeviceGuard g(c->device);
    unsigned long long *d_out = nullptr;
    CGS_TRY(dmalloc(&d_out, 4, m->stream));
    unsigned long long h[4] = {0, 0, 0, 0};
    const int rc = [&]() -> int {
        CGS_CUDA(cudaMemsetAsync(d_out, 0, 4 * sizeof(unsigned long long), m->stream));
        state_hash_kernel<<<c->num_sms * 4, 256, 0, m->stream>>>(m->d_nwk, (int64_t)c->n_regions * m->Kp, d_out);
        state_hash_kernel<<<1, 256, 0, m->stream>>>(m->d_nk, (int64_t)m->Kp, d_out + 2);
        CGS_CUDA(cudaGetLastError());
        m->launches += 2;
        CGS_CUDA(cgs_memcpy_sync(m->stream, h, d_out, sizeof h, cudaMemcpyDeviceToHost));
        return 0;
    }();
    dfree(d_out, m->stream);
    if (rc != 0) return rc;
    out4[0] = h[0]; out4[1] = h[1]; out4[2] = h[3]; out4[3] = h[2];  // hash(n_wk), sum(n_wk), sum(n_k), hash(n_k)
    return 0;
}

// ---- asynchronous peer-memory exchange (exchange_kernels.cuh) -----------------------------------------
static size_t arena_delta_bytes(const cgs_model *m) {
    return (size_t)m->corpus->n_regions * (size_t)m->Kp * 4;
}

int cgs_model_exchange_arena(cgs_model *m, void **ptr, size_t *bytes) {
    if (!m || !ptr || !bytes) CGS_FAIL("NULL argument");
    DeviceGuard g(m->corpus->device);
    if (!m->d_arena) {
        const size_t wk = arena_delta_bytes(m);
        m->arena_bytes = 2 * wk + kExchangeFlagWords * 4;
        CGS_CUDA(cudaMalloc((void **)&m->d_arena, m->arena_bytes));
        CGS_CUDA(cudaMemsetAsync(m->d_arena, 0, m->arena_bytes, m->stream));
        CGS_TRY(ensure_exchange_buffers(m));  // snap (zero: the first round then publishes the local counts)
        CGS_CUDA(cudaMemsetAsync(m->d_snap, 0, wk, m->stream));
        CGS_CUDA(cudaStreamSynchronize(m->stream));
    }
    *ptr = m->d_arena;
    *bytes = m->arena_bytes;
    return 0;
}

int cgs_model_exchange_open(cgs_model *m, void *const *peer_arena, int32_t n_ranks, int32_t rank) {
    if (!m || !peer_arena) CGS_FAIL("NULL argument");
    if (n_ranks < 1 || n_ranks > kMaxExchangeRanks || rank < 0 || rank >= n_ranks) CGS_FAIL("bad n_ranks / rank");
    DeviceGuard g(m->corpus->device);
    void *own = nullptr;
    size_t bytes = 0;
    CGS_TRY(cgs_model_exchange_arena(m, &own, &bytes));
    const size_t wk = arena_delta_bytes(m);
    memset(&m->xpeers, 0, sizeof m->xpeers);
    for (int q = 0; q < n_ranks; ++q) {
        char *base = (char *)(q == rank ? own : peer_arena[q]);
        if (!base) CGS_FAIL("NULL peer arena");
        m->xpeers.delta[q] = (int4 *)base;
        m->xpeers.total[q] = (int4 *)(base + wk);
        m->xpeers.flags[q] = (uint32_t *)(base + 2 * wk);
    }
    m->x_ranks = n_ranks;
    m->x_rank = rank;
    m->x_round = 0;
    if (!m->x_stream) {
        int lo = 0, hi = 0;
        CGS_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CGS_CUDA(cudaStreamCreateWithPriority(&m->x_stream, cudaStreamNonBlocking, hi));
        CGS_CUDA(cudaEventCreateWithFlags(&m->x_event, cudaEventDisableTiming));
        CGS_CUDA(cudaEventCreateWithFlags(&m->x_done, cudaEventDisableTiming));
    }
    return 0;
}

// rows [row0, row1) of the table (row1 < 0: all rows)
static void fill_exchange_params(cgs_model *m, ExchangeParams &p, int64_t row0 = 0, int64_t row1 = -1) {
    cgs_corpus *c = m->corpus;
    if (row1 < 0) row1 = c->n_regions;
    const int64_t off = row0 * m->Kp;  // in counts; a multiple of 4
    p.n_wk = m->d_nwk + off;
    p.snap = m->d_snap ? m->d_snap + off : nullptr;
    p.n_k = m->d_nk;
    p.peers = m->xpeers;
    for (int q = 0; q < m->x_ranks; ++q) {
        p.peers.delta[q] += off / 4;
        p.peers.total[q] += off / 4;
    }
    p.n_ranks = m->x_ranks;
    p.rank = m->x_rank;
    p.n4 = (row1 - row0) * m->Kp / 4;
    p.Kp4 = m->Kp / 4;
    p.round = ++m->x_round;
}

int cgs_model_exchange_round(cgs_model *m, int32_t n_ctas) {
    if (!m) CGS_FAIL("NULL model");
    if (m->x_ranks < 1) CGS_FAIL("cgs_model_exchange_open has not been called");
    cgs_corpus *c = m->corpus;
    DeviceGuard g(c->device);
    if (n_ctas <= 0) n_ctas = 64;
    n_ctas = std::min(n_ctas, c->num_sms);
    ExchangeParams p;
    fill_exchange_params(m, p);
    NvtxRange nvtx("cgs_model_exchange_round");
    // the round starts once everything enqueued on the model's stream so far has finished, and then runs BESIDE
    // whatever the caller enqueues there next (the next sweep)
    CGS_CUDA(cudaEventRecord(m->x_event, m->stream));
    CGS_CUDA(cudaStreamWaitEvent(m->x_stream, m->x_event, 0));
    exchange_round_kernel<<<n_ctas, kExchangeThreads, sizeof(int) * (size_t)m->Kp, m->x_stream>>>(p);
    CGS_CUDA(cudaGetLastError());
    CGS_CUDA(cudaEventRecord(m->x_done, m->x_stream));
    m->x_pending = true;
    m->launches += 1;
    m->x_launches += 1;
    return 0;
}

static int exchange_sync_rows(cgs_model *m, int32_t n_ctas, int64_t row0, int64_t row1);

int cgs_model_exchange_sync(cgs_model *m, int32_t n_ctas) {
    if (!m) CGS_FAIL("NULL model");
    if (m->x_pending_block >= 0) CGS_FAIL("overlapped sweeps are pending: call cgs_model_exchange_flush");
    return exchange_sync_rows(m, n_ctas, 0, -1);
}

int cgs_model_exchange_flush(cgs_model *m) {
    if (!m) CGS_FAIL("NULL model");
    if (m->x_pending_block < 0) return 0;
    const std::vector<int32_t> &bb = m->corpus->block_bounds;
    const int rc = exchange_sync_rows(m, 0, bb[(size_t)m->x_pending_block], bb[(size_t)m->x_pending_block + 1]);
    if (rc == 0) m->x_pending_block = -1;
    return rc;
}

// One sweep = one launch per region block.  exchange: while block b is sampled, the rows of the block sampled by the
// previous launch are exchanged inside the same launch.
static int sweep_region_blocks(cgs_model *m, int32_t n_sweeps, bool exchange, int32_t x_ctas) {
    cgs_corpus *c = m->corpus;
    const int nb = (int)c->block_bounds.size() - 1;
    if (nb < 1) CGS_FAIL("cgs_corpus_set_region_blocks / _split has not been called");
    if (m->det_parts > 0) CGS_FAIL("region-blocked sweeps are not available in deterministic mode");
    if (exchange && m->fp64) CGS_FAIL("overlapped sweeps need the fp32 conditional");
    DeviceGuard g(c->device);
    int rc = 0;
    for (int32_t s = 0; s < n_sweeps && rc == 0; ++s) {
        for (int b = 0; b < nb && rc == 0; ++b) {
            const bool fused = exchange && m->x_pending_block >= 0;
            if (fused) {
                const int pb = m->x_pending_block;
                if (pb == b && nb > 1) { rc = fail(__FILE__, __LINE__, "overlapped sweep out of step"); break; }
                if (pb == b) {  // a single block: nothing can overlap, exchange first
                    rc = cgs_model_exchange_flush(m);
                    if (rc != 0) break;
                } else {
                    fill_exchange_params(m, m->x_fused, c->block_bounds[(size_t)pb], c->block_bounds[(size_t)pb + 1]);
                    m->x_launches += 1;
                }
            }
            m->sweep_block = b;
            m->sweep_x_ctas = (fused && m->x_pending_block >= 0) ? x_ctas : 0;
            const int64_t before = m->sweeps;
            rc = sweep_impl(m, 1, 0, 1);
            m->sweeps = before;  // the blocks of one sweep share the Philox sweep index
            m->sweep_block = -1;
            m->sweep_x_ctas = 0;
            if (exchange) m->x_pending_block = b;
        }
        if (rc == 0) m->sweeps += 1;
    }
    return rc;
}

int cgs_model_sweep_blocked(cgs_model *m, int32_t n_sweeps) {
    if (!m) CGS_FAIL("NULL model");
    if (n_sweeps < 0) CGS_FAIL("n_sweeps < 0");
    return sweep_region_blocks(m, n_sweeps, false, 0);
}

int cgs_model_sweep_overlapped(cgs_model *m, int32_t n_sweeps, int32_t x_ctas) {
    if (!m) CGS_FAIL("NULL model");
    if (n_sweeps < 0) CGS_FAIL("n_sweeps < 0");
    if (m->x_ranks < 1) CGS_FAIL("cgs_model_exchange_open has not been called");
    cgs_corpus *c = m->corpus;
    // measured on 8 ranks (C3, K = 60, 128 MB table; profiles/r2_adlda_c3_n8_overlap.txt): 74 exchange blocks leave 0.02 ms
    // of the exchange exposed, 148 leave 0.10 ms, 296 leave 0.25 ms -- the blocks only have to finish before the
    // block sweep does, and every one of them takes a sampling block's place
    if (x_ctas <= 0) x_ctas = std::max(1, c->num_sms / 2);
    x_ctas = std::min(x_ctas, 4 * c->num_sms);
    return sweep_region_blocks(m, n_sweeps, true, x_ctas);
}

static int exchange_sync_rows(cgs_model *m, int32_t n_ctas, int64_t row0, int64_t row1) {
    if (m->x_ranks < 1) CGS_FAIL("cgs_model_exchange_open has not been called");
    cgs_corpus *c = m->corpus;
    DeviceGuard g(c->device);
    if (m->x_sync_blocks == 0) {
        int per_sm = 0;
        CGS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, exchange_sync_kernel, kExchangeThreads,
                                                                sizeof(int) * (size_t)m->Kp));
        m->x_sync_blocks = c->num_sms * std::max(1, std::min(per_sm, 4));
    }
    // every CTA waits at the cross-rank barriers: the grid must be resident at once
    n_ctas = n_ctas <= 0 ? m->x_sync_blocks : std::min(n_ctas, m->x_sync_blocks);
    ExchangeParams p;
    fill_exchange_params(m, p, row0, row1);
    NvtxRange nvtx("cgs_model_exchange_sync");
    exchange_sync_kernel<<<n_ctas, kExchangeThreads, sizeof(int) * (size_t)m->Kp, m->stream>>>(p);
    CGS_CUDA(cudaGetLastError());
    m->launches += 1;
    m->x_launches += 1;
    return 0;
}

static int exchange_local(cgs_model *const *models, int32_t n_models, int32_t n_ctas, int sync_variant);

int cgs_exchange_round_local(cgs_model *const *models, int32_t n_models, int32_t n_ctas) {
    return exchange_local(models, n_models, n_ctas, 0);
}

int cgs_exchange_sync_local(cgs_model *const *models, int32_t n_models, int32_t n_ctas) {
    return exchange_local(models, n_models, n_ctas, 1);
}

static int exchange_local(cgs_model *const *models, int32_t n_models, int32_t n_ctas, int sync_variant) {
    if (!models) CGS_FAIL("NULL argument");
    if (n_models < 1 || n_models > kMaxLocalRanks) CGS_FAIL("n_models out of range");
    for (int i = 0; i < n_models; ++i) {
        if (!models[i]) CGS_FAIL("NULL model");
        if (models[i]->x_ranks < 1) CGS_FAIL("cgs_model_exchange_open has not been called");
        if (models[i]->corpus->device != models[0]->corpus->device) CGS_FAIL("models must share one device");
        if (models[i]->Kp != models[0]->Kp) CGS_FAIL("models must share n_topics");
    }
    cgs_model *lead = models[0];
    DeviceGuard g(lead->corpus->device);
    if (n_ctas <= 0) n_ctas = 16;
    ExchangeParamsMulti a;
    memset(&a, 0, sizeof a);
    for (int i = 0; i < n_models; ++i) {
        CGS_CUDA(cudaEventRecord(models[i]->x_event, models[i]->stream));
        CGS_CUDA(cudaStreamWaitEvent(lead->x_stream, models[i]->x_event, 0));
        fill_exchange_params(models[i], a.r[i]);
    }
    void *args[] = {&a, &sync_variant};
    CGS_CUDA(cudaLaunchCooperativeKernel((const void *)exchange_round_multi_kernel, dim3((unsigned)n_ctas, (unsigned)n_models),
                                         dim3(kExchangeThreads), args, sizeof(int) * (size_t)lead->Kp, lead->x_stream));
    for (int i = 0; i < n_models; ++i) {
        CGS_CUDA(cudaEventRecord(models[i]->x_done, lead->x_stream));
        models[i]->x_pending = true;
    }
    lead->launches += 1;
    lead->x_launches += 1;
    return 0;
}

int cgs_model_exchange_wait(cgs_model *m) {
    if (!m) CGS_FAIL("NULL model");
    DeviceGuard g(m->corpus->device);
    if (m->x_pending) {
        CGS_CUDA(cudaStreamWaitEvent(m->stream, m->x_done, 0));
        m->x_pending = false;
    }
    return 0;
}

int cgs_model_exchange_status(cgs_model *m, int32_t *timeouts) {
    if (!m || !timeouts) CGS_FAIL("NULL argument");
    DeviceGuard g(m->corpus->device);
    *timeouts = 0;
    if (!m->d_arena) return 0;
    if (m->x_stream) CGS_CUDA(cudaStreamSynchronize(m->x_stream));
    uint32_t v = 0;
    const char *flags = m->d_arena + 2 * arena_delta_bytes(m);
    CGS_CUDA(cudaMemcpy(&v, flags + (2 * kMaxExchangeRanks + 2) * 4, 4, cudaMemcpyDeviceToHost));
    *timeouts = (int32_t)v;
    return 0;
}

int cgs_model_exchange_phase_times(cgs_model *m, double *us5) {
    if (!m || !us5) CGS_FAIL("NULL argument");
    DeviceGuard g(m->corpus->device);
    if (!m->d_arena) CGS_FAIL("no exchange arena");
    CGS_CUDA(cudaStreamSynchronize(m->stream));
    unsigned long long t[6];
    const char *flags = m->d_arena + 2 * arena_delta_bytes(m);
    CGS_CUDA(cudaMemcpy(t, flags + kExchangeStampWord * 4, sizeof t, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 5; ++i) us5[i] = (double)(t[i + 1] - t[i]) * 1e-3;
    return 0;
}

// ---- NCCL behind the ABI (cgs_allreduce_delta) -------------------------------------------
// The library does not link NCCL: the entry points are resolved at first use from the libnccl.so.2 the host
// process has already loaded (torch's) or that the loader finds.  Only the few signatures used are declared.
struct NcclId128 { char bytes[128]; };  // ncclUniqueId (passed by value to ncclCommInitRank)
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId128, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static std::mutex g_nccl_mutex;

static int nccl_api() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);  // shards driven by several host threads resolve it once
    if (g_nccl.lib) return 0;
    // the copy this process already uses, else the one the host names (pycistopic_b200/_lib.py points CGS_NCCL_LIBRARY at
    // the NCCL wheel torch will load, so that a later `import torch` meets its own library), else the loader's choice
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    const char *named = getenv("CGS_NCCL_LIBRARY");
    if (!h && named && named[0]) h = dlopen(named, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) CGS_FAIL(std::string("cannot load libnccl.so.2: ") + dlerror());
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.GetErrorString)
        CGS_FAIL("libnccl.so.2 lacks an expected entry point");
    g_nccl.lib = h;
    return 0;
}
#define CGS_NCCL(expr)                                                                                       \
    do {                                                                                                     \
        int _r = (expr);                                                                                     \
        if (_r != 0) return ::cgs::fail(__FILE__, __LINE__, std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); \
    } while (0)

int cgs_nccl_unique_id(void *id_out_128_bytes) {
    if (!id_out_128_bytes) CGS_FAIL("NULL argument");
    CGS_TRY(nccl_api());
    CGS_NCCL(g_nccl.GetUniqueId(id_out_128_bytes));
    return 0;
}

int cgs_nccl_comm_create(void **comm, int device, int32_t n_ranks, int32_t rank, const void *id_128_bytes) {
    if (!comm || !id_128_bytes) CGS_FAIL("NULL argument");
    CGS_TRY(nccl_api());
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    NcclId128 id;
    memcpy(&id, id_128_bytes, sizeof id);
    CGS_NCCL(g_nccl.CommInitRank(comm, n_ranks, id, rank));
    return 0;
}

int cgs_nccl_comm_destroy(void *comm) {
    if (!comm) return 0;
    CGS_TRY(nccl_api());
    CGS_NCCL(g_nccl.CommDestroy(comm));
    return 0;
}

int cgs_model_allreduce_delta(cgs_model *m, void *nccl_comm) {
    if (!m || !nccl_comm) CGS_FAIL("NULL argument");
    CGS_TRY(nccl_api());
    DeviceGuard g(m->corpus->device);
    NvtxRange nvtx("cgs_model_allreduce_delta");
    CGS_TRY(cgs_model_delta_compute(m));
    const size_t n = (size_t)m->corpus->n_regions * (size_t)m->Kp;
    // ncclInt32 = 2, ncclSum = 0 (nccl.h); in place on the delta buffer
    CGS_NCCL(g_nccl.AllReduce(m->d_delta, m->d_delta, n, 2, 0, nccl_comm, m->stream));
    CGS_TRY(cgs_model_delta_apply(m));
    return 0;
}

// ---- imputation --------------------------------------------------------------------------------
}  // extern "C"

struct cgs_impute_plan {
    int device = 0;
    int num_sms = kNumSMsFallback;
    int32_t K = 0, Kp = 0, C = 0;
    float *d_cells_raw = nullptr;           // K x C copy of cell_topic (cross-check path only)
    float *d_b_hi = nullptr, *d_b_lo = nullptr;  // C x Kp, TF32 split, K-major
    int64_t launches = 0;
};

namespace cgs {
static bool impute_use_simt() {
    const char *e = getenv("CGS_IMPUTE_SIMT");
    return e && e[0] == '1';
}

// Splits the region operand, runs the tensor-core kernel against the plan's cell operand.
static int impute_run(cgs_impute_plan *p, const float *d_A, int32_t R, float scale, int as_int32, void *d_out,
                      uint8_t *d_row_nonzero, cudaStream_t stream) {
    if (impute_use_simt()) {
        CGS_TRY(impute_simt_launch(d_A, p->d_cells_raw, R, p->K, p->C, scale, as_int32, d_out, d_row_nonzero, stream));
        p->launches += 1;
        return 0;
    }
    float *d_a_hi = nullptr, *d_a_lo = nullptr;
    const size_t n = (size_t)R * p->Kp;
    CGS_TRY(dmalloc(&d_a_hi, n, stream));
    CGS_TRY(dmalloc(&d_a_lo, n, stream));
    int rc = [&]() -> int {
        CGS_TRY(impute_prepare_regions(d_A, R, p->K, p->Kp, d_a_hi, d_a_lo, p->num_sms, stream));
        CGS_TRY(impute_umma_launch(d_a_hi, d_a_lo, p->d_b_hi, p->d_b_lo, R, p->K, p->Kp, p->C, scale, as_int32, d_out,
                                   d_row_nonzero, p->num_sms, stream));
        p->launches += 2;
        return 0;
    }();
    dfree(d_a_hi, stream);
    dfree(d_a_lo, stream);
    return rc;
}
}  // namespace cgs

extern "C" {

int cgs_impute_plan_create(cgs_impute_plan **out, int device, const float *cell_topic, int cell_topic_on_device,
                           int32_t n_topics, int32_t n_cells, void *stream_) {
    if (!out || !cell_topic) CGS_FAIL("NULL argument");
    if (n_topics <= 0 || n_cells <= 0) CGS_FAIL("empty operand");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    cudaStream_t stream = (cudaStream_t)stream_;
    cgs_impute_plan *p = new cgs_impute_plan();
    *out = nullptr;
    int rc = [&]() -> int {
        p->device = device;
        CGS_TRY(query_sms(device, &p->num_sms));
        CGS_TRY(configure_pool(device));
        p->K = n_topics;
        p->Kp = impute_padded_k(n_topics);
        p->C = n_cells;
        const size_t nb = (size_t)n_topics * n_cells;
        CGS_TRY(dmalloc(&p->d_cells_raw, nb, stream));
        CGS_CUDA(cudaMemcpyAsync(p->d_cells_raw, cell_topic, nb * 4,
                                 cell_topic_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
        CGS_TRY(dmalloc(&p->d_b_hi, (size_t)n_cells * p->Kp, stream));
        CGS_TRY(dmalloc(&p->d_b_lo, (size_t)n_cells * p->Kp, stream));
        CGS_TRY(impute_prepare_cells(p->d_cells_raw, p->K, p->C, p->Kp, p->d_b_hi, p->d_b_lo, stream));
        p->launches += 1;
        if (!impute_use_simt()) {  // the raw copy only serves the cross-check kernel
            dfree(p->d_cells_raw, stream);
            p->d_cells_raw = nullptr;
        }
        CGS_CUDA(cudaStreamSynchronize(stream));  // later chunks may run on other streams
        return 0;
    }();
    if (rc != 0) { cgs_impute_plan_destroy(p); return rc; }
    *out = p;
    return 0;
}

int cgs_impute_plan_destroy(cgs_impute_plan *p) {
    if (!p) return 0;
    DeviceGuard g(p->device);
    cudaDeviceSynchronize();
    dfree(p->d_cells_raw, 0); dfree(p->d_b_hi, 0); dfree(p->d_b_lo, 0);
    delete p;
    return 0;
}

int cgs_impute_plan_launch_count(const cgs_impute_plan *p, int64_t *launches) {
    if (!p || !launches) CGS_FAIL("NULL argument");
    *launches = p->launches;
    return 0;
}

int cgs_impute_plan_chunk_device(cgs_impute_plan *p, const float *d_topic_region, int32_t n_regions, float scale,
                                 int as_int32, void *d_out, uint8_t *d_row_nonzero, void *stream) {
    if (!p || !d_topic_region || !d_out) CGS_FAIL("NULL argument");
    if (n_regions <= 0) CGS_FAIL("empty operand");
    DeviceGuard g(p->device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    return impute_run(p, d_topic_region, n_regions, scale, as_int32, d_out, d_row_nonzero, (cudaStream_t)stream);
}

int cgs_impute_plan_chunk(cgs_impute_plan *p, const float *topic_region, int32_t n_regions, float scale, int as_int32,
                          void *out, uint8_t *row_nonzero) {
    if (!p || !topic_region || !out) CGS_FAIL("NULL argument");
    if (n_regions <= 0) CGS_FAIL("empty operand");
    DeviceGuard g(p->device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    float *d_a = nullptr;
    uint8_t *d_o = nullptr, *d_f = nullptr;
    const size_t na = (size_t)n_regions * p->K, no = (size_t)n_regions * p->C;
    int rc = [&]() -> int {
        CGS_TRY(dmalloc(&d_a, na, 0));
        CGS_TRY(dmalloc(&d_o, no * 4, 0));
        CGS_TRY(dmalloc(&d_f, (size_t)n_regions, 0));
        CGS_CUDA(cudaMemcpyAsync(d_a, topic_region, na * 4, cudaMemcpyHostToDevice, 0));
        CGS_TRY(impute_run(p, d_a, n_regions, scale, as_int32, d_o, d_f, 0));
        CGS_CUDA(cudaMemcpyAsync(out, d_o, no * 4, cudaMemcpyDeviceToHost, 0));
        if (row_nonzero) CGS_CUDA(cudaMemcpyAsync(row_nonzero, d_f, (size_t)n_regions, cudaMemcpyDeviceToHost, 0));
        CGS_CUDA(cudaStreamSynchronize(0));
        return 0;
    }();
    dfree(d_a, 0); dfree(d_o, 0); dfree(d_f, 0);
    return rc;
}

int cgs_impute_chunk_device(int device, const float *d_topic_region, const float *d_cell_topic, int32_t n_regions,
                            int32_t n_topics, int32_t n_cells, float scale, int as_int32, void *d_out,
                            uint8_t *d_row_nonzero, void *stream) {
    if (!d_topic_region || !d_cell_topic || !d_out) CGS_FAIL("NULL argument");
    if (n_regions <= 0 || n_topics <= 0 || n_cells <= 0) CGS_FAIL("empty operand");
    cgs_impute_plan *p = nullptr;
    CGS_TRY(cgs_impute_plan_create(&p, device, d_cell_topic, 1, n_topics, n_cells, stream));
    int rc = cgs_impute_plan_chunk_device(p, d_topic_region, n_regions, scale, as_int32, d_out, d_row_nonzero, stream);
    if (rc == 0) {
        DeviceGuard g(device);
        cudaStreamSynchronize((cudaStream_t)stream);
    }
    cgs_impute_plan_destroy(p);
    return rc;
}

int cgs_impute_chunk(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                     int32_t n_topics, int32_t n_cells, float scale, int as_int32, void *out, uint8_t *row_nonzero) {
    if (!topic_region || !cell_topic || !out) CGS_FAIL("NULL argument");
    if (n_regions <= 0 || n_topics <= 0 || n_cells <= 0) CGS_FAIL("empty operand");
    cgs_impute_plan *p = nullptr;
    CGS_TRY(cgs_impute_plan_create(&p, device, cell_topic, 0, n_topics, n_cells, nullptr));
    int rc = cgs_impute_plan_chunk(p, topic_region, n_regions, scale, as_int32, out, row_nonzero);
    cgs_impute_plan_destroy(p);
    return rc;
}

// ---- normalize_scores ---------------------------------------------------------------------------
int cgs_normalize_scores_device(int device, const void *d_x, int dtype, int64_t n_rows, int64_t n_cols,
                                double scale_factor, void *d_out, void *d_colsum_scratch, void *stream_) {
    if (!d_x || !d_out || !d_colsum_scratch) CGS_FAIL("NULL argument");
    if (n_rows <= 0 || n_cols <= 0) CGS_FAIL("empty matrix");
    if (!(scale_factor != 0.0)) CGS_FAIL("scale_factor must not be zero");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    cudaStream_t stream = (cudaStream_t)stream_;
    switch (dtype) {
        case 0: return normalize_launch<int32_t, double>((const int32_t *)d_x, n_rows, n_cols, scale_factor, d_colsum_scratch,
                                                         (double *)d_out, stream);
        case 1: return normalize_launch<float, float>((const float *)d_x, n_rows, n_cols, scale_factor, d_colsum_scratch,
                                                      (float *)d_out, stream);
        case 2: return normalize_launch<double, double>((const double *)d_x, n_rows, n_cols, scale_factor, d_colsum_scratch,
                                                        (double *)d_out, stream);
    }
    CGS_FAIL("dtype must be 0 (int32), 1 (float32) or 2 (float64)");
}

int cgs_normalize_scores(int device, const void *x, int dtype, int64_t n_rows, int64_t n_cols, double scale_factor,
                         void *out) {
    if (!x || !out) CGS_FAIL("NULL argument");
    if (n_rows <= 0 || n_cols <= 0) CGS_FAIL("empty matrix");
    if (dtype < 0 || dtype > 2) CGS_FAIL("dtype must be 0 (int32), 1 (float32) or 2 (float64)");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    CGS_TRY(configure_pool(device));
    const size_t in_elem = dtype == 2 ? 8 : 4, out_elem = dtype == 1 ? 4 : 8;
    const size_t n = (size_t)n_rows * (size_t)n_cols;
    uint8_t *d_x = nullptr, *d_o = nullptr, *d_s = nullptr;
    HostSource source;
    HostSink sink;
    cudaStream_t stream = nullptr;
    CGS_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int rc = [&]() -> int {
        CGS_TRY(dmalloc(&d_x, n * in_elem, stream));
        CGS_TRY(dmalloc(&d_o, n * out_elem, stream));
        CGS_TRY(dmalloc(&d_s, (size_t)n_cols * 8, stream));
        CGS_TRY(source.open(n * in_elem));
        CGS_TRY(sink.open(device, n * out_elem));
        CGS_TRY(source.copy(d_x, x, n * in_elem, stream));
        CGS_TRY(cgs_normalize_scores_device(device, d_x, dtype, n_rows, n_cols, scale_factor, d_o, d_s, stream));
        CGS_TRY(sink.copy(out, d_o, n * out_elem, stream));
        sink.finish();
        CGS_CUDA(cudaStreamSynchronize(stream));
        return 0;
    }();
    source.close(); sink.close();
    dfree(d_x, stream); dfree(d_o, stream); dfree(d_s, stream);
    cudaStreamDestroy(stream);
    return rc;
}


// ---- per-region mean / variance in front of find_highly_variable_features (diff_features.py:636-639) ----------
int cgs_row_mean_var_device(int device, int dtype, const void *d_x, int64_t ld, int64_t n_rows, int64_t n_cols,
                            float *d_mean, float *d_var, void *stream_) {
    if (!d_x || !d_mean || !d_var) CGS_FAIL("NULL argument");
    if (n_rows <= 0 || n_cols <= 0) CGS_FAIL("empty matrix");
    if (ld < n_cols) CGS_FAIL("row stride smaller than the number of columns");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    int sms = kNumSMsFallback;
    CGS_TRY(query_sms(device, &sms));
    cudaStream_t stream = (cudaStream_t)stream_;
    switch (dtype) {
        case 0: return row_mean_var_launch<int32_t>((const int32_t *)d_x, ld, n_rows, n_cols, d_mean, d_var, sms, stream);
        case 1: return row_mean_var_launch<float>((const float *)d_x, ld, n_rows, n_cols, d_mean, d_var, sms, stream);
        case 2: return row_mean_var_launch<double>((const double *)d_x, ld, n_rows, n_cols, d_mean, d_var, sms, stream);
    }
    CGS_FAIL("dtype must be 0 (int32), 1 (float32) or 2 (float64)");
}

int cgs_row_mean_var(int device, int dtype, const void *x, int64_t ld, int64_t n_rows, int64_t n_cols, float *mean,
                     float *var) {
    if (!x || !mean || !var) CGS_FAIL("NULL argument");
    if (n_rows <= 0 || n_cols <= 0) CGS_FAIL("empty matrix");
    if (dtype < 0 || dtype > 2) CGS_FAIL("dtype must be 0 (int32), 1 (float32) or 2 (float64)");
    if (ld < n_cols) CGS_FAIL("row stride smaller than the number of columns");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    CGS_TRY(configure_pool(device));
    const size_t elem = dtype == 2 ? 8 : 4;
    // rows per round: about 1 GB on the device, two buffers so the next upload overlaps the reduction
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n_rows, (int64_t)((1LL << 30) / ((size_t)n_cols * elem))));
    uint8_t *d_x[2] = {nullptr, nullptr};
    float *d_m = nullptr, *d_v = nullptr;
    HostSource source;
    cudaStream_t stream = nullptr;
    CGS_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int rc = [&]() -> int {
        CGS_TRY(source.open((size_t)n_rows * n_cols * elem));
        CGS_TRY(dmalloc(&d_x[0], (size_t)chunk * n_cols * elem, stream));
        CGS_TRY(dmalloc(&d_x[1], (size_t)chunk * n_cols * elem, stream));
        CGS_TRY(dmalloc(&d_m, (size_t)n_rows, stream));
        CGS_TRY(dmalloc(&d_v, (size_t)n_rows, stream));
        int which = 0;
        for (int64_t r0 = 0; r0 < n_rows; r0 += chunk, which ^= 1) {
            const int64_t n = std::min(chunk, n_rows - r0);
            if (ld == n_cols) {
                CGS_TRY(source.copy(d_x[which], (const uint8_t *)x + (size_t)r0 * ld * elem, (size_t)n * n_cols * elem, stream));
            } else {
                CGS_CUDA(cudaMemcpy2DAsync(d_x[which], (size_t)n_cols * elem, (const uint8_t *)x + (size_t)r0 * ld * elem,
                                           (size_t)ld * elem, (size_t)n_cols * elem, (size_t)n, cudaMemcpyHostToDevice, stream));
            }
            CGS_TRY(cgs_row_mean_var_device(device, dtype, d_x[which], n_cols, n, n_cols, d_m + r0, d_v + r0, stream));
        }
        CGS_CUDA(cudaMemcpyAsync(mean, d_m, sizeof(float) * (size_t)n_rows, cudaMemcpyDeviceToHost, stream));
        CGS_CUDA(cudaMemcpyAsync(var, d_v, sizeof(float) * (size_t)n_rows, cudaMemcpyDeviceToHost, stream));
        CGS_CUDA(cudaStreamSynchronize(stream));
        return 0;
    }();
    cudaStreamSynchronize(stream);
    source.close();
    dfree(d_x[0], stream); dfree(d_x[1], stream); dfree(d_m, stream); dfree(d_v, stream);
    cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream);
    return rc;
}

// ---- make_rankings (diff_features.py:274-337): per-cell rankings of the imputed matrix -----------------------
}  // extern "C"

template <typename T>
static int rank_columns_launch(const T *d_x, int64_t ld, int64_t n_rows, int64_t n_cols, uint64_t seed, int64_t first_col,
                               int32_t *d_out, int64_t ld_out, int device, cudaStream_t stream) {
    using K = typename RankKey<T>::type;
    auto kernel = rank_columns_kernel<T>;
    // clusters in flight: what the SMs can hold, but no more columns than L2 keeps resident (scores [+ rankings
    // for 8-byte scores] + the ping-pong pairs the passes touch: one for the two passes of small integers)
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = kRkCluster; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.gridDim = dim3(kRkCluster); cfg.blockDim = dim3(kRkThreads); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int max_clusters = 1, l2_bytes = 0;
    CGS_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg));
    CGS_CUDA(cudaDeviceGetAttribute(&l2_bytes, cudaDevAttrL2CacheSize, device));
    const int64_t per_value = std::is_same<T, int32_t>::value ? 12 : std::is_same<T, float>::value ? 20 : 36;
    int64_t n_clusters = (int64_t)(0.7 * (double)l2_bytes) / (n_rows * per_value);
    if (const char *e = getenv("CGS_RANK_CLUSTERS")) { if (atoll(e) > 0) n_clusters = atoll(e); }  // measurement knob
    n_clusters = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(n_clusters, std::max(1, max_clusters)), n_cols));
    const size_t n = (size_t)n_rows * (size_t)n_cols, scratch = (size_t)n_clusters * (size_t)n_rows;
    T *d_xT = nullptr;
    int32_t *d_rT = nullptr;
    K *d_k0 = nullptr, *d_k1 = nullptr;
    uint32_t *d_p0 = nullptr, *d_p1 = nullptr;
    int rc = [&]() -> int {
        CGS_TRY(dmalloc(&d_xT, n, stream));
        if (sizeof(T) == 8) CGS_TRY(dmalloc(&d_rT, n, stream));
        CGS_TRY(dmalloc(&d_k0, scratch, stream));
        CGS_TRY(dmalloc(&d_k1, scratch, stream));
        CGS_TRY(dmalloc(&d_p0, scratch, stream));
        CGS_TRY(dmalloc(&d_p1, scratch, stream));
        const int64_t tiles = ceil_div(n_rows, 32) * ceil_div(n_cols, 32);
        if (tiles > 2147483647LL) CGS_FAIL("matrix chunk too large for one launch");
        rk_transpose_kernel<T><<<(unsigned)tiles, dim3(32, 8), 0, stream>>>(d_x, ld, n_rows, n_cols, d_xT, n_rows);
        CGS_CUDA(cudaGetLastError());
        RankColsParams p;
        p.xT = d_xT; p.rT = d_rT;
        p.keys[0] = d_k0; p.keys[1] = d_k1; p.pay[0] = d_p0; p.pay[1] = d_p1;
        p.n_rows = (uint32_t)n_rows; p.n_cols = n_cols; p.seed = seed; p.first_col = first_col;
        p.half_bits = rk_half_bits((uint32_t)n_rows);
        cfg.gridDim = dim3((unsigned)(n_clusters * kRkCluster));
        CGS_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
        const int32_t *ranks = sizeof(T) == 8 ? d_rT : reinterpret_cast<const int32_t *>(d_xT);
        rk_transpose_kernel<int32_t><<<(unsigned)tiles, dim3(32, 8), 0, stream>>>(ranks, n_rows, n_cols, n_rows, d_out, ld_out);
        CGS_CUDA(cudaGetLastError());
        return 0;
    }();
    dfree(d_xT, stream); dfree(d_rT, stream); dfree(d_k0, stream); dfree(d_k1, stream); dfree(d_p0, stream); dfree(d_p1, stream);
    return rc;
}

extern "C" {

int cgs_rank_columns_device(int device, int dtype, const void *d_x, int64_t ld, int64_t n_rows, int64_t n_cols,
                            uint64_t seed, int64_t first_col, int32_t *d_out, int64_t ld_out, void *stream_) {
    if (!d_x || !d_out) CGS_FAIL("NULL argument");
    if (n_rows <= 0 || n_cols <= 0) CGS_FAIL("empty matrix");
    if (n_rows > 2147483647LL) CGS_FAIL("more than 2^31 - 1 rows");
    if (ld < n_cols || ld_out < n_cols) CGS_FAIL("row stride smaller than the number of columns");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    cudaStream_t stream = (cudaStream_t)stream_;
    switch (dtype) {
        case 0: return rank_columns_launch<int32_t>((const int32_t *)d_x, ld, n_rows, n_cols, seed, first_col, d_out, ld_out, device, stream);
        case 1: return rank_columns_launch<float>((const float *)d_x, ld, n_rows, n_cols, seed, first_col, d_out, ld_out, device, stream);
        case 2: return rank_columns_launch<double>((const double *)d_x, ld, n_rows, n_cols, seed, first_col, d_out, ld_out, device, stream);
    }
    CGS_FAIL("dtype must be 0 (int32), 1 (float32) or 2 (float64)");
}

int cgs_rank_columns(int device, int dtype, const void *x, int64_t ld, int64_t n_rows, int64_t n_cols, uint64_t seed,
                     int64_t first_col, int32_t *out, int64_t ld_out) {
    if (!x || !out) CGS_FAIL("NULL argument");
    if (n_rows <= 0 || n_cols <= 0) CGS_FAIL("empty matrix");
    if (dtype < 0 || dtype > 2) CGS_FAIL("dtype must be 0 (int32), 1 (float32) or 2 (float64)");
    if (n_rows > 2147483647LL) CGS_FAIL("more than 2^31 - 1 rows");
    if (ld < n_cols || ld_out < n_cols) CGS_FAIL("row stride smaller than the number of columns");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    CGS_TRY(configure_pool(device));
    const size_t elem = dtype == 2 ? 8 : 4;
    // columns per round: about 2^29 values on the device at a time (2-4 GB per staging buffer)
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n_cols, std::max<int64_t>(32, (1LL << 29) / n_rows)));
    uint8_t *d_x = nullptr;
    int32_t *d_o = nullptr;
    HostSource source;
    HostSink sink;
    cudaStream_t stream = nullptr;
    CGS_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    int rc = [&]() -> int {
        CGS_TRY(source.open((size_t)n_rows * n_cols * elem));
        CGS_TRY(sink.open(device, (size_t)n_rows * n_cols * 4));
        CGS_TRY(dmalloc(&d_x, (size_t)n_rows * (size_t)chunk * elem, stream));
        CGS_TRY(dmalloc(&d_o, (size_t)n_rows * (size_t)chunk, stream));
        for (int64_t c0 = 0; c0 < n_cols; c0 += chunk) {
            const int64_t nc = std::min(chunk, n_cols - c0);
            CGS_TRY(source.copy2d(d_x, (const uint8_t *)x + (size_t)c0 * elem, (size_t)ld * elem, (size_t)nc * elem,
                                  (size_t)n_rows, stream));
            CGS_TRY(cgs_rank_columns_device(device, dtype, d_x, nc, n_rows, nc, seed, first_col + c0, d_o, nc, stream));
            CGS_TRY(sink.copy2d(out + c0, (size_t)ld_out * 4, d_o, (size_t)nc * 4, (size_t)n_rows, stream));
            sink.finish();
            CGS_CUDA(cudaStreamSynchronize(stream));
        }
        return 0;
    }();
    cudaStreamSynchronize(stream);
    source.close(); sink.close();
    dfree(d_x, stream); dfree(d_o, stream);
    cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream);
    return rc;
}

}  // extern "C"

// ---- model-quality metrics on the device (lda_models.py:291-305, :332-344) ------------
extern "C" {

int cgs_corpus_export_cell_sizes(const cgs_corpus *c, int64_t *sizes) {
    if (!c || !sizes) CGS_FAIL("NULL argument");
    DeviceGuard g(c->device);
    if (c->n_cells == 0) return 0;
    int64_t *d_s = nullptr;
    CGS_TRY(dmalloc(&d_s, (size_t)c->n_cells, c->stream));
    cell_sizes_kernel<<<(unsigned)std::min<int64_t>(c->num_sms * 8, ceil_div(c->n_cells, 256)), 256, 0, c->stream>>>(
        c->d_cell_ptr, c->n_cells, d_s);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cgs_memcpy_sync(c->stream, sizes, d_s, sizeof(int64_t) * (size_t)c->n_cells, cudaMemcpyDeviceToHost);
    dfree(d_s, c->stream);
    CGS_CUDA(e);
    return 0;
}

int cgs_model_topic_gram(cgs_model *m, double *gram) {
    if (!m || !gram) CGS_FAIL("NULL argument");
    cgs_corpus *c = m->corpus;
    DeviceGuard g(c->device);
    const int64_t W = c->n_regions;
    const int32_t K = m->K, Kp = m->Kp;
    unsigned long long *d_g = nullptr;
    CGS_TRY(dmalloc(&d_g, (size_t)Kp * Kp, m->stream));
    std::vector<long long> S((size_t)Kp * Kp);
    std::vector<int32_t> nk(Kp);
    int rc = [&]() -> int {
        CGS_CUDA(cudaMemsetAsync(d_g, 0, sizeof(unsigned long long) * (size_t)Kp * Kp, m->stream));
        const int nt = Kp / 4, n_tiles = nt * (nt + 1) / 2;
        const int gy = (int)ceil_div(n_tiles, kGramThreads);
        const int64_t n_slabs = ceil_div(W, kGramRows);
        const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(n_slabs, std::max(1, c->num_sms * 8 / gy)));
        const size_t smem = sizeof(int32_t) * kGramRows * (size_t)Kp;
        topic_gram_kernel<<<dim3(gx, gy), kGramThreads, smem, m->stream>>>(m->d_nwk, W, Kp, n_tiles, d_g);
        CGS_CUDA(cudaGetLastError());
        m->launches += 1;
        CGS_CUDA(cudaMemcpyAsync(S.data(), d_g, sizeof(long long) * S.size(), cudaMemcpyDeviceToHost, m->stream));
        CGS_CUDA(cudaMemcpyAsync(nk.data(), m->d_nk, sizeof(int32_t) * (size_t)Kp, cudaMemcpyDeviceToHost, m->stream));
        CGS_CUDA(cudaStreamSynchronize(m->stream));
        return 0;
    }();
    dfree(d_g, m->stream);
    CGS_TRY(rc);
    // topic_word[k, w] = (n_wk[w,k] + eta) / (n_k[k] + eta W)  =>
    // gram[i][j] = (S_ij + eta (n_i + n_j) + eta^2 W) / ((n_i + eta W)(n_j + eta W)), S exact
    const double eta = m->eta, etaW = m->eta * (double)W;
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) {
            const long long sij = i <= j ? S[(size_t)i * Kp + j] : S[(size_t)j * Kp + i];
            const double num = (double)sij + eta * ((double)nk[i] + (double)nk[j]) + eta * eta * (double)W;
            gram[(size_t)i * K + j] = num / (((double)nk[i] + etaW) * ((double)nk[j] + etaW));
        }
    return 0;
}

int cgs_model_cell_mixture(cgs_model *m, double *mixture, double *length_sq) {
    if (!m || !mixture || !length_sq) CGS_FAIL("NULL argument");
    cgs_corpus *c = m->corpus;
    DeviceGuard g(c->device);
    const int32_t K = m->K, D = c->n_cells;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(c->num_sms * 2, ceil_div(D, 8)));
    double *d_p = nullptr;
    CGS_TRY(dmalloc(&d_p, (size_t)(blocks + 1) * (K + 1), m->stream));
    std::vector<double> h(K + 1);
    int rc = [&]() -> int {
        cell_mixture_kernel<<<blocks, dim3(32, 8), 0, m->stream>>>(m->d_ndk, c->d_cell_ptr, D, K, m->Kp, m->alpha, d_p);
        double *d_out = d_p + (size_t)blocks * (K + 1);
        column_sum_partials_kernel<<<(unsigned)ceil_div(K + 1, 128), 128, 0, m->stream>>>(d_p, blocks, K + 1, d_out);
        CGS_CUDA(cudaGetLastError());
        m->launches += 2;
        CGS_CUDA(cudaMemcpyAsync(h.data(), d_out, sizeof(double) * (size_t)(K + 1), cudaMemcpyDeviceToHost, m->stream));
        CGS_CUDA(cudaStreamSynchronize(m->stream));
        return 0;
    }();
    dfree(d_p, m->stream);
    CGS_TRY(rc);
    std::copy(h.begin(), h.begin() + K, mixture);
    *length_sq = h[K];
    return 0;
}

int cgs_model_top_regions(cgs_model *m, int32_t top_n, int32_t *top_regions, int32_t *top_counts) {
    if (!m || !top_regions) CGS_FAIL("NULL argument");
    cgs_corpus *c = m->corpus;
    if (top_n < 1 || top_n > 32) CGS_FAIL("top_n must be in [1, 32]");
    if (top_n > c->n_regions)
        CGS_FAIL("`top_n=" + std::to_string(top_n) + "` is larger than the vocabulary size of " +
                 std::to_string(c->n_regions) + " words");
    DeviceGuard g(c->device);
    const int32_t K = m->K;
    const int n_slabs = (int)ceil_div(c->n_regions, kTopSlab);
    unsigned long long *d_cand = nullptr;
    int32_t *d_top = nullptr;
    CGS_TRY(dmalloc(&d_cand, (size_t)n_slabs * K * 32, m->stream));
    int rc = dmalloc(&d_top, (size_t)2 * K * top_n, m->stream);
    std::vector<int32_t> h((size_t)2 * K * top_n);
    if (rc == 0) rc = [&]() -> int {
        top_regions_slab_kernel<<<dim3(n_slabs, (unsigned)ceil_div(K, kTopWarps)), kTopWarps * 32, 0, m->stream>>>(
            m->d_nwk, c->n_regions, K, m->Kp, d_cand);
        top_regions_merge_kernel<<<(unsigned)ceil_div(K, kTopWarps), kTopWarps * 32, 0, m->stream>>>(
            d_cand, n_slabs, K, top_n, d_top, d_top + (size_t)K * top_n);
        CGS_CUDA(cudaGetLastError());
        m->launches += 2;
        CGS_CUDA(cudaMemcpyAsync(h.data(), d_top, sizeof(int32_t) * h.size(), cudaMemcpyDeviceToHost, m->stream));
        CGS_CUDA(cudaStreamSynchronize(m->stream));
        return 0;
    }();
    dfree(d_cand, m->stream);
    dfree(d_top, m->stream);
    CGS_TRY(rc);
    std::copy(h.begin(), h.begin() + (size_t)K * top_n, top_regions);
    if (top_counts) std::copy(h.begin() + (size_t)K * top_n, h.end(), top_counts);
    return 0;
}

int cgs_model_codocument_frequencies(cgs_model *m, int32_t top_n, const int32_t *top_regions, int32_t *codf) {
    if (!m || !top_regions || !codf) CGS_FAIL("NULL argument");
    cgs_corpus *c = m->corpus;
    if (top_n < 1 || top_n > 32) CGS_FAIL("top_n must be in [1, 32]");
    DeviceGuard g(c->device);
    const int32_t K = m->K, D = c->n_cells, W = c->n_regions;
    const size_t n_ent = (size_t)K * top_n;
    // lookup tables: bitmap over regions, sorted distinct listed regions, their (topic << 5 | rank) entries
    std::vector<std::pair<int32_t, int32_t>> pairs(n_ent);
    for (int t = 0; t < K; ++t)
        for (int r = 0; r < top_n; ++r) {
            const int32_t w = top_regions[(size_t)t * top_n + r];
            if (w < 0 || w >= W) CGS_FAIL("top_regions holds an index outside the region axis");
            pairs[(size_t)t * top_n + r] = {w, (t << 5) | r};
        }
    std::sort(pairs.begin(), pairs.end());
    const size_t words = ((size_t)W + 31) / 32;
    // one upload: [bitmap words][u n_ent][ent_ptr n_ent + 1][ent n_ent]
    std::vector<int32_t> tab(words + 3 * n_ent + 1, 0);
    int32_t *bitmap = tab.data(), *u = bitmap + words, *ent_ptr = u + n_ent, *ent = ent_ptr + n_ent + 1;
    int32_t M = 0;
    for (size_t i = 0; i < n_ent; ++i) {
        const int32_t w = pairs[i].first;
        if (i == 0 || w != pairs[i - 1].first) { u[M] = w; ent_ptr[M] = (int32_t)i; ++M; }
        ent[i] = pairs[i].second;
        ((uint32_t *)bitmap)[w >> 5] |= 1u << (w & 31);
    }
    ent_ptr[M] = (int32_t)n_ent;
    int32_t *d_tab = nullptr, *d_codf = nullptr;
    uint32_t *d_masks = nullptr;
    int rc = [&]() -> int {
        CGS_TRY(dmalloc(&d_tab, tab.size(), m->stream));
        CGS_TRY(dmalloc(&d_codf, (size_t)K * top_n * top_n, m->stream));
        CGS_TRY(dmalloc(&d_masks, (size_t)std::max(1, D) * K, m->stream));
        CGS_CUDA(cudaMemcpyAsync(d_tab, tab.data(), sizeof(int32_t) * tab.size(), cudaMemcpyHostToDevice, m->stream));
        CGS_CUDA(cudaMemsetAsync(d_codf, 0, sizeof(int32_t) * (size_t)K * top_n * top_n, m->stream));
        if (D > 0) {
            const int mb = (int)std::max<int64_t>(1, std::min<int64_t>(c->num_sms * 8, ceil_div(D, kMaskWarps)));
            const size_t smem = sizeof(uint32_t) * kMaskWarps * (size_t)K;
            cell_masks_kernel<<<mb, kMaskWarps * 32, smem, m->stream>>>(
                c->d_cell_ptr, c->d_tok_region, D, K, (const uint32_t *)d_tab, d_tab + words, M,
                d_tab + words + n_ent, d_tab + words + 2 * n_ent + 1, d_masks);
            const int n_pairs = top_n * (top_n + 1) / 2;
            const int threads = std::max(256, (n_pairs + 31) / 32 * 32);
            const int slices = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(D, 4096), std::max(1, c->num_sms * 4 / K)));
            pair_count_kernel<<<dim3(K, slices), threads, 0, m->stream>>>(d_masks, D, K, top_n, d_codf);
            CGS_CUDA(cudaGetLastError());
            m->launches += 2;
        }
        CGS_CUDA(cudaMemcpyAsync(codf, d_codf, sizeof(int32_t) * (size_t)K * top_n * top_n, cudaMemcpyDeviceToHost, m->stream));
        CGS_CUDA(cudaStreamSynchronize(m->stream));
        return 0;
    }();
    dfree(d_tab, m->stream);
    dfree(d_codf, m->stream);
    dfree(d_masks, m->stream);
    return rc;
}

int cgs_model_loglik_compat(cgs_model *m, double alpha_raw, double eta_raw, double *ll) {
    if (!m || !ll) CGS_FAIL("NULL argument");
    cgs_corpus *c = m->corpus;
    DeviceGuard g(c->device);
    const int32_t K = m->K, Kp = m->Kp, D = c->n_cells;
    const int64_t W = c->n_regions;
    std::vector<int32_t> nk(Kp), row(Kp);
    int *d_idx = nullptr;  // [0] last region of the last non-empty topic, [1] last non-empty cell
    CGS_TRY(dmalloc(&d_idx, 2, m->stream));
    int h_idx[2] = {-1, -1};
    int32_t last_topic = -1, topic_count = 0;
    int64_t cell_ptr2[2] = {0, 0};
    int rc = [&]() -> int {
        CGS_CUDA(cudaMemsetAsync(d_idx, 0xff, sizeof(int) * 2, m->stream));
        CGS_CUDA(cgs_memcpy_sync(m->stream, nk.data(), m->d_nk, sizeof(int32_t) * (size_t)Kp, cudaMemcpyDeviceToHost));
        for (int k = K - 1; k >= 0; --k)
            if (nk[k] > 0) { last_topic = k; break; }
        const int blocks = c->num_sms * 4;
        if (last_topic >= 0) {
            last_positive_row_kernel<<<blocks, 256, 0, m->stream>>>(m->d_nwk, W, Kp, last_topic, d_idx);
            m->launches += 1;
        }
        if (D > 0) {
            last_nonempty_cell_kernel<<<blocks, 256, 0, m->stream>>>(c->d_cell_ptr, D, d_idx + 1);
            m->launches += 1;
        }
        CGS_CUDA(cudaGetLastError());
        CGS_CUDA(cgs_memcpy_sync(m->stream, h_idx, d_idx, sizeof(int) * 2, cudaMemcpyDeviceToHost));
        if (h_idx[0] >= 0)
            CGS_CUDA(cgs_memcpy_sync(m->stream, &topic_count, m->d_nwk + (size_t)h_idx[0] * Kp + last_topic,
                                     sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (h_idx[1] >= 0) {
            CGS_CUDA(cgs_memcpy_sync(m->stream, row.data(), m->d_ndk + (size_t)h_idx[1] * Kp, sizeof(int32_t) * (size_t)Kp,
                                     cudaMemcpyDeviceToHost));
            CGS_CUDA(cgs_memcpy_sync(m->stream, cell_ptr2, c->d_cell_ptr + h_idx[1], sizeof(int64_t) * 2,
                                     cudaMemcpyDeviceToHost));
        }
        return 0;
    }();
    dfree(d_idx, m->stream);
    CGS_TRY(rc);
    // utils.loglikelihood (utils.py:129-160) overwrites topic_ll / doc_ll inside its loops: what is left is the
    // lgamma of the LAST positive entry of the last non-empty row minus the row-total terms from that row on.
    const double etaW = eta_raw * (double)W, alphaK = alpha_raw * (double)K;
    double topic_ll = 0.0;
    int start = 0;
    if (last_topic >= 0) { topic_ll = std::lgamma((double)topic_count + eta_raw); start = last_topic; }
    for (int k = start; k < K; ++k) topic_ll -= std::lgamma(etaW + (double)nk[k]);
    double doc_ll = 0.0;
    int64_t dstart = 0;
    if (h_idx[1] >= 0) {
        int32_t last_k = -1;
        for (int k = K - 1; k >= 0; --k)
            if (row[k] > 0) { last_k = k; break; }
        if (last_k < 0) CGS_FAIL("n_dk row of a non-empty cell holds no count (model not initialised?)");
        doc_ll = std::lgamma((double)row[last_k] + alpha_raw);
        dstart = h_idx[1];
        doc_ll -= std::lgamma(alphaK + (double)(cell_ptr2[1] - cell_ptr2[0]));
        ++dstart;
    }
    const double lg_empty = std::lgamma(alphaK);
    for (int64_t d = dstart; d < D; ++d) doc_ll -= lg_empty;
    const double const_prior = ((double)K * std::lgamma(alpha_raw) - std::lgamma(alphaK)) * (double)D;
    const double const_ll = ((double)W * std::lgamma(eta_raw) - std::lgamma(etaW)) * (double)K;
    *ll = doc_ll - const_prior + topic_ll - const_ll;
    return 0;
}

}  // extern "C"

// ---- differential features: rank-sum test and log2 fold change per region (diff_features.py:839-1120) ----------
namespace cgs {
template <typename T>
static int ranksum_launch_t(const RankSide &fg, const RankSide &bg, int64_t n_rows, double *d_stat, double *d_pval,
                            double *d_l2fc, cudaStream_t stream, int num_sms) {
    using K = typename RankKey<T>::type;
    // the smaller group is sorted in shared memory, the larger one searches it
    const bool sort_fg = fg.n <= bg.n;
    const RankSide &sorted = sort_fg ? fg : bg, &other = sort_fg ? bg : fg;
    const int32_t cap = (int32_t)(128 * 1024 / sizeof(K));  // values per shared-memory piece (plus 1/32 of padding)
    int32_t piece = 2;
    while (piece < sorted.n && piece < cap) piece <<= 1;
    size_t smem = ((size_t)piece + (size_t)piece / 32 + 1) * sizeof(K);
    // integer rows of modest range are counted instead of searched (worth its fixed cost from ~2000 values on)
    int32_t hist_bins = 0;
    if (std::is_same<T, int32_t>::value && (int64_t)fg.n + bg.n >= 2048 && !getenv("CGS_RANKSUM_NO_HIST")) {
        hist_bins = kRankHistBins;
        smem = std::max(smem, (size_t)2 * hist_bins * sizeof(uint32_t));
    }
    auto kernel = ranksum_rows_kernel<T>;
    CGS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(136 * 1024)));
    int per_sm = 1;
    CGS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kRankThreads, smem));
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(n_rows, (int64_t)num_sms * std::max(1, per_sm)));
    kernel<<<blocks, kRankThreads, smem, stream>>>(sorted, other, sort_fg ? 1 : 0, n_rows, piece, hist_bins, d_stat, d_pval, d_l2fc);
    CGS_CUDA(cudaGetLastError());
    return 0;
}

static int ranksum_launch(int dtype, const RankSide &fg, const RankSide &bg, int64_t n_rows, double *d_stat,
                          double *d_pval, double *d_l2fc, cudaStream_t stream, int num_sms) {
    switch (dtype) {
        case 0: return ranksum_launch_t<int32_t>(fg, bg, n_rows, d_stat, d_pval, d_l2fc, stream, num_sms);
        case 1: return ranksum_launch_t<float>(fg, bg, n_rows, d_stat, d_pval, d_l2fc, stream, num_sms);
        case 2: return ranksum_launch_t<double>(fg, bg, n_rows, d_stat, d_pval, d_l2fc, stream, num_sms);
    }
    CGS_FAIL("dtype must be 0 (int32), 1 (float32) or 2 (float64)");
}
}  // namespace cgs

extern "C" {

int cgs_ranksum_rows_device(int device, int dtype, int64_t n_rows,
                            const void *d_fg, int64_t ld_fg, const int32_t *d_fg_cols, int32_t n_fg,
                            const void *d_bg, int64_t ld_bg, const int32_t *d_bg_cols, int32_t n_bg,
                            double *d_statistic, double *d_pvalue, double *d_log2fc, void *stream) {
    if (!d_fg || !d_bg || !d_statistic || !d_pvalue || !d_log2fc) CGS_FAIL("NULL argument");
    if (n_fg < 1 || n_bg < 1) CGS_FAIL("foreground and background need at least one cell each");
    if (n_rows <= 0) return 0;
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    int sms = kNumSMsFallback;
    CGS_TRY(query_sms(device, &sms));
    RankSide fg{d_fg, ld_fg, d_fg_cols, n_fg}, bg{d_bg, ld_bg, d_bg_cols, n_bg};
    return ranksum_launch(dtype, fg, bg, n_rows, d_statistic, d_pvalue, d_log2fc, (cudaStream_t)stream, sms);
}

int cgs_ranksum_rows(int device, int dtype, int64_t n_rows,
                     const void *fg, int64_t ld_fg, const int32_t *fg_cols, int32_t n_fg,
                     const void *bg, int64_t ld_bg, const int32_t *bg_cols, int32_t n_bg,
                     double *statistic, double *pvalue, double *log2fc) {
    if (!fg || !bg || !statistic || !pvalue || !log2fc) CGS_FAIL("NULL argument");
    if (dtype < 0 || dtype > 2) CGS_FAIL("dtype must be 0 (int32), 1 (float32) or 2 (float64)");
    if (n_fg < 1 || n_bg < 1) CGS_FAIL("foreground and background need at least one cell each");
    if (ld_fg < 1 || ld_bg < 1) CGS_FAIL("row strides must be positive");
    if (n_rows <= 0) return 0;
    if ((!fg_cols && n_fg > ld_fg) || (!bg_cols && n_bg > ld_bg)) CGS_FAIL("group is wider than its matrix");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    CGS_TRY(configure_pool(device));
    const size_t elem = dtype == 2 ? 8 : 4;
    const bool shared_matrix = fg == bg && ld_fg == ld_bg;
    // rows per round: about 1 GiB of matrix on the device at a time
    const int64_t per_row = (int64_t)elem * (ld_fg + (shared_matrix ? 0 : ld_bg));
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n_rows, ((int64_t)1 << 30) / std::max<int64_t>(1, per_row)));
    cudaStream_t stream = nullptr;
    CGS_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    uint8_t *d_a = nullptr, *d_b = nullptr;
    int32_t *d_fc = nullptr, *d_bc = nullptr;
    double *d_out = nullptr;
    HostSource source;
    int rc = [&]() -> int {
        CGS_TRY(source.open((size_t)n_rows * ld_fg * elem));
        CGS_TRY(dmalloc(&d_a, (size_t)chunk * ld_fg * elem, stream));
        if (!shared_matrix) CGS_TRY(dmalloc(&d_b, (size_t)chunk * ld_bg * elem, stream));
        CGS_TRY(dmalloc(&d_out, (size_t)chunk * 3, stream));
        if (fg_cols) {
            for (int32_t i = 0; i < n_fg; ++i)
                if (fg_cols[i] < 0 || fg_cols[i] >= ld_fg) CGS_FAIL("foreground column index out of bounds");
            CGS_TRY(dmalloc(&d_fc, (size_t)n_fg, stream));
            CGS_CUDA(cudaMemcpyAsync(d_fc, fg_cols, sizeof(int32_t) * (size_t)n_fg, cudaMemcpyHostToDevice, stream));
        }
        if (bg_cols) {
            for (int32_t i = 0; i < n_bg; ++i)
                if (bg_cols[i] < 0 || bg_cols[i] >= ld_bg) CGS_FAIL("background column index out of bounds");
            CGS_TRY(dmalloc(&d_bc, (size_t)n_bg, stream));
            CGS_CUDA(cudaMemcpyAsync(d_bc, bg_cols, sizeof(int32_t) * (size_t)n_bg, cudaMemcpyHostToDevice, stream));
        }
        for (int64_t r0 = 0; r0 < n_rows; r0 += chunk) {
            const int64_t nr = std::min(chunk, n_rows - r0);
            CGS_TRY(source.copy(d_a, (const uint8_t *)fg + (size_t)r0 * ld_fg * elem, (size_t)nr * ld_fg * elem, stream));
            if (!shared_matrix)
                CGS_TRY(source.copy(d_b, (const uint8_t *)bg + (size_t)r0 * ld_bg * elem, (size_t)nr * ld_bg * elem, stream));
            CGS_TRY(cgs_ranksum_rows_device(device, dtype, nr, d_a, ld_fg, d_fc, n_fg, shared_matrix ? d_a : d_b, ld_bg, d_bc,
                                            n_bg, d_out, d_out + chunk, d_out + 2 * chunk, stream));
            CGS_CUDA(cudaMemcpyAsync(statistic + r0, d_out, sizeof(double) * (size_t)nr, cudaMemcpyDeviceToHost, stream));
            CGS_CUDA(cudaMemcpyAsync(pvalue + r0, d_out + chunk, sizeof(double) * (size_t)nr, cudaMemcpyDeviceToHost, stream));
            CGS_CUDA(cudaMemcpyAsync(log2fc + r0, d_out + 2 * chunk, sizeof(double) * (size_t)nr, cudaMemcpyDeviceToHost, stream));
            CGS_CUDA(cudaStreamSynchronize(stream));
        }
        return 0;
    }();
    cudaStreamSynchronize(stream);
    source.close();
    dfree(d_a, stream); dfree(d_b, stream); dfree(d_out, stream); dfree(d_fc, stream); dfree(d_bc, stream);
    cudaStreamDestroy(stream);
    return rc;
}

}  // extern "C"

// ---- fused consumers of the imputed matrix: the product is recomputed chunk by chunk on the
// device (a 20 000 x 200 000 chunk costs 4 ms) and consumed there, so the regions x cells matrix -- 800 GB at atlas
// scale -- never exists, neither in HBM nor on the host.
namespace cgs {
struct ImputeStream {
    cgs_impute_plan *plan = nullptr;
    cudaStream_t stream = nullptr;
    float *d_a = nullptr;      // R x K topic_region
    uint8_t *d_chunk = nullptr;  // chunk_rows x C, 4 bytes per element
    uint8_t *d_flags = nullptr;  // R
    int32_t R = 0, K = 0, C = 0, chunk = 0;
    float scale = 1.0f;
    int as_int32 = 0;

    int open(int device, const float *topic_region, const float *cell_topic, int32_t R_, int32_t K_, int32_t C_,
             float scale_, int as_int32_, int32_t chunk_rows) {
        R = R_; K = K_; C = C_; scale = scale_; as_int32 = as_int32_;
        chunk = std::max<int32_t>(1, std::min(chunk_rows > 0 ? chunk_rows : 20000, R));
        CGS_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CGS_TRY(cgs_impute_plan_create(&plan, device, cell_topic, 0, K, C, stream));
        CGS_TRY(dmalloc(&d_a, (size_t)R * K, stream));
        CGS_TRY(dmalloc(&d_chunk, (size_t)chunk * C * 4, stream));
        CGS_TRY(dmalloc(&d_flags, (size_t)R, stream));
        CGS_CUDA(cudaMemcpyAsync(d_a, topic_region, sizeof(float) * (size_t)R * K, cudaMemcpyHostToDevice, stream));
        return 0;
    }
    // imputes rows [r0, r0 + n) into d_chunk (and their non-zero flags into d_flags + r0)
    int run(int32_t r0, int32_t n) {
        return cgs_impute_plan_chunk_device(plan, d_a + (size_t)r0 * K, n, scale, as_int32, d_chunk, d_flags + r0, stream);
    }
    void close() {
        dfree(d_a, stream); dfree(d_chunk, stream); dfree(d_flags, stream);
        if (plan) cgs_impute_plan_destroy(plan);
        if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
    }
};
}  // namespace cgs

extern "C" {

}  // extern "C"

// totals_in != NULL: the per-cell totals of the WHOLE matrix (this call handles a shard of its regions); the first
// pass then only recomputes the chunks for the row flags
static int impute_normalized_impl(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                                  int32_t n_topics, int32_t n_cells, float impute_scale, int as_int32, double norm_scale,
                                  int32_t chunk_rows, void *out, uint8_t *row_nonzero, void *column_totals,
                                  const void *totals_in, int64_t *n_kept) {
    if (!topic_region || !cell_topic || !out || !n_kept) CGS_FAIL("NULL argument");
    if (n_regions <= 0 || n_topics <= 0 || n_cells <= 0) CGS_FAIL("empty operand");
    if (!(norm_scale != 0.0)) CGS_FAIL("scale_factor must not be zero");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    CGS_TRY(configure_pool(device));
    ImputeStream is;
    uint8_t *d_norm[2] = {nullptr, nullptr}, *d_tot = nullptr;  // two normalised chunks: one computes while one leaves
    HostSink sink;
    const size_t out_elem = as_int32 ? 8 : 4;
    const int64_t C = n_cells;
    std::vector<uint8_t> flags((size_t)n_regions);
    int rc = [&]() -> int {
        CGS_TRY(is.open(device, topic_region, cell_topic, n_regions, n_topics, n_cells, impute_scale, as_int32, chunk_rows));
        CGS_TRY(sink.open(device, (size_t)n_regions * C * out_elem));
        CGS_TRY(dmalloc(&d_norm[0], (size_t)is.chunk * C * out_elem, is.stream));
        CGS_TRY(dmalloc(&d_norm[1], (size_t)is.chunk * C * out_elem, is.stream));
        CGS_TRY(dmalloc(&d_tot, (size_t)C * 8, is.stream));
        CGS_CUDA(cudaMemsetAsync(d_tot, 0, (size_t)C * 8, is.stream));
        // pass 1: per-cell totals over every region (all-zero regions add nothing) and the row flags
        for (int32_t r0 = 0; r0 < n_regions; r0 += is.chunk) {
            const int32_t n = std::min(is.chunk, n_regions - r0);
            CGS_TRY(is.run(r0, n));
            if (totals_in) continue;
            dim3 grid((unsigned)ceil_div(C, kNormCols), (unsigned)ceil_div(n, kNormSlab));
            if (as_int32) colsum_kernel<int32_t><<<grid, kNormCols, 0, is.stream>>>((const int32_t *)is.d_chunk, n, C, (long long *)d_tot);
            else colsum_kernel<float><<<grid, kNormCols, 0, is.stream>>>((const float *)is.d_chunk, n, C, (double *)d_tot);
            CGS_CUDA(cudaGetLastError());
        }
        if (totals_in) CGS_CUDA(cudaMemcpyAsync(d_tot, totals_in, (size_t)C * 8, cudaMemcpyHostToDevice, is.stream));
        CGS_CUDA(cudaMemcpyAsync(flags.data(), is.d_flags, (size_t)n_regions, cudaMemcpyDeviceToHost, is.stream));
        if (column_totals) CGS_CUDA(cudaMemcpyAsync(column_totals, d_tot, (size_t)C * 8, cudaMemcpyDeviceToHost, is.stream));
        CGS_CUDA(cudaStreamSynchronize(is.stream));
        // pass 2: recompute, normalise, and hand back only the regions that have a non-zero entry
        int64_t kept = 0;
        int which = 0;
        for (int32_t r0 = 0; r0 < n_regions; r0 += is.chunk, which ^= 1) {
            const int32_t n = std::min(is.chunk, n_regions - r0);
            CGS_TRY(is.run(r0, n));
            CGS_TRY(sink.wait_sources(which, is.stream));  // the chunk before the previous one has left d_norm[which]
            dim3 grid((unsigned)ceil_div(C, kNormCols), (unsigned)ceil_div(n, kNormSlab));
            if (as_int32)
                normalize_log1p_kernel<int32_t, double><<<grid, kNormCols, 0, is.stream>>>(
                    (const int32_t *)is.d_chunk, n, C, (const long long *)d_tot, norm_scale, (double *)d_norm[which]);
            else
                normalize_log1p_kernel<float, float><<<grid, kNormCols, 0, is.stream>>>(
                    (const float *)is.d_chunk, n, C, (const double *)d_tot, norm_scale, (float *)d_norm[which]);
            CGS_CUDA(cudaGetLastError());
            for (int32_t i = 0; i < n;) {  // runs of consecutive kept rows
                if (!flags[(size_t)r0 + i]) { ++i; continue; }
                int32_t j = i;
                while (j < n && flags[(size_t)r0 + j]) ++j;
                CGS_TRY(sink.copy((uint8_t *)out + (size_t)kept * C * out_elem, d_norm[which] + (size_t)i * C * out_elem,
                                  (size_t)(j - i) * C * out_elem, is.stream));
                kept += j - i;
                i = j;
            }
            CGS_TRY(sink.mark_sources(which));
        }
        sink.finish();
        CGS_CUDA(cudaStreamSynchronize(is.stream));
        *n_kept = kept;
        if (row_nonzero) memcpy(row_nonzero, flags.data(), (size_t)n_regions);
        return 0;
    }();
    sink.close();
    dfree(d_norm[0], is.stream); dfree(d_norm[1], is.stream); dfree(d_tot, is.stream);
    is.close();
    return rc;
}

extern "C" {

int cgs_impute_normalized(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                          int32_t n_topics, int32_t n_cells, float impute_scale, int as_int32, double norm_scale,
                          int32_t chunk_rows, void *out, uint8_t *row_nonzero, void *column_totals, int64_t *n_kept) {
    return impute_normalized_impl(device, topic_region, cell_topic, n_regions, n_topics, n_cells, impute_scale, as_int32,
                                  norm_scale, chunk_rows, out, row_nonzero, column_totals, nullptr, n_kept);
}

int cgs_impute_normalized_shard(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                                int32_t n_topics, int32_t n_cells, float impute_scale, int as_int32, double norm_scale,
                                int32_t chunk_rows, const void *column_totals, void *out, uint8_t *row_nonzero,
                                int64_t *n_kept) {
    if (!column_totals) CGS_FAIL("NULL argument");
    return impute_normalized_impl(device, topic_region, cell_topic, n_regions, n_topics, n_cells, impute_scale, as_int32,
                                  norm_scale, chunk_rows, out, row_nonzero, nullptr, column_totals, n_kept);
}

int cgs_impute_rows(int device, const float *topic_region, const float *cell_topic, int32_t n_regions, int32_t n_topics,
                    int32_t n_cells, float impute_scale, int as_int32, int32_t chunk_rows, const uint8_t *keep, void *out,
                    int64_t *n_written) {
    if (!topic_region || !cell_topic || !out) CGS_FAIL("NULL argument");
    if (n_regions <= 0 || n_topics <= 0 || n_cells <= 0) CGS_FAIL("empty operand");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    CGS_TRY(configure_pool(device));
    ImputeStream is;
    HostSink sink;
    uint8_t *d_chunk2 = nullptr;  // second product buffer: chunk i + 1 is computed while chunk i travels to the host
    const size_t row_bytes = (size_t)n_cells * 4;
    int64_t written = 0;
    int rc = [&]() -> int {
        CGS_TRY(is.open(device, topic_region, cell_topic, n_regions, n_topics, n_cells, impute_scale, as_int32, chunk_rows));
        CGS_TRY(sink.open(device, (size_t)n_regions * row_bytes));
        CGS_TRY(dmalloc(&d_chunk2, (size_t)is.chunk * row_bytes, is.stream));
        uint8_t *bufs[2] = {is.d_chunk, d_chunk2};
        int which = 0;
        for (int32_t r0 = 0; r0 < n_regions; r0 += is.chunk, which ^= 1) {
            const int32_t n = std::min(is.chunk, n_regions - r0);
            CGS_TRY(sink.wait_sources(which, is.stream));  // the chunk before the previous one has left this buffer
            CGS_TRY(cgs_impute_plan_chunk_device(is.plan, is.d_a + (size_t)r0 * n_topics, n, impute_scale, as_int32,
                                                 bufs[which], is.d_flags + r0, is.stream));
            for (int32_t i = 0; i < n;) {  // runs of consecutive rows to keep go straight to their final place
                if (keep && !keep[(size_t)r0 + i]) { ++i; continue; }
                int32_t j = i;
                while (j < n && (!keep || keep[(size_t)r0 + j])) ++j;
                CGS_TRY(sink.copy((uint8_t *)out + (size_t)written * row_bytes, bufs[which] + (size_t)i * row_bytes,
                                  (size_t)(j - i) * row_bytes, is.stream));
                written += j - i;
                i = j;
            }
            CGS_TRY(sink.mark_sources(which));
        }
        sink.finish();
        CGS_CUDA(cudaStreamSynchronize(is.stream));
        return 0;
    }();
    sink.close();
    dfree(d_chunk2, is.stream);
    is.close();
    if (n_written) *n_written = written;
    return rc;
}

int cgs_impute_column_totals(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                             int32_t n_topics, int32_t n_cells, float impute_scale, int as_int32, int32_t chunk_rows,
                             void *column_totals, uint8_t *row_nonzero) {
    if (!topic_region || !cell_topic || !column_totals) CGS_FAIL("NULL argument");
    if (n_regions <= 0 || n_topics <= 0 || n_cells <= 0) CGS_FAIL("empty operand");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    CGS_TRY(configure_pool(device));
    ImputeStream is;
    uint8_t *d_tot = nullptr;
    const int64_t C = n_cells;
    int rc = [&]() -> int {
        CGS_TRY(is.open(device, topic_region, cell_topic, n_regions, n_topics, n_cells, impute_scale, as_int32, chunk_rows));
        CGS_TRY(dmalloc(&d_tot, (size_t)C * 8, is.stream));
        CGS_CUDA(cudaMemsetAsync(d_tot, 0, (size_t)C * 8, is.stream));
        for (int32_t r0 = 0; r0 < n_regions; r0 += is.chunk) {
            const int32_t n = std::min(is.chunk, n_regions - r0);
            CGS_TRY(is.run(r0, n));
            dim3 grid((unsigned)ceil_div(C, kNormCols), (unsigned)ceil_div(n, kNormSlab));
            if (as_int32) colsum_kernel<int32_t><<<grid, kNormCols, 0, is.stream>>>((const int32_t *)is.d_chunk, n, C, (long long *)d_tot);
            else colsum_kernel<float><<<grid, kNormCols, 0, is.stream>>>((const float *)is.d_chunk, n, C, (double *)d_tot);
            CGS_CUDA(cudaGetLastError());
        }
        CGS_CUDA(cudaMemcpyAsync(column_totals, d_tot, (size_t)C * 8, cudaMemcpyDeviceToHost, is.stream));
        if (row_nonzero) CGS_CUDA(cudaMemcpyAsync(row_nonzero, is.d_flags, (size_t)n_regions, cudaMemcpyDeviceToHost, is.stream));
        CGS_CUDA(cudaStreamSynchronize(is.stream));
        return 0;
    }();
    dfree(d_tot, is.stream);
    is.close();
    return rc;
}

int cgs_impute_ranksum(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                       int32_t n_topics, int32_t n_cells, float impute_scale, int as_int32, int32_t chunk_rows,
                       int32_t n_contrasts, const int32_t *fg_cols, const int64_t *fg_offsets,
                       const int32_t *bg_cols, const int64_t *bg_offsets,
                       double *statistic, double *pvalue, double *log2fc, uint8_t *row_nonzero) {
    if (!topic_region || !cell_topic || !fg_cols || !fg_offsets || !bg_cols || !bg_offsets || !statistic || !pvalue || !log2fc)
        CGS_FAIL("NULL argument");
    if (n_regions <= 0 || n_topics <= 0 || n_cells <= 0) CGS_FAIL("empty operand");
    if (n_contrasts < 1) CGS_FAIL("need at least one contrast");
    for (int32_t k = 0; k < n_contrasts; ++k)
        if (fg_offsets[k + 1] <= fg_offsets[k] || bg_offsets[k + 1] <= bg_offsets[k])
            CGS_FAIL("foreground and background need at least one cell each");
    const int64_t n_fg_all = fg_offsets[n_contrasts], n_bg_all = bg_offsets[n_contrasts];
    for (int64_t i = 0; i < n_fg_all; ++i)
        if (fg_cols[i] < 0 || fg_cols[i] >= n_cells) CGS_FAIL("foreground column index out of bounds");
    for (int64_t i = 0; i < n_bg_all; ++i)
        if (bg_cols[i] < 0 || bg_cols[i] >= n_cells) CGS_FAIL("background column index out of bounds");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    CGS_TRY(configure_pool(device));
    int sms = kNumSMsFallback;
    CGS_TRY(query_sms(device, &sms));
    ImputeStream is;
    int32_t *d_fg = nullptr, *d_bg = nullptr;
    double *d_res = nullptr;  // [3][n_contrasts][chunk]
    int rc = [&]() -> int {
        CGS_TRY(is.open(device, topic_region, cell_topic, n_regions, n_topics, n_cells, impute_scale, as_int32, chunk_rows));
        CGS_TRY(dmalloc(&d_fg, (size_t)n_fg_all, is.stream));
        CGS_TRY(dmalloc(&d_bg, (size_t)n_bg_all, is.stream));
        CGS_TRY(dmalloc(&d_res, (size_t)3 * n_contrasts * is.chunk, is.stream));
        CGS_CUDA(cudaMemcpyAsync(d_fg, fg_cols, sizeof(int32_t) * (size_t)n_fg_all, cudaMemcpyHostToDevice, is.stream));
        CGS_CUDA(cudaMemcpyAsync(d_bg, bg_cols, sizeof(int32_t) * (size_t)n_bg_all, cudaMemcpyHostToDevice, is.stream));
        const int dtype = as_int32 ? 0 : 1;
        for (int32_t r0 = 0; r0 < n_regions; r0 += is.chunk) {
            const int32_t n = std::min(is.chunk, n_regions - r0);
            CGS_TRY(is.run(r0, n));  // one product per chunk, every contrast reads it from HBM / L2
            for (int32_t k = 0; k < n_contrasts; ++k) {
                RankSide fg{is.d_chunk, n_cells, d_fg + fg_offsets[k], (int32_t)(fg_offsets[k + 1] - fg_offsets[k])};
                RankSide bg{is.d_chunk, n_cells, d_bg + bg_offsets[k], (int32_t)(bg_offsets[k + 1] - bg_offsets[k])};
                double *base = d_res + (size_t)k * is.chunk;
                const size_t plane = (size_t)n_contrasts * is.chunk;
                CGS_TRY(ranksum_launch(dtype, fg, bg, n, base, base + plane, base + 2 * plane, is.stream, sms));
            }
            for (int32_t k = 0; k < n_contrasts; ++k) {
                const double *base = d_res + (size_t)k * is.chunk;
                const size_t plane = (size_t)n_contrasts * is.chunk, o = (size_t)k * n_regions + r0;
                CGS_CUDA(cudaMemcpyAsync(statistic + o, base, sizeof(double) * n, cudaMemcpyDeviceToHost, is.stream));
                CGS_CUDA(cudaMemcpyAsync(pvalue + o, base + plane, sizeof(double) * n, cudaMemcpyDeviceToHost, is.stream));
                CGS_CUDA(cudaMemcpyAsync(log2fc + o, base + 2 * plane, sizeof(double) * n, cudaMemcpyDeviceToHost, is.stream));
            }
            CGS_CUDA(cudaStreamSynchronize(is.stream));
        }
        if (row_nonzero) {
            CGS_CUDA(cgs_memcpy_sync(is.stream, row_nonzero, is.d_flags, (size_t)n_regions, cudaMemcpyDeviceToHost));
        }
        return 0;
    }();
    dfree(d_fg, is.stream); dfree(d_bg, is.stream); dfree(d_res, is.stream);
    is.close();
    return rc;
}



int cgs_ranksum_contrasts(int device, int dtype, int64_t n_rows, const void *x, int64_t ld, int32_t n_contrasts,
                          const int32_t *fg_cols, const int64_t *fg_offsets, const int32_t *bg_cols,
                          const int64_t *bg_offsets, double *statistic, double *pvalue, double *log2fc) {
    if (!x || !fg_cols || !fg_offsets || !bg_cols || !bg_offsets || !statistic || !pvalue || !log2fc) CGS_FAIL("NULL argument");
    if (dtype < 0 || dtype > 2) CGS_FAIL("dtype must be 0 (int32), 1 (float32) or 2 (float64)");
    if (n_contrasts < 1) CGS_FAIL("need at least one contrast");
    if (ld < 1) CGS_FAIL("row stride must be positive");
    if (n_rows <= 0) return 0;
    for (int32_t k = 0; k < n_contrasts; ++k)
        if (fg_offsets[k + 1] <= fg_offsets[k] || bg_offsets[k + 1] <= bg_offsets[k])
            CGS_FAIL("foreground and background need at least one cell each");
    const int64_t n_fg_all = fg_offsets[n_contrasts], n_bg_all = bg_offsets[n_contrasts];
    for (int64_t i = 0; i < n_fg_all; ++i)
        if (fg_cols[i] < 0 || fg_cols[i] >= ld) CGS_FAIL("foreground column index out of bounds");
    for (int64_t i = 0; i < n_bg_all; ++i)
        if (bg_cols[i] < 0 || bg_cols[i] >= ld) CGS_FAIL("background column index out of bounds");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    CGS_TRY(configure_pool(device));
    int sms = kNumSMsFallback;
    CGS_TRY(query_sms(device, &sms));
    const size_t elem = dtype == 2 ? 8 : 4;
    // rows per round: about 1 GiB of matrix, two buffers: round i + 1 uploads while round i is tested
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n_rows, ((int64_t)1 << 30) / std::max<int64_t>(1, (int64_t)elem * ld)));
    cudaStream_t stream = nullptr, up = nullptr;
    CGS_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CGS_CUDA(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
    HostSource source;
    uint8_t *d_x[2] = {nullptr, nullptr};
    int32_t *d_fg = nullptr, *d_bg = nullptr;
    double *d_res = nullptr;  // [3][n_contrasts][chunk]
    cudaEvent_t uploaded[2] = {nullptr, nullptr}, tested[2] = {nullptr, nullptr};
    int rc = [&]() -> int {
        CGS_TRY(source.open((size_t)n_rows * ld * elem));
        for (int b = 0; b < 2; ++b) {
            CGS_TRY(dmalloc(&d_x[b], (size_t)chunk * ld * elem, stream));
            CGS_CUDA(cudaEventCreateWithFlags(&uploaded[b], cudaEventDisableTiming));
            CGS_CUDA(cudaEventCreateWithFlags(&tested[b], cudaEventDisableTiming));
        }
        CGS_TRY(dmalloc(&d_fg, (size_t)n_fg_all, stream));
        CGS_TRY(dmalloc(&d_bg, (size_t)n_bg_all, stream));
        CGS_TRY(dmalloc(&d_res, (size_t)3 * n_contrasts * chunk, stream));
        CGS_CUDA(cudaMemcpyAsync(d_fg, fg_cols, sizeof(int32_t) * (size_t)n_fg_all, cudaMemcpyHostToDevice, stream));
        CGS_CUDA(cudaMemcpyAsync(d_bg, bg_cols, sizeof(int32_t) * (size_t)n_bg_all, cudaMemcpyHostToDevice, stream));
        CGS_CUDA(cudaStreamSynchronize(stream));  // the buffers exist before the upload stream touches them
        auto upload = [&](int64_t r0, int b) -> int {
            const int64_t nr = std::min(chunk, n_rows - r0);
            CGS_CUDA(cudaStreamWaitEvent(up, tested[b], 0));  // the round that used this buffer is done with it
            CGS_TRY(source.copy(d_x[b], (const uint8_t *)x + (size_t)r0 * ld * elem, (size_t)nr * ld * elem, up));
            CGS_CUDA(cudaEventRecord(uploaded[b], up));
            return 0;
        };
        CGS_TRY(upload(0, 0));
        int which = 0;
        const size_t plane = (size_t)n_contrasts * chunk;
        for (int64_t r0 = 0; r0 < n_rows; r0 += chunk, which ^= 1) {
            const int64_t nr = std::min(chunk, n_rows - r0);
            CGS_CUDA(cudaStreamWaitEvent(stream, uploaded[which], 0));
            for (int32_t k = 0; k < n_contrasts; ++k) {  // every contrast reads the round while it is in HBM / L2
                RankSide fg{d_x[which], ld, d_fg + fg_offsets[k], (int32_t)(fg_offsets[k + 1] - fg_offsets[k])};
                RankSide bg{d_x[which], ld, d_bg + bg_offsets[k], (int32_t)(bg_offsets[k + 1] - bg_offsets[k])};
                double *base = d_res + (size_t)k * chunk;
                CGS_TRY(ranksum_launch(dtype, fg, bg, nr, base, base + plane, base + 2 * plane, stream, sms));
            }
            CGS_CUDA(cudaEventRecord(tested[which], stream));
            if (r0 + chunk < n_rows) CGS_TRY(upload(r0 + chunk, which ^ 1));  // overlaps the tests just launched
            for (int32_t k = 0; k < n_contrasts; ++k) {
                const double *base = d_res + (size_t)k * chunk;
                const size_t o = (size_t)k * n_rows + r0;
                CGS_CUDA(cudaMemcpyAsync(statistic + o, base, sizeof(double) * nr, cudaMemcpyDeviceToHost, stream));
                CGS_CUDA(cudaMemcpyAsync(pvalue + o, base + plane, sizeof(double) * nr, cudaMemcpyDeviceToHost, stream));
                CGS_CUDA(cudaMemcpyAsync(log2fc + o, base + 2 * plane, sizeof(double) * nr, cudaMemcpyDeviceToHost, stream));
            }
            CGS_CUDA(cudaStreamSynchronize(stream));  // d_res is reused by the next round
        }
        CGS_CUDA(cudaStreamSynchronize(up));
        return 0;
    }();
    cudaStreamSynchronize(up); cudaStreamSynchronize(stream);
    source.close();
    for (int b = 0; b < 2; ++b) {
        dfree(d_x[b], stream);
        if (uploaded[b]) cudaEventDestroy(uploaded[b]);
        if (tested[b]) cudaEventDestroy(tested[b]);
    }
    dfree(d_fg, stream); dfree(d_bg, stream); dfree(d_res, stream);
    cudaStreamSynchronize(stream);
    cudaStreamDestroy(up); cudaStreamDestroy(stream);
    return rc;
}

}  // extern "C"

static int impute_row_stats_impl(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                                 int32_t n_topics, int32_t n_cells, float impute_scale, int as_int32, double norm_scale,
                                 int32_t chunk_rows, const void *totals_in, float *mean, float *var, uint8_t *row_nonzero) {
    if (!topic_region || !cell_topic || !mean || !var) CGS_FAIL("NULL argument");
    if (n_regions <= 0 || n_topics <= 0 || n_cells <= 0) CGS_FAIL("empty operand");
    if (norm_scale < 0.0) CGS_FAIL("scale_factor must not be negative (0 = statistics of the imputed values themselves)");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    CGS_TRY(configure_pool(device));
    int sms = kNumSMsFallback;
    CGS_TRY(query_sms(device, &sms));
    const bool normalise = norm_scale > 0.0;
    ImputeStream is;
    uint8_t *d_norm = nullptr, *d_tot = nullptr;
    float *d_m = nullptr, *d_v = nullptr;
    const size_t out_elem = as_int32 ? 8 : 4;
    const int64_t C = n_cells;
    int rc = [&]() -> int {
        CGS_TRY(is.open(device, topic_region, cell_topic, n_regions, n_topics, n_cells, impute_scale, as_int32, chunk_rows));
        CGS_TRY(dmalloc(&d_m, (size_t)n_regions, is.stream));
        CGS_TRY(dmalloc(&d_v, (size_t)n_regions, is.stream));
        if (normalise) {
            CGS_TRY(dmalloc(&d_norm, (size_t)is.chunk * C * out_elem, is.stream));
            CGS_TRY(dmalloc(&d_tot, (size_t)C * 8, is.stream));
            CGS_CUDA(cudaMemsetAsync(d_tot, 0, (size_t)C * 8, is.stream));
            if (totals_in) CGS_CUDA(cudaMemcpyAsync(d_tot, totals_in, (size_t)C * 8, cudaMemcpyHostToDevice, is.stream));
            for (int32_t r0 = 0; r0 < n_regions && !totals_in; r0 += is.chunk) {  // pass 1: per-cell totals
                const int32_t n = std::min(is.chunk, n_regions - r0);
                CGS_TRY(is.run(r0, n));
                dim3 grid((unsigned)ceil_div(C, kNormCols), (unsigned)ceil_div(n, kNormSlab));
                if (as_int32) colsum_kernel<int32_t><<<grid, kNormCols, 0, is.stream>>>((const int32_t *)is.d_chunk, n, C, (long long *)d_tot);
                else colsum_kernel<float><<<grid, kNormCols, 0, is.stream>>>((const float *)is.d_chunk, n, C, (double *)d_tot);
                CGS_CUDA(cudaGetLastError());
            }
        }
        for (int32_t r0 = 0; r0 < n_regions; r0 += is.chunk) {  // (re)compute, normalise, reduce every region
            const int32_t n = std::min(is.chunk, n_regions - r0);
            CGS_TRY(is.run(r0, n));
            if (normalise) {
                dim3 grid((unsigned)ceil_div(C, kNormCols), (unsigned)ceil_div(n, kNormSlab));
                if (as_int32) {
                    normalize_log1p_kernel<int32_t, double><<<grid, kNormCols, 0, is.stream>>>(
                        (const int32_t *)is.d_chunk, n, C, (const long long *)d_tot, norm_scale, (double *)d_norm);
                    CGS_CUDA(cudaGetLastError());
                    CGS_TRY(row_mean_var_launch<double>((const double *)d_norm, C, n, C, d_m + r0, d_v + r0, sms, is.stream));
                } else {
                    normalize_log1p_kernel<float, float><<<grid, kNormCols, 0, is.stream>>>(
                        (const float *)is.d_chunk, n, C, (const double *)d_tot, norm_scale, (float *)d_norm);
                    CGS_CUDA(cudaGetLastError());
                    CGS_TRY(row_mean_var_launch<float>((const float *)d_norm, C, n, C, d_m + r0, d_v + r0, sms, is.stream));
                }
            } else if (as_int32) {
                CGS_TRY(row_mean_var_launch<int32_t>((const int32_t *)is.d_chunk, C, n, C, d_m + r0, d_v + r0, sms, is.stream));
            } else {
                CGS_TRY(row_mean_var_launch<float>((const float *)is.d_chunk, C, n, C, d_m + r0, d_v + r0, sms, is.stream));
            }
        }
        CGS_CUDA(cudaMemcpyAsync(mean, d_m, sizeof(float) * (size_t)n_regions, cudaMemcpyDeviceToHost, is.stream));
        CGS_CUDA(cudaMemcpyAsync(var, d_v, sizeof(float) * (size_t)n_regions, cudaMemcpyDeviceToHost, is.stream));
        if (row_nonzero) CGS_CUDA(cudaMemcpyAsync(row_nonzero, is.d_flags, (size_t)n_regions, cudaMemcpyDeviceToHost, is.stream));
        CGS_CUDA(cudaStreamSynchronize(is.stream));
        return 0;
    }();
    dfree(d_norm, is.stream); dfree(d_tot, is.stream); dfree(d_m, is.stream); dfree(d_v, is.stream);
    is.close();
    return rc;
}

extern "C" {

int cgs_impute_row_stats(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                         int32_t n_topics, int32_t n_cells, float impute_scale, int as_int32, double norm_scale,
                         int32_t chunk_rows, float *mean, float *var, uint8_t *row_nonzero) {
    return impute_row_stats_impl(device, topic_region, cell_topic, n_regions, n_topics, n_cells, impute_scale, as_int32,
                                 norm_scale, chunk_rows, nullptr, mean, var, row_nonzero);
}

int cgs_impute_row_stats_shard(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                               int32_t n_topics, int32_t n_cells, float impute_scale, int as_int32, double norm_scale,
                               int32_t chunk_rows, const void *column_totals, float *mean, float *var,
                               uint8_t *row_nonzero) {
    if (!column_totals) CGS_FAIL("NULL argument");
    if (!(norm_scale > 0.0)) CGS_FAIL("scale_factor must be positive when per-cell totals are given");
    return impute_row_stats_impl(device, topic_region, cell_topic, n_regions, n_topics, n_cells, impute_scale, as_int32,
                                 norm_scale, chunk_rows, column_totals, mean, var, row_nonzero);
}

int cgs_impute_rankings(int device, const float *topic_region, const float *cell_topic, int32_t n_regions,
                        int32_t n_topics, int32_t n_cells, float impute_scale, int as_int32, uint64_t seed,
                        int32_t cell_block, int64_t first_cell, int32_t *out, int64_t ld_out) {
    if (!topic_region || !cell_topic || !out) CGS_FAIL("NULL argument");
    if (n_regions <= 0 || n_topics <= 0 || n_cells <= 0) CGS_FAIL("empty operand");
    if (ld_out < n_cells) CGS_FAIL("row stride smaller than the number of cells");
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    CGS_TRY(configure_pool(device));
    const int64_t R = n_regions, K = n_topics, C = n_cells;
    // cells per round: the whole regions x block product (4 B per value) and its rankings live on the device
    const int32_t block = (int32_t)std::max<int64_t>(1, std::min<int64_t>(C, cell_block > 0 ? cell_block : std::max<int64_t>(32, (1LL << 29) / R)));
    cudaStream_t stream = nullptr;
    CGS_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    float *d_a = nullptr;
    uint8_t *d_prod = nullptr, *d_flags = nullptr;
    int32_t *d_rank = nullptr;
    cgs_impute_plan *plan = nullptr;
    std::vector<float> cells((size_t)K * block);
    HostSink sink;
    int rc = [&]() -> int {
        CGS_TRY(sink.open(device, (size_t)R * C * 4));
        CGS_TRY(dmalloc(&d_a, (size_t)R * K, stream));
        CGS_TRY(dmalloc(&d_prod, (size_t)R * block * 4, stream));
        CGS_TRY(dmalloc(&d_rank, (size_t)R * block, stream));
        CGS_TRY(dmalloc(&d_flags, (size_t)R, stream));
        CGS_CUDA(cudaMemcpyAsync(d_a, topic_region, sizeof(float) * (size_t)R * K, cudaMemcpyHostToDevice, stream));
        for (int64_t c0 = 0; c0 < C; c0 += block) {
            const int32_t nc = (int32_t)std::min<int64_t>(block, C - c0);
            for (int64_t k = 0; k < K; ++k) memcpy(cells.data() + k * nc, cell_topic + k * C + c0, sizeof(float) * (size_t)nc);
            CGS_TRY(cgs_impute_plan_create(&plan, device, cells.data(), 0, n_topics, nc, stream));
            CGS_TRY(cgs_impute_plan_chunk_device(plan, d_a, n_regions, impute_scale, as_int32, d_prod, d_flags, stream));
            CGS_TRY(cgs_rank_columns_device(device, as_int32 ? 0 : 1, d_prod, nc, R, nc, seed, first_cell + c0, d_rank, nc, stream));
            CGS_TRY(sink.copy2d(out + c0, (size_t)ld_out * 4, d_rank, (size_t)nc * 4, (size_t)R, stream));
            sink.finish();
            CGS_CUDA(cudaStreamSynchronize(stream));
            cgs_impute_plan_destroy(plan);
            plan = nullptr;
        }
        return 0;
    }();
    if (plan) cgs_impute_plan_destroy(plan);
    cudaStreamSynchronize(stream);
    sink.close();
    dfree(d_a, stream); dfree(d_prod, stream); dfree(d_rank, stream); dfree(d_flags, stream);
    cudaStreamSynchronize(stream);
    cudaStreamDestroy(stream);
    return rc;
}

}  // extern "C"

// ---- count matrix from fragments (cistopic_class.py:868-889) ---------------------------------
struct cgs_fragment_counts {
    int device = 0;
    int32_t n_regions = 0, n_cells = 0;
    int64_t nnz = 0, n_hits = 0;
    cudaStream_t stream = nullptr;
    int64_t *d_indptr = nullptr;   // R + 1
    int32_t *d_indices = nullptr;  // nnz, cells ascending inside a region
    int32_t *d_counts = nullptr;   // nnz
};

extern "C" {

int cgs_fragment_counts_destroy(cgs_fragment_counts *h) {
    if (!h) return 0;
    DeviceGuard g(h->device);
    dfree(h->d_indptr, h->stream); dfree(h->d_indices, h->stream); dfree(h->d_counts, h->stream);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    delete h;
    return 0;
}

int cgs_fragment_counts_create(cgs_fragment_counts **out, int device, int32_t n_chrom, const int32_t *region_chrom_ptr,
                               const int32_t *region_start, const int32_t *region_end, int64_t n_fragments,
                               const int32_t *frag_chrom, const int32_t *frag_start, const int32_t *frag_end,
                               const int32_t *frag_cell, int32_t n_cells) {
    if (!out || !region_chrom_ptr || !region_start || !region_end) CGS_FAIL("NULL argument");
    if (n_fragments > 0 && (!frag_chrom || !frag_start || !frag_end || !frag_cell)) CGS_FAIL("NULL argument");
    if (n_chrom < 1 || n_cells < 1 || n_fragments < 0) CGS_FAIL("empty input");
    const int32_t R = region_chrom_ptr[n_chrom];
    if (region_chrom_ptr[0] != 0 || R < 1) CGS_FAIL("region_chrom_ptr must start at 0 and cover at least one region");
    std::vector<int32_t> maxlen((size_t)n_chrom, 0);
    for (int32_t c = 0; c < n_chrom; ++c) {
        if (region_chrom_ptr[c + 1] < region_chrom_ptr[c]) CGS_FAIL("region_chrom_ptr must be non-decreasing");
        for (int32_t j = region_chrom_ptr[c]; j < region_chrom_ptr[c + 1]; ++j) {
            if (j > region_chrom_ptr[c] && region_start[j] < region_start[j - 1])
                CGS_FAIL("regions must be sorted by start inside every chromosome");
            if (region_end[j] < region_start[j]) CGS_FAIL("region with end < start");
            maxlen[c] = std::max(maxlen[c], region_end[j] - region_start[j]);
        }
    }
    DeviceGuard g(device);
    if (!g.ok) CGS_FAIL("cannot select CUDA device");
    CGS_TRY(configure_pool(device));
    int sms = kNumSMsFallback;
    CGS_TRY(query_sms(device, &sms));
    cgs_fragment_counts *h = new cgs_fragment_counts();
    *out = nullptr;
    h->device = device; h->n_regions = R; h->n_cells = n_cells;
    int32_t *d_rs = nullptr, *d_re = nullptr, *d_cp = nullptr, *d_ml = nullptr, *d_f = nullptr, *d_raw = nullptr;
    unsigned long long *d_fill = nullptr, *d_nnz = nullptr;
    int64_t *d_raw_ptr = nullptr;
    int rc = [&]() -> int {
        CGS_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        cudaStream_t st = h->stream;
        const size_t F = (size_t)n_fragments;
        CGS_TRY(dmalloc(&d_rs, (size_t)R, st)); CGS_TRY(dmalloc(&d_re, (size_t)R, st));
        CGS_TRY(dmalloc(&d_cp, (size_t)n_chrom + 1, st)); CGS_TRY(dmalloc(&d_ml, (size_t)n_chrom, st));
        CGS_TRY(dmalloc(&d_f, 4 * F, st));
        CGS_TRY(dmalloc(&d_fill, (size_t)R, st)); CGS_TRY(dmalloc(&d_nnz, (size_t)R, st));
        CGS_TRY(dmalloc(&d_raw_ptr, (size_t)R + 1, st));
        CGS_CUDA(cudaMemcpyAsync(d_rs, region_start, sizeof(int32_t) * (size_t)R, cudaMemcpyHostToDevice, st));
        CGS_CUDA(cudaMemcpyAsync(d_re, region_end, sizeof(int32_t) * (size_t)R, cudaMemcpyHostToDevice, st));
        CGS_CUDA(cudaMemcpyAsync(d_cp, region_chrom_ptr, sizeof(int32_t) * ((size_t)n_chrom + 1), cudaMemcpyHostToDevice, st));
        CGS_CUDA(cudaMemcpyAsync(d_ml, maxlen.data(), sizeof(int32_t) * (size_t)n_chrom, cudaMemcpyHostToDevice, st));
        const int32_t *src[4] = {frag_chrom, frag_start, frag_end, frag_cell};
        for (int k = 0; k < 4 && F > 0; ++k)
            CGS_CUDA(cudaMemcpyAsync(d_f + k * F, src[k], sizeof(int32_t) * F, cudaMemcpyHostToDevice, st));
        CGS_CUDA(cudaMemsetAsync(d_fill, 0, sizeof(unsigned long long) * (size_t)R, st));
        FragRegions rg{d_rs, d_re, d_cp, d_ml, n_chrom};
        const int blocks = sms * 8;
        fragment_hits_kernel<false><<<blocks, 256, 0, st>>>(d_f, d_f + F, d_f + 2 * F, d_f + 3 * F, n_fragments, n_cells, rg,
                                                            d_fill, nullptr, nullptr);
        CGS_CUDA(cudaGetLastError());
        exclusive_scan_i64_kernel<<<1, 1024, 0, st>>>(d_fill, d_raw_ptr, R);
        CGS_CUDA(cudaGetLastError());
        CGS_CUDA(cgs_memcpy_sync(st, &h->n_hits, d_raw_ptr + R, sizeof(int64_t), cudaMemcpyDeviceToHost));
        CGS_TRY(dmalloc(&d_raw, (size_t)h->n_hits, st));
        CGS_CUDA(cudaMemsetAsync(d_fill, 0, sizeof(unsigned long long) * (size_t)R, st));
        fragment_hits_kernel<true><<<blocks, 256, 0, st>>>(d_f, d_f + F, d_f + 2 * F, d_f + 3 * F, n_fragments, n_cells, rg,
                                                           d_fill, d_raw_ptr, d_raw);
        CGS_CUDA(cudaGetLastError());
        // per-region histogram over a shared-memory window of the cell axis
        int max_smem = 0;
        CGS_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        const int32_t window = (int32_t)std::max<int64_t>(1, std::min<int64_t>(n_cells, (max_smem - 1024) / 4));
        const size_t smem = (size_t)window * 4;
        CGS_CUDA(cudaFuncSetAttribute(region_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CGS_CUDA(cudaFuncSetAttribute(region_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per_sm = smem > 100 * 1024 ? 1 : (smem > 48 * 1024 ? 2 : 4);
        const int rblocks = std::max(1, std::min(R, sms * per_sm));
        region_rows_kernel<false><<<rblocks, 512, smem, st>>>(d_raw, d_raw_ptr, R, n_cells, window, d_nnz, nullptr, nullptr, nullptr);
        CGS_CUDA(cudaGetLastError());
        CGS_TRY(dmalloc(&h->d_indptr, (size_t)R + 1, st));
        exclusive_scan_i64_kernel<<<1, 1024, 0, st>>>(d_nnz, h->d_indptr, R);
        CGS_CUDA(cudaGetLastError());
        CGS_CUDA(cgs_memcpy_sync(st, &h->nnz, h->d_indptr + R, sizeof(int64_t), cudaMemcpyDeviceToHost));
        CGS_TRY(dmalloc(&h->d_indices, (size_t)h->nnz, st));
        CGS_TRY(dmalloc(&h->d_counts, (size_t)h->nnz, st));
        region_rows_kernel<true><<<rblocks, 512, smem, st>>>(d_raw, d_raw_ptr, R, n_cells, window, nullptr, h->d_indptr,
                                                            h->d_indices, h->d_counts);
        CGS_CUDA(cudaGetLastError());
        CGS_CUDA(cudaStreamSynchronize(st));
        return 0;
    }();
    dfree(d_rs, h->stream); dfree(d_re, h->stream); dfree(d_cp, h->stream); dfree(d_ml, h->stream); dfree(d_f, h->stream);
    dfree(d_raw, h->stream); dfree(d_fill, h->stream); dfree(d_nnz, h->stream); dfree(d_raw_ptr, h->stream);
    if (rc != 0) { cgs_fragment_counts_destroy(h); return rc; }
    *out = h;
    return 0;
}

int cgs_fragment_counts_dims(const cgs_fragment_counts *h, int32_t *n_regions, int32_t *n_cells, int64_t *nnz,
                             int64_t *n_hits) {
    if (!h) CGS_FAIL("NULL handle");
    if (n_regions) *n_regions = h->n_regions;
    if (n_cells) *n_cells = h->n_cells;
    if (nnz) *nnz = h->nnz;
    if (n_hits) *n_hits = h->n_hits;
    return 0;
}

int cgs_fragment_counts_export(const cgs_fragment_counts *h, int64_t *indptr, int32_t *indices, int32_t *counts) {
    if (!h) CGS_FAIL("NULL handle");
    DeviceGuard g(h->device);
    if (indptr)
        CGS_CUDA(cudaMemcpyAsync(indptr, h->d_indptr, sizeof(int64_t) * ((size_t)h->n_regions + 1), cudaMemcpyDeviceToHost, h->stream));
    if (indices && h->nnz)
        CGS_CUDA(cudaMemcpyAsync(indices, h->d_indices, sizeof(int32_t) * (size_t)h->nnz, cudaMemcpyDeviceToHost, h->stream));
    if (counts && h->nnz)
        CGS_CUDA(cudaMemcpyAsync(counts, h->d_counts, sizeof(int32_t) * (size_t)h->nnz, cudaMemcpyDeviceToHost, h->stream));
    CGS_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

}  // extern "C"
