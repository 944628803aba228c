// exchange_kernels.cuh -- AD-LDA count exchange over peer memory, fused into ONE kernel per round and
// built to run BESIDE the sweep kernel (SURVEY.md section 8e ii; the reference's analogue is Mallet's
// per-iteration merge of per-thread count copies, lda_models.py:572-573), plus the device-side
// consistency checks bench.py and the tests run on the count tables.
//
// Every rank keeps a full replica of n_wk (W x Kp int32) that its own sweep kernel updates with live
// reductions.  `snap` holds, per entry, the value of n_wk at the moment the entry was last captured.
// One exchange ROUND on rank r (all ranks run the same round number at about the same time):
//
//   phase 1  capture   d = n_wk - snap;  delta_r = d;  snap = n_wk          (all entries, local HBM;
//                      n_wk is read while the sampler keeps changing it: every change lands in exactly
//                      one captured delta, this round's or the next one's)
//   barrier A          (flags in peer memory: every rank's delta is complete and visible)
//   phase 2  reduce    for the 1/R slice of entries rank r owns: t = sum_q delta_q (loads over NVLink
//                      from every peer's delta buffer), then total_q = t for every q (stores over
//                      NVLink) -- reduce-scatter + all-gather, 2 (R-1)/R x 4WKp bytes per direction
//   barrier B
//   phase 3  apply     x = total_r - delta_r (what the OTHER ranks changed);  n_wk += x with no-return
//                      global reductions (the sampler is running);  snap += x;  n_k += column sums of x
//
// Invariants (tests/test_gpu_exchange.py): counts stay exact integers; on every rank
// n_wk = own tokens (live) + every other rank's tokens as of its last capture, hence >= 0 entry-wise;
// after a round with no sampler running all replicas are identical and equal the histogram of all z.
//
// Cross-rank barriers are flags in the peers' arenas (st.release.sys / ld.acquire.sys); the kernels of all
// ranks must be co-resident, which holds because each runs on its own GPU (or, in the single-GPU tests,
// because R x n_ctas CTAs fit on the device).
#pragma once
#include "common.cuh"

namespace cgs {

constexpr int kMaxExchangeRanks = 16;
constexpr int kExchangeThreads = 256;  // 256 threads x <= 64 registers: a CTA takes the room of two sweep CTAs
constexpr int kExchangeStampWord = 64;   // 6 x uint64 phase timestamps of the last synchronous exchange (8-byte aligned)
constexpr int kExchangeFlagWords = 256;  // per arena: [2][16] arrival flags, [2] local CTA counters, [1] timeouts

struct ExchangePeers {
    int4 *delta[kMaxExchangeRanks];
    int4 *total[kMaxExchangeRanks];
    uint32_t *flags[kMaxExchangeRanks];
};

struct ExchangeParams {
    int32_t *n_wk;
    int32_t *snap;
    int32_t *n_k;
    ExchangePeers peers;
    int32_t n_ranks, rank;
    int64_t n4;      // W * Kp / 4
    int32_t Kp4;     // Kp / 4
    uint32_t round;  // 1, 2, 3, ... (flags only ever grow)
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// data written by another GPU (or read from another GPU's memory): never served from this SM's L1
__device__ __forceinline__ int4 ld_sys_v4(const int4 *p) {
    int4 v;
    asm volatile("ld.relaxed.sys.global.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_add_global_i32(int32_t *p, int32_t v) {
    asm volatile("red.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

constexpr uint64_t kExchangeTimeoutNs = 4000000000ull;  // 4 s
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// All CTAs of all ranks.  `phase` selects the flag row; the last CTA of the rank to arrive publishes the
// round number in every peer's arena, then every CTA waits until all ranks have published.
__device__ __forceinline__ void exchange_barrier(const ExchangeParams &p, int phase, unsigned n_ctas) {
    __syncthreads();
    uint32_t *mine = p.peers.flags[p.rank];
    if (threadIdx.x == 0) {
        // Every CTA publishes its writes at GPU scope and arrives on the rank's counter; only the LAST CTA pays for the
        // system-scope fence, which by cumulativity covers everything the other CTAs published before they arrived
        // (a system-scope fence in each of the ~600 CTAs of a stand-alone exchange cost ~0.02 ms more per exchange).
        __threadfence();
        uint32_t *counter = mine + 2 * kMaxExchangeRanks + phase;
        const unsigned prev = atomicAdd(counter, 1u);
        if (prev == n_ctas - 1) {
            __threadfence();
            *counter = 0;  // next use is in the next round, after both barriers of this one
            __threadfence_system();  // ONE system-scope fence, then the flags (a release store per peer would fence R times)
            for (int j = 0; j < p.n_ranks; ++j) {
                const int q = (p.rank + 1 + j) % p.n_ranks;
                st_relaxed_sys_u32(p.peers.flags[q] + phase * kMaxExchangeRanks + p.rank, p.round);
            }
        }
    }
    if ((int)threadIdx.x < p.n_ranks) {
        const uint32_t *f = mine + phase * kMaxExchangeRanks + threadIdx.x;
        // rounds are compared modulo 2^32: "reached" = not behind.  A peer that never shows up (a rank that died, or
        // round kernels that cannot be co-resident) must not hang the device: after kExchangeTimeoutNs the wait gives
        // up and counts a timeout, which cgs_model_exchange_status reports (the tables are then meaningless).
        const uint64_t t0 = global_timer_ns();
        while ((int32_t)(ld_acquire_sys_u32(f) - p.round) < 0) {
            __nanosleep(200);
            if (global_timer_ns() - t0 > kExchangeTimeoutNs) {
                atomicAdd(mine + 2 * kMaxExchangeRanks + 2, 1u);
                break;
            }
        }
    }
    __syncthreads();
}

// One round of one rank, executed by n_ctas CTAs of which this is number `cta`.
__device__ __forceinline__ void exchange_round_body(const ExchangeParams &p, unsigned cta, unsigned n_ctas, int *x_colsum) {
    const int Kp = p.Kp4 * 4;
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) x_colsum[k] = 0;
    const int64_t stride = (int64_t)n_ctas * blockDim.x;
    const int64_t tid0 = (int64_t)cta * blockDim.x + threadIdx.x;
    int4 *n_wk = reinterpret_cast<int4 *>(p.n_wk);
    int4 *snap = reinterpret_cast<int4 *>(p.snap);
    int4 *my_delta = p.peers.delta[p.rank];
    const int4 *my_total = p.peers.total[p.rank];

    // ---- phase 1: capture -------------------------------------------------------------------------
    // (four independent entries per thread and iteration: the few CTAs this kernel may take need the loads in flight)
    for (int64_t i0 = tid0; i0 < p.n4; i0 += 4 * stride) {
        int4 a[4], sn[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < p.n4) { a[u] = __ldcg(n_wk + i); sn[u] = __ldcs(snap + i); }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < p.n4) {
                __stcg(my_delta + i, make_int4(a[u].x - sn[u].x, a[u].y - sn[u].y, a[u].z - sn[u].z, a[u].w - sn[u].w));
                __stcs(snap + i, a[u]);
            }
        }
    }
    exchange_barrier(p, 0, n_ctas);

    // ---- phase 2: reduce the slice this rank owns, hand the sums to everybody -----------------------
    {
        const int R = p.n_ranks;
        const int64_t lo = p.n4 * p.rank / R, hi = p.n4 * (p.rank + 1) / R;
        for (int64_t i0 = lo + tid0; i0 < hi; i0 += 4 * stride) {  // four entries in flight per thread, one peer at a time
            int4 t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] = make_int4(0, 0, 0, 0);
            for (int j = 0; j < R; ++j) {
                const int q = (p.rank + 1 + j) % R;  // staggered peers
                int4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t i = i0 + u * stride;
                    v[u] = i < hi ? ld_sys_v4(p.peers.delta[q] + i) : make_int4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) { t[u].x += v[u].x; t[u].y += v[u].y; t[u].z += v[u].z; t[u].w += v[u].w; }
            }
            for (int j = 0; j < R; ++j) {
                const int q = (p.rank + 1 + j) % R;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t i = i0 + u * stride;
                    if (i < hi) __stcg(p.peers.total[q] + i, t[u]);
                }
            }
        }
    }
    exchange_barrier(p, 1, n_ctas);

    // ---- phase 3: apply what the other ranks changed ------------------------------------------------
    for (int64_t i0 = tid0; i0 < p.n4; i0 += 4 * stride) {
        int4 t[4], d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < p.n4) { t[u] = ld_sys_v4(my_total + i); d[u] = __ldcg(my_delta + i); }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (i >= p.n4) continue;
            const int4 x = make_int4(t[u].x - d[u].x, t[u].y - d[u].y, t[u].z - d[u].z, t[u].w - d[u].w);
            if ((x.x | x.y | x.z | x.w) != 0) {
                int32_t *e = p.n_wk + i * 4;
                const int col = (int)(i % p.Kp4) * 4;
                if (x.x) { red_add_global_i32(e + 0, x.x); atomicAdd(&x_colsum[col + 0], x.x); }
                if (x.y) { red_add_global_i32(e + 1, x.y); atomicAdd(&x_colsum[col + 1], x.y); }
                if (x.z) { red_add_global_i32(e + 2, x.z); atomicAdd(&x_colsum[col + 2], x.z); }
                if (x.w) { red_add_global_i32(e + 3, x.w); atomicAdd(&x_colsum[col + 3], x.w); }
                int4 sn = __ldcs(snap + i);
                sn.x += x.x; sn.y += x.y; sn.z += x.z; sn.w += x.w;
                __stcs(snap + i, sn);
            }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < Kp; k += blockDim.x)
        if (x_colsum[k]) red_add_global_i32(p.n_k + k, x_colsum[k]);
}

// ---- synchronous variant: the cheapest exchange when nothing overlaps it ---------------------------------------------
// Between two sweeps (no sampler running) a rank needs no separate snapshot: the receive buffer `total` of the arena
// holds G, the global table as of the last exchange (all zero before the first one), and the round is
//   phase 1   delta = n_wk - G                                         (read 2, write 1 table)
//   barrier
//   phase 2   owner of a slice: G' = G + sum_q delta_q, written into EVERY rank's receive buffer over NVLink
//   barrier
//   phase 3   n_k += column sums of (G' - n_wk);  n_wk = G'             (read 2, write <= 1 table)
// on any contiguous range of ROWS of the table (the pointers of ExchangeParams are offset by the host), i.e. 6 table
// passes of local traffic instead of the 8 of delta_compute + all-reduce + delta_apply, and 2 (R-1)/R
// tables over NVLink.  Statistically identical to the all-reduce scheme (every rank samples a sweep against the
// global counts of the previous exchange plus its own live changes).
__device__ __forceinline__ void exchange_sync_body(const ExchangeParams &p, unsigned cta, unsigned n_ctas, int *x_colsum) {
    const int Kp = p.Kp4 * 4;
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) x_colsum[k] = 0;
    const int64_t stride = (int64_t)n_ctas * blockDim.x;
    const int64_t tid0 = (int64_t)cta * blockDim.x + threadIdx.x;
    int4 *n_wk = reinterpret_cast<int4 *>(p.n_wk);
    int4 *my_delta = p.peers.delta[p.rank];
    int4 *my_g = p.peers.total[p.rank];
    // phase boundaries as seen by CTA 0 (nanoseconds of %globaltimer), read by cgs_model_exchange_phase_times
    unsigned long long *stamps = reinterpret_cast<unsigned long long *>(p.peers.flags[p.rank] + kExchangeStampWord);
    const bool stamper = cta == 0 && threadIdx.x == 0;
    if (stamper) stamps[0] = global_timer_ns();

    for (int64_t i0 = tid0; i0 < p.n4; i0 += 4 * stride) {
        int4 a[4], g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < p.n4) { a[u] = __ldcs(n_wk + i); g[u] = __ldcs(my_g + i); }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < p.n4) __stcg(my_delta + i, make_int4(a[u].x - g[u].x, a[u].y - g[u].y, a[u].z - g[u].z, a[u].w - g[u].w));
        }
    }
    if (stamper) stamps[1] = global_timer_ns();
    exchange_barrier(p, 0, n_ctas);
    if (stamper) stamps[2] = global_timer_ns();
    {
        const int R = p.n_ranks;
        const int64_t lo = p.n4 * p.rank / R, hi = p.n4 * (p.rank + 1) / R;
        // Four entries per thread and iteration, one peer at a time: with a single entry per thread the loop is bound by
        // the NVLink round trip (measured 366 GB/s per direction on 2 ranks, profiles/r2_exchange_phases.txt); four
        // independent 16-byte loads per thread keep ~10 MB in flight per GPU, several times latency x bandwidth.
        for (int64_t i0 = lo + tid0; i0 < hi; i0 += 4 * stride) {
            int4 t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = i0 + u * stride;
                if (i < hi) t[u] = __ldcs(my_g + i);
            }
            for (int j = 0; j < R; ++j) {
                // staggered: rank r starts with r + 1, so the R ranks pull from R different peers at any moment
                const int q = (p.rank + 1 + j) % R;
                int4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t i = i0 + u * stride;
                    if (i < hi) v[u] = ld_sys_v4(p.peers.delta[q] + i);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) { t[u].x += v[u].x; t[u].y += v[u].y; t[u].z += v[u].z; t[u].w += v[u].w; }
            }
            for (int j = 0; j < R; ++j) {
                const int q = (p.rank + 1 + j) % R;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t i = i0 + u * stride;
                    if (i < hi) __stcg(p.peers.total[q] + i, t[u]);
                }
            }
        }
    }
    if (stamper) stamps[3] = global_timer_ns();
    exchange_barrier(p, 1, n_ctas);
    if (stamper) stamps[4] = global_timer_ns();
    // n_k moves by the column sums of what the OTHER ranks changed (G' - n_wk): exact, and valid while a sampler that
    // works on other rows keeps adding its own changes to n_k
    for (int64_t i0 = tid0; i0 < p.n4; i0 += 4 * stride) {
        int4 g[4], o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < p.n4) { g[u] = ld_sys_v4(my_g + i); o[u] = __ldcg(n_wk + i); }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (i >= p.n4) continue;
            const int4 x = make_int4(g[u].x - o[u].x, g[u].y - o[u].y, g[u].z - o[u].z, g[u].w - o[u].w);
            if ((x.x | x.y | x.z | x.w) == 0) continue;
            __stcg(n_wk + i, g[u]);  // stays in L2 for the sweep that follows
            const int col = (int)(i % p.Kp4) * 4;
            if (x.x) atomicAdd(&x_colsum[col + 0], x.x);
            if (x.y) atomicAdd(&x_colsum[col + 1], x.y);
            if (x.z) atomicAdd(&x_colsum[col + 2], x.z);
            if (x.w) atomicAdd(&x_colsum[col + 3], x.w);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < Kp; k += blockDim.x)
        if (x_colsum[k]) red_add_global_i32(p.n_k + k, x_colsum[k]);
    if (stamper) stamps[5] = global_timer_ns();
}

__global__ void __launch_bounds__(kExchangeThreads, 4)
exchange_sync_kernel(const ExchangeParams p) {
    extern __shared__ int x_colsum[];
    exchange_sync_body(p, blockIdx.x, gridDim.x, x_colsum);
}

// Production: one launch per rank and round, each rank on its own GPU.
__global__ void __launch_bounds__(kExchangeThreads, 4)
exchange_round_kernel(const ExchangeParams p) {
    extern __shared__ int x_colsum[];  // Kp ints
    exchange_round_body(p, blockIdx.x, gridDim.x, x_colsum);
}

// Several ranks that share ONE GPU (tests, and a host that packs more shards than it has GPUs): kernels that wait
// for each other must not be separate launches on one device -- nothing guarantees they run at the same time
// (B200_PROFILING.md) -- so all ranks' rounds are ONE cooperative launch: grid = (n_ctas, n_ranks_here), row y
// executes rank y's round; co-residency of the whole grid is what a cooperative launch guarantees.
constexpr int kMaxLocalRanks = 8;
struct ExchangeParamsMulti {
    ExchangeParams r[kMaxLocalRanks];
};
__global__ void __launch_bounds__(kExchangeThreads, 4)
exchange_round_multi_kernel(const ExchangeParamsMulti a, int sync_variant) {
    extern __shared__ int x_colsum[];
    if (sync_variant) exchange_sync_body(a.r[blockIdx.y], blockIdx.x, gridDim.x, x_colsum);
    else exchange_round_body(a.r[blockIdx.y], blockIdx.x, gridDim.x, x_colsum);
}

// ---- consistency checks ---------------------------------------------------------------------------
// One CTA per cell (grid-stride): the cell's topic histogram from z must equal its n_dk row; every token
// adds one to hist_wk[region][topic] (zeroed by the caller).  mismatch[0] += differing n_dk entries.
__global__ void recount_kernel(const int64_t *__restrict__ cell_ptr, const int32_t *__restrict__ tok_region,
                               const uint16_t *__restrict__ z, int32_t n_cells, int32_t K, int32_t Kp,
                               const int32_t *__restrict__ n_dk, int32_t *__restrict__ hist_wk,
                               unsigned long long *__restrict__ mismatch) {
    extern __shared__ int rc_hist[];
    for (int32_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        for (int k = threadIdx.x; k < Kp; k += blockDim.x) rc_hist[k] = 0;
        __syncthreads();
        const int64_t b = cell_ptr[cell], e = cell_ptr[cell + 1];
        for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
            const int k = z[i];
            if (k < K) {
                atomicAdd(&rc_hist[k], 1);
                red_add_global_i32(hist_wk + (int64_t)tok_region[i] * Kp + k, 1);
            } else {
                atomicAdd(mismatch, 1ull);
            }
        }
        __syncthreads();
        int bad = 0;
        for (int k = threadIdx.x; k < Kp; k += blockDim.x)
            if (rc_hist[k] != n_dk[(int64_t)cell * Kp + k]) ++bad;
        if (bad) atomicAdd(mismatch, (unsigned long long)bad);
        __syncthreads();
    }
}

// mismatch[1] += entries where a != b;  colsum[k] += sum over rows of a[., k] (int64, zeroed by the caller)
__global__ void compare_counts_kernel(const int32_t *__restrict__ a, const int32_t *__restrict__ b, int64_t n,
                                      int32_t Kp, unsigned long long *__restrict__ mismatch,
                                      long long *__restrict__ colsum) {
    extern __shared__ long long cc_col[];
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) cc_col[k] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int32_t va = a[i];
        if (va != b[i]) ++bad;
        if (va) atomicAdd((unsigned long long *)&cc_col[i % Kp], (unsigned long long)(long long)va);
    }
    if (bad) atomicAdd(mismatch + 1, bad);
    __syncthreads();
    for (int k = threadIdx.x; k < Kp; k += blockDim.x)
        if (cc_col[k]) atomicAdd((unsigned long long *)&colsum[k], (unsigned long long)cc_col[k]);
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

// out[0] += sum_i mix64(i) * v[i]  (mod 2^64: a position-dependent hash that can be reduced in any order),
// out[1] += sum_i v[i]
__global__ void state_hash_kernel(const int32_t *__restrict__ v, int64_t n, unsigned long long *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long h = 0, s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long x = v[i];
        h += mix64((uint64_t)i + 0x9E3779B97F4A7C15ull) * (unsigned long long)x;
        s += (unsigned long long)x;
    }
    for (int off = 16; off > 0; off >>= 1) {
        h += __shfl_down_sync(0xffffffffu, h, off);
        s += __shfl_down_sync(0xffffffffu, s, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out + 0, h);
        atomicAdd(out + 1, s);
    }
}

}  // namespace cgs
