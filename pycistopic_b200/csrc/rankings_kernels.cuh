// rankings_kernels.cuh -- per-cell rankings of the imputed matrix on the device:
// CistopicImputedFeatures.make_rankings of the reference (src/pycisTopic/diff_features.py:274-337).
//
// The reference ranks every column (cell) of the regions x cells matrix by DESCENDING score, 0 = best, and breaks
// ties at random:  perm = rng.permutation(R);  ranking = perm[(-scores)[perm].argsort()].argsort()  (:294-306).
// Same structure here: the column is read in the order of a keyed pseudo-random bijection pi of [0, R) (4-round
// Feistel network with cycle walking, keyed by (seed, cell index)) and then sorted by score with a STABLE sort, so
// equal scores end up in pi's order; the final position of a row is its ranking.  Wherever a column has no ties the
// result is the reference's bit for bit; among ties any order is a valid draw (NumPy's own tie order comes from
// PCG64 + an unstable introsort and is not reproducible outside NumPy), ours is a pure function of (seed, cell).
//
// One cluster of 8 CTAs owns a column.  The sort is an LSD radix sort on order-preserving integer keys, 8 bits per pass; an
// OR / AND reduction over the keys finds the digits in which the column varies at all, and only those are sorted
// (imputed accessibility is trunc(1e6 p): two passes).  A pass is the classic counting sort: digit histogram,
// exclusive scan, then tiles of 4 096 items whose in-tile ranks come from a warp match (8 ballots) + per-warp counters (stable:
// items are taken in order, warp by warp).  The matrix is transposed into column-major staging first, so the
// histogram reads and the ping-pong buffers are coalesced and a column (4 B x R) stays L2 resident while it is
// being sorted.  Integer / byte work throughout; bound by L2 and shared-memory traffic, not by the tensor cores.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "ranksum_kernels.cuh"  // rank_key, RankKey

namespace cgs {

constexpr int kRkThreads = 512;
constexpr int kRkWarps = kRkThreads / 32;
constexpr int kRkItems = 8;                        // consecutive 32-item rounds per warp and tile
constexpr int kRkTile = kRkThreads * kRkItems;     // 4 096 items

__host__ __device__ __forceinline__ uint32_t rk_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ uint32_t rk_column_key(uint64_t seed, uint64_t col) {
    uint32_t k = rk_mix32((uint32_t)seed ^ 0x9e3779b9u);
    k = rk_mix32(k ^ (uint32_t)(seed >> 32));
    k = rk_mix32(k ^ (uint32_t)col);
    return rk_mix32(k + (uint32_t)(col >> 32));
}
// smallest h >= 1 with 2^(2h) >= n
__host__ __device__ __forceinline__ int rk_half_bits(uint32_t n) {
    int h = 1;
    while (h < 16 && (1ull << (2 * h)) < (unsigned long long)n) ++h;
    return h;
}
// pi(t): a bijection of [0, n) -- 4 Feistel rounds on 2h bits, walked until the image falls inside [0, n)
__host__ __device__ __forceinline__ uint32_t rk_permute(uint32_t t, uint32_t n, int half_bits, uint32_t key) {
    const uint32_t mask = (1u << half_bits) - 1u;
    uint32_t x = t;
    do {
        uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
        for (uint32_t i = 1; i <= 4; ++i) {
            const uint32_t f = rk_mix32(r ^ (key + 0x9e3779b9u * i)) & mask;
            const uint32_t nl = r;
            r = l ^ f;
            l = nl;
        }
        x = (l << half_bits) | r;
    } while (x >= n);
    return x;
}

// ascending key <=> descending score; NaN scores rank last (NumPy sorts them to the end of (-scores).argsort())
__device__ __forceinline__ uint32_t rk_desc_key(int32_t v) { return ~rank_key(v); }
__device__ __forceinline__ uint32_t rk_desc_key(float v) { return v != v ? 0xffffffffu : rank_key(-v); }
__device__ __forceinline__ unsigned long long rk_desc_key(double v) { return v != v ? ~0ull : rank_key(-v); }

// out[c][r] = in[r][c] for a tile of 32 x 32; in: n_rows x n_cols (row stride ld_in), out: n_cols x n_rows (ld_out)
template <typename T>
__global__ void __launch_bounds__(256)
rk_transpose_kernel(const T *__restrict__ in, int64_t ld_in, int64_t n_rows, int64_t n_cols, T *__restrict__ out,
                    int64_t ld_out) {
    __shared__ T tile[32][33];
    const int64_t tiles_r = (n_rows + 31) / 32;
    const int64_t r0 = ((int64_t)blockIdx.x % tiles_r) * 32, c0 = ((int64_t)blockIdx.x / tiles_r) * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int64_t r = r0 + i, c = c0 + threadIdx.x;
        if (r < n_rows && c < n_cols) tile[i][threadIdx.x] = in[r * ld_in + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int64_t c = c0 + i, r = r0 + threadIdx.x;
        if (r < n_rows && c < n_cols) out[c * ld_out + r] = tile[threadIdx.x][i];
    }
}

// lanes of `active` (all of which call this) that hold the same 8-bit digit: eight ballots instead of match.any,
// whose cost grows with the number of distinct values in the warp
__device__ __forceinline__ unsigned rk_match_digit(unsigned active, uint32_t dg) {
    unsigned peers = active;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (dg >> b) & 1u;
        const unsigned bal = __ballot_sync(active, bit);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

constexpr int kRkCluster = 8;  // CTAs that share one column (portable cluster size)

struct RankColsParams {
    void *xT;            // n_cols x n_rows scores, one column of the matrix per row; 4-byte scores are replaced by
                         // their rankings in place (a row is read and then written by the same thread, last pass only)
    int32_t *rT;         // n_cols x n_rows rankings when the scores are 8 bytes wide, else unused
    void *keys[2];       // n_clusters x n_rows keys each (ping-pong)
    uint32_t *pay[2];    // n_clusters x n_rows row indices each
    uint32_t n_rows;
    int64_t n_cols;
    uint64_t seed;
    int64_t first_col;   // cell index of column 0 (keys the permutation)
    int half_bits;
};

// One CLUSTER of kRkCluster CTAs owns a column: CTA g sorts the g-th contiguous slice of the current sequence, the
// slices' digit histograms are exchanged through distributed shared memory, and the scatter goes to the column's
// ping-pong buffer in global memory.  The launch keeps only as many clusters in flight as L2 holds columns
// (scores + one ping-pong pair = 12 B per value): with a whole CTA grid of independent columns the pi-ordered
// gathers and the final rank scatter missed L2 and moved 33x the algorithmic bytes through HBM (ncu, r1).
template <typename T>
__global__ void __launch_bounds__(kRkThreads, 2)
rank_columns_kernel(RankColsParams p) {
    using K = typename RankKey<T>::type;
    constexpr int kDigits = (int)sizeof(K);
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ uint32_t warp_hist[kRkWarps][256];
    __shared__ uint32_t local_hist[256];   // this CTA's slice; read by the whole cluster
    __shared__ uint32_t bin_offset[256];
    __shared__ uint32_t scan_tot[8];
    __shared__ K red_or[kRkWarps], red_and[kRkWarps];
    __shared__ K cta_or_and[2];            // read by the whole cluster
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t n = p.n_rows;
    const unsigned g = cluster.block_rank();
    const int64_t cluster_id = blockIdx.x / kRkCluster, n_clusters = gridDim.x / kRkCluster;
    // slice of the sequence this CTA sorts: whole tiles
    const uint32_t slice = (uint32_t)(((((uint64_t)n + kRkCluster - 1) / kRkCluster + kRkTile - 1) / kRkTile) * kRkTile);
    const uint32_t s0 = (uint32_t)min((uint64_t)n, (uint64_t)g * slice), s1 = (uint32_t)min((uint64_t)n, (uint64_t)(g + 1) * slice);
    K *const kb0 = static_cast<K *>(p.keys[0]) + (size_t)cluster_id * n, *const kb1 = static_cast<K *>(p.keys[1]) + (size_t)cluster_id * n;
    uint32_t *const pb0 = p.pay[0] + (size_t)cluster_id * n, *const pb1 = p.pay[1] + (size_t)cluster_id * n;

    for (int64_t col = cluster_id; col < p.n_cols; col += n_clusters) {
        const T *x = static_cast<const T *>(p.xT) + (size_t)col * n;
        int32_t *out = sizeof(T) == 4 ? reinterpret_cast<int32_t *>(p.xT) + (size_t)col * n : p.rT + (size_t)col * n;
        const uint32_t ckey = rk_column_key(p.seed, (uint64_t)(p.first_col + col));

        // digits in which the keys of this column differ at all
        K k_or = 0, k_and = ~(K)0;
        for (uint32_t i = s0 + threadIdx.x; i < s1; i += kRkThreads) {
            const K k = rk_desc_key(x[i]);
            k_or |= k; k_and &= k;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            k_or |= __shfl_xor_sync(0xffffffffu, k_or, off);
            k_and &= __shfl_xor_sync(0xffffffffu, k_and, off);
        }
        if (lane == 0) { red_or[wid] = k_or; red_and[wid] = k_and; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 0; w < kRkWarps; ++w) { k_or |= red_or[w]; k_and &= red_and[w]; }
            cta_or_and[0] = k_or; cta_or_and[1] = k_and;
        }
        cluster.sync();
        k_or = 0; k_and = ~(K)0;
        for (unsigned c = 0; c < kRkCluster; ++c) {
            const K *remote = cluster.map_shared_rank(cta_or_and, c);
            k_or |= remote[0]; k_and &= remote[1];
        }
        const K varying = k_or ^ k_and;
        int passes_left = 0;
        for (int d = 0; d < kDigits; ++d) passes_left += ((varying >> (8 * d)) & 0xff) ? 1 : 0;
        if (passes_left == 0) {  // a constant column: the ranking is the permutation itself
            cluster.sync();      // every CTA has read the scores of its slice before any of them is overwritten
            for (uint32_t t = s0 + threadIdx.x; t < s1; t += kRkThreads) out[rk_permute(t, n, p.half_bits, ckey)] = (int32_t)t;
            continue;
        }

        const K *ksrc = nullptr;  // NULL: the column itself, read through pi
        const uint32_t *psrc = nullptr;
        K *kdst = kb0;
        uint32_t *pdst = pb0;
        for (int d = 0; d < kDigits; ++d) {
            if (!((varying >> (8 * d)) & 0xff)) continue;
            const bool last = --passes_left == 0;
            const int shift = 8 * d;
            // digit histogram of this CTA's slice of the current sequence
            if (threadIdx.x < 256) local_hist[threadIdx.x] = 0u;
            __syncthreads();
            for (uint32_t i0 = s0; i0 < s1; i0 += kRkThreads) {
                const uint32_t i = i0 + threadIdx.x;
                const bool ok = i < s1;
                const unsigned active = __ballot_sync(0xffffffffu, ok);
                if (ok) {
                    const K k = ksrc ? ksrc[i] : rk_desc_key(x[rk_permute(i, n, p.half_bits, ckey)]);
                    const uint32_t dg = (uint32_t)(k >> shift) & 0xffu;
                    const unsigned peers = rk_match_digit(active, dg);
                    if (lane == __ffs(peers) - 1) atomicAdd(&local_hist[dg], (uint32_t)__popc(peers));
                }
            }
            cluster.sync();
            // bin start (exclusive scan of the cluster's totals) + what the lower slices put into the bin
            uint32_t cnt = 0, incl = 0, before = 0;
            if (threadIdx.x < 256) {
                for (unsigned c = 0; c < kRkCluster; ++c) {
                    const uint32_t h = cluster.map_shared_rank(local_hist, c)[threadIdx.x];
                    cnt += h;
                    before += c < g ? h : 0u;
                }
                incl = cnt;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl += t;
                }
                if (lane == 31) scan_tot[wid] = incl;
            }
            __syncthreads();
            if (threadIdx.x < 256) {
                uint32_t base = 0;
                for (int w = 0; w < wid; ++w) base += scan_tot[w];
                bin_offset[threadIdx.x] = base + incl - cnt + before;
            }
            __syncthreads();

            for (uint32_t tile0 = s0; tile0 < s1; tile0 += kRkTile) {
                K key[kRkItems];
                uint32_t pay[kRkItems], local[kRkItems];
                bool ok[kRkItems];
#pragma unroll
                for (int r = 0; r < kRkItems; ++r) {
                    const uint32_t idx = tile0 + wid * (32 * kRkItems) + r * 32 + lane;
                    ok[r] = idx < s1;
                    key[r] = 0; pay[r] = 0;
                    if (ok[r]) {
                        if (!ksrc) {
                            pay[r] = rk_permute(idx, n, p.half_bits, ckey);
                            key[r] = rk_desc_key(x[pay[r]]);
                        } else {
                            key[r] = ksrc[idx];
                            pay[r] = psrc[idx];
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) warp_hist[wid][lane + 32 * j] = 0u;
                __syncwarp();
#pragma unroll
                for (int r = 0; r < kRkItems; ++r) {
                    const unsigned vmask = __ballot_sync(0xffffffffu, ok[r]);
                    if (ok[r]) {
                        const uint32_t dg = (uint32_t)(key[r] >> shift) & 0xffu;
                        const unsigned peers = rk_match_digit(vmask, dg);
                        const int leader = __ffs(peers) - 1;
                        uint32_t base = 0;
                        if (lane == leader) {
                            base = warp_hist[wid][dg];
                            warp_hist[wid][dg] = base + (uint32_t)__popc(peers);
                        }
                        base = __shfl_sync(vmask, base, leader);
                        local[r] = base + (uint32_t)__popc(peers & lt_mask);
                    }
                    __syncwarp();
                }
                __syncthreads();
                if (threadIdx.x < 256) {  // per-warp counts -> start of every warp's run inside the bin
                    uint32_t run = bin_offset[threadIdx.x];
#pragma unroll
                    for (int w = 0; w < kRkWarps; ++w) {
                        const uint32_t c = warp_hist[w][threadIdx.x];
                        warp_hist[w][threadIdx.x] = run;
                        run += c;
                    }
                    bin_offset[threadIdx.x] = run;
                }
                __syncthreads();
#pragma unroll
                for (int r = 0; r < kRkItems; ++r) {
                    if (ok[r]) {
                        const uint32_t dest = warp_hist[wid][(uint32_t)(key[r] >> shift) & 0xffu] + local[r];
                        if (last) {
                            out[pay[r]] = (int32_t)dest;
                        } else {
                            kdst[dest] = key[r];
                            pdst[dest] = pay[r];
                        }
                    }
                }
                __syncwarp();
            }
            __threadfence();
            cluster.sync();  // the pass's global writes are visible to the whole cluster; local_hist may be reused
            ksrc = kdst; psrc = pdst;
            kdst = kdst == kb0 ? kb1 : kb0;
            pdst = pdst == pb0 ? pb1 : pb0;
        }
    }
    cluster.sync();  // no CTA leaves while its shared memory may still be read
}

}  // namespace cgs
