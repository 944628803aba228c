// ranksum_kernels.cuh -- differential features on the device: per region (row) the Wilcoxon
// rank-sum test of scipy.stats.ranksums between foreground and background cells and the log2 fold change,
// i.e. get_wilcox_test_pvalues + get_log2_fc of the reference (src/pycisTopic/diff_features.py:968-995,
// :1095-1120, called from markers() :839-965).
//
// The rank sum needs no sort of the pooled sample:  s_fg = n1 (n1 + 1) / 2 + U,
//   U = #{(x in fg, y in bg): x > y} + 1/2 #{x == y}   (average ranks for ties, exactly what rankdata gives),
// and U is additive over elements of either group.  One CTA owns a row: it sorts the SMALLER group in shared
// memory (bitonic, in pieces when it does not fit) and every element of the other group does a lower- and an
// upper-bound search in it.  All counting is integer (2U in uint64), so the statistic is the exact one; only
// the final z / erfc are floating point (fp64).  HBM-bound in the limit: every value of the row is read once
// per piece (4 or 8 bytes per value); L2 keeps the index lists.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace cgs {

constexpr int kRankThreads = 512;
constexpr int kRankHistBins = 8192;  // 2 x 32 KB of shared-memory counters

// order-preserving keys; equal values <=> equal keys (-0.0 is folded onto +0.0, NaNs sort last)
__device__ __forceinline__ uint32_t rank_key(int32_t v) { return (uint32_t)v ^ 0x80000000u; }
__device__ __forceinline__ uint32_t rank_key(float v) {
    const uint32_t b = __float_as_uint(v + 0.0f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ unsigned long long rank_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
template <typename T> struct RankKey { using type = uint32_t; };
template <> struct RankKey<double> { using type = unsigned long long; };

// Shared-memory layout: logical index i lives at i + (i >> 5).  A binary search over a power-of-two array probes
// addresses that are multiples of 32 words apart -- all in ONE bank (measured: the unskewed search ran 1.8x slower
// than a plain mid-point search); one pad word per 32 spreads them over the banks.
__device__ __forceinline__ int rank_phys(int i) { return i + (i >> 5); }

template <typename K>
__device__ __forceinline__ void bitonic_sort_shared(K *keys, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < (n_pow2 >> 1); i += blockDim.x) {
                const int a = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const int b = a | j;
                const K ka = keys[rank_phys(a)], kb = keys[rank_phys(b)];
                const bool up = (a & k) == 0;
                if ((ka > kb) == up) { keys[rank_phys(a)] = kb; keys[rank_phys(b)] = ka; }
            }
            __syncthreads();
        }
    }
}

// Branch-free lower bounds (#keys < key) of kIlp keys at once over the padded power-of-two array (padding = the
// largest key): the step loop is outermost, so the kIlp dependent chains of one thread overlap their latencies.
constexpr int kRankIlp = 4;
template <typename K>
__device__ __forceinline__ void lower_bounds_shared(const K *keys, int n_pow2, const K (&key)[kRankIlp],
                                                    int (&lb)[kRankIlp]) {
#pragma unroll
    for (int u = 0; u < kRankIlp; ++u) lb[u] = 0;
    for (int step = n_pow2 >> 1; step > 0; step >>= 1) {
#pragma unroll
        for (int u = 0; u < kRankIlp; ++u) lb[u] += keys[rank_phys(lb[u] + step - 1)] < key[u] ? step : 0;
    }
#pragma unroll
    for (int u = 0; u < kRankIlp; ++u) lb[u] += keys[rank_phys(lb[u])] < key[u] ? 1 : 0;
}

// #keys <= key, searched only for values that tie with a staged one
template <typename K>
__device__ __forceinline__ int upper_bound_shared(const K *keys, int n_pow2, K key) {
    int hi = 0;
    for (int step = n_pow2 >> 1; step > 0; step >>= 1) hi += keys[rank_phys(hi + step - 1)] <= key ? step : 0;
    return hi + (keys[rank_phys(hi)] <= key ? 1 : 0);
}

// ---- counting path for integer rows of modest range ------------------------------------------------------------
// Imputed accessibility is trunc(1e6 p): small integers.  When max - min of a row is below `hist_bins`, the two
// groups are histogrammed in shared memory (native shared atomics) and U = sum_v fg[v] (bg_below[v] + bg[v] / 2)
// falls out of one prefix pass over the bins: ~10 instructions per value instead of ~100 for the searches.
constexpr int kRankHistIlp = 8;  // independent loads in flight per thread (a 10 000-cell row is 20 values per thread)
// f(value) over one group of a row; the index loads, then the value loads, are issued back to back
template <typename F>
__device__ __forceinline__ void rank_for_each(const int32_t *row, const int32_t *cols, int n, F f) {
    for (int j0 = threadIdx.x; j0 < n; j0 += kRankHistIlp * kRankThreads) {
        int c[kRankHistIlp], v[kRankHistIlp];
        bool ok[kRankHistIlp];
#pragma unroll
        for (int u = 0; u < kRankHistIlp; ++u) {
            const int j = j0 + u * kRankThreads;
            ok[u] = j < n;
            c[u] = ok[u] ? (cols ? cols[j] : j) : 0;
        }
#pragma unroll
        for (int u = 0; u < kRankHistIlp; ++u) v[u] = ok[u] ? row[c[u]] : 0;
#pragma unroll
        for (int u = 0; u < kRankHistIlp; ++u)
            if (ok[u]) f(v[u]);
    }
}

struct RankSide {
    const void *x;        // matrix holding this group's values
    int64_t ld;           // its row stride in elements
    const int32_t *cols;  // column of every member, or NULL for 0 .. n-1
    int32_t n;
};

// `sorted` is the group staged in shared memory, `other` the one that searches; sorted_is_fg tells which is which.
template <typename T>
__global__ void __launch_bounds__(kRankThreads)
ranksum_rows_kernel(RankSide sorted, RankSide other, int sorted_is_fg, int64_t n_rows, int32_t piece, int32_t hist_bins,
                    double *__restrict__ statistic, double *__restrict__ pvalue, double *__restrict__ log2fc) {
    using K = typename RankKey<T>::type;
    extern __shared__ __align__(16) unsigned char rank_smem[];
    K *keys = reinterpret_cast<K *>(rank_smem);
    __shared__ double red_d[2][kRankThreads / 32];
    __shared__ unsigned long long red_u[kRankThreads / 32];
    const T *xs = static_cast<const T *>(sorted.x);
    const T *xo = static_cast<const T *>(other.x);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const T *row_s = xs + r * sorted.ld;
        const T *row_o = xo + r * other.ld;
        double sum_s = 0.0, sum_o = 0.0;
        unsigned long long u2 = 0ull;  // 2 U
        bool counted = false;
        if constexpr (std::is_same<T, int32_t>::value) {
            if (hist_bins > 0) {
                __shared__ int range_s[2][kRankThreads / 32];
                int vmin = 2147483647, vmax = -2147483647 - 1;
                auto range = [&](int v) { vmin = min(vmin, v); vmax = max(vmax, v); };
                rank_for_each(row_s, sorted.cols, sorted.n, range);
                rank_for_each(row_o, other.cols, other.n, range);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    vmin = min(vmin, __shfl_xor_sync(0xffffffffu, vmin, off));
                    vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, off));
                }
                if (lane == 0) { range_s[0][wid] = vmin; range_s[1][wid] = vmax; }
                __syncthreads();
                for (int w = 0; w < kRankThreads / 32; ++w) { vmin = min(vmin, range_s[0][w]); vmax = max(vmax, range_s[1][w]); }
                __syncthreads();
                if ((long long)vmax - (long long)vmin < (long long)hist_bins) {  // uniform over the CTA
                    uint32_t *hs = reinterpret_cast<uint32_t *>(rank_smem), *ho = hs + hist_bins;
                    const int nb = vmax - vmin + 1;
                    for (int i = threadIdx.x; i < nb; i += blockDim.x) { hs[i] = 0u; ho[i] = 0u; }
                    __syncthreads();
                    long long isum_s = 0, isum_o = 0;  // exact; fits: |v| < 2^31, n < 2^31
                    rank_for_each(row_s, sorted.cols, sorted.n, [&](int v) { isum_s += v; atomicAdd(hs + (v - vmin), 1u); });
                    rank_for_each(row_o, other.cols, other.n, [&](int v) { isum_o += v; atomicAdd(ho + (v - vmin), 1u); });
                    sum_s = (double)isum_s; sum_o = (double)isum_o;
                    __syncthreads();
                    const uint32_t *hf = sorted_is_fg ? hs : ho, *hb = sorted_is_fg ? ho : hs;
                    const int per = (nb + blockDim.x - 1) / blockDim.x;  // a contiguous slice of bins per thread
                    const int b0 = min(nb, (int)threadIdx.x * per), b1 = min(nb, b0 + per);
                    uint32_t mine = 0;
                    for (int i = b0; i < b1; ++i) mine += hb[i];
                    uint32_t incl = mine;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
                        if (lane >= off) incl += t;
                    }
                    __shared__ uint32_t warp_tot[kRankThreads / 32];
                    if (lane == 31) warp_tot[wid] = incl;
                    __syncthreads();
                    unsigned long long below = incl - mine;  // background values below this thread's first bin
                    for (int w = 0; w < wid; ++w) below += warp_tot[w];
                    for (int i = b0; i < b1; ++i) {
                        const unsigned long long bg = hb[i];
                        u2 += (unsigned long long)hf[i] * (2ull * below + bg);
                        below += bg;
                    }
                    counted = true;
                    __syncthreads();
                }
            }
        }
        for (int32_t p0 = 0; p0 < sorted.n && !counted; p0 += piece) {
            const int cnt = min(piece, sorted.n - p0);
            int n_pow2 = 2;
            while (n_pow2 < cnt) n_pow2 <<= 1;
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                K key = ~(K)0;
                if (i < cnt) {
                    const T v = row_s[sorted.cols ? sorted.cols[p0 + i] : p0 + i];
                    sum_s += (double)v;
                    key = rank_key(v);
                }
                keys[rank_phys(i)] = key;
            }
            __syncthreads();
            bitonic_sort_shared(keys, n_pow2);
            for (int j0 = 0; j0 < other.n; j0 += kRankIlp * kRankThreads) {
                K key[kRankIlp];
                bool ok[kRankIlp];
                int lb[kRankIlp];
#pragma unroll
                for (int u = 0; u < kRankIlp; ++u) {
                    const int j = j0 + u * kRankThreads + threadIdx.x;
                    ok[u] = j < other.n;
                    const T v = ok[u] ? row_o[other.cols ? other.cols[j] : j] : T(0);
                    if (p0 == 0 && ok[u]) sum_o += (double)v;
                    key[u] = rank_key(v);
                }
                lower_bounds_shared(keys, n_pow2, key, lb);
#pragma unroll
                for (int u = 0; u < kRankIlp; ++u) {
                    const int lo = min(lb[u], cnt);  // a value equal to the padding key must not count the padding
                    int ub = lo;
                    if (lo < cnt && keys[rank_phys(lo)] == key[u]) ub = min(upper_bound_shared(keys, n_pow2, key[u]), cnt);
                    // sorted = fg: this bg value is beaten by the cnt - ub larger fg values;
                    // sorted = bg: this fg value beats the lo smaller bg values; ties count one half either way
                    if (ok[u])
                        u2 += (unsigned long long)(sorted_is_fg ? 2 * (cnt - ub) : 2 * lo) + (unsigned long long)(ub - lo);
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sum_s += __shfl_down_sync(0xffffffffu, sum_s, off);
            sum_o += __shfl_down_sync(0xffffffffu, sum_o, off);
            u2 += __shfl_down_sync(0xffffffffu, u2, off);
        }
        if (lane == 0) { red_d[0][wid] = sum_s; red_d[1][wid] = sum_o; red_u[wid] = u2; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double ts = 0.0, to = 0.0;
            unsigned long long tu = 0ull;
            for (int w = 0; w < kRankThreads / 32; ++w) { ts += red_d[0][w]; to += red_d[1][w]; tu += red_u[w]; }
            const double n1 = sorted_is_fg ? (double)sorted.n : (double)other.n;
            const double n2 = sorted_is_fg ? (double)other.n : (double)sorted.n;
            const double s = 0.5 * n1 * (n1 + 1.0) + 0.5 * (double)tu;  // rank sum of the foreground
            const double expected = n1 * (n1 + n2 + 1.0) / 2.0;
            const double z = (s - expected) / sqrt(n1 * n2 * (n1 + n2 + 1.0) / 12.0);
            statistic[r] = z;
            pvalue[r] = erfc(fabs(z) * 0.70710678118654752440);  // 2 norm.sf(|z|)
            const double mean_fg = (sorted_is_fg ? ts : to) / n1, mean_bg = (sorted_is_fg ? to : ts) / n2;
            log2fc[r] = log2((mean_fg + 1e-12) / (mean_bg + 1e-12));
        }
        __syncthreads();
    }
}

}  // namespace cgs
