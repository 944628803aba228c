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
#include "common.cuh"

namespace cgs {

constexpr int kRankThreads = 512;

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

struct RankSide {
    const void *x;        // matrix holding this group's values
    int64_t ld;           // its row stride in elements
    const int32_t *cols;  // column of every member, or NULL for 0 .. n-1
    int32_t n;
};

// `sorted` is the group staged in shared memory, `other` the one that searches; sorted_is_fg tells which is which.
template <typename T>
__global__ void __launch_bounds__(kRankThreads)
ranksum_rows_kernel(RankSide sorted, RankSide other, int sorted_is_fg, int64_t n_rows, int32_t piece,
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
        for (int32_t p0 = 0; p0 < sorted.n; p0 += piece) {
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
