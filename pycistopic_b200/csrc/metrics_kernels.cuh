// metrics_kernels.cuh -- ingredients of the four model-quality metrics of run_cgs_model
// (reference src/pycisTopic/lda_models.py:291-305, :332-344) computed from the device-resident
// counts, so that neither the K x W distributions nor the token matrix have to visit the host:
//   * topic_gram_kernel      exact int64 Gram  S[i][j] = sum_w n_wk[w,i] n_wk[w,j]   (Arun 2010 singular
//                            values and Cao-Juan 2009 cosines are functions of topic_word topic_word^T)
//   * cell_mixture_kernel    sum_d nd_d doc_topic[d,k] and sum_d nd_d^2   (Arun's cm2, tmtoolkit's
//                            marginal_topic_distrib)
//   * top_regions_*          the top_n regions of every topic (Mimno 2011: argsort(...)[:, :-(n+1):-1])
//   * cell_masks_kernel / pair_count_kernel   document and co-document frequencies of those regions from
//                            the cell-major token list (one bit per (cell, topic, rank))
//   * last_positive_* kernels   the entries that survive utils.loglikelihood's overwritten loops
//                            (utils.py:129-160)
// All integer results are exact; fp64 sums are two-stage and deterministic.
#pragma once
#include "common.cuh"

namespace cgs {

// ---- integer Gram of n_wk ---------------------------------------------------------------------
constexpr int kGramRows = 16;      // regions per shared-memory slab
constexpr int kGramThreads = 128;  // one 4 x 4 tile of topic pairs per thread

// Upper-triangle 4 x 4 tile number t (row-major over ti <= tj) -> (ti, tj); nt = tiles per side.
__device__ __forceinline__ void gram_tile_coords(int t, int nt, int &ti, int &tj) {
    int row = 0, len = nt;
    while (t >= len) { t -= len; ++row; --len; }
    ti = row;
    tj = row + t;
}

__global__ void __launch_bounds__(kGramThreads)
topic_gram_kernel(const int32_t *__restrict__ n_wk, int64_t W, int32_t Kp, int32_t n_tiles,
                  unsigned long long *__restrict__ gram /* Kp x Kp, zeroed */) {
    extern __shared__ int32_t slab[];  // kGramRows x Kp
    const int nt = Kp / 4;
    const int tile = blockIdx.y * kGramThreads + threadIdx.x;
    int ti = 0, tj = 0;
    const bool active = tile < n_tiles;
    if (active) gram_tile_coords(tile, nt, ti, tj);
    long long acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0;
    const int64_t n_slabs = (W + kGramRows - 1) / kGramRows;
    const int slab_vec = kGramRows * Kp / 4;  // int4 per slab
    for (int64_t s = blockIdx.x; s < n_slabs; s += gridDim.x) {
        const int64_t row0 = s * kGramRows;
        const int64_t valid = min((int64_t)kGramRows, W - row0) * Kp / 4;
        const int4 *src = reinterpret_cast<const int4 *>(n_wk + row0 * Kp);
        for (int i = threadIdx.x; i < slab_vec; i += kGramThreads)
            reinterpret_cast<int4 *>(slab)[i] = i < valid ? __ldg(src + i) : make_int4(0, 0, 0, 0);
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int r = 0; r < kGramRows; ++r) {
                const int4 a = *reinterpret_cast<const int4 *>(slab + r * Kp + 4 * ti);
                const int4 b = *reinterpret_cast<const int4 *>(slab + r * Kp + 4 * tj);
                const int av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] += (long long)av[x] * (long long)bv[y];
            }
        }
        __syncthreads();
    }
    if (active) {
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y)
                if (acc[x][y] != 0)
                    atomicAdd(gram + (size_t)(4 * ti + x) * Kp + 4 * tj + y, (unsigned long long)acc[x][y]);
    }
}

// ---- cell-side mixture --------------------------------------------------------------------------
// partials[b][k] = sum over the block's cells of nd_d (n_dk[d,k] + alpha) / (nd_d + alpha K), k < K;
// partials[b][K] = sum nd_d^2.  blockDim = (32, 8): x = topic lane, y = cell lane.
constexpr int kMixMaxPerLane = 16;  // 32 x 16 = 512 topics

__global__ void __launch_bounds__(256)
cell_mixture_kernel(const int32_t *__restrict__ n_dk, const int64_t *__restrict__ cell_ptr, int32_t D, int32_t K,
                    int32_t Kp, double alpha, double *__restrict__ partials) {
    __shared__ double red[8][33];
    double acc[kMixMaxPerLane];
#pragma unroll
    for (int j = 0; j < kMixMaxPerLane; ++j) acc[j] = 0.0;
    double len_sq = 0.0;
    const int per_lane = (K + 31) / 32;
    for (int64_t d = (int64_t)blockIdx.x * 8 + threadIdx.y; d < D; d += (int64_t)gridDim.x * 8) {
        const double nd = (double)(cell_ptr[d + 1] - cell_ptr[d]);
        const double scale = nd / (nd + alpha * (double)K);
        if (threadIdx.x == 0) len_sq += nd * nd;
#pragma unroll
        for (int j = 0; j < kMixMaxPerLane; ++j) {
            const int k = threadIdx.x + 32 * j;
            if (j < per_lane && k < K) acc[j] += ((double)n_dk[d * Kp + k] + alpha) * scale;
        }
    }
    double *out = partials + (size_t)blockIdx.x * (K + 1);
#pragma unroll
    for (int j = 0; j < kMixMaxPerLane; ++j) {
        if (j >= per_lane) break;  // uniform over the block
        red[threadIdx.y][threadIdx.x] = acc[j];
        __syncthreads();
        if (threadIdx.y == 0) {
            double s = 0.0;
#pragma unroll
            for (int y = 0; y < 8; ++y) s += red[y][threadIdx.x];
            const int k = threadIdx.x + 32 * j;
            if (k < K) out[k] = s;
        }
        __syncthreads();
    }
    red[threadIdx.y][threadIdx.x] = len_sq;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        double s = 0.0;
#pragma unroll
        for (int y = 0; y < 8; ++y) s += red[y][0];
        out[K] = s;
    }
}

// out[k] = sum_b partials[b][k], k <= K (serial over blocks: deterministic)
__global__ void column_sum_partials_kernel(const double *__restrict__ partials, int32_t n_blocks, int32_t n_cols,
                                           double *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_cols) return;
    double s = 0.0;
    for (int b = 0; b < n_blocks; ++b) s += partials[(size_t)b * n_cols + k];
    out[k] = s;
}

// ---- top regions of every topic -----------------------------------------------------------------
// Order: count descending, ties by region index descending -- the order of a stable ascending argsort
// read backwards (numpy's default argsort leaves the order of ties unspecified).  A warp keeps the 32
// best keys, one per lane, sorted; key = (count << 32 | region) + 1, 0 = empty.
constexpr int kTopSlab = 4096;  // regions per CTA in the first stage
constexpr int kTopWarps = 8;

__device__ __forceinline__ unsigned long long top_key(int32_t count, int64_t region) {
    return (((unsigned long long)(uint32_t)count << 32) | (unsigned long long)(uint32_t)region) + 1ull;
}

__device__ __forceinline__ void top_insert(unsigned long long &mine, unsigned long long cand, int lane) {
    const unsigned bigger = __ballot_sync(0xffffffffu, mine > cand);
    const int pos = __popc(bigger);  // keys are sorted descending, so the bigger ones are lanes [0, pos)
    const unsigned long long up = __shfl_up_sync(0xffffffffu, mine, 1);
    if (lane == pos) mine = cand;
    else if (lane > pos) mine = up;
}

// Feeds the keys held by the lanes (one each, 0 = nothing) into the sorted list.
__device__ __forceinline__ void top_offer(unsigned long long &mine, unsigned long long key, int lane) {
    unsigned long long thr = __shfl_sync(0xffffffffu, mine, 31);
    unsigned pending = __ballot_sync(0xffffffffu, key > thr);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const unsigned long long cand = __shfl_sync(0xffffffffu, key, src);
        if (cand > thr) {
            top_insert(mine, cand, lane);
            thr = __shfl_sync(0xffffffffu, mine, 31);
        }
    }
}

// grid (n_slabs, ceil(K / kTopWarps)); cand is [n_slabs][K][32]
__global__ void __launch_bounds__(kTopWarps * 32)
top_regions_slab_kernel(const int32_t *__restrict__ n_wk, int64_t W, int32_t K, int32_t Kp,
                        unsigned long long *__restrict__ cand) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.y * kTopWarps + (threadIdx.x >> 5);
    if (k >= K) return;
    const int64_t r0 = (int64_t)blockIdx.x * kTopSlab;
    const int64_t r1 = min(W, r0 + kTopSlab);
    unsigned long long mine = 0ull;
    for (int64_t base = r0; base < r1; base += 32) {
        const int64_t r = base + lane;
        const unsigned long long key = r < r1 ? top_key(__ldg(n_wk + r * Kp + k), r) : 0ull;
        top_offer(mine, key, lane);
    }
    cand[((size_t)blockIdx.x * K + k) * 32 + lane] = mine;
}

// one warp per topic: merge the slab lists, write the first top_n regions (and their counts)
__global__ void __launch_bounds__(kTopWarps * 32)
top_regions_merge_kernel(const unsigned long long *__restrict__ cand, int32_t n_slabs, int32_t K, int32_t top_n,
                         int32_t *__restrict__ top_regions, int32_t *__restrict__ top_counts) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * kTopWarps + (threadIdx.x >> 5);
    if (k >= K) return;
    unsigned long long mine = 0ull;
    for (int s = 0; s < n_slabs; ++s) top_offer(mine, cand[((size_t)s * K + k) * 32 + lane], lane);
    if (lane < top_n) {
        const unsigned long long key = mine - 1ull;  // lists hold >= top_n entries because W >= top_n
        top_regions[k * top_n + lane] = (int32_t)(uint32_t)(key & 0xffffffffull);
        top_counts[k * top_n + lane] = (int32_t)(uint32_t)(key >> 32);
    }
}

// ---- document / co-document frequencies of the top regions ----------------------------------------
// The host turns the K x top_n lists into: a bitmap over the region axis (bit set = region is in some list),
// the sorted distinct listed regions u[0..M), and for each of them the (topic << 5 | rank) entries
// ent[ent_ptr[i] .. ent_ptr[i+1]).  One warp walks one cell's tokens and sets bit `rank` of
// masks[cell][topic] for every listed region the cell contains.
constexpr int kMaskWarps = 8;

__global__ void __launch_bounds__(kMaskWarps * 32)
cell_masks_kernel(const int64_t *__restrict__ cell_ptr, const int32_t *__restrict__ tok_region, int32_t D, int32_t K,
                  const uint32_t *__restrict__ bitmap, const int32_t *__restrict__ u, int32_t M,
                  const int32_t *__restrict__ ent_ptr, const int32_t *__restrict__ ent,
                  uint32_t *__restrict__ masks /* D x K */) {
    extern __shared__ uint32_t mask_s[];  // kMaskWarps x K
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t *mine = mask_s + (size_t)wid * K;
    for (int64_t d = (int64_t)blockIdx.x * kMaskWarps + wid; d < D; d += (int64_t)gridDim.x * kMaskWarps) {
        for (int k = lane; k < K; k += 32) mine[k] = 0u;
        __syncwarp();
        const int64_t b = cell_ptr[d], e = cell_ptr[d + 1];
        for (int64_t i = b + lane; i < e; i += 32) {
            const int32_t w = __ldg(tok_region + i);
            if ((__ldg(bitmap + (w >> 5)) >> (w & 31)) & 1u) {
                int lo = 0, hi = M - 1;  // w is in u: plain binary search
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (__ldg(u + mid) < w) lo = mid + 1; else hi = mid;
                }
                for (int j = __ldg(ent_ptr + lo); j < __ldg(ent_ptr + lo + 1); ++j) {
                    const int32_t en = __ldg(ent + j);
                    atomicOr(mine + (en >> 5), 1u << (en & 31));
                }
            }
        }
        __syncwarp();
        for (int k = lane; k < K; k += 32) masks[(size_t)d * K + k] = mine[k];
        __syncwarp();
    }
}

// grid (K, cell slices); blockDim = top_n (top_n + 1) / 2 rounded up to a warp (<= 544); thread p owns the
// pair (a >= b); codf[t][a][b] = codf[t][b][a] = #cells with both bits (a == b: the document frequency).
constexpr int kPairMaxThreads = 544;
__global__ void __launch_bounds__(kPairMaxThreads)
pair_count_kernel(const uint32_t *__restrict__ masks, int32_t D, int32_t K, int32_t top_n,
                  int32_t *__restrict__ codf /* K x top_n x top_n, zeroed */) {
    __shared__ uint32_t ms[256];
    const int t = blockIdx.x;
    int a = 0, b = 0;
    const int n_pairs = top_n * (top_n + 1) / 2;
    const bool active = (int)threadIdx.x < n_pairs;
    if (active) {
        int p = threadIdx.x;
        while (p > a) { p -= a + 1; ++a; }
        b = p;
    }
    const int64_t per = ((int64_t)D + gridDim.y - 1) / gridDim.y;
    const int64_t d0 = (int64_t)blockIdx.y * per, d1 = min((int64_t)D, d0 + per);
    int32_t count = 0;
    for (int64_t base = d0; base < d1; base += 256) {
        const int64_t d = base + threadIdx.x;
        if (threadIdx.x < 256) ms[threadIdx.x] = d < d1 ? masks[(size_t)d * K + t] : 0u;
        __syncthreads();
        if (active) {
#pragma unroll 8
            for (int i = 0; i < 256; ++i) {
                const uint32_t m = ms[i];
                count += (int32_t)((m >> a) & (m >> b) & 1u);
            }
        }
        __syncthreads();
    }
    if (active && count) {
        atomicAdd(codf + ((size_t)t * top_n + a) * top_n + b, count);
        if (a != b) atomicAdd(codf + ((size_t)t * top_n + b) * top_n + a, count);
    }
}

// ---- what survives utils.loglikelihood's overwritten loops ------------------------------------------
// *out = max over rows r with counts[r * stride + col] > 0 of r (out initialised to -1)
__global__ void last_positive_row_kernel(const int32_t *__restrict__ counts, int64_t rows, int32_t stride, int32_t col,
                                         int *__restrict__ out) {
    int best = -1;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x)
        if (counts[r * stride + col] > 0) best = (int)r;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, off));
    if ((threadIdx.x & 31) == 0 && best >= 0) atomicMax(out, best);
}

// *out = index of the last cell that holds a token (-1 if none)
__global__ void last_nonempty_cell_kernel(const int64_t *__restrict__ cell_ptr, int32_t D, int *__restrict__ out) {
    int best = -1;
    for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < D; d += (int64_t)gridDim.x * blockDim.x)
        if (cell_ptr[d + 1] > cell_ptr[d]) best = (int)d;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, off));
    if ((threadIdx.x & 31) == 0 && best >= 0) atomicMax(out, best);
}

__global__ void cell_sizes_kernel(const int64_t *__restrict__ cell_ptr, int32_t D, int64_t *__restrict__ sizes) {
    for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < D; d += (int64_t)gridDim.x * blockDim.x)
        sizes[d] = cell_ptr[d + 1] - cell_ptr[d];
}

}  // namespace cgs
