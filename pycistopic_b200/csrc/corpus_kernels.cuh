// corpus_kernels.cuh -- token-list construction on the device.
//
// Replaces, for pycisTopic's binary matrix, the host-side
//   sparse.csr_matrix(cistopic_obj.binary_matrix.transpose())   (lda_models.py:151)
//   lda.utils.matrix_to_lists(X)                                
// The result is the cell-major token stream: cell_ptr int64[D+1] and tok_region int32[N]
// with the regions of every cell in ascending order -- bit-identical to what scipy + the
// `lda` package produce.
//
// Method (HBM-bound integer work, no sort library): (1) count tokens per cell, (2) exclusive
// scan, (3) scatter region ids into each cell's segment in arbitrary order with one global
// atomic per token, (4) one CTA per cell canonicalises its segment through a shared-memory
// BITMAP over the region axis (227 KB of smem = 1.8 M regions per window): set a bit per
// token, then emit the set bits in order with a block-wide prefix popcount.  The bitmap also
// removes duplicates, which is the right semantics for binary accessibility data.
#pragma once
#include "common.cuh"

namespace cgs {

// ---- (1) per-cell counts from a regions x cells CSR -----------------------------------------
__global__ void count_cells_kernel(const int32_t *__restrict__ cell_idx, int64_t nnz,
                                   int32_t cell_begin, int32_t cell_end,
                                   unsigned long long *__restrict__ raw_count) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += stride) {
        int32_t c = cell_idx[j];
        if (c >= cell_begin && c < cell_end) atomicAdd(&raw_count[c - cell_begin], 1ull);
    }
}

// ---- (2) exclusive scan of int64 counts (single CTA; D is at most a few 1e5..1e6) -----------
__global__ void exclusive_scan_i64_kernel(const unsigned long long *__restrict__ in,
                                          int64_t *__restrict__ out, int32_t n) {
    __shared__ unsigned long long warp_sums[32];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int32_t base = 0; base < n; base += blockDim.x) {
        int32_t i = base + threadIdx.x;
        unsigned long long v = (i < n) ? in[i] : 0ull;
        unsigned long long s = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, s, off);
            if (lane >= off) s += t;
        }
        if (lane == 31) warp_sums[wid] = s;
        __syncthreads();
        if (wid == 0) {
            unsigned long long ws = (lane < nw) ? warp_sums[lane] : 0ull;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                unsigned long long t = __shfl_up_sync(0xffffffffu, ws, off);
                if (lane >= off) ws += t;
            }
            warp_sums[lane] = ws;  // inclusive over warps
        }
        __syncthreads();
        unsigned long long carry = carry_s;
        unsigned long long prefix = carry + (wid > 0 ? warp_sums[wid - 1] : 0ull) + (s - v);
        if (i < n) out[i] = (int64_t)prefix;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = carry + warp_sums[nw - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = (int64_t)carry_s;
}

// ---- (3) scatter region ids into the owning cell's segment (one warp per region row) --------
__global__ void scatter_rows_kernel(const int64_t *__restrict__ row_ptr,
                                    const int32_t *__restrict__ cell_idx, int32_t n_regions,
                                    int32_t cell_begin, int32_t cell_end,
                                    const int64_t *__restrict__ raw_ptr,
                                    unsigned long long *__restrict__ fill,
                                    int32_t *__restrict__ tmp_region) {
    const int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_regions; r += nwarps) {
        int64_t b = row_ptr[r], e = row_ptr[r + 1];
        for (int64_t j = b + lane; j < e; j += 32) {
            int32_t c = cell_idx[j];
            if (c >= cell_begin && c < cell_end) {
                unsigned long long pos = atomicAdd(&fill[c - cell_begin], 1ull);
                tmp_region[raw_ptr[c - cell_begin] + (int64_t)pos] = (int32_t)r;
            }
        }
    }
}

// ---- (4) per-cell bitmap canonicalisation: sort + unique ------------------------------------
// One CTA per cell (grid-stride).  `words` = number of 32-bit bitmap words in shared memory.
// out segment starts at raw_ptr[cell]; uniq_count[cell] receives the number of unique regions.
__global__ void bitmap_sort_cells_kernel(const int32_t *__restrict__ tmp_region,
                                         const int64_t *__restrict__ raw_ptr, int32_t n_cells,
                                         int32_t n_regions, int32_t words,
                                         int32_t *__restrict__ out_region,
                                         unsigned long long *__restrict__ uniq_count,
                                         int *__restrict__ bad_index_flag) {
    extern __shared__ uint32_t bitmap[];
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t running_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int64_t window = (int64_t)words * 32;

    for (int32_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        const int64_t b = raw_ptr[cell], e = raw_ptr[cell + 1];
        if (threadIdx.x == 0) running_s = 0;
        for (int64_t w0 = 0; w0 < n_regions; w0 += window) {
            const int32_t nbits = (int32_t)min(window, (int64_t)n_regions - w0);
            const int32_t nwords = (nbits + 31) >> 5;
            for (int32_t i = threadIdx.x; i < nwords; i += blockDim.x) bitmap[i] = 0u;
            __syncthreads();
            for (int64_t j = b + threadIdx.x; j < e; j += blockDim.x) {
                int32_t r = tmp_region[j];
                if (r < 0 || r >= n_regions) { *bad_index_flag = 1; continue; }
                int64_t off = (int64_t)r - w0;
                if (off >= 0 && off < nbits) atomicOr(&bitmap[off >> 5], 1u << (off & 31));
            }
            __syncthreads();
            // each thread owns a contiguous slice of words
            const int32_t per = (nwords + blockDim.x - 1) / blockDim.x;
            const int32_t ws = min(nwords, (int32_t)threadIdx.x * per);
            const int32_t we = min(nwords, ws + per);
            uint32_t cnt = 0;
            for (int32_t i = ws; i < we; ++i) cnt += __popc(bitmap[i]);
            uint32_t s = cnt;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, s, off);
                if (lane >= off) s += t;
            }
            if (lane == 31) warp_sums[wid] = s;
            __syncthreads();
            if (wid == 0) {
                uint32_t v = (lane < nw) ? warp_sums[lane] : 0u;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    uint32_t t = __shfl_up_sync(0xffffffffu, v, off);
                    if (lane >= off) v += t;
                }
                warp_sums[lane] = v;
            }
            __syncthreads();
            const uint32_t running = running_s;
            int64_t pos = b + running + (wid > 0 ? warp_sums[wid - 1] : 0u) + (s - cnt);
            for (int32_t i = ws; i < we; ++i) {
                uint32_t m = bitmap[i];
                while (m) {
                    int bit = __ffs(m) - 1;
                    m &= m - 1;
                    out_region[pos++] = (int32_t)(w0 + (int64_t)i * 32 + bit);
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) running_s = running + warp_sums[nw - 1];
            __syncthreads();
        }
        if (threadIdx.x == 0) uniq_count[cell] = running_s;
        __syncthreads();
    }
}

// ---- compaction when duplicates were removed --------------------------------------------------
__global__ void compact_cells_kernel(const int32_t *__restrict__ src, const int64_t *__restrict__ raw_ptr,
                                     const int64_t *__restrict__ cell_ptr, int32_t n_cells,
                                     int32_t *__restrict__ dst) {
    for (int32_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        const int64_t sb = raw_ptr[cell], db = cell_ptr[cell];
        const int64_t n = cell_ptr[cell + 1] - db;
        for (int64_t j = threadIdx.x; j < n; j += blockDim.x) dst[db + j] = src[sb + j];
    }
}

// ---- per-cell raw counts from a cells x regions CSR row pointer -------------------------------
__global__ void diff_ptr_kernel(const int64_t *__restrict__ ptr, int32_t n,
                                unsigned long long *__restrict__ out) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (unsigned long long)(ptr[i + 1] - ptr[i]);
}

// ---- DS export: cell id of every token --------------------------------------------------------
__global__ void expand_cells_kernel(const int64_t *__restrict__ cell_ptr, int32_t n_cells,
                                    int32_t *__restrict__ DS) {
    for (int32_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        const int64_t b = cell_ptr[cell], e = cell_ptr[cell + 1];
        for (int64_t j = b + threadIdx.x; j < e; j += blockDim.x) DS[j] = cell;
    }
}

// ---- cell schedule: order cells by descending token count (longest first) --------------------
// Counting sort on a log-ish bucket key would lose exact ordering; D is small enough that the
// host sorts the D counts once (corpus build), see corpus.cu.

}  // namespace cgs
