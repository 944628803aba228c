// stats_kernels.cuh -- log-likelihood (lda._lda._loglikelihood) and the
// reference-layout exports (lda final normalisation; DataFrame layouts of
// lda_models.py:345-357).  fp64 throughout, deterministic two-stage reductions.
#pragma once
#include "common.cuh"

namespace cgs {

constexpr int kReduceThreads = 256;

__device__ __forceinline__ double block_sum(double v, double *smem) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = (lane < (blockDim.x >> 5)) ? smem[lane] : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) r += __shfl_down_sync(0xffffffffu, r, off);
    }
    __syncthreads();
    return r;  // valid in thread 0
}

// sum over entries c > 0 of lgamma(prior + c) - lgamma(prior); one partial per block.
__global__ void lgamma_sum_kernel(const int32_t *__restrict__ counts, int64_t n, double prior,
                                  double *__restrict__ partials) {
    __shared__ double smem[32];
    const double lg_prior = lgamma(prior);
    double acc = 0.0;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int32_t c = counts[i];
        if (c > 0) acc += lgamma(prior + (double)c) - lg_prior;
    }
    const double s = block_sum(acc, smem);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// sum_d lgamma(alpha K) - lgamma(alpha K + nd_d), nd_d from the cell pointer.
__global__ void cell_norm_sum_kernel(const int64_t *__restrict__ cell_ptr, int32_t n_cells,
                                     double alphaK, double *__restrict__ partials) {
    __shared__ double smem[32];
    const double lg = lgamma(alphaK);
    double acc = 0.0;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < n_cells; d += stride)
        acc += lg - lgamma(alphaK + (double)(cell_ptr[d + 1] - cell_ptr[d]));
    const double s = block_sum(acc, smem);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// K lgamma(eta W) - sum_k lgamma(eta W + n_k); single block.
__global__ void topic_norm_sum_kernel(const int32_t *__restrict__ n_k, int32_t K, double etaW,
                                      double *__restrict__ out) {
    __shared__ double smem[32];
    double acc = 0.0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) acc += lgamma(etaW) - lgamma(etaW + (double)n_k[k]);
    const double s = block_sum(acc, smem);
    if (threadIdx.x == 0) *out = s;
}

__global__ void final_sum_kernel(const double *__restrict__ partials, int32_t n, double *__restrict__ out) {
    __shared__ double smem[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
    const double s = block_sum(acc, smem);
    if (threadIdx.x == 0) *out = s;
}

// ---- exports ---------------------------------------------------------------------------------
// src is rows x Kp (int32); dst is K x rows (int32): the `lda` package's nzw layout.
__global__ void transpose_counts_kernel(const int32_t *__restrict__ src, int64_t rows, int32_t K, int32_t Kp,
                                        int32_t *__restrict__ dst) {
    __shared__ int32_t tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int k0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = r0 + j;
        const int k = k0 + threadIdx.x;
        tile[j][threadIdx.x] = (r < rows && k < K) ? src[r * Kp + k] : 0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int k = k0 + j;
        const int64_t r = r0 + threadIdx.x;
        if (k < K && r < rows) dst[(int64_t)k * rows + r] = tile[threadIdx.x][j];
    }
}

// src rows x Kp int32 -> dst rows x K int32 (drop the padding)
__global__ void unpad_counts_kernel(const int32_t *__restrict__ src, int64_t rows, int32_t K, int32_t Kp,
                                    int32_t *__restrict__ dst) {
    int64_t n = rows * K;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t r = i / K;
        const int k = (int)(i - r * K);
        dst[i] = src[r * Kp + k];
    }
}

// topic_word[k, w] = (n_wk[w,k] + eta) / (n_k[k] + eta W)
// transposed != 0 : dst is K x W (lda's topic_word_);  == 0 : dst is W x K (topic_region frame)
__global__ void topic_word_kernel(const int32_t *__restrict__ n_wk, const int32_t *__restrict__ n_k,
                                  int64_t W, int32_t K, int32_t Kp, double eta, int transposed,
                                  double *__restrict__ dst) {
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int k0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = r0 + j;
        const int k = k0 + threadIdx.x;
        double v = 0.0;
        if (r < W && k < K) v = ((double)n_wk[r * Kp + k] + eta) / ((double)n_k[k] + eta * (double)W);
        if (transposed) tile[j][threadIdx.x] = v;
        else if (r < W && k < K) dst[r * K + k] = v;
    }
    if (!transposed) return;
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int k = k0 + j;
        const int64_t r = r0 + threadIdx.x;
        if (k < K && r < W) dst[(int64_t)k * W + r] = tile[threadIdx.x][j];
    }
}

// doc_topic[d, k] = (n_dk[d,k] + alpha) / (nd_d + alpha K)
// transposed != 0 : dst is K x D (cell_topic frame);  == 0 : dst is D x K (lda's doc_topic_)
__global__ void doc_topic_kernel(const int32_t *__restrict__ n_dk, const int64_t *__restrict__ cell_ptr,
                                 int64_t D, int32_t K, int32_t Kp, double alpha, int transposed,
                                 double *__restrict__ dst) {
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int k0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = r0 + j;
        const int k = k0 + threadIdx.x;
        double v = 0.0;
        if (r < D && k < K) {
            const double nd = (double)(cell_ptr[r + 1] - cell_ptr[r]);
            v = ((double)n_dk[r * Kp + k] + alpha) / (nd + alpha * (double)K);
        }
        if (transposed) tile[j][threadIdx.x] = v;
        else if (r < D && k < K) dst[r * K + k] = v;
    }
    if (!transposed) return;
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int k = k0 + j;
        const int64_t r = r0 + threadIdx.x;
        if (k < K && r < D) dst[(int64_t)k * D + r] = tile[threadIdx.x][j];
    }
}

}  // namespace cgs
