// impute_kernels.cuh -- imputed accessibility chunk product (replaces the fp32 BLAS product
// and its epilogue in diff_features.py:443-486):
//     chunk = topic_region[r0:r1] @ cell_topic ; chunk *= float32(scale) ; chunk.astype(int32)
//     keep  = rows with any non-zero
// Round-1 kernel: fp32 CUDA-core tiles with the fused epilogue (scale, truncate toward zero,
// row-non-zero flags).  The tcgen05 (3xTF32) kernel replaces the main loop; the epilogue and
// the ABI stay.
#pragma once
#include "common.cuh"

namespace cgs {

constexpr int kImpTile = 64;   // output tile is 64 regions x 64 cells
constexpr int kImpK = 16;      // topics per shared-memory stage

template <bool AS_INT>
__global__ void __launch_bounds__(256) impute_simt_kernel(const float *__restrict__ A,  // R x K
                                                          const float *__restrict__ B,  // K x C
                                                          int32_t R, int32_t K, int32_t C, float scale,
                                                          void *__restrict__ out, uint8_t *__restrict__ row_nonzero) {
    __shared__ float As[kImpK][kImpTile + 1];
    __shared__ float Bs[kImpK][kImpTile];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 x 4 outputs each
    const int64_t r0 = (int64_t)blockIdx.y * kImpTile;
    const int64_t c0 = (int64_t)blockIdx.x * kImpTile;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < K; k0 += kImpK) {
        for (int e = threadIdx.x; e < kImpTile * kImpK; e += 256) {
            const int rr = e / kImpK, kk = e % kImpK;
            const int64_t r = r0 + rr;
            As[kk][rr] = (r < R && k0 + kk < K) ? A[r * K + k0 + kk] : 0.0f;
        }
        for (int e = threadIdx.x; e < kImpTile * kImpK; e += 256) {
            const int kk = e / kImpTile, cc = e % kImpTile;
            const int64_t c = c0 + cc;
            Bs[kk][cc] = (c < C && k0 + kk < K) ? B[(int64_t)(k0 + kk) * C + c] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kImpK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = r0 + ty * 4 + i;
        if (r >= R) continue;
        bool any = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t c = c0 + tx * 4 + j;
            if (c >= C) continue;
            if (AS_INT) {
                const int v = __float2int_rz(acc[i][j] * scale);
                reinterpret_cast<int32_t *>(out)[r * C + c] = v;
                any |= (v != 0);
            } else {
                const float v = (scale != 1.0f) ? acc[i][j] * scale : acc[i][j];
                reinterpret_cast<float *>(out)[r * C + c] = v;
                any |= (v != 0.0f);
            }
        }
        if (any && row_nonzero) row_nonzero[r] = 1;
    }
}

static int impute_launch(int device, const float *d_A, const float *d_B, int32_t R, int32_t K, int32_t C, float scale,
                         int as_int32, void *d_out, uint8_t *d_row_nonzero, cudaStream_t stream) {
    (void)device;
    if (d_row_nonzero) CGS_CUDA(cudaMemsetAsync(d_row_nonzero, 0, (size_t)R, stream));
    dim3 grid((unsigned)ceil_div(C, kImpTile), (unsigned)ceil_div(R, kImpTile));
    if (as_int32)
        impute_simt_kernel<true><<<grid, 256, 0, stream>>>(d_A, d_B, R, K, C, scale, d_out, d_row_nonzero);
    else
        impute_simt_kernel<false><<<grid, 256, 0, stream>>>(d_A, d_B, R, K, C, scale, d_out, d_row_nonzero);
    CGS_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cgs
