// normalize_kernels.cuh -- normalize_scores (diff_features.py:511-574, inner function :540-548):
//     normalized = X / (colsum(X) / scale_factor) ; log1p in place
// on an imputed-accessibility matrix X (regions x cells, row-major).  Two HBM-bound passes:
//   colsum_kernel        reads X once (4 B per element), per-cell totals: int64 for int32 input (exact, as
//                        NumPy's int64 accumulation), fp64 for floating input
//   normalize_log1p_kernel  reads X again, writes log1p(x / (colsum / scale)) -- fp64 arithmetic in the
//                        reference's operation order, so int32 and fp64 inputs reproduce NumPy to the last
//                        ulps of log1p; fp32 input is rounded to fp32 at the end like NumPy's fp32 result
#pragma once
#include <type_traits>

#include "common.cuh"

namespace cgs {

constexpr int kNormCols = 256;   // cells per CTA (one per thread: coalesced rows)
constexpr int kNormSlab = 512;   // regions per CTA in the column-sum pass

template <typename T> struct NormAcc { using type = double; };
template <> struct NormAcc<int32_t> { using type = long long; };

// colsum[c] += sum over this CTA's row slab; colsum must be zeroed first
template <typename T>
__global__ void __launch_bounds__(kNormCols) colsum_kernel(const T *__restrict__ x, int64_t R, int64_t C,
                                                           typename NormAcc<T>::type *__restrict__ colsum) {
    using A = typename NormAcc<T>::type;
    const int64_t c = (int64_t)blockIdx.x * kNormCols + threadIdx.x;
    if (c >= C) return;
    const int64_t r0 = (int64_t)blockIdx.y * kNormSlab;
    const int64_t r1 = r0 + kNormSlab < R ? r0 + kNormSlab : R;
    A acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    int64_t r = r0;
    for (; r + 3 < r1; r += 4) {  // four independent loads in flight
        acc0 += (A)x[r * C + c];
        acc1 += (A)x[(r + 1) * C + c];
        acc2 += (A)x[(r + 2) * C + c];
        acc3 += (A)x[(r + 3) * C + c];
    }
    for (; r < r1; ++r) acc0 += (A)x[r * C + c];
    const A s = (acc0 + acc1) + (acc2 + acc3);
    if (sizeof(A) == 8 && std::is_same<A, long long>::value)
        atomicAdd(reinterpret_cast<unsigned long long *>(colsum) + c, (unsigned long long)s);
    else
        atomicAdd(reinterpret_cast<double *>(colsum) + c, (double)s);
}

// out = log1p(x / (colsum / scale)); OUT = double (int32 / fp64 input) or float (fp32 input)
template <typename T, typename OUT>
__global__ void __launch_bounds__(kNormCols) normalize_log1p_kernel(const T *__restrict__ x, int64_t R, int64_t C,
                                                                    const typename NormAcc<T>::type *__restrict__ colsum,
                                                                    double scale, OUT *__restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * kNormCols + threadIdx.x;
    if (c >= C) return;
    const double d = (double)colsum[c] / scale;  // np.sum(X, axis=0) / scale_factor
    const int64_t r0 = (int64_t)blockIdx.y * kNormSlab;
    const int64_t r1 = r0 + kNormSlab < R ? r0 + kNormSlab : R;
    for (int64_t r = r0; r < r1; ++r) {
        const double q = (double)x[r * C + c] / d;
        out[r * C + c] = (OUT)log1p(q);
    }
}

template <typename T, typename OUT>
static int normalize_launch(const T *d_x, int64_t R, int64_t C, double scale, void *d_colsum, OUT *d_out,
                            cudaStream_t stream) {
    using A = typename NormAcc<T>::type;
    CGS_CUDA(cudaMemsetAsync(d_colsum, 0, sizeof(A) * (size_t)C, stream));
    dim3 grid((unsigned)ceil_div(C, kNormCols), (unsigned)ceil_div(R, kNormSlab));
    colsum_kernel<T><<<grid, kNormCols, 0, stream>>>(d_x, R, C, reinterpret_cast<A *>(d_colsum));
    CGS_CUDA(cudaGetLastError());
    normalize_log1p_kernel<T, OUT><<<grid, kNormCols, 0, stream>>>(d_x, R, C, reinterpret_cast<const A *>(d_colsum), scale,
                                                                  d_out);
    CGS_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cgs
