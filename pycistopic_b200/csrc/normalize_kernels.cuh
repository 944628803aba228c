// normalize_kernels.cuh -- normalize_scores (diff_features.py:511-574, inner function :540-548):
//     normalized = X / (colsum(X) / scale_factor) ; log1p in place
// on an imputed-accessibility matrix X (regions x cells, row-major).  Two HBM-bound passes:
//   colsum_kernel        reads X once (4 B per element), per-cell totals: int64 for int32 input (exact, as
//                        NumPy's int64 accumulation), fp64 for floating input
//   normalize_log1p_kernel  reads X again, writes log1p(x / (colsum / scale)) -- fp64 arithmetic in the
//                        reference's operation order, so int32 and fp64 inputs reproduce NumPy to the last
//                        ulps of log1p; fp32 input is rounded to fp32 at the end like NumPy's fp32 result
#pragma once
#include <algorithm>
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
    // A zero numerator sends the fp64 division down its slow path (a subroutine of ~100 instructions), and a truncated
    // imputation is mostly zeros: log1p(0 / d) is that same zero exactly for d > 0, so those values skip both.
    const bool d_pos = d > 0.0;
    for (int64_t r = r0; r < r1; ++r) {
        const T v = x[r * C + c];
        if (v == T(0) && d_pos) {
            out[r * C + c] = (OUT)v;  // keeps the sign of a floating-point zero, as log1p(-0 / d) does
            continue;
        }
        const double q = (double)v / d;
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

// ---- per-region mean and variance over the cells: the reduction in front of find_highly_variable_features
// (diff_features.py:636-639: np.mean(mat, axis=1, dtype=np.float32), np.var(mat, axis=1, dtype=np.float32)) ----
// One CTA per row, two passes over it (the second one hits L2): mean, then the mean of the squared deviations,
// NumPy's own two-pass definition.  NumPy's casts are kept -- values enter the mean as float32; deviations are
// taken in the promoted type (float32 for float32 input, float64 otherwise) and enter the second sum as float32
// -- but both sums are accumulated in fp64 instead of NumPy's pairwise float32, so the result is the correctly
// rounded float32 of what NumPy approximates (they differ by a few float32 ulps).  HBM-bound: 4 or 8 B per value.
constexpr int kRowStatThreads = 256;
constexpr int kRowStatIlp = 8;

__device__ __forceinline__ double rowstat_block_sum(double v, double *red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();  // red may still be read from the previous call
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kRowStatThreads / 32; ++w) t += red[w];
    return t;
}

template <typename T>
__global__ void __launch_bounds__(kRowStatThreads)
row_mean_var_kernel(const T *__restrict__ x, int64_t ld, int64_t n_rows, int64_t n_cols, float *__restrict__ mean,
                    float *__restrict__ var) {
    __shared__ double red[kRowStatThreads / 32];
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const T *row = x + r * ld;
        double acc = 0.0;
        for (int64_t c0 = threadIdx.x; c0 < n_cols; c0 += (int64_t)kRowStatIlp * kRowStatThreads) {
            T v[kRowStatIlp];
#pragma unroll
            for (int u = 0; u < kRowStatIlp; ++u) {
                const int64_t c = c0 + (int64_t)u * kRowStatThreads;
                v[u] = c < n_cols ? row[c] : T(0);
            }
#pragma unroll
            for (int u = 0; u < kRowStatIlp; ++u) acc += (double)(float)v[u];
        }
        const float m = (float)(rowstat_block_sum(acc, red) / (double)n_cols);
        acc = 0.0;
        for (int64_t c0 = threadIdx.x; c0 < n_cols; c0 += (int64_t)kRowStatIlp * kRowStatThreads) {
            T v[kRowStatIlp];
            bool ok[kRowStatIlp];
#pragma unroll
            for (int u = 0; u < kRowStatIlp; ++u) {
                const int64_t c = c0 + (int64_t)u * kRowStatThreads;
                ok[u] = c < n_cols;
                v[u] = ok[u] ? row[c] : T(0);
            }
#pragma unroll
            for (int u = 0; u < kRowStatIlp; ++u) {
                float sq;
                if (std::is_same<T, float>::value) {
                    const float d = __fsub_rn((float)v[u], m);
                    sq = __fmul_rn(d, d);
                } else {
                    const double d = (double)v[u] - (double)m;
                    sq = (float)(d * d);
                }
                if (ok[u]) acc += (double)sq;
            }
        }
        const float s2 = (float)(rowstat_block_sum(acc, red) / (double)n_cols);
        if (threadIdx.x == 0) { mean[r] = m; var[r] = s2; }
    }
}

template <typename T>
static int row_mean_var_launch(const T *d_x, int64_t ld, int64_t n_rows, int64_t n_cols, float *d_mean, float *d_var,
                               int num_sms, cudaStream_t stream) {
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(n_rows, (int64_t)num_sms * 8));
    row_mean_var_kernel<T><<<blocks, kRowStatThreads, 0, stream>>>(d_x, ld, n_rows, n_cols, d_mean, d_var);
    CGS_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cgs
