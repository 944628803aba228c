// common.cuh -- error plumbing, small device helpers, Philox4x32-10.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

namespace cgs {

extern thread_local std::string g_last_error;

inline int fail(const char *file, int line, const std::string &msg) {
    char buf[64];
    snprintf(buf, sizeof buf, " [%s:%d]", file, line);
    g_last_error = msg + buf;
    return 1;
}

#define CGS_FAIL(msg) return ::cgs::fail(__FILE__, __LINE__, (msg))
#define CGS_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return ::cgs::fail(__FILE__, __LINE__,                                              \
                               std::string(#expr) + ": " + cudaGetErrorString(_e));             \
    } while (0)
#define CGS_TRY(expr)                                                                           \
    do {                                                                                        \
        int _rc = (expr);                                                                       \
        if (_rc != 0) return _rc;                                                               \
    } while (0)

constexpr int kNumSMsFallback = 148;

// ---- Philox4x32-10 (Salmon et al., SC'11), counter = (c0, c1, c2, c3), key = (k0, k1) ------
struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2,
                                                       uint32_t &c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                           uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// uniform in [0, 1): 24 random bits
__host__ __device__ __forceinline__ float u01_from_bits(uint32_t b) {
    return (float)(b >> 8) * (1.0f / 16777216.0f);
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace cgs
