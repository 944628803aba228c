// impute_kernels.cuh -- imputed accessibility chunk product (replaces the fp32 BLAS product
// and its epilogue in diff_features.py:443-486):
//     chunk = topic_region[r0:r1] @ cell_topic ; chunk *= float32(scale) ; chunk.astype(int32)
//     keep  = rows with any non-zero
//
// The product runs on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in
// tensor memory).  TF32 keeps 11 significant bits, the reference computes in fp32, so every
// operand is split as x = hi + lo with hi = x truncated to TF32 (exact) and lo = x - hi (exact in
// fp32), and the tile accumulates hi*hi + lo*hi + hi*lo: the dropped lo*lo term and the TF32
// truncation of lo are both ~2^-21 relative ("3xTF32").  Operand preparation
// (impute_split_rows_kernel / impute_transpose_split_kernel) writes the four K-major arrays
//     A_hi, A_lo : regions x Kp      B_hi, B_lo : cells x Kp      Kp = K rounded up to 32
// so that every operand tile is a plain 2-D TMA box with the 128-byte swizzle the UMMA shared
// memory descriptors expect.
//
// Kernels: impute_umma_pair_kernel (default: CTA pairs, tcgen05.mma.cta_group::2, 256 x 256 tile per pair,
// 3-stage 64 KB ring per CTA) and impute_umma_kernel (one CTA, 128 x 256 tile, 2-stage 96 KB ring), both
// persistent with one CTA per SM.  Roles, single-CTA numbers:
//   warp 0   TMA producer: per 32-wide K block four boxes (A_hi, A_lo 128 x 32; B_hi, B_lo 256 x 32)
//            into a 2-stage shared-memory ring, completion on an mbarrier (expect_tx)
//   warp 1   one elected thread issues 12 tcgen05.mma per K block (4 k-steps x 3 operand pairs)
//            into one of two 256-column TMEM accumulators; tcgen05.commit releases the stage
//            and, after the last K block, hands the accumulator to the epilogue
//   warp 2   allocates / frees the 512 TMEM columns
//   warps 4-11 epilogue (two warps per TMEM lane quarter, 128 columns each): tcgen05.ld 32 lanes x
//            16 columns (next block in flight) -> * scale -> truncate toward zero (FP32-pipe add in
//            round-toward-zero mode; cvt.rzi only for values outside [0, 2^23)) -> swizzled 32 x 16
//            staging tile (two per warp) -> 16-byte coalesced stores (or one TMA store per block;
//            4-byte stores when n_cells is not a multiple of 4 and rows are not 16-byte aligned).
//            Row-non-zero flags come from the same registers
// Tiles are walked in groups of 8 neighbouring cell blocks, cell block fastest inside a group, so that
// the CTAs running at the same time share a handful of cell-operand tiles in L2 while the (small) region
// operand stays L2 resident.
//
// The CUDA-core kernel below it (impute_simt_kernel) is the cross-check used by the GPU tests
// (CGS_IMPUTE_SIMT=1); it is not a fallback: the tensor-core path handles every shape.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace cgs {

// ---------------------------------------------------------------------------------------------
// CUDA-core cross-check kernel
// ---------------------------------------------------------------------------------------------
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

// ---------------------------------------------------------------------------------------------
// Operand preparation: hi / lo split into K-major, Kp-padded arrays
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// A (R x K, row-major) -> A_hi, A_lo (R x Kp)
__global__ void impute_split_rows_kernel(const float *__restrict__ A, int64_t R, int32_t K, int32_t Kp,
                                         float *__restrict__ hi, float *__restrict__ lo) {
    const int64_t n = R * Kp;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const int64_t r = e / Kp;
        const int k = (int)(e - r * Kp);
        const float x = k < K ? A[r * K + k] : 0.0f;
        const float h = tf32_hi(x);
        hi[e] = h;
        lo[e] = x - h;
    }
}

// B (K x C, row-major) -> B_hi, B_lo (C x Kp): 32 x 32 shared-memory transpose
__global__ void impute_transpose_split_kernel(const float *__restrict__ B, int32_t K, int64_t C, int32_t Kp,
                                              float *__restrict__ hi, float *__restrict__ lo) {
    __shared__ float t[32][33];
    const int64_t c0 = (int64_t)blockIdx.x * 32;
    const int k0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int k = k0 + i;
        const int64_t c = c0 + threadIdx.x;
        t[i][threadIdx.x] = (k < K && c < C) ? B[(int64_t)k * C + c] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t c = c0 + i;
        const int k = k0 + threadIdx.x;
        if (c < C && k < Kp) {
            const float x = t[threadIdx.x][i];
            const float h = tf32_hi(x);
            hi[c * Kp + k] = h;
            lo[c * Kp + k] = x - h;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Tensor-core kernel
// ---------------------------------------------------------------------------------------------
namespace imp {
constexpr int BM = 128;     // regions per tile = UMMA M
constexpr int BN = 256;     // cells per tile   = UMMA N
constexpr int BK = 32;      // fp32 per 128-byte swizzle row
constexpr int UK = 8;       // K of one kind::tf32 instruction (32 bytes)
constexpr int STAGES = 2;
constexpr int A_BYTES = BM * BK * 4;                       // 16 KB
constexpr int B_BYTES = BN * BK * 4;                       // 32 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;     // A_hi | A_lo | B_hi | B_lo = 96 KB
constexpr int EPI_WARPS = 8;                               // two per TMEM lane quarter
constexpr int EPI_TILE_BYTES = 2 * 32 * 16 * 4;            // two 32 x 16 staging tiles (2 KB each) per epilogue warp
constexpr int EPI_BYTES = EPI_WARPS * EPI_TILE_BYTES;
constexpr int BAR_BYTES = 128;
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
constexpr int THREADS = 128 + 32 * EPI_WARPS;
constexpr int TMEM_COLS = 512;  // two BN-column accumulators

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
                 "r"(c1)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read_but_one() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, one CTA
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused with swizzle), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell), bits [46,48)
    d |= (uint64_t)2 << 61;                         // layout type SWIZZLE_128B, bits [61,64)
    return d;
}
// kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4)                 // c_format = F32
           | (2u << 7)               // a_format = TF32
           | (2u << 10)              // b_format = TF32
           | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld_32x32_issue(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_32x16_issue(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    tmem_ld_32x32_issue(taddr, v);
    tmem_ld_wait();
}

// Tile walk: groups of kTileGroupN neighbouring cell blocks; inside a group the region blocks advance slowest
// and the cell blocks fastest.  The CTAs that run at the same time then (a) share a handful of cell-operand
// tiles in L2 and (b) write kTileGroupN KB of every output row back to back -- a column-block-at-a-time walk
// writes isolated 1 KB pieces 800 KB apart and tops out at 4.5 TB/s of the 7.5 TB/s a plain fill reaches.
#ifndef CGS_IMPUTE_GROUP_N
#define CGS_IMPUTE_GROUP_N 8
#endif
constexpr int kTileGroupN = CGS_IMPUTE_GROUP_N;
__device__ __forceinline__ void tile_coords(int64_t t, int n_m, int n_n, int &m_idx, int &n_idx) {
    const int64_t per_group = (int64_t)n_m * kTileGroupN;
    const int g = (int)(t / per_group);
    const int r = (int)(t - (int64_t)g * per_group);
    const int gn = min(kTileGroupN, n_n - g * kTileGroupN);
    m_idx = r / gn;
    n_idx = g * kTileGroupN + r - m_idx * gn;
}

// ---- CTA-pair (cta_group::2) variants --------------------------------------------------------
constexpr int STAGES2 = 3;
constexpr int B_HALF_BYTES = (BN / 2) * BK * 4;                  // each CTA of the pair stages half of the cell tile
constexpr int STAGE2_BYTES = 2 * A_BYTES + 2 * B_HALF_BYTES;     // 64 KB
constexpr int SMEM2_BYTES = 1024 + STAGES2 * STAGE2_BYTES + EPI_BYTES + BAR_BYTES;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;                   // shared::cluster address of the same offset in CTA 0

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 rem;\n\t"
        "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [rem];\n\t}" ::"r"(bar),
        "r"(cta)
        : "memory");
}
// TMA load issued by either CTA of a pair; the bytes are accounted on CTA 0's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(dst),
        "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
}  // namespace imp

// Epilogue of one 32 x 16 accumulator block held in v (lane = row): scale, convert, stage, store.
// `reload_taddr` is the block's TMEM address (the exact-conversion path reads it again); `buf` is one of
// the warp's two 2 KB staging tiles (32 rows x 64 bytes, 64-byte swizzle).
// STORE: 0 = 4-byte coalesced stores (rows not 16-byte aligned), 1 = TMA store (the warp's two staging tiles
// alternate), 2 = 16-byte coalesced stores, 4 = TMA store that always uses the same staging tile
template <bool AS_INT, int STORE>
__device__ __forceinline__ void impute_epi_chunk(uint32_t (&v)[16], uint32_t reload_taddr, float scale, bool &any,
                                                 uint32_t buf, int lane, const CUtensorMap *tm_out, int col0, int row0,
                                                 int R, int C, void *out) {
    using namespace imp;
    if (AS_INT) {
        // cvt.rzi.s32.f32 runs on the special-function pipe; for 0 <= y < 2^23 (every imputed count: a
        // probability times 10^6) truncation is an FP32-pipe add in round-toward-zero mode:
        // y +rz 2^23 = 2^23 + trunc(y), whose low 23 bits are the integer.  Anything else (negative, huge,
        // NaN -- visible as an unsigned bit pattern >= 0x4B000000) takes the exact cvt path.
        uint32_t maxbits = 0, orbits = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float y = __uint_as_float(v[j]) * scale;
            maxbits = max(maxbits, __float_as_uint(y));
            const uint32_t iv = __float_as_uint(__fadd_rz(y, 8388608.0f)) - 0x4B000000u;
            orbits |= iv;
            v[j] = iv;
        }
        if (__any_sync(0xffffffffu, maxbits >= 0x4B000000u)) {
            tmem_ld_32x16_issue(reload_taddr, v);
            tmem_ld_wait();
            orbits = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t iv = (uint32_t)__float2int_rz(__uint_as_float(v[j]) * scale);
                orbits |= iv;
                v[j] = iv;
            }
        }
        any |= (orbits != 0);
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float x = __uint_as_float(v[j]);
            const float fv = (scale != 1.0f) ? x * scale : x;
            any |= (fv != 0.0f);
            v[j] = __float_as_uint(fv);
        }
    }
    if (STORE == 1 || STORE == 4) {
        // the store that last used THIS staging tile must have finished reading it: two stores ago when the
        // tiles alternate, the previous one otherwise
        if (lane == 0) {
            if (STORE == 1) tma_store_wait_read_but_one();
            else tma_store_wait_read();
        }
        __syncwarp();
        // row `lane`, four 16-byte pieces, 64-byte swizzle (matches the store's tensor map)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t addr = buf + (uint32_t)lane * 64u + (uint32_t)((c ^ ((lane >> 1) & 3)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[4 * c]), "r"(v[4 * c + 1]),
                         "r"(v[4 * c + 2]), "r"(v[4 * c + 3])
                         : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) tma_store_2d(tm_out, buf, col0, row0);
    } else if (STORE == 2) {
        // same swizzled staging tile, read back as 16-byte pieces: a warp instruction stores 8 rows x 64 B
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t addr = buf + (uint32_t)lane * 64u + (uint32_t)((c ^ ((lane >> 1) & 3)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[4 * c]), "r"(v[4 * c + 1]),
                         "r"(v[4 * c + 2]), "r"(v[4 * c + 3])
                         : "memory");
        }
        __syncwarp();
        const int piece = lane & 3;
        const int col = col0 + piece * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = i * 8 + (lane >> 2);
            const int row = row0 + r;
            uint4 o;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                         : "r"(buf + (uint32_t)r * 64u + (uint32_t)((piece ^ ((r >> 1) & 3)) << 4)));
            if (row < R && col < C)  // C % 4 == 0: a 16-byte piece is inside the row or outside
                *reinterpret_cast<uint4 *>(reinterpret_cast<uint32_t *>(out) + (int64_t)row * C + col) = o;
        }
    } else {
        // 32 x 16 transpose through the 2 KB staging tile (column index rotated by the row), then two rows per
        // warp instruction: 64 contiguous bytes per row
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(buf + (uint32_t)(lane * 16 + ((j + lane) & 15)) * 4u), "r"(v[j])
                         : "memory");
        }
        __syncwarp();
        const int col = col0 + (lane & 15);
        if (col < C) {
#pragma unroll 8
            for (int r2 = 0; r2 < 16; ++r2) {
                const int r = r2 * 2 + (lane >> 4);
                const int row = row0 + r;
                if (row < R) {
                    uint32_t o;
                    asm volatile("ld.shared.b32 %0, [%1];"
                                 : "=r"(o)
                                 : "r"(buf + (uint32_t)(r * 16 + (((lane & 15) + r) & 15)) * 4u));
                    reinterpret_cast<uint32_t *>(out)[(int64_t)row * C + col] = o;
                }
            }
        }
        __syncwarp();
    }
}

// STORE = 5 ("wide"): the same epilogue in 32 x 32 blocks.  tcgen05.ld.32x32b.x32 leaves every lane 32 consecutive
// columns of its row -- a full 128-byte line -- so after the staging tile (32 rows x 128 bytes, 16-byte pieces
// XOR-swizzled by the row, conflict-free both ways) a warp instruction stores FOUR complete lines where the 32 x 16
// blocks store eight half lines: half the wavefronts through the L1 / LSU data path per byte written, which is the unit
// round 1's ncu showed busiest in this epilogue (77 % at K = 8).  One 4 KB staging tile per warp; the TMEM load of
// block i + 1 is in flight while block i is converted and stored.
template <bool AS_INT>
__device__ __forceinline__ void impute_epi_chunk_wide(uint32_t (&v)[32], uint32_t reload_taddr, float scale, bool &any,
                                                      uint32_t buf, int lane, int col0, int row0, int R, int C, void *out) {
    using namespace imp;
    if (AS_INT) {
        uint32_t maxbits = 0, orbits = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float y = __uint_as_float(v[j]) * scale;
            maxbits = max(maxbits, __float_as_uint(y));
            const uint32_t iv = __float_as_uint(__fadd_rz(y, 8388608.0f)) - 0x4B000000u;
            orbits |= iv;
            v[j] = iv;
        }
        if (__any_sync(0xffffffffu, maxbits >= 0x4B000000u)) {  // negative, huge or NaN somewhere: the exact conversion
            tmem_ld_32x32(reload_taddr, v);
            orbits = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const uint32_t iv = (uint32_t)__float2int_rz(__uint_as_float(v[j]) * scale);
                orbits |= iv;
                v[j] = iv;
            }
        }
        any |= (orbits != 0);
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(v[j]);
            const float fv = (scale != 1.0f) ? x * scale : x;
            any |= (fv != 0.0f);
            v[j] = __float_as_uint(fv);
        }
    }
    __syncwarp();  // the previous block's reads of the staging tile are done
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint32_t addr = buf + (uint32_t)lane * 128u + (uint32_t)((c ^ (lane & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[4 * c]), "r"(v[4 * c + 1]),
                     "r"(v[4 * c + 2]), "r"(v[4 * c + 3])
                     : "memory");
    }
    __syncwarp();
    const int piece = lane & 7;
    const int col = col0 + piece * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + (lane >> 3);
        const int row = row0 + r;
        uint4 o;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                     : "r"(buf + (uint32_t)r * 128u + (uint32_t)((piece ^ (r & 7)) << 4)));
        if (row < R && col < C)  // C % 4 == 0: a 16-byte piece is inside the row or outside
            *reinterpret_cast<uint4 *>(reinterpret_cast<uint32_t *>(out) + (int64_t)row * C + col) = o;
    }
}

template <bool AS_INT>
__device__ __forceinline__ void impute_epi_tile_wide(uint32_t taddr0, int n_blocks16, float scale, bool &any,
                                                     uint32_t stage_buf, int lane, int col0, int row0, int R, int C,
                                                     void *out) {
    using namespace imp;
    const int n_blocks = (n_blocks16 + 1) >> 1;  // 32-column blocks (columns past C are masked piece by piece)
    uint32_t va[32], vb[32];
    if (n_blocks > 0) tmem_ld_32x32_issue(taddr0, va);
#pragma unroll
    for (int ch = 0; ch < BN / 64; ++ch) {
        if (ch < n_blocks) {  // warp-uniform
            tmem_ld_wait();
            if (ch & 1) {
                if (ch + 1 < n_blocks) tmem_ld_32x32_issue(taddr0 + (uint32_t)(ch + 1) * 32u, va);
                impute_epi_chunk_wide<AS_INT>(vb, taddr0 + (uint32_t)ch * 32u, scale, any, stage_buf, lane, col0 + ch * 32,
                                              row0, R, C, out);
            } else {
                if (ch + 1 < n_blocks) tmem_ld_32x32_issue(taddr0 + (uint32_t)(ch + 1) * 32u, vb);
                impute_epi_chunk_wide<AS_INT>(va, taddr0 + (uint32_t)ch * 32u, scale, any, stage_buf, lane, col0 + ch * 32,
                                              row0, R, C, out);
            }
        }
    }
}

// The 32 rows x 128 columns an epilogue warp owns in a tile: eight 32 x 16 blocks.  The TMEM load of block
// i + 1 is in flight while block i is converted, and the two staging tiles alternate so that a TMA store
// reads one while the next block is written into the other.
// STORE as above, or 3 = even blocks through 16-byte stores (staging tile 0), odd blocks through TMA stores
// (staging tile 1): the load/store data path and the TMA engine each carry half of the output
template <bool AS_INT, int STORE>
__device__ __forceinline__ void impute_epi_tile(uint32_t taddr0, int n_blocks, float scale, bool &any, uint32_t stage_buf,
                                                int lane, const CUtensorMap *tm_out, int col0, int row0, int R, int C,
                                                void *out) {
    using namespace imp;
    if constexpr (STORE == 5) {
        impute_epi_tile_wide<AS_INT>(taddr0, n_blocks, scale, any, stage_buf, lane, col0, row0, R, C, out);
        return;
    }
    uint32_t va[16], vb[16];
    if (n_blocks > 0) tmem_ld_32x16_issue(taddr0, va);
#pragma unroll
    for (int ch = 0; ch < BN / 32; ++ch) {
        if (ch < n_blocks) {  // warp-uniform
            tmem_ld_wait();
            const uint32_t buf = stage_buf + (uint32_t)(ch & 1) * (EPI_TILE_BYTES / 2);
            if (ch & 1) {
                if (ch + 1 < n_blocks) tmem_ld_32x16_issue(taddr0 + (uint32_t)(ch + 1) * 16u, va);
                impute_epi_chunk<AS_INT, (STORE == 3 ? 4 : STORE)>(vb, taddr0 + (uint32_t)ch * 16u, scale, any, buf, lane,
                                                                   tm_out, col0 + ch * 16, row0, R, C, out);
            } else {
                if (ch + 1 < n_blocks) tmem_ld_32x16_issue(taddr0 + (uint32_t)(ch + 1) * 16u, vb);
                impute_epi_chunk<AS_INT, (STORE == 3 ? 2 : STORE)>(va, taddr0 + (uint32_t)ch * 16u, scale, any, buf, lane,
                                                                   tm_out, col0 + ch * 16, row0, R, C, out);
            }
        }
    }
}

template <bool AS_INT, int STORE>
__global__ void __launch_bounds__(imp::THREADS, 1)
impute_umma_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                   const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                   const __grid_constant__ CUtensorMap tm_out, int32_t R, int32_t C, int32_t n_ksteps, float scale,
                   void *__restrict__ out, uint8_t *__restrict__ row_nonzero) {
    using namespace imp;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // swizzle-128B tiles need 1024-byte alignment
    uint8_t *smem = smem_raw + (base - raw);
    const uint32_t s_epi = base + STAGES * STAGE_BYTES;
    const uint32_t s_bar = s_epi + EPI_BYTES;
    // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base address
    const uint32_t bar_full = s_bar, bar_empty = s_bar + 8 * STAGES, bar_tfull = s_bar + 16 * STAGES,
                   bar_tempty = s_bar + 16 * STAGES + 16;
    const uint32_t tmem_slot_addr = s_bar + 16 * STAGES + 32;
    volatile uint32_t *tmem_slot =
        reinterpret_cast<volatile uint32_t *>(smem + STAGES * STAGE_BYTES + EPI_BYTES + 16 * STAGES + 32);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (warp == 1 && lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, EPI_WARPS);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot_addr),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_m = (R + BM - 1) / BM;
    const int n_n = (C + BN - 1) / BN;
    const int64_t n_tiles = (int64_t)n_m * n_n;
    const int n_kblocks = (n_ksteps + BK / UK - 1) / (BK / UK);

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                int mi, ni;
                tile_coords(t, n_m, n_n, mi, ni);
                const int m0 = mi * BM, n0 = ni * BN;
                for (int kb = 0; kb < n_kblocks; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                    const uint32_t full = bar_full + 8 * stage;
                    const uint32_t st = base + stage * STAGE_BYTES;
                    mbar_expect_tx(full, STAGE_BYTES);
                    tma_load_2d(st, &tm_a_hi, full, kb * BK, m0);
                    tma_load_2d(st + A_BYTES, &tm_a_lo, full, kb * BK, m0);
                    tma_load_2d(st + 2 * A_BYTES, &tm_b_hi, full, kb * BK, n0);
                    tma_load_2d(st + 2 * A_BYTES + B_BYTES, &tm_b_lo, full, kb * BK, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < n_kblocks; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t st = base + stage * STAGE_BYTES;
                    const int ks = min(BK / UK, n_ksteps - kb * (BK / UK));  // the zero padding of K is not multiplied
                    for (int k = 0; k < ks; ++k) {
                        const uint32_t koff = (uint32_t)(k * UK * 4);
                        const uint64_t a_hi = umma_desc_k_sw128(st + koff);
                        const uint64_t a_lo = umma_desc_k_sw128(st + A_BYTES + koff);
                        const uint64_t b_hi = umma_desc_k_sw128(st + 2 * A_BYTES + koff);
                        const uint64_t b_lo = umma_desc_k_sw128(st + 2 * A_BYTES + B_BYTES + koff);
                        tc_mma_tf32(d_tmem, a_hi, b_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                        tc_mma_tf32(d_tmem, a_lo, b_hi, idesc, 1u);
                        tc_mma_tf32(d_tmem, a_hi, b_lo, idesc, 1u);
                    }
                    tc_commit(bar_empty + 8 * stage);  // stage free once these MMAs have read it
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                tc_commit(bar_tfull + 8 * acc);  // accumulator complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue ====================================
        const int q = warp & 3;             // TMEM lane quarter this warp may read
        const int half = (warp - 4) >> 2;   // which 128 columns of the tile
        const uint32_t stage_buf = s_epi + (uint32_t)(warp - 4) * EPI_TILE_BYTES;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            int mi, ni;
            tile_coords(t, n_m, n_n, mi, ni);
            const int m0 = mi * BM, n0 = ni * BN;
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            const int row0 = m0 + q * 32;
            bool any = false;
            {
                const int colbase = n0 + half * (BN / 2);
                const int n_blocks = colbase >= C ? 0 : min(BN / 32, (C - colbase + 15) / 16);
                impute_epi_tile<AS_INT, STORE>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2)),
                                                   n_blocks, scale, any, stage_buf, lane, &tm_out, colbase, row0, R, C, out);
            }
            // every lane of this warp has finished reading the accumulator
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
            // (the two warps of a lane quarter see different columns: either may set the flag)
            if (any && row_nonzero && row0 + lane < R) row_nonzero[row0 + lane] = 1;
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if ((STORE == 1 || STORE == 3) && lane == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

// CTA-pair version: two CTAs on neighbouring SMs compute one 256 x 256 tile with tcgen05.mma.cta_group::2
// (M = 256).  Each CTA stages its own 128 regions and HALF of the 256 cells per K block (64 KB instead of
// 96 KB, so the ring has three stages and a third less operand traffic comes out of L2); the leader CTA
// issues the MMAs, which read both CTAs' shared memory and write each CTA's 128 rows into its own TMEM.
template <bool AS_INT, int STORE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(imp::THREADS, 1)
impute_umma_pair_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                        const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                        const __grid_constant__ CUtensorMap tm_out, int32_t R, int32_t C, int32_t n_ksteps, float scale,
                        void *__restrict__ out, uint8_t *__restrict__ row_nonzero) {
    using namespace imp;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    const uint32_t s_epi = base + STAGES2 * STAGE2_BYTES;
    const uint32_t s_bar = s_epi + EPI_BYTES;
    // full[3] | empty[3] | tmem_full[2] | tmem_empty[2] | TMEM base address
    const uint32_t bar_full = s_bar, bar_empty = s_bar + 8 * STAGES2, bar_tfull = s_bar + 16 * STAGES2,
                   bar_tempty = s_bar + 16 * STAGES2 + 16;
    const uint32_t tmem_slot_addr = s_bar + 16 * STAGES2 + 32;
    volatile uint32_t *tmem_slot =
        reinterpret_cast<volatile uint32_t *>(smem + STAGES2 * STAGE2_BYTES + EPI_BYTES + 16 * STAGES2 + 32);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();  // 0 = leader
    const bool leader = rank == 0;

    if (warp == 1 && lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES2; ++s) {
            mbar_init(bar_full + 8 * s, 2);   // one arrival per CTA of the pair (the leader's carries the byte count)
            mbar_init(bar_empty + 8 * s, 1);  // the leader's commit, multicast to both CTAs
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 2 * EPI_WARPS);  // every epilogue warp of both CTAs
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot_addr),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_m2 = (R + 2 * BM - 1) / (2 * BM);
    const int n_n = (C + BN - 1) / BN;
    const int64_t n_tiles = (int64_t)n_m2 * n_n;
    const int n_kblocks = (n_ksteps + BK / UK - 1) / (BK / UK);
    const int64_t t_first = blockIdx.x >> 1, t_step = gridDim.x >> 1;

    if (warp == 0) {
        // ================================ TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t t = t_first; t < n_tiles; t += t_step) {
                int mi, ni;
                tile_coords(t, n_m2, n_n, mi, ni);
                const int m0 = mi * 2 * BM + (int)rank * BM;
                const int n0 = ni * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < n_kblocks; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                    const uint32_t full = bar_full + 8 * stage;
                    const uint32_t st = base + stage * STAGE2_BYTES;
                    tma_load_2d_pair(st, &tm_a_hi, full, kb * BK, m0);
                    tma_load_2d_pair(st + A_BYTES, &tm_a_lo, full, kb * BK, m0);
                    tma_load_2d_pair(st + 2 * A_BYTES, &tm_b_hi, full, kb * BK, n0);
                    tma_load_2d_pair(st + 2 * A_BYTES + B_HALF_BYTES, &tm_b_lo, full, kb * BK, n0);
                    if (leader) mbar_expect_tx(full, 2 * STAGE2_BYTES);
                    else mbar_arrive_cluster(full, 0);
                    if (++stage == STAGES2) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (leader CTA only) =================
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = umma_idesc_tf32(2 * BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int64_t t = t_first; t < n_tiles; t += t_step) {
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < n_kblocks; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t st = base + stage * STAGE2_BYTES;
                    const int ks = min(BK / UK, n_ksteps - kb * (BK / UK));
                    for (int k = 0; k < ks; ++k) {
                        const uint32_t koff = (uint32_t)(k * UK * 4);
                        const uint64_t a_hi = umma_desc_k_sw128(st + koff);
                        const uint64_t a_lo = umma_desc_k_sw128(st + A_BYTES + koff);
                        const uint64_t b_hi = umma_desc_k_sw128(st + 2 * A_BYTES + koff);
                        const uint64_t b_lo = umma_desc_k_sw128(st + 2 * A_BYTES + B_HALF_BYTES + koff);
                        tc_mma_tf32_pair(d_tmem, a_hi, b_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                        tc_mma_tf32_pair(d_tmem, a_lo, b_hi, idesc, 1u);
                        tc_mma_tf32_pair(d_tmem, a_hi, b_lo, idesc, 1u);
                    }
                    tc_commit_pair(bar_empty + 8 * stage);
                    if (++stage == STAGES2) { stage = 0; phase ^= 1u; }
                }
                tc_commit_pair(bar_tfull + 8 * acc);
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ================================ epilogue (own 128 rows, all 256 columns) ======
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const uint32_t stage_buf = s_epi + (uint32_t)(warp - 4) * EPI_TILE_BYTES;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t t = t_first; t < n_tiles; t += t_step) {
            int mi, ni;
            tile_coords(t, n_m2, n_n, mi, ni);
            const int m0 = mi * 2 * BM + (int)rank * BM, n0 = ni * BN;
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            const int row0 = m0 + q * 32;
            bool any = false;
            {
                const int colbase = n0 + half * (BN / 2);
                const int n_blocks = colbase >= C ? 0 : min(BN / 32, (C - colbase + 15) / 16);
                impute_epi_tile<AS_INT, STORE>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2)),
                                                   n_blocks, scale, any, stage_buf, lane, &tm_out, colbase, row0, R, C, out);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(bar_tempty + 8 * acc, 0);  // the leader issues the next MMAs
            if (any && row_nonzero && row0 + lane < R) row_nonzero[row0 + lane] = 1;
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if ((STORE == 1 || STORE == 3) && lane == 0) tma_store_wait_all();
    }

    // both CTAs must be done with each other's shared memory and barriers before either leaves
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*cgs_PFN_tensorMapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                                 CUtensorMapFloatOOBfill);

// The driver entry point is resolved through the runtime so that the library has no link-time
// dependency on libcuda (it must load on machines without a driver, e.g. for the symbol checks).
static int get_tensor_map_encoder(cgs_PFN_tensorMapEncodeTiled *fn) {
    static cgs_PFN_tensorMapEncodeTiled cached = nullptr;
    if (!cached) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CGS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !p) CGS_FAIL("cuTensorMapEncodeTiled is not available in this driver");
        cached = reinterpret_cast<cgs_PFN_tensorMapEncodeTiled>(p);
    }
    *fn = cached;
    return 0;
}

// 2-D map over a K-major fp32 array rows x Kp, box = box_rows x 32 (one 128-byte swizzle row wide)
static int make_operand_map(CUtensorMap *map, const float *ptr, int64_t rows, int32_t Kp, int box_rows) {
    cgs_PFN_tensorMapEncodeTiled enc;
    CGS_TRY(get_tensor_map_encoder(&enc));
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Kp * 4};
    cuuint32_t box[2] = {(cuuint32_t)imp::BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) CGS_FAIL("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return 0;
}

static inline int32_t impute_padded_k(int32_t K) { return (K + imp::BK - 1) / imp::BK * imp::BK; }

static int impute_prepare_cells(const float *d_B, int32_t K, int64_t C, int32_t Kp, float *d_hi, float *d_lo,
                                cudaStream_t stream) {
    dim3 grid((unsigned)ceil_div(C, 32), (unsigned)(Kp / 32)), block(32, 8);
    impute_transpose_split_kernel<<<grid, block, 0, stream>>>(d_B, K, C, Kp, d_hi, d_lo);
    CGS_CUDA(cudaGetLastError());
    return 0;
}

static int impute_prepare_regions(const float *d_A, int64_t R, int32_t K, int32_t Kp, float *d_hi, float *d_lo,
                                  int num_sms, cudaStream_t stream) {
    impute_split_rows_kernel<<<num_sms * 8, 256, 0, stream>>>(d_A, R, K, Kp, d_hi, d_lo);
    CGS_CUDA(cudaGetLastError());
    return 0;
}

// 2-D map over the output chunk (rows x C, 4-byte elements), box = 32 rows x 16 columns, 64-byte swizzle
static int make_output_map(CUtensorMap *map, void *ptr, int64_t rows, int64_t C) {
    cgs_PFN_tensorMapEncodeTiled enc;
    CGS_TRY(get_tensor_map_encoder(&enc));
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)C * 4};
    cuuint32_t box[2] = {16, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) CGS_FAIL("cuTensorMapEncodeTiled (output) failed (" + std::to_string((int)r) + ")");
    return 0;
}

template <bool AS_INT, int STORE>
static int impute_umma_launch_t(const CUtensorMap &ma_hi, const CUtensorMap &ma_lo, const CUtensorMap &mb_hi,
                                const CUtensorMap &mb_lo, const CUtensorMap &mo, int32_t R, int32_t C, int32_t n_ksteps,
                                float scale, void *d_out, uint8_t *d_row_nonzero, int grid, cudaStream_t stream) {
    CGS_CUDA(cudaFuncSetAttribute(impute_umma_kernel<AS_INT, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  imp::SMEM_BYTES));
    impute_umma_kernel<AS_INT, STORE><<<grid, imp::THREADS, imp::SMEM_BYTES, stream>>>(
        ma_hi, ma_lo, mb_hi, mb_lo, mo, R, C, n_ksteps, scale, d_out, d_row_nonzero);
    CGS_CUDA(cudaGetLastError());
    return 0;
}

template <bool AS_INT, int STORE>
static int impute_umma_pair_launch_t(const CUtensorMap &ma_hi, const CUtensorMap &ma_lo, const CUtensorMap &mb_hi,
                                     const CUtensorMap &mb_lo, const CUtensorMap &mo, int32_t R, int32_t C, int32_t n_ksteps,
                                     float scale, void *d_out, uint8_t *d_row_nonzero, int grid, cudaStream_t stream) {
    CGS_CUDA(cudaFuncSetAttribute(impute_umma_pair_kernel<AS_INT, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  imp::SMEM2_BYTES));
    impute_umma_pair_kernel<AS_INT, STORE><<<grid, imp::THREADS, imp::SMEM2_BYTES, stream>>>(
        ma_hi, ma_lo, mb_hi, mb_lo, mo, R, C, n_ksteps, scale, d_out, d_row_nonzero);
    CGS_CUDA(cudaGetLastError());
    return 0;
}

// Measured on B200 (20 000 x K x 200 000, GB/s written): TMA stores 4560 / 4280 / 3970 for K = 8 / 50 / 100,
// 16-byte coalesced stores 4960 / 4690 / 3890; neither path is the limiter.
constexpr int kImputeDefaultStore = 2;

// out = trunc(scale * A B) from the prepared operands
static int impute_umma_launch(const float *d_a_hi, const float *d_a_lo, const float *d_b_hi, const float *d_b_lo, int32_t R,
                              int32_t K, int32_t Kp, int32_t C, float scale, int as_int32, void *d_out,
                              uint8_t *d_row_nonzero, int num_sms, cudaStream_t stream) {
    // CTA pairs (cta_group::2) by default; CGS_IMPUTE_1CTA=1 selects the single-CTA kernel
    const char *e1 = getenv("CGS_IMPUTE_1CTA");
    const bool pair = !(e1 && e1[0] == '1') && num_sms >= 2;
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo, mo;
    CGS_TRY(make_operand_map(&ma_hi, d_a_hi, R, Kp, imp::BM));
    CGS_TRY(make_operand_map(&ma_lo, d_a_lo, R, Kp, imp::BM));
    CGS_TRY(make_operand_map(&mb_hi, d_b_hi, C, Kp, pair ? imp::BN / 2 : imp::BN));
    CGS_TRY(make_operand_map(&mb_lo, d_b_lo, C, Kp, pair ? imp::BN / 2 : imp::BN));
    // Output path: 16-byte aligned rows allow TMA stores and 16-byte coalesced stores.  CGS_IMPUTE_STORE
    // (tuning knob) = scalar | tma | vec | mix overrides the default.
    const bool aligned = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 15) == 0);
    int store = aligned ? kImputeDefaultStore : 0;
    if (const char *e2 = getenv("CGS_IMPUTE_STORE")) {
        const std::string v(e2);
        if (v == "scalar") store = 0;
        else if (aligned && v == "tma") store = 1;
        else if (aligned && v == "vec") store = 2;
        else if (aligned && v == "mix") store = 3;
        else if (aligned && v == "wide") store = 5;
    }
    if (store == 1 || store == 3) CGS_TRY(make_output_map(&mo, d_out, R, C));
    else mo = ma_hi;  // unused
    if (d_row_nonzero) CGS_CUDA(cudaMemsetAsync(d_row_nonzero, 0, (size_t)R, stream));
    const int32_t n_ksteps = (K + imp::UK - 1) / imp::UK;
    int64_t tiles;
    int grid;
    if (pair) {
        tiles = ceil_div(R, 2 * imp::BM) * ceil_div(C, imp::BN);
        grid = 2 * (int)std::max<int64_t>(1, std::min<int64_t>(tiles, num_sms / 2));
    } else {
        tiles = ceil_div(R, imp::BM) * ceil_div(C, imp::BN);
        grid = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, num_sms));
    }
#define CGS_IMPUTE_DISPATCH(AS_INT_, STORE_)                                                                               \
    (pair ? impute_umma_pair_launch_t<AS_INT_, STORE_>(ma_hi, ma_lo, mb_hi, mb_lo, mo, R, C, n_ksteps, scale, d_out,         \
                                                       d_row_nonzero, grid, stream)                                         \
          : impute_umma_launch_t<AS_INT_, STORE_>(ma_hi, ma_lo, mb_hi, mb_lo, mo, R, C, n_ksteps, scale, d_out,              \
                                                  d_row_nonzero, grid, stream))
    if (as_int32) {
        switch (store) {
            case 0: return CGS_IMPUTE_DISPATCH(true, 0);
            case 1: return CGS_IMPUTE_DISPATCH(true, 1);
            case 2: return CGS_IMPUTE_DISPATCH(true, 2);
            case 5: return CGS_IMPUTE_DISPATCH(true, 5);
            default: return CGS_IMPUTE_DISPATCH(true, 3);
        }
    }
    switch (store) {
        case 0: return CGS_IMPUTE_DISPATCH(false, 0);
        case 1: return CGS_IMPUTE_DISPATCH(false, 1);
        case 2: return CGS_IMPUTE_DISPATCH(false, 2);
        case 5: return CGS_IMPUTE_DISPATCH(false, 5);
        default: return CGS_IMPUTE_DISPATCH(false, 3);
    }
#undef CGS_IMPUTE_DISPATCH
}

static int impute_simt_launch(const float *d_A, const float *d_B, int32_t R, int32_t K, int32_t C, float scale, int as_int32,
                              void *d_out, uint8_t *d_row_nonzero, cudaStream_t stream) {
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
