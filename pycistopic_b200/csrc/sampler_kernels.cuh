// sampler_kernels.cuh -- the collapsed-Gibbs sweep (replaces lda._lda._sample_topics,
// SURVEY.md A6; call site lda_models.py:287) and the count bookkeeping around it
// (lda.LDA._initialize).
//
// Data layout in HBM
//   tok_region int32[N]     region id of every token, cell-major, regions ascending per cell
//   z          uint16[N]    current topic of every token
//   n_wk       int32[W*Kp]  region-major topic counts, Kp = K rounded up to 4 (16-byte rows)
//   n_dk       int32[D*Kp]  cell-major topic counts
//   n_k        int32[Kp]    topic totals
//
// Sweep kernel.  Persistent CTAs pull cells from a longest-first queue; a cell is owned by
// exactly one CTA for the sweep, so its n_dk row lives in shared memory (as integer counts
// and as the pre-multiplied factor f[k] = (n_dk[k] + alpha) / (n_k[k] + eta*W)).  Inside the
// CTA every warp takes 32 consecutive tokens at a time: one coalesced load of their region
// ids and topics, one Philox4x32-10 draw per token (counter = global token index, sweep).
// A token is handled by TPT lanes; each lane owns CH 16-byte chunks (4 topics each) of the
// token's n_wk row, read with 128-bit loads one step ahead of use.  The conditional
//   p(k) = (n_wk[w,k] - [k==z] + eta) * (n_dk[d,k] - [k==z] + alpha) / (n_k[k] - [k==z] + eta*W)
// is accumulated as a per-lane running sum, scanned across the TPT lanes with shuffles,
// and inverted with one ballot + an in-lane count.  A token that changes topic issues two
// no-return global reductions on n_wk (live counts: every other token on the GPU sees the
// update immediately, which is what keeps the chain close to the sequential reference) and
// shared-memory atomics for the cell's row; n_k deltas are aggregated per CTA and flushed once
// per cell.
#pragma once
#include "common.cuh"

namespace cgs {

constexpr int kSweepThreads = 256;  // upper bound; the launch uses 32..256
constexpr int kMaxTopics = 512;  // TPT=32 lanes x CH=4 chunks x 4 topics

struct SweepParams {
    const int64_t *cell_ptr;
    const int32_t *tok_region;
    uint16_t *z;
    int32_t *n_wk;
    int32_t *n_dk;
    int32_t *n_k;
    const int32_t *cell_order;
    int32_t *work_counter;
    int32_t n_cells;
    int32_t K;
    int32_t Kp;
    float alpha;
    float eta;
    float etaW;
    uint32_t seed_lo;
    uint32_t seed_hi;
    uint32_t sweep;
    int32_t refresh_every;  // a warp re-synchronises the topic totals every this many steps
    int32_t part;           // this launch handles queue slots part, part + n_parts, ... (sub-sweep)
    int32_t n_parts;
    int64_t token_offset;
};

__device__ __forceinline__ int4 ld_row_chunk(const int32_t *p) {
    // n_wk rows are read once per token and written by reductions from every SM: keep them out
    // of L1 (L2 is the coherence point for the reductions).
    int4 r;
    asm volatile("ld.global.cg.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ int32_t ld_volatile_i32(const int32_t *p) {
    int32_t v;
    asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void red_add_i32(int32_t *p, int32_t v) {
    asm volatile("red.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int TPT>
__device__ __forceinline__ int log2_tpt() {
    return TPT == 1 ? 0 : TPT == 2 ? 1 : TPT == 4 ? 2 : TPT == 8 ? 3 : TPT == 16 ? 4 : 5;
}

// (c + oc) mod CH for c, oc in [0, CH)
template <int CH>
__device__ __forceinline__ int rot_chunk(int c, int oc) {
    if (CH == 1) return 0;
    if (CH == 2) return (c + oc) & 1;
    if (CH == 4) return (c + oc) & 3;
    const int t = c + oc;
    return t >= CH ? t - CH : t;
}

template <int TPT, int CH>
__global__ void __launch_bounds__(kSweepThreads, (CH <= 2 ? 4 : 3)) sweep_kernel(const SweepParams p) {
    constexpr int KS = TPT * CH * 4;  // topic slots covered by one token group
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ __align__(16) float f_s[KS];    // (n_dk + alpha) / (n_k + eta W)
    __shared__ __align__(16) float inv_s[KS];  // 1 / (n_k + eta W)
    __shared__ __align__(16) float den_s[KS];  // n_k + eta W
    __shared__ __align__(16) int ndk_s[KS];
    __shared__ __align__(16) int dnk_s[KS];
    __shared__ int cell_slot_s;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const int sub = lane & (TPT - 1);
    const int gbase = lane & ~(TPT - 1);
    const int K = p.K, Kp = p.Kp;
    const float alpha = p.alpha, eta = p.eta;

    for (;;) {
        if (tid == 0) cell_slot_s = atomicAdd(p.work_counter, 1);
        __syncthreads();
        const int slot = cell_slot_s * p.n_parts + p.part;
        if (slot >= p.n_cells) break;
        const int cell = p.cell_order[slot];
        const int64_t t0 = p.cell_ptr[cell], t1 = p.cell_ptr[cell + 1];
        int32_t *ndk_row = p.n_dk + (int64_t)cell * Kp;

        for (int k = tid; k < KS; k += blockDim.x) {
            float den = 1.0f, inv = 0.0f, f = 0.0f;
            int nd = 0;
            if (k < K) {
                den = (float)ld_volatile_i32(p.n_k + k) + p.etaW;
                inv = 1.0f / den;
                nd = ndk_row[k];
                f = ((float)nd + alpha) * inv;
            }
            den_s[k] = den; inv_s[k] = inv; f_s[k] = f; ndk_s[k] = nd; dnk_s[k] = 0;
        }
        __syncthreads();

        const int64_t nbatch = (t1 - t0 + 31) >> 5;
        // Token lists are sorted by region and concurrently scheduled cells advance in lockstep, so
        // without this every token of a popular region would be sampled at the same moment against
        // the same stale n_wk row.  Start each cell at its own pseudo-random batch and wrap.
        uint32_t h = (uint32_t)cell * 0x9E3779B1u ^ (p.sweep * 0x85EBCA77u) ^ p.seed_lo;
        h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
        const int64_t rot = nbatch > 0 ? (int64_t)(h % (uint32_t)nbatch) : 0;
        int since_refresh = 0;
        for (int64_t bb = warp; bb < nbatch; bb += nwarps) {
            int64_t b = bb + rot;
            if (b >= nbatch) b -= nbatch;
            const int64_t base = t0 + (b << 5);
            const int64_t i = base + lane;
            const bool valid = i < t1;
            int w_j = -1, zo_j = 0;
            if (valid) {
                w_j = __ldg(p.tok_region + i);
                zo_j = (int)p.z[i];
            }
            const uint64_t gi = (uint64_t)(p.token_offset + i);
            const Philox4 rnd = philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), p.sweep, 0u,
                                              p.seed_lo, p.seed_hi);
            const float u_j = u01_from_bits(rnd.x);
            int zn_j = zo_j;

            const int rem = (int)min((int64_t)32, t1 - base);
            const int nsteps = rem < TPT ? rem : TPT;

            // Chunk order is rotated per token so that the chunk holding the token's own topic is
            // always register chunk 0 of its lane: the own-token exclusion then patches 4 slots
            // instead of all CH*4 (any fixed order of the topics gives the same distribution).
            int4 rcur[CH], rnext[CH];
            int w_cur = __shfl_sync(FULL, w_j, gbase);
            int zo_cur = __shfl_sync(FULL, zo_j, gbase);
            {
                const int oc = (zo_cur >> 2) >> log2_tpt<TPT>();
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int chunk = sub + TPT * rot_chunk<CH>(c, oc);
                    rcur[c] = make_int4(0, 0, 0, 0);
                    if (w_cur >= 0 && chunk * 4 < Kp)
                        rcur[c] = ld_row_chunk(p.n_wk + (int64_t)w_cur * Kp + chunk * 4);
                }
            }

            for (int s = 0; s < nsteps; ++s) {
                // ---- keep the topic totals live: publish this CTA's n_k deltas, re-read n_k ----
                if (++since_refresh > p.refresh_every) {
                    since_refresh = 1;
                    for (int k = lane; k < K; k += 32) {
                        const int d = atomicExch(&dnk_s[k], 0);
                        if (d != 0) red_add_i32(p.n_k + k, d);
                        const float den = (float)ld_volatile_i32(p.n_k + k) + p.etaW;
                        const float inv = __frcp_rn(den);
                        den_s[k] = den;
                        inv_s[k] = inv;
                        f_s[k] = ((float)ndk_s[k] + alpha) * inv;
                    }
                    __syncwarp();
                }
                // ---- prefetch the rows of the next step --------------------------------------
                int w_next = -1, zo_next = 0;
                if (TPT > 1) {
                    const int src = gbase + ((s + 1) & (TPT - 1));
                    w_next = __shfl_sync(FULL, w_j, src);
                    zo_next = __shfl_sync(FULL, zo_j, src);
                    if (s + 1 >= nsteps) w_next = -1;
                    const int ocn = (zo_next >> 2) >> log2_tpt<TPT>();
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const int chunk = sub + TPT * rot_chunk<CH>(c, ocn);
                        rnext[c] = make_int4(0, 0, 0, 0);
                        if (w_next >= 0 && chunk * 4 < Kp)
                            rnext[c] = ld_row_chunk(p.n_wk + (int64_t)w_next * Kp + chunk * 4);
                    }
                }
                const int w = w_cur;
                const int zo = zo_cur;
                const float u = (TPT > 1) ? __shfl_sync(FULL, u_j, gbase + s) : u_j;

                // ---- own-token exclusion: component of z_old in chunk 0 of its lane (or -1) -----
                const int ochunk = zo >> 2;
                const int oc = ochunk >> log2_tpt<TPT>();
                const int hitq = (sub == (ochunk & (TPT - 1))) ? (zo & 3) : -1;
                const float fpp = __fdividef((float)(ndk_s[zo] - 1) + alpha, den_s[zo] - 1.0f);

                // ---- running sum of p(k) over this lane's slots -------------------------------
                float cum[CH * 4];
                float acc = 0.0f;
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int chunk = sub + TPT * rot_chunk<CH>(c, oc);
                    const float4 fv = *reinterpret_cast<const float4 *>(&f_s[chunk * 4]);
                    int nn[4] = {rcur[c].x, rcur[c].y, rcur[c].z, rcur[c].w};
                    float ff[4] = {fv.x, fv.y, fv.z, fv.w};
                    if (c == 0) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (hitq == q) { nn[q] -= 1; ff[q] = fpp; }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        acc = fmaf((float)nn[q] + eta, ff[q], acc);
                        cum[c * 4 + q] = acc;
                    }
                }

                // ---- scan over the TPT lanes of the token, draw, invert ------------------------
                float incl = acc;
                if (TPT > 1) {
#pragma unroll
                    for (int off = 1; off < TPT; off <<= 1) {
                        const float t = __shfl_up_sync(FULL, incl, off, TPT);
                        if (sub >= off) incl += t;
                    }
                }
                const float total = (TPT > 1) ? __shfl_sync(FULL, incl, TPT - 1, TPT) : incl;
                const float tloc = u * total - (incl - acc);  // threshold relative to this lane
                int cnt = 0;
#pragma unroll
                for (int j = 0; j < CH * 4; ++j) cnt += (cum[j] <= tloc) ? 1 : 0;
                const int cand = (sub + TPT * rot_chunk<CH>(cnt >> 2, oc)) * 4 + (cnt & 3);
                int zn;
                if (TPT > 1) {
                    const unsigned bal = __ballot_sync(FULL, cnt < CH * 4);
                    const unsigned gm = (TPT == 32) ? bal : ((bal >> gbase) & ((1u << TPT) - 1u));
                    const int src = gm ? (gbase + __ffs(gm) - 1) : lane;
                    zn = __shfl_sync(FULL, cand, src);
                    if (!gm) zn = zo;
                } else {
                    zn = cand;
                }
                if (zn >= K) zn = zo;  // rounding guard: padded slots carry p = 0

                // ---- apply the move (one lane per token) ---------------------------------------
                if (sub == (s & (TPT - 1)) && w >= 0 && zn != zo) {
                    int32_t *row = p.n_wk + (int64_t)w * Kp;
                    red_add_i32(row + zo, -1);
                    red_add_i32(row + zn, 1);
                    const int o0 = atomicAdd(&ndk_s[zo], -1);
                    const int o1 = atomicAdd(&ndk_s[zn], 1);
                    f_s[zo] = ((float)(o0 - 1) + alpha) * inv_s[zo];
                    f_s[zn] = ((float)(o1 + 1) + alpha) * inv_s[zn];
                    atomicAdd(&dnk_s[zo], -1);
                    atomicAdd(&dnk_s[zn], 1);
                    zn_j = zn;
                }
                if (TPT > 1) {
                    w_cur = w_next;
                    zo_cur = zo_next;
#pragma unroll
                    for (int c = 0; c < CH; ++c) rcur[c] = rnext[c];
                }
            }
            if (valid) p.z[i] = (uint16_t)zn_j;
        }
        __syncthreads();
        for (int k = tid; k < K; k += blockDim.x) {
            ndk_row[k] = ndk_s[k];
            const int d = dnk_s[k];
            if (d != 0) red_add_i32(p.n_k + k, d);
        }
        __syncthreads();
    }
}

// ---- z initialisation and count rebuild ------------------------------------------------------
__global__ void init_z_kernel(uint16_t *__restrict__ z, int64_t n, int64_t global_offset, int32_t K) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        z[i] = (uint16_t)((global_offset + i) % K);
}

__global__ void z_from_i32_kernel(const int32_t *__restrict__ zin, uint16_t *__restrict__ z, int64_t n,
                                  int32_t K, int *__restrict__ bad_flag) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int32_t v = zin[i];
        if (v < 0 || v >= K) { *bad_flag = 1; v = 0; }
        z[i] = (uint16_t)v;
    }
}

__global__ void z_to_i32_kernel(const uint16_t *__restrict__ z, int32_t *__restrict__ zout, int64_t n) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) zout[i] = z[i];
}

// One CTA per cell (grid-stride): histogram of the cell's topics in shared memory, one global
// reduction per token on n_wk.  All tables must be zeroed first.
__global__ void count_kernel(const int64_t *__restrict__ cell_ptr, const int32_t *__restrict__ tok_region,
                             const uint16_t *__restrict__ z, int32_t n_cells, int32_t K, int32_t Kp,
                             int32_t *__restrict__ n_wk, int32_t *__restrict__ n_dk, int32_t *__restrict__ n_k) {
    extern __shared__ int hist[];
    for (int32_t cell = blockIdx.x; cell < n_cells; cell += gridDim.x) {
        for (int k = threadIdx.x; k < Kp; k += blockDim.x) hist[k] = 0;
        __syncthreads();
        const int64_t b = cell_ptr[cell], e = cell_ptr[cell + 1];
        for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
            const int k = z[i];
            atomicAdd(&hist[k], 1);
            red_add_i32(n_wk + (int64_t)tok_region[i] * Kp + k, 1);
        }
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const int h = hist[k];
            n_dk[(int64_t)cell * Kp + k] = h;
            if (h) red_add_i32(n_k + k, h);
        }
        __syncthreads();
    }
}

// ---- AD-LDA exchange kernels -----------------------------------------------------------------
__global__ void delta_compute_kernel(const int4 *__restrict__ n_wk, const int4 *__restrict__ snap,
                                     int4 *__restrict__ delta, int64_t n4) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int4 a = n_wk[i], b = snap[i];
        delta[i] = make_int4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
    }
}

constexpr int kMaxPeers = 16;
struct PeerPtrs {
    const int4 *p[kMaxPeers];
};

// n_wk = snap + sum_r delta_r ; snap = n_wk ; n_k += column sums.  With n_ranks == 1 and
// peers.p[0] == local delta this is the plain "apply" after a library all-reduce.
// Kp4 = Kp / 4 (row length in int4 units); n_k must be zeroed first.
__global__ void delta_apply_kernel(int4 *__restrict__ n_wk, int4 *__restrict__ snap, const PeerPtrs peers,
                                   int32_t n_ranks, int64_t n4, int32_t Kp4, int32_t *__restrict__ n_k) {
    extern __shared__ int colsum[];  // Kp ints
    const int Kp = Kp4 * 4;
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) colsum[k] = 0;
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        int4 acc = snap[i];
        for (int r = 0; r < n_ranks; ++r) {
            const int4 d = peers.p[r][i];
            acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
        }
        n_wk[i] = acc;
        snap[i] = acc;
        const int col = (int)(i % Kp4) * 4;
        if (acc.x) atomicAdd(&colsum[col + 0], acc.x);
        if (acc.y) atomicAdd(&colsum[col + 1], acc.y);
        if (acc.z) atomicAdd(&colsum[col + 2], acc.z);
        if (acc.w) atomicAdd(&colsum[col + 3], acc.w);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < Kp; k += blockDim.x)
        if (colsum[k]) red_add_i32(n_k + k, colsum[k]);
}

}  // namespace cgs
