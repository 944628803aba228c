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
// and as the pre-multiplied factor f[k] = (n_dk[k] + alpha) / (n_k[k] + eta*W)).  A warp walks
// its share of the cell in batches of 32 consecutive tokens (one coalesced load of region ids
// and topics, issued one batch ahead) grouped four at a time so that ONE Philox4x32-10 call
// per lane serves four batches.  A token is handled by TPT lanes; a warp step therefore
// samples 32/TPT tokens.  Each lane owns CH 16-byte chunks (4 topics each) of the token's
// n_wk row, read with 128-bit loads one step ahead of use into a register ping-pong buffer
// (the step loop is unrolled, so the buffers never move).  The conditional
//   p(k) = (n_wk[w,k] - [k==z] + eta) * (n_dk[d,k] - [k==z] + alpha) / (n_k[k] - [k==z] + eta*W)
// is accumulated as a per-lane running sum, scanned across the TPT lanes with shuffles,
// and inverted with one ballot + an in-lane count.  A token that changes topic is applied by
// two lanes of its group (one per topic): a no-return global reduction on n_wk (live counts:
// every other token on the GPU sees the update immediately, which is what keeps the chain
// close to the sequential reference) and a shared-memory atomic on the cell's row; n_k
// deltas are the difference between the row and its last published copy, flushed every few
// batches and once per cell.
#pragma once
#include "common.cuh"
#include "exchange_kernels.cuh"

namespace cgs {

constexpr int kSweepThreads = 512;  // upper bound (16 warps per cell); the launch uses 32..512
constexpr int kMaxTopics = 512;     // TPT * CH * 4 slots, TPT <= 32, CH <= 8
constexpr int kMaxChunksPerLane = 8;

struct SweepParams {
    const int64_t *cell_ptr;
    const int32_t *tok_region;
    uint16_t *z;
    int32_t *n_wk;               // receives the count updates (no-return reductions)
    const int32_t *n_wk_read;    // rows are READ from here: n_wk itself (live counts, default) or a copy taken
                                 // at the start of the sweep part (deterministic mode, cgs_model_set_deterministic)
    int64_t nwk_write_delta;     // (char *)n_wk - (char *)n_wk_read
    int32_t *n_dk;
    int32_t *n_k;
    const int32_t *n_k_read;     // same for the topic totals
    const int32_t *cell_order;
    const int64_t *cell_lo;      // per cell: first and one-past-last token this launch samples -- cell_ptr and
    const int64_t *cell_hi;      // cell_ptr + 1 for whole cells, or the bounds of one REGION BLOCK of every cell
    int32_t x_ctas;              // fused launches: blocks [0, x_ctas) run the n_wk exchange, the rest sample
    int32_t *work_counter;
    int32_t n_cells;
    int32_t K;
    int32_t Kp;
    float alpha;
    float eta;
    float etaW;
    double alpha64, eta64, etaW64;  // the same three for the fp64 A/B build
    uint32_t seed_lo;
    uint32_t seed_hi;
    uint32_t sweep;
    int32_t refresh_every;  // a warp re-synchronises the topic totals every this many batches
    int32_t part;           // this launch handles queue slots part, part + n_parts, ... (sub-sweep)
    int32_t n_parts;
    int64_t token_offset;
    unsigned long long *visit_counter;  // diagnostics: += tokens sampled (NULL in production launches)
    uint64_t rng_index_mask;            // Philox counter = token index & mask (see cgs_model_set_rng_period)
};

// ---- memory primitives ----------------------------------------------------------------------
// n_wk rows are read once per token and written by reductions from every SM: keep them out of
// L1 (L2 is the coherence point for the reductions).
__device__ __forceinline__ int32_t ld_volatile_i32(const int32_t *p) {
    int32_t v;
    asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void red_add_i32(int32_t *p, int32_t v) {
    asm volatile("red.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Shared memory is addressed with 32-bit window addresses computed once per CTA.
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_f(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_i(uint32_t a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_i(uint32_t a, int v) {
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void reds_add(uint32_t a, int v) {
    asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ int atoms_exch(uint32_t a, int v) {
    int o;
    asm volatile("atom.shared.exch.b32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory");
    return o;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ---- the arithmetic type of the conditional ------------------------------------------------------
// float in production.  double is the A/B build the parity tests use (cgs_model_set_conditional_precision):
// same schedule, same random bits, same integer bookkeeping -- only the arithmetic of p(k), its running sum and
// the inversion are carried in the reference's precision (lda._lda._sample_topics works in fp64).
template <typename R> struct Real;
template <> struct Real<float> {
    static __device__ __forceinline__ void lds4(uint32_t a, float (&v)[4]) {
        const float4 t = lds_f4(a);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ float lds(uint32_t a) { return lds_f(a); }
    static __device__ __forceinline__ void sts(uint32_t a, float v) { sts_f(a, v); }
    static __device__ __forceinline__ float rcp_fast(float x) { return rcp_approx(x); }
    static __device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
};
template <> struct Real<double> {
    static __device__ __forceinline__ void lds4(uint32_t a, double (&v)[4]) {
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(a));
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[2]), "=d"(v[3]) : "r"(a + 16));
    }
    static __device__ __forceinline__ double lds(uint32_t a) {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
        return v;
    }
    static __device__ __forceinline__ void sts(uint32_t a, double v) {
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
    }
    static __device__ __forceinline__ double rcp_fast(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000ll); }
};

// Shared-memory layout of a cell: f | inv | den (KS reals each), then ndk | pub (KS ints each), then the queue slot.
template <int TPT, int CH, typename R> struct CellSmem {
    static constexpr int KS = TPT * CH * 4;
    static constexpr int RB = (int)sizeof(R);
    static constexpr uint32_t F = 0, INV = KS * RB, DEN = 2 * KS * RB, NDK = 3 * KS * RB, PUB = NDK + KS * 4,
                              SLOT = PUB + KS * 4;
    static constexpr int WORDS = (int)(SLOT / 4) + 4;
};

template <int TPT>
__host__ __device__ constexpr int log2_tpt() {
    return TPT == 1 ? 0 : TPT == 2 ? 1 : TPT == 4 ? 2 : TPT == 8 ? 3 : TPT == 16 ? 4 : 5;
}

// (c + oc) mod CH for c, oc in [0, CH)
template <int CH>
__device__ __forceinline__ int rot_chunk(int c, int oc) {
    if (CH == 1) return 0;
    if (CH == 2) return (c + oc) & 1;
    if (CH == 4) return (c + oc) & 3;
    if (CH == 8) return (c + oc) & 7;
    const int t = c + oc;
    return t >= CH ? t - CH : t;
}

// Number of leading entries of the non-decreasing array cum[0..N) that are <= t: bisection
// over registers (select trees), about N + 2 log2 N instructions instead of 2.5 N.
template <int N, typename R>
__device__ __forceinline__ int count_le_sorted(const R (&cum)[N], R t) {
    constexpr int P = N <= 4 ? 4 : N <= 8 ? 8 : N <= 16 ? 16 : 32;
    R a[P];
#pragma unroll
    for (int j = 0; j < P; ++j) a[j] = j < N ? cum[j] : Real<R>::inf();
    int cnt = 0;
#pragma unroll
    for (int S = P / 2; S >= 1; S >>= 1) {
        const bool ge = a[S - 1] <= t;
        if (ge) cnt += S;
#pragma unroll
        for (int j = 0; j + 1 < S; ++j) a[j] = ge ? a[j + S] : a[j];
        if (S == 1 && ge) {  // the last entry itself
            // (a[1] is the entry after the bisection point: counts when it is <= t as well)
        }
    }
    // after the loop cnt is the index of the first entry > t within the first P - 1 entries; the
    // entry P - 1 is never tested by the bisection above
    return cnt;
}

// A token travels between lanes as two words: its region id and (22 random bits | padding flag |
// topic).
__device__ __forceinline__ int pack_token(uint32_t bits, int zo, bool valid) {
    return (int)((bits & 0xfffffc00u) | (valid ? 0u : 0x200u) | (uint32_t)zo);
}

// One warp step: samples the 32/TPT tokens (w, pk) whose n_wk rows sit in `buf`, and -- as soon
// as the rows have been consumed -- requests the rows of the next group (w_next, pk_next) into
// the same registers, so the loads fly while this step scans, inverts and applies.  Returns the
// new topic of the group's token (same value in all TPT lanes of the group).  sm = shared-memory
// window address of f; inv, den, ndk follow at multiples of KS words.
template <int TPT, int CH, typename R>
__device__ __forceinline__ int sweep_step(int4 (&buf)[CH], const char *nwk_lane, uint32_t rowbytes, int lim, int w,
                                          int pk, int w_next, int pk_next, uint32_t sm, int sub, int gbase, int lane,
                                          R alpha, R eta, int K, int64_t wdelta) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int LOG = log2_tpt<TPT>();
    using L = CellSmem<TPT, CH, R>;
    constexpr int RB = L::RB;
    const int zo = pk & 0x1ff;
    const bool live = (pk & 0x200) == 0;

    // ---- own-token exclusion: component of z_old in register chunk 0 of its lane (or -1) ----------
    const int ochunk = zo >> 2;
    const int oc = ochunk >> LOG;
    const int hitq = (sub == (ochunk & (TPT - 1))) ? (zo & 3) : -1;
    const R fpp = ((R)(lds_i(sm + L::NDK + zo * 4) - 1) + alpha) * Real<R>::rcp_fast(Real<R>::lds(sm + L::DEN + zo * RB) - (R)1);

    // ---- running sum of p(k) over this lane's slots ------------------------------------------
    R cum[CH * 4];
    R acc = (R)0;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int rc = rot_chunk<CH>(c, oc);
        R ff[4];
        Real<R>::lds4(sm + L::F + (sub * 4 + rc * (TPT * 4)) * RB, ff);
        const int4 nv = buf[c];
        R x[4] = {(R)nv.x + eta, (R)nv.y + eta, (R)nv.z + eta, (R)nv.w + eta};
        if (c == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (hitq == q) { x[q] -= (R)1; ff[q] = fpp; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            acc = fma(x[q], ff[q], acc);
            cum[c * 4 + q] = acc;
        }
    }

    // ---- request the next group's rows (chunk order rotated so its own topic lands in chunk 0) ----
    {
        const char *row_n = nwk_lane + (uint64_t)((uint32_t)w_next) * rowbytes;
        const int ocn = ((pk_next & 0x1ff) >> 2) >> LOG;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int rc = rot_chunk<CH>(c, ocn);
            // chunks past the end of the row are neither fetched nor cleared: their slots carry f = 0
            if (rc < lim) buf[c] = __ldcg(reinterpret_cast<const int4 *>(row_n + rc * (TPT * 16)));
        }
    }

    // ---- butterfly over the TPT lanes of the token: exclusive prefix and total -----------------
    R excl = (R)0, total = acc;
    if (TPT > 1) {
#pragma unroll
        for (int off = 1; off < TPT; off <<= 1) {
            const R t = __shfl_xor_sync(FULL, total, off);
            if (sub & off) excl += t;
            total += t;
        }
    }
    const R u = (R)((uint32_t)pk >> 10) * (R)(1.0 / 4194304.0);
    const R tloc = u * total - excl;  // threshold relative to this lane
    // cum[CH*4 - 1] == acc; a lane whose acc <= tloc holds no crossing
    const bool mine = !(acc <= tloc);
    const int cnt = count_le_sorted<CH * 4, R>(cum, tloc);
    const int cand = (sub + TPT * rot_chunk<CH>(cnt >> 2, oc)) * 4 + (cnt & 3);
    int zn;
    if (TPT > 1) {
        const unsigned bal = __ballot_sync(FULL, mine);
        const unsigned gm = (TPT == 32) ? bal : ((bal >> gbase) & ((1u << TPT) - 1u));
        const int src = gm ? (gbase + __ffs(gm) - 1) : lane;
        zn = __shfl_sync(FULL, cand, src);
        if (!gm) zn = zo;
    } else {
        zn = mine ? cand : zo;
    }
    if (zn >= K || !live) zn = zo;  // rounding guard: padded slots carry p = 0

    // ---- apply the move -------------------------------------------------------------------------
    // The cell's row is updated with integer reductions (exact, commutative); the factor f[k] is then rebuilt
    // from a FRESH read of the count, so that when several token groups of the warp (or several warps of the
    // CTA) move into or out of the same topic in the same step every writer stores the value of the latest
    // count it can see instead of the one its own atomic returned.
    if (zn != zo) {
        const char *row = nwk_lane + (uint64_t)((uint32_t)w) * rowbytes + wdelta;
        if (TPT > 1) {
            // lane 0 of the group retires the old topic, lane 1 installs the new one
            if (sub < 2) {
                const int k = sub ? zn : zo;
                const int sg = sub ? 1 : -1;
                red_add_i32((int32_t *)const_cast<char *>(row + (k * 4 - sub * 16)), sg);
                reds_add(sm + L::NDK + k * 4, sg);
                Real<R>::sts(sm + L::F + k * RB, ((R)lds_i(sm + L::NDK + k * 4) + alpha) * Real<R>::lds(sm + L::INV + k * RB));
            }
        } else {
            int32_t *rowi = (int32_t *)const_cast<char *>(row);
            red_add_i32(rowi + zo, -1);
            red_add_i32(rowi + zn, 1);
            reds_add(sm + L::NDK + zo * 4, -1);
            reds_add(sm + L::NDK + zn * 4, 1);
            Real<R>::sts(sm + L::F + zo * RB, ((R)lds_i(sm + L::NDK + zo * 4) + alpha) * Real<R>::lds(sm + L::INV + zo * RB));
            Real<R>::sts(sm + L::F + zn * RB, ((R)lds_i(sm + L::NDK + zn * 4) + alpha) * Real<R>::lds(sm + L::INV + zn * RB));
        }
    }
    return zn;
}

// Shapes whose steps are too short to hide the token stream's latency behind (see the kernel): one or two lanes
// per token and at most three chunks per lane (K <= 24; beyond that the registers it costs spill).
template <int TPT, int CH> constexpr bool kDeepTokenPrefetch = (TPT <= 2 && CH <= 3);

// Register budget (measured on B200: occupancy beats the few spills it costs; profiles/r1_shape_scan_*.txt):
// 40 warps per SM at 48 registers with one chunk per lane (K <= 8: the deeper token prefetch costs 7 registers, which
// at the former 64-register budget took a fifth of the resident warps away -- profiles/r2_sweep_k2_deep_v1_ncu_summary.csv;
// with two chunks per lane the 48-register budget measured 3-4 % SLOWER than 64, profiles/r2_small_k_variants.txt),
// 32 warps at 64 registers up to 4 chunks, 24 warps at 80 registers for 5-6 chunks, 16 warps beyond.  A cell may take
// up to 8 warps (256 threads) for CH = 1, 16 (512) for CH 2-4, 12 (384) for CH 5-6, 8 (256) beyond.
template <int CH, typename R = float, bool XCH = false> struct SweepBounds {
    static constexpr int kThreads = CH <= 1 ? 256 : CH <= 4 ? 512 : CH <= 6 ? 384 : 256;
    // the fp64 A/B build trades occupancy for registers; the exchange blocks of a fused launch want 64
    static constexpr int kMinBlocks = sizeof(R) == 8 ? 1 : (CH <= 1 && !XCH) ? 5 : (CH <= 1 ? 4 : 2);
};

// XCH = true: the launch that overlaps the AD-LDA exchange with sampling.  Tokens are sorted by region inside a cell, so
// a sweep can be cut at region indices into B region-block sweeps -- first every cell's tokens of block 0, then of block
// 1, ... -- and while one block is sampled, the rows of n_wk that belong to the PREVIOUS block (just changed, untouched by
// this launch) are exchanged between the ranks: blocks [0, x_ctas) of the grid run the synchronous exchange
// on that row range (exchange_sync_body: delta, reduce-scatter + all-gather over the peers' arenas, apply), all other
// blocks sample.  One launch does both, so the exchange blocks are resident from the first cycle (they come first in
// the grid) and the NVLink round trips, the cross-rank barriers and the table passes hide behind the sampling.  Every
// row is still exchanged exactly once per sweep, right after it was sampled: the statistics are those of the
// synchronous scheme (a rank sees the others' counts of a row as of the previous sweep).
template <int TPT, int CH, typename R = float, bool XCH = false>
__global__ void __launch_bounds__(SweepBounds<CH, R, XCH>::kThreads, SweepBounds<CH, R, XCH>::kMinBlocks)
sweep_kernel(const SweepParams p, const ExchangeParams xp) {
    using L = CellSmem<TPT, CH, R>;
    constexpr int KS = L::KS;  // topic slots covered by one token group
    constexpr int RB = L::RB;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ __align__(16) uint32_t smem_w[L::WORDS];
    if constexpr (XCH) {
        if ((int)blockIdx.x < p.x_ctas) {
            exchange_sync_body(xp, blockIdx.x, (unsigned)p.x_ctas, reinterpret_cast<int *>(smem_w));
            return;
        }
    }
    const uint32_t sm = smem_u32(smem_w);
    const uint32_t s_f = sm + L::F, s_inv = sm + L::INV, s_den = sm + L::DEN, s_ndk = sm + L::NDK, s_pub = sm + L::PUB;
    const uint32_t s_slot = sm + L::SLOT;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(FULL, tid >> 5, 0);  // warp-uniform by construction
    const int nwarps = blockDim.x >> 5;
    const int sub = lane & (TPT - 1);
    const int gbase = lane & ~(TPT - 1);
    const int K = p.K, Kp = p.Kp;
    const R alpha = sizeof(R) == 8 ? (R)p.alpha64 : (R)p.alpha, eta = sizeof(R) == 8 ? (R)p.eta64 : (R)p.eta;
    const R etaW = sizeof(R) == 8 ? (R)p.etaW64 : (R)p.etaW;
    const uint32_t rowbytes = (uint32_t)Kp * 4u;
    // register chunks of this lane that hold topics (the row stride Kp may be padded beyond them)
    const int lim = (((K + 3) >> 2) - sub + TPT - 1) >> log2_tpt<TPT>();
    const char *nwk_lane = reinterpret_cast<const char *>(p.n_wk_read) + sub * 16;

    int4 buf[CH];  // the n_wk row chunks of the group sampled next
#pragma unroll
    for (int c = 0; c < CH; ++c) buf[c] = make_int4(0, 0, 0, 0);

    for (;;) {
        if (tid == 0) sts_i(s_slot, atomicAdd(p.work_counter, 1));
        __syncthreads();
        const int slot = lds_i(s_slot) * p.n_parts + p.part;
        if (slot >= p.n_cells) break;
        const int cell = p.cell_order[slot];
        const int64_t t0 = p.cell_lo[cell], t1 = p.cell_hi[cell];
        int32_t *ndk_row = p.n_dk + (int64_t)cell * Kp;

        for (int k = tid; k < KS; k += blockDim.x) {
            R den = (R)1, inv = (R)0, f = (R)0;
            int nd = 0;
            if (k < K) {
                den = (R)ld_volatile_i32(p.n_k_read + k) + etaW;
                inv = (R)1 / den;
                nd = ndk_row[k];
                f = ((R)nd + alpha) * inv;
            }
            Real<R>::sts(s_den + k * RB, den); Real<R>::sts(s_inv + k * RB, inv); Real<R>::sts(s_f + k * RB, f);
            sts_i(s_ndk + k * 4, nd); sts_i(s_pub + k * 4, nd);
        }
        __syncthreads();

        // A warp's share of the cell.  The cell's batches (32 tokens) are walked from a per-cell
        // pseudo-random start batch `rot`, wrapping at the end: token lists are sorted by region and
        // concurrently scheduled cells advance in lockstep, so without the rotation every token of a
        // popular region would be sampled at the same moment against the same stale n_wk row.  The
        // rotated sequence [rot, nbatch) ++ [0, rot) is cut into super-batches of up to four
        // consecutive batches (never across the wrap) that the warps take round-robin.
        const int ntok = (int)(t1 - t0);
        const int nbatch = (ntok + 31) >> 5;
        int rot = 0;
        if (nbatch > 0) {
            // The start is the SAME in every sweep: a start that changes per sweep turns the systematic scan of
            // the reference into a partly random scan, which measurably slows the transient (C1, K = 50,
            // sequential CPU emulation, 16 seeds: -0.8 % at sweep 20 with a per-sweep start, -0.3 % with a
            // fixed one, which is the level of independent random numbers alone).
            uint32_t h = (uint32_t)cell * 0x9E3779B1u ^ p.seed_lo;
            h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
            rot = (int)(h % (uint32_t)nbatch);
        }
        const int nsup1 = (nbatch - rot + 3) >> 2;        // super-batches of the segment [rot, nbatch)
        const int nsuper = nsup1 + ((rot + 3) >> 2);      // ... plus those of [0, rot)
        const int32_t *tok_cell = p.tok_region + t0;
        uint16_t *z_cell = p.z + t0;
        const uint64_t gi0 = (uint64_t)(p.token_offset + t0) + (uint64_t)lane;

        // Position of a batch in this warp's walk over the cell's rotated batch sequence.
        struct Pos {
            int sidx, off, bq, nbq;
            bool more;
        };
        auto start_super = [&](Pos &q) {  // q.sidx is set: locate its super-batch
            q.more = q.sidx < nsuper;
            q.bq = 0;
            if (q.more) {
                const int b0 = q.sidx < nsup1 ? rot + 4 * q.sidx : 4 * (q.sidx - nsup1);
                q.nbq = min(4, (q.sidx < nsup1 ? nbatch : rot) - b0);
                q.off = b0 << 5;
            }
        };
        auto advance = [&](Pos &q) {
            q.off += 32;
            q.bq += 1;
            if (q.bq == q.nbq) {
                q.sidx += nwarps;
                start_super(q);
            }
        };

        if constexpr (kDeepTokenPrefetch<TPT, CH>) {
            // ---- K <= 16 (one or two lanes per token): a step is so short that the chain "token ids (streamed from
            // HBM) -> row address -> row (L2) -> sample" is what a warp waits for (ncu, round 1: long-scoreboard 13.5 of
            // 24 stall cycles per issue at K = 2).  The token ids and topics are therefore requested TWO batches ahead:
            // they have a whole step to arrive before the step that turns them into row requests.
            Pos pn;  // position of the batch after the current one
            pn.sidx = warp; pn.off = 0; pn.nbq = 0;
            start_super(pn);
            int off = pn.off;
            bool more = pn.more;
            int w_j = 0, pk_j = 0;
            int w_c = 0, pk_c = 0x200;  // the group sampled by the next step
            Philox4 rnd = {0u, 0u, 0u, 0u};
            int w_n = 0, z_n = 0, pkp_n = 0x200;  // next batch: region ids, topics (in flight), random bits | validity
            if (more) {
                const bool valid = (off + lane) < ntok;
                const int ic = valid ? off + lane : ntok - 1;
                w_j = __ldg(tok_cell + ic);
                const uint64_t gi = (gi0 + (uint64_t)off) & p.rng_index_mask;
                rnd = philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), p.sweep, 0u, p.seed_lo, p.seed_hi);
                pk_j = pack_token(rnd.x, (int)z_cell[ic], valid);
                w_c = (TPT > 1) ? __shfl_sync(FULL, w_j, 0, TPT) : w_j;
                pk_c = (TPT > 1) ? __shfl_sync(FULL, pk_j, 0, TPT) : pk_j;
                const char *row = nwk_lane + (uint64_t)((uint32_t)w_c) * rowbytes;
                const int oc = ((pk_c & 0x1ff) >> 2) >> log2_tpt<TPT>();
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int rc = rot_chunk<CH>(c, oc);
                    if (rc < lim) buf[c] = __ldcg(reinterpret_cast<const int4 *>(row + rc * (TPT * 16)));
                }
                advance(pn);
                if (pn.more) {
                    const bool valid_n = (pn.off + lane) < ntok;
                    const int icn = valid_n ? pn.off + lane : ntok - 1;
                    w_n = __ldg(tok_cell + icn);
                    z_n = (int)z_cell[icn];
                    if (pn.bq == 0) {
                        const uint64_t gn = (gi0 + (uint64_t)pn.off) & p.rng_index_mask;
                        rnd = philox4x32_10((uint32_t)gn, (uint32_t)(gn >> 32), p.sweep, 0u, p.seed_lo, p.seed_hi);
                    }
                    const uint32_t bits = pn.bq == 0 ? rnd.x : pn.bq == 1 ? rnd.y : pn.bq == 2 ? rnd.z : rnd.w;
                    pkp_n = pack_token(bits, 0, valid_n);
                }
            }
            int since_refresh = 0;
            while (more) {
                // ---- two batches ahead: request region ids and topics ----------------------------------
                const int off_n = pn.off;
                const bool more_n = pn.more;
                int w_nn = 0, z_nn = 0, pkp_nn = 0x200;
                advance(pn);  // pn is now the batch after the next one
                if (more_n && pn.more) {
                    const bool valid_nn = (pn.off + lane) < ntok;
                    const int icnn = valid_nn ? pn.off + lane : ntok - 1;
                    w_nn = __ldg(tok_cell + icnn);
                    z_nn = (int)z_cell[icnn];
                    if (pn.bq == 0) {  // one Philox call per lane serves the four batches of a super-batch
                        const uint64_t gn = (gi0 + (uint64_t)pn.off) & p.rng_index_mask;
                        rnd = philox4x32_10((uint32_t)gn, (uint32_t)(gn >> 32), p.sweep, 0u, p.seed_lo, p.seed_hi);
                    }
                    const uint32_t bits = pn.bq == 0 ? rnd.x : pn.bq == 1 ? rnd.y : pn.bq == 2 ? rnd.z : rnd.w;
                    pkp_nn = pack_token(bits, 0, valid_nn);
                }
                // the next batch's topics were requested one step ago
                if (!more_n) { w_n = w_j; pkp_n = 0x200; z_n = 0; }
                const int pk_n = pkp_n | z_n;

                // ---- keep the topic totals live: publish this cell's n_k deltas, re-read n_k ---------
                if (++since_refresh > p.refresh_every) {
                    since_refresh = 1;
                    for (int k = lane; k < K; k += 32) {
                        const int cur = lds_i(s_ndk + k * 4);
                        const int d = cur - atoms_exch(s_pub + k * 4, cur);
                        if (d != 0) red_add_i32(p.n_k + k, d);
                        const R den = (R)ld_volatile_i32(p.n_k_read + k) + etaW;
                        const R inv = Real<R>::rcp(den);
                        Real<R>::sts(s_den + k * RB, den);
                        Real<R>::sts(s_inv + k * RB, inv);
                        Real<R>::sts(s_f + k * RB, ((R)lds_i(s_ndk + k * 4) + alpha) * inv);
                    }
                    __syncwarp();
                }

                int zn_j = pk_j & 0x1ff;
                if (TPT == 1) {
                    // one token per lane: the lane's own next-batch token is the next group
                    zn_j = sweep_step<TPT, CH, R>(buf, nwk_lane, rowbytes, lim, w_j, pk_j, w_n, pk_n, sm, sub, gbase, lane,
                                               alpha, eta, K, p.nwk_write_delta);
                } else {
                    constexpr int UNR = (TPT <= 8) ? TPT : 2;
#pragma unroll UNR
                    for (int s = 0; s < TPT; ++s) {
                        // the last step of a batch requests the first group of the next batch
                        const bool last = (s + 1 >= TPT);
                        const int src = last ? 0 : s + 1;
                        const int w_x = __shfl_sync(FULL, last ? w_n : w_j, src, TPT);
                        const int pk_x = __shfl_sync(FULL, last ? pk_n : pk_j, src, TPT);
                        const int zn = sweep_step<TPT, CH, R>(buf, nwk_lane, rowbytes, lim, w_c, pk_c, w_x, pk_x, sm, sub,
                                                           gbase, lane, alpha, eta, K, p.nwk_write_delta);
                        if (sub == s) zn_j = zn;
                        w_c = w_x;
                        pk_c = pk_x;
                    }
                }
                if ((pk_j & 0x200) == 0) z_cell[off + lane] = (uint16_t)zn_j;
                if (p.visit_counter) {
                    const unsigned live = __ballot_sync(FULL, (pk_j & 0x200) == 0);
                    if (lane == 0) atomicAdd(p.visit_counter, (unsigned long long)__popc(live));
                }
                off = off_n; more = more_n;
                w_j = w_n; pk_j = pk_n;
                w_n = w_nn; z_n = z_nn; pkp_n = pkp_nn;
                if (!more_n) pn.more = false;
            }
        } else {
            int sidx = warp;  // super-batch index in the rotated sequence
            int off = 0;      // first token of the current batch, relative to the cell
            int nbq = 0;      // batches in the current super-batch
            bool more = sidx < nsuper;
            if (more) {
                const int b0 = sidx < nsup1 ? rot + 4 * sidx : 4 * (sidx - nsup1);
                nbq = min(4, (sidx < nsup1 ? nbatch : rot) - b0);
                off = b0 << 5;
            }
            int w_j = 0, pk_j = 0;
            int w_c = 0, pk_c = 0x200;  // the group sampled by the next step
            Philox4 rnd = {0u, 0u, 0u, 0u};
            int bq = 0;  // batch inside the super-batch
            if (more) {
                const bool valid = (off + lane) < ntok;
                const int ic = valid ? off + lane : ntok - 1;
                w_j = __ldg(tok_cell + ic);
                const uint64_t gi = (gi0 + (uint64_t)off) & p.rng_index_mask;
                rnd = philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), p.sweep, 0u, p.seed_lo, p.seed_hi);
                pk_j = pack_token(rnd.x, (int)z_cell[ic], valid);
                w_c = (TPT > 1) ? __shfl_sync(FULL, w_j, 0, TPT) : w_j;
                pk_c = (TPT > 1) ? __shfl_sync(FULL, pk_j, 0, TPT) : pk_j;
                const char *row = nwk_lane + (uint64_t)((uint32_t)w_c) * rowbytes;
                const int oc = ((pk_c & 0x1ff) >> 2) >> log2_tpt<TPT>();
    #pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int rc = rot_chunk<CH>(c, oc);
                    if (rc < lim) buf[c] = __ldcg(reinterpret_cast<const int4 *>(row + rc * (TPT * 16)));
                }
            }
            int since_refresh = 0;
            while (more) {
                // ---- the batch after this one (its tokens feed the last step's prefetch) ---------------
                int sidx_n = sidx, off_n = off + 32, bq_n = bq + 1;
                if (bq_n == nbq) {
                    bq_n = 0;
                    sidx_n = sidx + nwarps;
                    const int b0 = sidx_n < nsup1 ? rot + 4 * sidx_n : 4 * (sidx_n - nsup1);
                    nbq = min(4, (sidx_n < nsup1 ? nbatch : rot) - b0);
                    off_n = b0 << 5;
                }
                const bool more_n = sidx_n < nsuper;
                int w_n = w_j, pk_n = pk_j | 0x200;
                if (more_n) {
                    const bool valid_n = (off_n + lane) < ntok;
                    const int icn = valid_n ? off_n + lane : ntok - 1;
                    w_n = __ldg(tok_cell + icn);
                    if (bq_n == 0) {  // one Philox call per lane serves the four batches of a super-batch
                        const uint64_t gi = (gi0 + (uint64_t)off_n) & p.rng_index_mask;
                        rnd = philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), p.sweep, 0u, p.seed_lo, p.seed_hi);
                    }
                    const uint32_t bits = bq_n == 0 ? rnd.x : bq_n == 1 ? rnd.y : bq_n == 2 ? rnd.z : rnd.w;
                    pk_n = pack_token(bits, (int)z_cell[icn], valid_n);
                }

                // ---- keep the topic totals live: publish this cell's n_k deltas, re-read n_k ---------
                if (++since_refresh > p.refresh_every) {
                    since_refresh = 1;
                    for (int k = lane; k < K; k += 32) {
                        const int cur = lds_i(s_ndk + k * 4);
                        const int d = cur - atoms_exch(s_pub + k * 4, cur);
                        if (d != 0) red_add_i32(p.n_k + k, d);
                        const R den = (R)ld_volatile_i32(p.n_k_read + k) + etaW;
                        const R inv = Real<R>::rcp(den);
                        Real<R>::sts(s_den + k * RB, den);
                        Real<R>::sts(s_inv + k * RB, inv);
                        Real<R>::sts(s_f + k * RB, ((R)lds_i(s_ndk + k * 4) + alpha) * inv);
                    }
                    __syncwarp();
                }

                int zn_j = pk_j & 0x1ff;
                if (TPT == 1) {
                    // one token per lane: the lane's own next-batch token is the next group
                    zn_j = sweep_step<TPT, CH, R>(buf, nwk_lane, rowbytes, lim, w_j, pk_j, w_n, pk_n, sm, sub, gbase, lane,
                                               alpha, eta, K, p.nwk_write_delta);
                } else {
                    constexpr int UNR = (TPT <= 8) ? TPT : 2;
    #pragma unroll UNR
                    for (int s = 0; s < TPT; ++s) {
                        // the last step of a batch requests the first group of the next batch
                        const bool last = (s + 1 >= TPT);
                        const int src = last ? 0 : s + 1;
                        const int w_x = __shfl_sync(FULL, last ? w_n : w_j, src, TPT);
                        const int pk_x = __shfl_sync(FULL, last ? pk_n : pk_j, src, TPT);
                        const int zn = sweep_step<TPT, CH, R>(buf, nwk_lane, rowbytes, lim, w_c, pk_c, w_x, pk_x, sm, sub,
                                                           gbase, lane, alpha, eta, K, p.nwk_write_delta);
                        if (sub == s) zn_j = zn;
                        w_c = w_x;
                        pk_c = pk_x;
                    }
                }
                if ((pk_j & 0x200) == 0) z_cell[off + lane] = (uint16_t)zn_j;
                if (p.visit_counter) {
                    const unsigned live = __ballot_sync(FULL, (pk_j & 0x200) == 0);
                    if (lane == 0) atomicAdd(p.visit_counter, (unsigned long long)__popc(live));
                }
                sidx = sidx_n; off = off_n; bq = bq_n; more = more_n;
                w_j = w_n; pk_j = pk_n;
            }
        }
        __syncthreads();
        for (int k = tid; k < K; k += blockDim.x) {
            const int cur = lds_i(s_ndk + k * 4);
            ndk_row[k] = cur;
            const int d = cur - lds_i(s_pub + k * 4);
            if (d != 0) red_add_i32(p.n_k + k, d);
        }
        __syncthreads();
    }
}

// ---- deterministic sweep (cgs_model_set_deterministic) --------------------------------------------------------------
// Same sampling step as sweep_kernel (sweep_step), arranged so that NOTHING a warp reads depends on another warp's or
// another CTA's timing:
//   * n_wk rows and n_k come from copies taken when the sub-sweep starts (p.n_wk_read / p.n_k_read); updates go to the live
//     tables as integer reductions, which commute;
//   * a cell is owned by one CTA and worked on in ROUNDS: in a round every warp samples one batch of 32 tokens against its
//     PRIVATE copy of the cell row (counts and f, refreshed from the cell's shared row when the round starts; the warp sees
//     its own moves at once, exactly as in the default kernel), then the warps add what they changed to the shared row
//     with integer atomics; two barriers per round.  What a warp sees of the others is one round old -- the same order of
//     in-cell staleness as the default kernel's steps in flight, but at points fixed by the data, not by the scheduler.
// Hence z after a sweep is a pure function of (z before, tables before, seed), for any grid size or scheduling order.
template <int TPT, int CH, typename R = float>
__global__ void __launch_bounds__(SweepBounds<CH, float, true>::kThreads, sizeof(R) == 8 ? 1 : 2)
sweep_det_kernel(const SweepParams p) {
    using L = CellSmem<TPT, CH, R>;
    constexpr int RB = L::RB;
    constexpr int KS = L::KS;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int kMaxWarps = SweepBounds<CH, float, true>::kThreads / 32;
    constexpr uint32_t REGION = (uint32_t)L::WORDS * 4u;  // one private copy of the cell layout per warp
    extern __shared__ __align__(16) uint32_t det_smem[];   // [nwarps] private regions, then shared ndk[KS], first[KS], slot
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = __shfl_sync(FULL, tid >> 5, 0);
    const int nwarps = blockDim.x >> 5;
    const uint32_t sm0 = smem_u32(det_smem);
    const uint32_t sm = sm0 + (uint32_t)warp * REGION;                 // this warp's private region
    const uint32_t s_shared = sm0 + (uint32_t)nwarps * REGION;         // the cell's row, authoritative between rounds
    const uint32_t s_first = s_shared + KS * 4;                        // the row as loaded (for the n_k update)
    const uint32_t s_slot = s_first + KS * 4;
    (void)kMaxWarps;
    const int sub = lane & (TPT - 1);
    const int gbase = lane & ~(TPT - 1);
    const int K = p.K, Kp = p.Kp;
    const R alpha = sizeof(R) == 8 ? (R)p.alpha64 : (R)p.alpha, eta = sizeof(R) == 8 ? (R)p.eta64 : (R)p.eta;
    const R etaW = sizeof(R) == 8 ? (R)p.etaW64 : (R)p.etaW;
    const uint32_t rowbytes = (uint32_t)Kp * 4u;
    const int lim = (((K + 3) >> 2) - sub + TPT - 1) >> log2_tpt<TPT>();
    const char *nwk_lane = reinterpret_cast<const char *>(p.n_wk_read) + sub * 16;
    int4 buf[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) buf[c] = make_int4(0, 0, 0, 0);

    for (;;) {
        if (tid == 0) sts_i(s_slot, atomicAdd(p.work_counter, 1));
        __syncthreads();
        const int slot = lds_i(s_slot) * p.n_parts + p.part;
        if (slot >= p.n_cells) break;
        const int cell = p.cell_order[slot];
        const int64_t t0 = p.cell_lo[cell], t1 = p.cell_hi[cell];
        int32_t *ndk_row = p.n_dk + (int64_t)cell * Kp;
        for (int k = tid; k < KS; k += blockDim.x) {
            const int nd = k < K ? ndk_row[k] : 0;
            sts_i(s_shared + k * 4, nd);
            sts_i(s_first + k * 4, nd);
        }
        // the totals are those of the copy for the whole cell: every warp keeps them in its own region
        for (int k = lane; k < KS; k += 32) {
            R den = (R)1, inv = (R)0;
            if (k < K) {
                den = (R)p.n_k_read[k] + etaW;
                inv = (R)1 / den;
            }
            Real<R>::sts(sm + L::DEN + k * RB, den);
            Real<R>::sts(sm + L::INV + k * RB, inv);
        }
        __syncthreads();

        const int ntok = (int)(t1 - t0);
        const int nbatch = (ntok + 31) >> 5;
        int rot = 0;
        if (nbatch > 0) {
            uint32_t h = (uint32_t)cell * 0x9E3779B1u ^ p.seed_lo;
            h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
            rot = (int)(h % (uint32_t)nbatch);
        }
        const int32_t *tok_cell = p.tok_region + t0;
        uint16_t *z_cell = p.z + t0;
        const uint64_t gi0 = (uint64_t)(p.token_offset + t0) + (uint64_t)lane;
        const int rounds = (nbatch + nwarps - 1) / nwarps;
        for (int r = 0; r < rounds; ++r) {
            // ---- refresh the private copy of the cell row --------------------------------------------------
            for (int k = lane; k < KS; k += 32) {
                const int nd = lds_i(s_shared + k * 4);
                sts_i(sm + L::NDK + k * 4, nd);
                sts_i(sm + L::PUB + k * 4, nd);
                Real<R>::sts(sm + L::F + k * RB, k < K ? ((R)nd + alpha) * Real<R>::lds(sm + L::INV + k * RB) : (R)0);
            }
            __syncwarp();
            const int b = r * nwarps + warp;  // this warp's batch of the round, in the cell's rotated order
            if (b < nbatch) {
                int bb = b + rot;
                if (bb >= nbatch) bb -= nbatch;
                const int off = bb << 5;
                const bool valid = (off + lane) < ntok;
                const int ic = valid ? off + lane : ntok - 1;
                const int w_j = __ldg(tok_cell + ic);
                const uint64_t gi = (gi0 + (uint64_t)off) & p.rng_index_mask;
                const Philox4 rnd = philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), p.sweep, 1u, p.seed_lo, p.seed_hi);
                const int pk_j = pack_token(rnd.x, (int)z_cell[ic], valid);
                int w_c = (TPT > 1) ? __shfl_sync(FULL, w_j, 0, TPT) : w_j;
                int pk_c = (TPT > 1) ? __shfl_sync(FULL, pk_j, 0, TPT) : pk_j;
                {
                    const char *row = nwk_lane + (uint64_t)((uint32_t)w_c) * rowbytes;
                    const int oc = ((pk_c & 0x1ff) >> 2) >> log2_tpt<TPT>();
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const int rc = rot_chunk<CH>(c, oc);
                        if (rc < lim) buf[c] = __ldcg(reinterpret_cast<const int4 *>(row + rc * (TPT * 16)));
                    }
                }
                int zn_j = pk_j & 0x1ff;
                if (TPT == 1) {
                    zn_j = sweep_step<TPT, CH, R>(buf, nwk_lane, rowbytes, lim, w_j, pk_j, w_j, pk_j | 0x200, sm, sub, gbase,
                                                      lane, alpha, eta, K, p.nwk_write_delta);
                } else {
                    constexpr int UNR = (TPT <= 8) ? TPT : 2;
#pragma unroll UNR
                    for (int s = 0; s < TPT; ++s) {
                        const bool last = (s + 1 >= TPT);
                        const int src = last ? 0 : s + 1;
                        const int w_x = __shfl_sync(FULL, w_j, src, TPT);
                        int pk_x = __shfl_sync(FULL, pk_j, src, TPT);
                        if (last) pk_x |= 0x200;  // nothing to prefetch across the round barrier
                        const int zn = sweep_step<TPT, CH, R>(buf, nwk_lane, rowbytes, lim, w_c, pk_c, w_x, pk_x, sm, sub,
                                                                  gbase, lane, alpha, eta, K, p.nwk_write_delta);
                        if (sub == s) zn_j = zn;
                        w_c = w_x;
                        pk_c = pk_x;
                    }
                }
                if (valid) z_cell[off + lane] = (uint16_t)zn_j;
                if (p.visit_counter) {
                    const unsigned live = __ballot_sync(FULL, valid);
                    if (lane == 0) atomicAdd(p.visit_counter, (unsigned long long)__popc(live));
                }
            }
            __syncthreads();  // every warp has taken its copy of the shared row before anyone changes it ...
            for (int k = lane; k < K; k += 32) {
                const int d = lds_i(sm + L::NDK + k * 4) - lds_i(sm + L::PUB + k * 4);
                if (d != 0) reds_add(s_shared + k * 4, d);
            }
            __syncthreads();  // ... and every change is in before the next round copies it
        }
        for (int k = tid; k < K; k += blockDim.x) {
            const int cur = lds_i(s_shared + k * 4);
            ndk_row[k] = cur;
            const int d = cur - lds_i(s_first + k * 4);
            if (d != 0) red_add_i32(p.n_k + k, d);
        }
        __syncthreads();
    }
}

// cell_mid[d] = index of the first token of cell d whose region is >= region_split (tokens are sorted by region)
__global__ void cell_split_kernel(const int64_t *__restrict__ cell_ptr, const int32_t *__restrict__ tok_region,
                                  int32_t n_cells, int32_t region_split, int64_t *__restrict__ cell_mid) {
    // (one launch per block boundary of cgs_corpus_set_region_blocks)
    const int32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_cells) return;
    int64_t lo = cell_ptr[d], hi = cell_ptr[d + 1];
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (tok_region[mid] < region_split) lo = mid + 1;
        else hi = mid;
    }
    cell_mid[d] = lo;
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
