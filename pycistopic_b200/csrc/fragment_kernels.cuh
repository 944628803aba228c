// fragment_kernels.cuh -- the count matrix upstream of the path: fragments-in-regions counting,
// i.e. `regions.join(fragments)` + `groupby(["Name", "regionID"]).size().unstack()` of
// create_cistopic_object_from_fragments (reference src/pycisTopic/cistopic_class.py:868-889), as a regions x cells
// CSR built on the device.  HBM-bound integer work, bit-exact, no sort library (same plan as corpus_kernels.cuh):
//   (1) fragment_hits_kernel<false>  every fragment finds the regions it overlaps (binary search on the region
//       starts of its chromosome, bounded by the longest region) and counts one hit per region;
//   (2) exclusive scan of the per-region hit counts;
//   (3) fragment_hits_kernel<true>   the same walk scatters the fragment's cell into the region's segment;
//   (4) region_rows_kernel<false/true>  one CTA per region histograms its segment over a shared-memory window of
//       the cell axis (uint32 per cell), counts the distinct cells, and on the second launch emits
//       (cell, count) in ascending cell order -- the CSR row.
// Overlap is pyranges' (half-open intervals): fragment.Start < region.End and region.Start < fragment.End.
#pragma once
#include "common.cuh"

namespace cgs {

struct FragRegions {
    const int32_t *start;         // [R] sorted by (chromosome, start)
    const int32_t *end;           // [R]
    const int32_t *chrom_ptr;     // [n_chrom + 1] region range of every chromosome
    const int32_t *chrom_maxlen;  // [n_chrom] longest region (end - start) of the chromosome
    int32_t n_chrom;
};

template <bool EMIT>
__global__ void fragment_hits_kernel(const int32_t *__restrict__ f_chrom, const int32_t *__restrict__ f_start,
                                     const int32_t *__restrict__ f_end, const int32_t *__restrict__ f_cell,
                                     int64_t n_fragments, int32_t n_cells, FragRegions rg,
                                     unsigned long long *__restrict__ row_fill /* [R] hit counters / cursors */,
                                     const int64_t *__restrict__ raw_ptr, int32_t *__restrict__ raw_cell) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_fragments; i += stride) {
        const int32_t c = f_chrom[i], cell = f_cell[i];
        if (c < 0 || c >= rg.n_chrom || cell < 0 || cell >= n_cells) continue;  // chromosome without regions / barcode not kept
        const int32_t s = f_start[i], e = f_end[i];
        int32_t lo = rg.chrom_ptr[c];
        const int32_t hi = rg.chrom_ptr[c + 1];
        // a region that overlaps ends after s, so it starts after s - maxlen: first region with start > s - maxlen
        const int64_t key = (int64_t)s - (int64_t)rg.chrom_maxlen[c];
        int32_t top = hi;
        while (lo < top) {
            const int32_t mid = lo + ((top - lo) >> 1);
            if ((int64_t)rg.start[mid] <= key) lo = mid + 1; else top = mid;
        }
        for (int32_t j = lo; j < hi && rg.start[j] < e; ++j) {
            if (rg.end[j] > s) {
                const unsigned long long pos = atomicAdd(&row_fill[j], 1ull);
                if (EMIT) raw_cell[raw_ptr[j] + (int64_t)pos] = cell;
            }
        }
    }
}

// One CTA per region (grid-stride).  `window` cells of uint32 counters in shared memory.
// EMIT == false: nnz_row[r] = number of distinct cells hit.   EMIT == true: write the CSR row at out_ptr[r].
template <bool EMIT>
__global__ void __launch_bounds__(512)
region_rows_kernel(const int32_t *__restrict__ raw_cell, const int64_t *__restrict__ raw_ptr, int32_t n_regions,
                   int32_t n_cells, int32_t window, unsigned long long *__restrict__ nnz_row,
                   const int64_t *__restrict__ out_ptr, int32_t *__restrict__ out_cell, int32_t *__restrict__ out_count) {
    extern __shared__ uint32_t frag_hist[];
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t running_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int32_t r = blockIdx.x; r < n_regions; r += gridDim.x) {
        const int64_t b = raw_ptr[r], e = raw_ptr[r + 1];
        if (b == e) {
            if (!EMIT && threadIdx.x == 0) nnz_row[r] = 0ull;
            continue;
        }
        if (threadIdx.x == 0) running_s = 0;
        for (int32_t w0 = 0; w0 < n_cells; w0 += window) {
            const int32_t nbins = min(window, n_cells - w0);
            for (int32_t i = threadIdx.x; i < nbins; i += blockDim.x) frag_hist[i] = 0u;
            __syncthreads();
            for (int64_t j = b + threadIdx.x; j < e; j += blockDim.x) {
                const int32_t off = raw_cell[j] - w0;
                if (off >= 0 && off < nbins) atomicAdd(&frag_hist[off], 1u);
            }
            __syncthreads();
            const int32_t per = (nbins + blockDim.x - 1) / blockDim.x;  // a contiguous slice of bins per thread
            const int32_t bs = min(nbins, (int32_t)threadIdx.x * per), be = min(nbins, bs + per);
            uint32_t cnt = 0;
            for (int32_t i = bs; i < be; ++i) cnt += frag_hist[i] != 0u;
            uint32_t s = cnt;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, s, off);
                if (lane >= off) s += t;
            }
            if (lane == 31) warp_sums[wid] = s;
            __syncthreads();
            if (wid == 0) {
                uint32_t v = lane < nw ? warp_sums[lane] : 0u;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, v, off);
                    if (lane >= off) v += t;
                }
                warp_sums[lane] = v;
            }
            __syncthreads();
            const uint32_t running = running_s;
            if (EMIT) {
                int64_t pos = out_ptr[r] + running + (wid > 0 ? warp_sums[wid - 1] : 0u) + (s - cnt);
                for (int32_t i = bs; i < be; ++i) {
                    const uint32_t h = frag_hist[i];
                    if (h) { out_cell[pos] = w0 + i; out_count[pos] = (int32_t)h; ++pos; }
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) running_s = running + warp_sums[nw - 1];
            __syncthreads();
        }
        if (!EMIT && threadIdx.x == 0) nnz_row[r] = running_s;
        __syncthreads();
    }
}

}  // namespace cgs
