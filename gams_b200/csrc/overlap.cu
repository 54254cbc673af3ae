// overlap.cu — Kernel C: batched range-overlap counting against sorted interval endpoints.
//
// Replaces the per-query walk of locate.rs:111-141: gams::count_rg (src/libs/utils.rs:24-36) ->
// rust_lapper::Lapper::count (BITS): count = lower_bound(starts, stop) - lower_bound(stops, start+1),
// and gams::find_one_idx (utils.rs:7-22) -> Lapper::find(start, stop).next().
//
// Bound: HBM.  Algorithmic traffic 12 B/query (segment, start, stop) + 4 B/result + 8 B/interval.
#include "common.cuh"

namespace gams {

// #elements < key in sorted a[0..n)
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t *__restrict__ a, uint32_t n, uint32_t key) {
    uint32_t lo = 0, len = n;
    while (len > 0) {
        const uint32_t half = len >> 1;
        const bool right = __ldg(a + lo + half) < key;
        lo = right ? lo + half + 1 : lo;
        len = right ? len - half - 1 : half;
    }
    return lo;
}

// v1: one thread per query, both searches in global memory (L2/L1 absorb the top tree levels).
__global__ void __launch_bounds__(256)
count_overlaps_direct_kernel(const uint32_t *__restrict__ starts, const uint32_t *__restrict__ stops,
                             const uint64_t *__restrict__ iv_off, int64_t n_seg, int64_t n_q,
                             const int32_t *__restrict__ q_seg, const uint32_t *__restrict__ q_start,
                             const uint32_t *__restrict__ q_stop, uint32_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_q; q += stride) {
        const int32_t s = q_seg[q];
        uint32_t r = 0;
        if (s >= 0 && s < n_seg) {
            const uint64_t o = iv_off[s];
            const uint32_t n = (uint32_t)(iv_off[s + 1] - o);
            const uint32_t last = lower_bound_u32(starts + o, n, q_stop[q]);
            const uint32_t first = lower_bound_u32(stops + o, n, q_start[q] + 1u);
            r = last - first;
        }
        out[q] = r;
    }
}

int32_t launch_count_overlaps(gams_cuda_count_plan *p, cudaStream_t st, const uint32_t *d_starts,
                              const uint32_t *d_stops, int64_t n_q, const int32_t *d_q_seg,
                              const uint32_t *d_q_start, const uint32_t *d_q_stop, int32_t grouped,
                              uint32_t *d_out) {
    (void)grouped;
    if (n_q == 0) return GAMS_CUDA_OK;
    int64_t blocks = (n_q + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    count_overlaps_direct_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_starts, d_stops, p->d_iv_off, p->n_seg, n_q,
                                                                   d_q_seg, d_q_start, d_q_stop, d_out);
    ++p->launches;
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

// Lapper::find(start, stop).next(): intervals of a chromosome sorted by (start, stop);
// off = lower_bound(iv.start, start - max_len saturating); scan forward; first iv with
// iv.start < stop && iv.stop > start; give up once iv.start >= stop.
__global__ void __launch_bounds__(256)
locate_ctg_kernel(int64_t n_chr, const uint32_t *__restrict__ ctg_start, const uint32_t *__restrict__ ctg_stop,
                  const uint64_t *__restrict__ chr_off, const uint32_t *__restrict__ chr_maxlen, int64_t n_q,
                  const int32_t *__restrict__ q_chr, const uint32_t *__restrict__ q_start,
                  const uint32_t *__restrict__ q_stop, int32_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_q; q += stride) {
        const int32_t c = q_chr[q];
        int32_t r = -1;
        if (c >= 0 && c < n_chr) {
            const uint64_t o = chr_off[c];
            const uint32_t n = (uint32_t)(chr_off[c + 1] - o);
            const uint32_t qs = q_start[q], qe = q_stop[q];
            const uint32_t ml = chr_maxlen[c];
            const uint32_t key = qs >= ml ? qs - ml : 0u;
            for (uint32_t k = lower_bound_u32(ctg_start + o, n, key); k < n; ++k) {
                const uint32_t s = ctg_start[o + k];
                if (s < qe && ctg_stop[o + k] > qs) {
                    r = (int32_t)(o + k);
                    break;
                }
                if (s >= qe) break;
            }
        }
        out[q] = r;
    }
}

int32_t launch_locate_ctg(cudaStream_t st, int64_t n_chr, const uint32_t *d_ctg_start, const uint32_t *d_ctg_stop,
                          const uint64_t *d_chr_off, const uint32_t *d_chr_maxlen, int64_t n_q,
                          const int32_t *d_q_chr, const uint32_t *d_q_start, const uint32_t *d_q_stop,
                          int32_t *d_out) {
    if (n_q == 0) return GAMS_CUDA_OK;
    int64_t blocks = (n_q + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    locate_ctg_kernel<<<(unsigned)blocks, 256, 0, st>>>(n_chr, d_ctg_start, d_ctg_stop, d_chr_off, d_chr_maxlen, n_q,
                                                        d_q_chr, d_q_start, d_q_stop, d_out);
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

}  // namespace gams
