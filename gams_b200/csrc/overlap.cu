// overlap.cu — Kernel C: batched range-overlap counting against sorted interval endpoints.
//
// Replaces the per-query walk of locate.rs:111-141: gams::count_rg (src/libs/utils.rs:24-36) ->
// rust_lapper::Lapper::count (BITS): count = lower_bound(starts, stop) - lower_bound(stops, start+1),
// and gams::find_one_idx (utils.rs:7-22) -> Lapper::find(start, stop).next().
//
// Bound: HBM.  Algorithmic traffic 12 B/query (segment, start, stop) + 4 B/result + 8 B/interval.
//
// Design.  A binary search over a 16 K-entry array costs 14 dependent probes per endpoint; on 10^8
// queries in arbitrary order that is 125 GB of DRAM sectors for 2.4 GB of useful bytes (profiles/r1a).
// Instead every sorted endpoint array gets a direct-address rank table, built once per index
// (build_rank_kernel): the segment's coordinate range [lo, hi] is cut into <= n equal power-of-two
// buckets and cum[b] = #elements < lo + (b << shift).  lower_bound(key) is then cum[b] plus a scan of the
// one or two elements of bucket b = (key - lo) >> shift: two sectors per endpoint, independent of n,
// no sorting or bucketing of the queries, results written straight in query order.
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

// ---- rank tables --------------------------------------------------------------------------------
// Table of segment s starts at entry iv_off[s] + 2 s and has nb + 1 <= n + 1 entries.
__global__ void __launch_bounds__(256)
rank_params_kernel(const uint32_t *__restrict__ starts, const uint32_t *__restrict__ stops,
                   const uint64_t *__restrict__ iv_off, int64_t n_seg, SegRank *__restrict__ par) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const uint64_t o = iv_off[s];
    const uint32_t n = (uint32_t)(iv_off[s + 1] - o);
    SegRank p;
    p.off = o;
    p.n = n;
    p.pad = 0;
    for (int which = 0; which < 2; ++which) {
        const uint32_t *a = (which ? stops : starts) + o;
        uint32_t lo = 0, hi = 0, shift = 0;
        if (n) {
            lo = a[0];
            hi = a[n - 1];
            while ((((uint64_t)(hi - lo)) >> shift) + 1 > (uint64_t)n) ++shift;  // at most n buckets
        }
        p.lo[which] = lo;
        p.hi[which] = hi;
        p.shift[which] = shift;
    }
    par[s] = p;
}

__global__ void __launch_bounds__(256)
build_rank_kernel(const uint32_t *__restrict__ starts, const uint32_t *__restrict__ stops,
                  const uint64_t *__restrict__ iv_off, int64_t n_seg, const SegRank *__restrict__ par,
                  uint64_t n_entries, uint32_t *__restrict__ cum_starts, uint32_t *__restrict__ cum_stops) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_entries; e += stride) {
        int64_t lo = 0, hi = n_seg;  // last segment with iv_off[s] + 2 s <= e
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (iv_off[mid] + 2 * (uint64_t)mid <= e) lo = mid; else hi = mid;
        }
        const SegRank p = par[lo];
        const uint64_t b = e - (p.off + 2 * (uint64_t)lo);
        for (int which = 0; which < 2; ++which) {
            const uint32_t *a = (which ? stops : starts) + p.off;
            const uint64_t edge = (uint64_t)p.lo[which] + (b << p.shift[which]);
            const uint32_t v = (p.n == 0) ? 0u : (edge > (uint64_t)p.hi[which]) ? p.n : lower_bound_u32(a, p.n, (uint32_t)edge);
            (which ? cum_stops : cum_starts)[e] = v;
        }
    }
}

// #elements < key through the rank table
__device__ __forceinline__ uint32_t ranked_lower_bound(const uint32_t *__restrict__ a, const uint32_t *__restrict__ cum,
                                                       uint32_t n, uint32_t lo, uint32_t hi, uint32_t shift, uint32_t key) {
    if (n == 0 || key <= lo) return 0;
    if (key > hi) return n;
    const uint32_t b = (key - lo) >> shift;
    const uint2 c = make_uint2(__ldg(cum + b), __ldg(cum + b + 1));
    uint32_t j = c.x;
    if (c.y - c.x <= 8u) {
        uint32_t r = c.x;
        for (; j < c.y; ++j) r += (__ldg(a + j) < key);
        return r;
    }
    return c.x + lower_bound_u32(a + c.x, c.y - c.x, key);
}

__global__ void __launch_bounds__(256)
count_overlaps_ranked_kernel(const uint32_t *__restrict__ starts, const uint32_t *__restrict__ stops,
                             const SegRank *__restrict__ par, const uint32_t *__restrict__ cum_starts,
                             const uint32_t *__restrict__ cum_stops, int64_t n_seg, int64_t n_q,
                             const int32_t *__restrict__ q_seg, const uint32_t *__restrict__ q_start,
                             const uint32_t *__restrict__ q_stop, uint32_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_q; q += stride) {
        const int32_t s = q_seg[q];
        uint32_t r = 0;
        if (s >= 0 && s < n_seg) {
            const SegRank p = par[s];
            const uint64_t t = p.off + 2 * (uint64_t)s;
            const uint32_t last = ranked_lower_bound(starts + p.off, cum_starts + t, p.n, p.lo[0], p.hi[0], p.shift[0], q_stop[q]);
            const uint32_t first = ranked_lower_bound(stops + p.off, cum_stops + t, p.n, p.lo[1], p.hi[1], p.shift[1], q_start[q] + 1u);
            r = last - first;
        }
        out[q] = r;
    }
}

// v1 (kept for A/B measurements, GAMS_CUDA_COUNT_MODE=1): one thread per query, two binary searches.
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

int32_t build_count_index(gams_cuda_count_plan *p, cudaStream_t st, const uint32_t *d_starts, const uint32_t *d_stops) {
    const uint64_t n_entries = p->total_iv + 2 * (uint64_t)p->n_seg;
    int32_t rc;
    if ((rc = p->tmp[0].reserve(sizeof(SegRank) * (size_t)std::max<int64_t>(p->n_seg, 1)))) return rc;
    if ((rc = p->tmp[1].reserve(sizeof(uint32_t) * std::max<uint64_t>(n_entries, 1)))) return rc;
    if ((rc = p->tmp[2].reserve(sizeof(uint32_t) * std::max<uint64_t>(n_entries, 1)))) return rc;
    if (p->n_seg) {
        rank_params_kernel<<<(unsigned)((p->n_seg + 255) / 256), 256, 0, st>>>(d_starts, d_stops, p->d_iv_off, p->n_seg,
                                                                              p->tmp[0].as<SegRank>());
        uint64_t blocks = (n_entries + 255) / 256;
        if (blocks > 148ull * 32) blocks = 148ull * 32;
        build_rank_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_starts, d_stops, p->d_iv_off, p->n_seg, p->tmp[0].as<SegRank>(),
                                                            n_entries, p->tmp[1].as<uint32_t>(), p->tmp[2].as<uint32_t>());
        p->launches += 2;
        CUDA_TRY(cudaGetLastError());
    }
    p->idx_starts = d_starts;
    p->idx_stops = d_stops;
    p->idx_valid = true;
    return GAMS_CUDA_OK;
}

int32_t launch_count_overlaps(gams_cuda_count_plan *p, cudaStream_t st, const uint32_t *d_starts,
                              const uint32_t *d_stops, int64_t n_q, const int32_t *d_q_seg,
                              const uint32_t *d_q_start, const uint32_t *d_q_stop, int32_t grouped,
                              uint32_t *d_out) {
    (void)grouped;  // the ranked kernel needs no ordering; kept in the ABI as a promise callers may give
    if (n_q == 0) return GAMS_CUDA_OK;
    int64_t blocks = (n_q + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (p->mode == 1) {
        count_overlaps_direct_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_starts, d_stops, p->d_iv_off, p->n_seg, n_q,
                                                                       d_q_seg, d_q_start, d_q_stop, d_out);
    } else {
        if (!p->idx_valid || p->idx_starts != d_starts || p->idx_stops != d_stops) {
            int32_t rc = build_count_index(p, st, d_starts, d_stops);
            if (rc) return rc;
        }
        count_overlaps_ranked_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_starts, d_stops, p->tmp[0].as<SegRank>(),
                                                                       p->tmp[1].as<uint32_t>(), p->tmp[2].as<uint32_t>(),
                                                                       p->n_seg, n_q, d_q_seg, d_q_start, d_q_stop, d_out);
    }
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
