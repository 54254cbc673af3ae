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
// Instead every sorted endpoint array gets direct-address rank nodes, built once per index
// (build_rank_kernel): the segment's coordinate range [lo, hi] is cut into equal power-of-two buckets and
// bucket b = (key - lo) >> shift owns one 32-byte node holding the count of elements before it and the
// bucket's elements as 16-bit offsets.  lower_bound(key) reads ONE sector per endpoint, independent of n;
// no sorting or bucketing of the queries is needed and results are written straight in query order.
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

// ---- rank nodes -----------------------------------------------------------------------------------
// One 32-byte node (one DRAM sector) per bucket: cum = #elements of the segment before the bucket, then up
// to 14 sorted 16-bit offsets (element - bucket base; 0xFFFF = empty slot).  lower_bound(key) = cum +
// #offsets < key - base: ONE sector per endpoint.  Buckets are 2^shift wide with shift <= 15 chosen so a
// segment of n elements has at most n/4 + 1 of them (4..8 elements per bucket on average; the table is
// 8 B/element at worst).  A bucket with more than 14 elements sets NODE_OVERFLOW and is searched in the raw
// array between its cum and the next node's; a segment too sparse for 16-bit offsets (shift would exceed
// 15) sets shift = RANK_PLAIN and is binary-searched directly.
// Nodes of segment s start at (iv_off[s] >> 2) + 2 s; the node after the last bucket holds cum = n.
constexpr uint32_t NODE_OVERFLOW = 0x80000000u;
constexpr uint32_t RANK_PLAIN = 0xFFu;

struct alignas(32) RankNode {
    uint32_t cum;
    uint16_t key[14];
};

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
    p.pad2[0] = p.pad2[1] = 0;
    for (int which = 0; which < 2; ++which) {
        const uint32_t *a = (which ? stops : starts) + o;
        uint32_t lo = 0, hi = 0, shift = 0;
        if (n) {
            lo = a[0];
            hi = a[n - 1];
            const uint64_t max_buckets = (uint64_t)(n >> 2) + 1;
            while ((((uint64_t)(hi - lo)) >> shift) + 1 > max_buckets) ++shift;
            if (shift > 15) shift = RANK_PLAIN;
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
                  uint64_t n_nodes, RankNode *__restrict__ node_starts, RankNode *__restrict__ node_stops) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_nodes; e += stride) {
        int64_t lo = 0, hi = n_seg;  // last segment whose node range starts at or before e
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if ((iv_off[mid] >> 2) + 2 * (uint64_t)mid <= e) lo = mid; else hi = mid;
        }
        const SegRank p = par[lo];
        const uint64_t b = e - ((p.off >> 2) + 2 * (uint64_t)lo);
        for (int which = 0; which < 2; ++which) {
            RankNode nd;
            nd.cum = p.n;
#pragma unroll
            for (int k = 0; k < 14; ++k) nd.key[k] = 0xFFFFu;
            const uint32_t sh = p.shift[which];
            if (p.n && sh != RANK_PLAIN) {
                const uint32_t *a = (which ? stops : starts) + p.off;
                const uint64_t edge0 = (uint64_t)p.lo[which] + (b << sh), edge1 = edge0 + (1ull << sh);
                const uint32_t j0 = (edge0 > (uint64_t)p.hi[which]) ? p.n : lower_bound_u32(a, p.n, (uint32_t)edge0);
                const uint32_t j1 = (edge1 > (uint64_t)p.hi[which]) ? p.n : lower_bound_u32(a, p.n, (uint32_t)edge1);
                nd.cum = j0;
                if (j1 - j0 > 14u) nd.cum |= NODE_OVERFLOW;
                else
                    for (uint32_t j = j0; j < j1; ++j) nd.key[j - j0] = (uint16_t)(a[j] - (uint32_t)edge0);
            }
            (which ? node_stops : node_starts)[e] = nd;
        }
    }
}

// One 32-byte node in one instruction: Blackwell's 256-bit global load (LDG.E.256, PTX ld.global.v8.b32,
// sm_100+), read-only path, not allocated in L1 (each node is used once per query).
struct Node8 { uint32_t w[8]; };
__device__ __forceinline__ Node8 ld_node(const void *p) {
    Node8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}

// #elements < key through the rank nodes of one endpoint array of one segment
__device__ __forceinline__ uint32_t ranked_lower_bound(const uint32_t *__restrict__ a, const RankNode *__restrict__ nodes,
                                                       uint32_t n, uint32_t lo, uint32_t hi, uint32_t shift, uint32_t key) {
    if (n == 0 || key <= lo) return 0;
    if (key > hi) return n;
    if (shift == RANK_PLAIN) return lower_bound_u32(a, n, key);
    const uint32_t d = key - lo, b = d >> shift, koff = d & ((1u << shift) - 1u);
    const Node8 nd = ld_node(nodes + b);
    if (nd.w[0] & NODE_OVERFLOW) {
        const uint32_t j0 = nd.w[0] & ~NODE_OVERFLOW, j1 = __ldg(&nodes[b + 1].cum) & ~NODE_OVERFLOW;
        return j0 + lower_bound_u32(a + j0, j1 - j0, key);
    }
    // Two 16-bit offsets per word, koff <= 0x7FFF, valid offsets <= 0x7FFF, empty slots 0xFFFF:
    // (h | 0x8000) - koff keeps bit 15 exactly when h >= koff and never borrows across the halves.
    const uint32_t kk = koff * 0x00010001u;
    uint32_t ge = 0;  // bit (15 - k) / (31 - k): offset of word k is >= koff
#pragma unroll
    for (int k = 0; k < 7; ++k) ge |= (((nd.w[k + 1] | 0x80008000u) - kk) & 0x80008000u) >> k;
    return nd.w[0] + 14u - (uint32_t)__popc(ge);
}

__global__ void __launch_bounds__(256)
count_overlaps_ranked_kernel(const uint32_t *__restrict__ starts, const uint32_t *__restrict__ stops,
                             const SegRank *__restrict__ par, const RankNode *__restrict__ node_starts,
                             const RankNode *__restrict__ node_stops, int64_t n_seg, int64_t n_q,
                             const int32_t *__restrict__ q_seg, const uint32_t *__restrict__ q_start,
                             const uint32_t *__restrict__ q_stop, uint32_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_q; q += stride) {
        const int32_t s = q_seg[q];
        const uint32_t qs = q_start[q], qe = q_stop[q];
        uint32_t r = 0;
        if (s >= 0 && s < n_seg) {
            const uint4 *pp = reinterpret_cast<const uint4 *>(par + s);
            const uint4 p0 = __ldg(pp), p1 = __ldg(pp + 1), p2 = __ldg(pp + 2);  // off,n,pad | lo0 lo1 hi0 hi1 | sh0 sh1 - -
            const uint64_t off = (uint64_t)p0.x | ((uint64_t)p0.y << 32);
            const uint32_t n = p0.z;
            const uint64_t t = (off >> 2) + 2 * (uint64_t)s;
            const uint32_t last = ranked_lower_bound(starts + off, node_starts + t, n, p1.x, p1.z, p2.x, qe);
            const uint32_t first = ranked_lower_bound(stops + off, node_stops + t, n, p1.y, p1.w, p2.y, qs + 1u);
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
    const uint64_t n_entries = (p->total_iv >> 2) + 2 * (uint64_t)p->n_seg + 2;
    int32_t rc;
    if ((rc = p->tmp[0].reserve(sizeof(SegRank) * (size_t)std::max<int64_t>(p->n_seg, 1)))) return rc;
    if ((rc = p->tmp[1].reserve(sizeof(RankNode) * n_entries))) return rc;
    if ((rc = p->tmp[2].reserve(sizeof(RankNode) * n_entries))) return rc;
    if (p->n_seg) {
        rank_params_kernel<<<(unsigned)((p->n_seg + 255) / 256), 256, 0, st>>>(d_starts, d_stops, p->d_iv_off, p->n_seg,
                                                                              p->tmp[0].as<SegRank>());
        uint64_t blocks = (n_entries + 255) / 256;
        if (blocks > 148ull * 32) blocks = 148ull * 32;
        build_rank_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_starts, d_stops, p->d_iv_off, p->n_seg, p->tmp[0].as<SegRank>(),
                                                            n_entries, p->tmp[1].as<RankNode>(), p->tmp[2].as<RankNode>());
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
                                                                       p->tmp[1].as<RankNode>(), p->tmp[2].as<RankNode>(),
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
