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
#include "overlap_index.cuh"

namespace gams {

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

// ---- staged kernel: queries grouped by segment ------------------------------------------------------
// When the caller promises q_seg is non-decreasing (grouped != 0: the order the reference processes ranges in,
// per ctg out of read_range's BTreeMap, locate.rs:104-141) every probe becomes a shared-memory access.  Each
// sorted endpoint array of each segment has a COMPACT form, built once per index (build_compact_kernel):
//   cum2[b] (u32) = cum[b] | cum[b+1] << 16,  cum[b] = #elements < lo + (b << l),   b = 0 .. nb-1 (width 2^l, l <= 7)
//   low[i]  (u8)  = (a[i] - lo) & (2^l - 1),                                         i = 0 .. n-1
//   lower_bound(key) = cum[b] + #{ i in [cum[b], cum[b+1]) : low[i] < rem },  b = (key - lo) >> l, rem = (key - lo) & (2^l - 1)
// about 4 B per bucket + 1 B per element (~3 B per interval endpoint instead of 4 B raw): a 16 K-interval ctg is
// 94 KB for both arrays, so two CTAs per SM each hold one segment.  A work item is (segment, chunk of <= ST_CHUNK of
// its queries): the CTA copies the segment's compact form into shared memory (coalesced 16-byte loads) and answers
// its queries with one cum2 word and an 8-byte window of low[] per endpoint, compared SIMD-in-register.
// DRAM traffic: the query stream once + ~6 B per interval (neighbouring chunks of a segment hit L2).
// Arrays that do not fit the form (l > 7, n > ST_MAX_N) are binary-searched in global memory by the same CTA.
// one CTA per segment: parameters and compact form of both endpoint arrays
__global__ void __launch_bounds__(256)
build_compact_kernel(const uint32_t *__restrict__ starts, const uint32_t *__restrict__ stops,
                     const uint64_t *__restrict__ iv_off, int64_t n_seg, const uint64_t *__restrict__ cofs,
                     SegCompact *__restrict__ par, uint8_t *__restrict__ buf, uint32_t *__restrict__ max_bytes) {
    for (int64_t s = blockIdx.x; s < n_seg; s += gridDim.x) {
        const uint64_t off = iv_off[s];
        const uint32_t n = (uint32_t)min((uint64_t)0xFFFFFFFFu, iv_off[s + 1] - off);
        // point intervals (rg `chr:p` -> [p, p+1), redis.rs:291-295): the stops are the starts shifted by one, one table serves both
        int same = 1;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) same &= (stops[off + i] == starts[off + i] + 1u) ? 1 : 0;
        const bool alias = __syncthreads_and(same) != 0 && n > 0;
        uint32_t seg_bytes = 0;
        bool par_ok_a = false;
        if (alias) {
            // E8: one 8-byte entry per bucket = cum (15 bits) | overflow flag (bit 15) | up to six low values (0xFF = empty);
            // both endpoints of a query are answered by ONE 64-bit shared-memory load each
            const uint32_t *a = starts + off;
            SegCompact p;
            p.n = n, p.lo = a[0], p.hi = a[n - 1], p.fmt = 1, p.low_off = 0;
            const uint32_t range = p.hi - p.lo, cap = st_nb_cap(n);
            uint32_t l = 0;
            while (l < 32u && (range >> l) + 1u > cap) ++l;
            if (n <= (uint32_t)ST_MAX_N && l <= 7) {
                p.l = l;
                p.nb = (range >> l) + 1u;
                p.bytes = (8u * (p.nb + 1u) + 15u) & ~15u;
                uint2 *ent = reinterpret_cast<uint2 *>(buf + cofs[s]);
                uint16_t *ent16 = reinterpret_cast<uint16_t *>(ent);
                uint8_t *ent8 = reinterpret_cast<uint8_t *>(ent);
                for (uint32_t b = threadIdx.x; b < p.bytes / 8u; b += blockDim.x) ent[b] = make_uint2(0xFFFF0000u | n, 0xFFFFFFFFu);
                __syncthreads();
                const uint32_t mask = (1u << l) - 1u;
                for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {          // cum of every bucket
                    const uint32_t b = (a[i] - p.lo) >> l;
                    const uint32_t pb = i ? ((a[i - 1] - p.lo) >> l) + 1u : 0u;
                    for (uint32_t bb = pb; bb <= b; ++bb) ent16[4 * bb] = (uint16_t)i;
                }
                __syncthreads();                                                  // (entries past the last bucket keep cum = n)
                for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {          // low values into the bucket's slots
                    const uint32_t d = a[i] - p.lo, b = d >> l, slot = i - (ent16[4 * b] & 0x7FFFu);
                    if (slot < 6u) ent8[8 * b + 2 + slot] = (uint8_t)(d & mask);
                }
                __syncthreads();
                for (uint32_t b = threadIdx.x; b < p.nb; b += blockDim.x)         // more than six elements: flag the bucket
                    if ((uint32_t)(ent16[4 * b + 4] & 0x7FFFu) - (uint32_t)(ent16[4 * b] & 0x7FFFu) > 6u) ent16[4 * b] |= 0x8000u;
                if (threadIdx.x == 0) {
                    par[2 * s] = p;
                    p.l = ST_ALIAS;
                    par[2 * s + 1] = p;
                    atomicMax(max_bytes, p.bytes);
                }
                __syncthreads();
                continue;
            }
        }
        for (int which = 0; which < 2; ++which) {
            const uint32_t *a = (which ? stops : starts) + off;
            SegCompact p;
            p.n = n;
            p.lo = n ? a[0] : 0u;
            p.hi = n ? a[n - 1] : 0u;
            p.fmt = 0;
            const uint32_t range = p.hi - p.lo, cap = st_nb_cap(n);
            uint32_t l = 0;
            while (l < 32u && (range >> l) + 1u > cap) ++l;
            const bool ok = n > 0 && n <= (uint32_t)ST_MAX_N && l <= 7;
            p.l = ok ? l : ST_PLAIN;
            p.nb = ok ? (range >> l) + 1u : 0u;
            p.low_off = ok ? ((4u * (p.nb + 1u) + 15u) & ~15u) : 0u;
            p.bytes = ok ? p.low_off + ((n + 16u + 15u) & ~15u) : 0u;
            if (which == 1 && alias && par_ok_a) {   // the stops go through the starts' table
                p.l = ST_ALIAS;
                p.nb = 0, p.low_off = 0, p.bytes = 0;
                if (threadIdx.x == 0) par[2 * s + 1] = p;
                continue;
            }
            if (which == 0) par_ok_a = ok;
            seg_bytes += p.bytes;
            if (threadIdx.x == 0) par[2 * s + which] = p;
            if (!ok) continue;
            uint8_t *base = buf + cofs[s] + (which ? st_arr_cap_bytes(n) : 0);
            uint16_t *cum16 = reinterpret_cast<uint16_t *>(base);     // cum[b] lives in halves 2b and 2b-1
            uint8_t *low = base + p.low_off;
            const uint32_t mask = (1u << l) - 1u;
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                const uint32_t d = a[i] - p.lo, b = d >> l;
                low[i] = (uint8_t)(d & mask);
                const uint32_t pb = i ? ((a[i - 1] - p.lo) >> l) + 1u : 0u;   // buckets (b_prev, b] start at element i
                for (uint32_t bb = pb; bb <= b; ++bb) {
                    cum16[2 * bb] = (uint16_t)i;
                    if (bb) cum16[2 * bb - 1] = (uint16_t)i;
                }
                if (i == n - 1) {                                             // b == nb - 1: cum[nb] = n
                    cum16[2 * p.nb - 1] = (uint16_t)n;
                    cum16[2 * p.nb] = (uint16_t)n;
                    cum16[2 * p.nb + 1] = (uint16_t)n;
                }
            }
            for (uint32_t i = n + threadIdx.x; i < ((n + 16u + 15u) & ~15u); i += blockDim.x) low[i] = 0xFFu;  // never < rem
        }
        if (threadIdx.x == 0) atomicMax(max_bytes, seg_bytes);
        __syncthreads();
    }
}

// first query of every segment in the grouped stream: qoff[s] = #queries with q_seg < s, s = 0 .. n_seg.
// Four consecutive queries per thread (one 16-byte load when q_seg is 16-byte aligned); a boundary is where the
// clamped segment id changes.
__device__ __forceinline__ void seg_bound_step(int64_t q, int32_t prev, int32_t cur, int64_t n_seg, int64_t *__restrict__ qoff) {
    if (prev != cur) {
        const int64_t a = min((int64_t)max(prev, -1), n_seg), b = min((int64_t)max(cur, -1), n_seg);
        for (int64_t s = a + 1; s <= b; ++s) qoff[s] = q;
    }
}
__global__ void __launch_bounds__(256)
seg_query_bounds_kernel(const int32_t *__restrict__ q_seg, int64_t n_q, int64_t n_seg, int64_t *__restrict__ qoff) {
    const bool vec = (reinterpret_cast<uintptr_t>(q_seg) & 15u) == 0;
    const int64_t n4 = (n_q + 3) >> 2, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += stride) {
        const int64_t q = g << 2;
        int32_t v[4];
        if (vec && q + 4 <= n_q) {
            const int4 t = __ldg(reinterpret_cast<const int4 *>(q_seg + q));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = q + k < n_q ? q_seg[q + k] : 0x7FFFFFFF;   // past the end: "segment" n_seg
        }
        const int32_t prev = q ? __ldg(q_seg + q - 1) : -1;
        seg_bound_step(q, prev, v[0], n_seg, qoff);
#pragma unroll
        for (int k = 1; k < 4; ++k) seg_bound_step(q + k, v[k - 1], v[k], n_seg, qoff);
        if (q + 4 >= n_q && q + 4 > 0) {   // the thread holding the last query closes the table
            const int32_t last = v[(n_q - 1) & 3];
            if (((n_q - 1) >> 2) == g) seg_bound_step(n_q, last, 0x7FFFFFFF, n_seg, qoff);
        }
    }
    if (n_q == 0 && blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t s = 0; s <= n_seg; ++s) qoff[s] = 0;
}

// work items: segment s is cut into ceil(nq_s / ST_CHUNK) items; ioff[s] = first item of segment s, item_seg[i] = segment
// of item i.  One CTA: every thread takes a contiguous run of segments (all its loads are in flight at once), one block scan.
constexpr int SI_THREADS = 1024;
constexpr int SI_PER = 8;            // segments per thread and round
__global__ void __launch_bounds__(SI_THREADS)
seg_items_kernel(const int64_t *__restrict__ qoff, int64_t n_seg, int64_t *__restrict__ ioff, int32_t *__restrict__ item_seg) {
    __shared__ int64_t wsum[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_seg; base += (int64_t)SI_THREADS * SI_PER) {
        const int64_t s0 = base + (int64_t)threadIdx.x * SI_PER;
        int64_t b[SI_PER + 1];
#pragma unroll
        for (int k = 0; k <= SI_PER; ++k) b[k] = qoff[min(s0 + k, n_seg)];
        int64_t v[SI_PER], tot = 0;
#pragma unroll
        for (int k = 0; k < SI_PER; ++k) {
            v[k] = (b[k + 1] - b[k] + ST_CHUNK - 1) / ST_CHUNK;   // 0 past the last segment
            tot += v[k];
        }
        int64_t inc = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if ((threadIdx.x & 31) >= d) inc += t;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
        __syncthreads();
        int64_t before = carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += wsum[w];
        int64_t at = before + inc - tot;
#pragma unroll
        for (int k = 0; k < SI_PER; ++k) {
            if (s0 + k < n_seg) {
                ioff[s0 + k] = at;
                for (int64_t i = 0; i < v[k]; ++i) item_seg[at + i] = (int32_t)(s0 + k);
            }
            at += v[k];
        }
        __syncthreads();
        if (threadIdx.x == SI_THREADS - 1) carry = before + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) ioff[n_seg] = carry;
}

#ifndef ST_MIN_CTAS
#define ST_MIN_CTAS 2
#endif
__global__ void __launch_bounds__(ST_THREADS, ST_MIN_CTAS)
count_overlaps_staged_kernel(const uint32_t *__restrict__ starts, const uint32_t *__restrict__ stops,
                             const uint64_t *__restrict__ iv_off, int64_t n_seg, int64_t n_q,
                             const int64_t *__restrict__ qoff, const int64_t *__restrict__ ioff,
                             const int32_t *__restrict__ item_seg, const SegCompact *__restrict__ par, const uint64_t *__restrict__ cofs,
                             const uint8_t *__restrict__ cbuf, const uint32_t *__restrict__ q_start,
                             const uint32_t *__restrict__ q_stop, uint32_t *__restrict__ out) {
    extern __shared__ __align__(16) uint8_t st_smem[];
    const int64_t n_items = ioff[n_seg];
    const int64_t item = blockIdx.x;
    if (item >= n_items) {
        // the CTAs past the last item zero the results of queries outside every segment (utils.rs:29-32: count 0)
        const int64_t v_lo = qoff[0], v_hi = qoff[n_seg], n_bad = v_lo + (n_q - v_hi);
        const int64_t stride = (int64_t)(gridDim.x - n_items) * ST_THREADS;
        for (int64_t k = (item - n_items) * ST_THREADS + threadIdx.x; k < n_bad; k += stride) out[k < v_lo ? k : v_hi + (k - v_lo)] = 0u;
        return;
    }
    const int64_t s = item_seg[item];
    const int64_t q0 = qoff[s] + (item - ioff[s]) * ST_CHUNK, q1 = min(q0 + (int64_t)ST_CHUNK, qoff[s + 1]);
    const uint64_t off = iv_off[s];
    SegCompact A = par[2 * s], B = par[2 * s + 1];
    // staging copies A.bytes + B.bytes; a handful of queries is cheaper by plain binary search (2 * 14 probes each)
    if ((uint64_t)(q1 - q0) * 64u < (uint64_t)A.bytes + B.bytes) A.l = B.l = ST_PLAIN;
    const bool alias = B.l == ST_ALIAS;   // lower_bound(stops, k) = lower_bound(starts, k - 1)
    if (alias) B = A;
    uint8_t *smA = st_smem, *smB = st_smem + ((A.l == ST_PLAIN || alias) ? 0u : A.bytes);
    {
        const uint8_t *gA = cbuf + cofs[s], *gB = gA + st_arr_cap_bytes(A.n);
        if (A.l != ST_PLAIN)
            for (uint32_t i = threadIdx.x * 16u; i < A.bytes; i += ST_THREADS * 16u)
                *reinterpret_cast<uint4 *>(smA + i) = __ldg(reinterpret_cast<const uint4 *>(gA + i));
        if (B.l != ST_PLAIN && !alias)
            for (uint32_t i = threadIdx.x * 16u; i < B.bytes; i += ST_THREADS * 16u)
                *reinterpret_cast<uint4 *>(smB + i) = __ldg(reinterpret_cast<const uint4 *>(gB + i));
    }
    __syncthreads();
    if (A.n == 0) {   // an empty index counts nothing (and has no lo / hi to clamp to)
        for (int64_t q = q0 + threadIdx.x; q < q1; q += ST_THREADS) out[q] = 0u;
        return;
    }
    constexpr int U = 4;
    const uint32_t nq = (uint32_t)(q1 - q0);
    if (alias && A.fmt == 1u) {
        const uint2 *ent = reinterpret_cast<const uint2 *>(smA);
        const uint32_t *__restrict__ a = starts + off;
        // whole groups of U * ST_THREADS queries: fixed offsets from three running pointers, no bounds checks
        const uint32_t *__restrict__ ps = q_start + q0 + threadIdx.x, *__restrict__ pe = q_stop + q0 + threadIdx.x;
        uint32_t *__restrict__ po = out + q0 + threadIdx.x;
        uint32_t left = nq;
        for (; left >= (uint32_t)(U * ST_THREADS); left -= U * ST_THREADS, ps += U * ST_THREADS, pe += U * ST_THREADS, po += U * ST_THREADS) {
            uint32_t qs[U], qe[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                qs[u] = __ldg(ps + u * ST_THREADS);
                qe[u] = __ldg(pe + u * ST_THREADS);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                // lower_bound(stops, qs + 1) = lower_bound(starts, qs); qs + 1 wrapping to 0 counts nothing, as in the general path
                const uint32_t last = e8_lower_bound(A, a, ent, qe[u]);
                const uint32_t first = (qs[u] + 1u) ? e8_lower_bound(A, a, ent, qs[u]) : 0u;
                po[u * ST_THREADS] = last - first;
            }
        }
        for (uint32_t q = threadIdx.x; q < left; q += ST_THREADS) {
            const uint32_t qs = __ldg(ps + (q - threadIdx.x)), qe = __ldg(pe + (q - threadIdx.x));
            const uint32_t last = e8_lower_bound(A, a, ent, qe);
            const uint32_t first = (qs + 1u) ? e8_lower_bound(A, a, ent, qs) : 0u;
            po[q - threadIdx.x] = last - first;
        }
        return;
    }
    const uint32_t *__restrict__ qsp = q_start + q0, *__restrict__ qep = q_stop + q0;
    uint32_t *__restrict__ op = out + q0;
    for (uint32_t qb = threadIdx.x; qb < nq; qb += U * ST_THREADS) {
        uint32_t qs[U], qe[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t q = min(qb + u * ST_THREADS, nq - 1u);   // clamped: the tail repeats the last query
            qs[u] = qsp[q];
            qe[u] = qep[q];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t last = staged_lower_bound(A, starts + off, smA, qe[u]);
            const uint32_t kb = qs[u] + 1u;
            const uint32_t first = alias ? (kb ? staged_lower_bound(B, starts + off, smB, kb - 1u) : 0u)
                                         : staged_lower_bound(B, stops + off, smB, kb);
            const uint32_t q = qb + u * ST_THREADS;
            if (q < nq) op[q] = last - first;
        }
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

// Compact per-segment index of the staged kernel: offsets are sized on the host from the interval counts alone.
int32_t build_compact_index(gams_cuda_count_plan *p, cudaStream_t st, const uint32_t *d_starts, const uint32_t *d_stops) {
    // function attributes belong to the current device: set with every index build (a process-wide flag would leave the
    // second device of a multi-device context at the 48 KB default)
    CUDA_TRY(cudaFuncSetAttribute(count_overlaps_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
    p->part_valid = false;
    std::vector<uint64_t> cofs((size_t)p->n_seg + 1, 0);
    for (int64_t s = 0; s < p->n_seg; ++s) {
        const uint64_t n = p->h_iv_off[s + 1] - p->h_iv_off[s];
        cofs[s + 1] = cofs[s] + 2 * st_arr_cap_bytes((uint32_t)std::min<uint64_t>(n, 0xFFFFFFFFull));
    }
    p->cpar_off = (sizeof(uint64_t) * cofs.size() + 15) & ~(size_t)15;
    int32_t rc;
    if ((rc = p->tmp[4].reserve(p->cpar_off + sizeof(SegCompact) * 2 * (size_t)std::max<int64_t>(p->n_seg, 1)))) return rc;
    if ((rc = p->tmp[5].reserve(cofs.back() + 64))) return rc;
    CUDA_TRY(cudaMemcpyAsync(p->tmp[4].p, cofs.data(), sizeof(uint64_t) * cofs.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));   // cofs is a local
    if (p->n_seg) {
        const int64_t blocks = std::min<int64_t>(p->n_seg, 148 * 16);
        uint32_t *d_max = reinterpret_cast<uint32_t *>(p->tmp[5].as<uint8_t>() + ((cofs.back() + 15) & ~15ull));
        CUDA_TRY(cudaMemsetAsync(d_max, 0, sizeof(uint32_t), st));
        build_compact_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_starts, d_stops, p->d_iv_off, p->n_seg, p->tmp[4].as<uint64_t>(),
                                                               reinterpret_cast<SegCompact *>(p->tmp[4].as<uint8_t>() + p->cpar_off),
                                                               p->tmp[5].as<uint8_t>(), d_max);
        ++p->launches;
        CUDA_TRY(cudaGetLastError());
        // the largest staged segment decides the shared memory of the count launches (and so how many CTAs share an SM)
        uint32_t h_max = 0;
        CUDA_TRY(cudaMemcpyAsync(&h_max, d_max, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        p->st_smem = std::min<size_t>(ST_SMEM, std::max<size_t>(h_max, 1024));
    }
    p->cidx_starts = d_starts;
    p->cidx_stops = d_stops;
    p->cidx_valid = true;
    return GAMS_CUDA_OK;
}

int32_t launch_count_overlaps(gams_cuda_count_plan *p, cudaStream_t st, const uint32_t *d_starts,
                              const uint32_t *d_stops, int64_t n_q, const int32_t *d_q_seg,
                              const uint32_t *d_q_start, const uint32_t *d_q_stop, int32_t grouped,
                              uint32_t *d_out) {
    if (n_q == 0) return GAMS_CUDA_OK;
    if (grouped && p->mode == 0) {
        // staged path: q_seg is non-decreasing; each segment's compact index is staged in shared memory
        int32_t rc;
        if (!p->cidx_valid || p->cidx_starts != d_starts || p->cidx_stops != d_stops) {
            if ((rc = build_compact_index(p, st, d_starts, d_stops))) return rc;
        }
        // items <= n_q / ST_CHUNK + n_seg; the CTAs beyond the actual count zero the out-of-index queries
        const int64_t grid = n_q / ST_CHUNK + p->n_seg + 8;
        if ((rc = p->tmp[3].reserve(sizeof(int64_t) * 2 * (size_t)(p->n_seg + 1) + sizeof(int32_t) * (size_t)grid))) return rc;
        int64_t *qoff = p->tmp[3].as<int64_t>(), *ioff = qoff + (p->n_seg + 1);
        int32_t *item_seg = reinterpret_cast<int32_t *>(ioff + (p->n_seg + 1));
        int64_t b0 = (n_q + 1 + 255) / 256;
        if (b0 > 148 * 32) b0 = 148 * 32;
        seg_query_bounds_kernel<<<(unsigned)b0, 256, 0, st>>>(d_q_seg, n_q, p->n_seg, qoff);
        seg_items_kernel<<<1, SI_THREADS, 0, st>>>(qoff, p->n_seg, ioff, item_seg);
        if (grid >= (1ll << 31)) {
            set_error("count_overlaps: too many segments for one launch");
            return GAMS_CUDA_EINVAL;
        }
        const uint64_t *d_cofs = p->tmp[4].as<uint64_t>();
        const SegCompact *d_par = reinterpret_cast<const SegCompact *>(p->tmp[4].as<uint8_t>() + p->cpar_off);
        count_overlaps_staged_kernel<<<(unsigned)grid, ST_THREADS, p->st_smem, st>>>(d_starts, d_stops, p->d_iv_off, p->n_seg, n_q, qoff,
                                                                                 ioff, item_seg, d_par, d_cofs, p->tmp[5].as<uint8_t>(),
                                                                                 d_q_start, d_q_stop, d_out);
        p->launches += 3;
        CUDA_TRY(cudaGetLastError());
        return GAMS_CUDA_OK;
    }
    if (!grouped && count_part_eligible(p, n_q)) {
        // arbitrary order, enough queries to pay for it: coarse partition, lookups through L2, un-permute (overlap_part.cu)
        bool done = false;
        int32_t rc = launch_count_overlaps_part(p, st, d_starts, d_stops, n_q, d_q_seg, d_q_start, d_q_stop, d_out, &done);
        if (rc || done) return rc;
    }
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
