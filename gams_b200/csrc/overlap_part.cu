// overlap_part.cu — Kernel C for queries in ARBITRARY order (the order `gams locate --count -f` walks its input file in,
// src/cmd_gams/locate.rs:111-141): coarse partition -> lookup through L2 -> un-permute.
//
// Why.  count = lower_bound(starts_seg, stop) - lower_bound(stops_seg, start + 1) (utils.rs:24-36 -> Lapper::count) touches two
// random places of an index of 10^8 intervals per query.  Answered in input order every touch is a DRAM access, and DRAM hands
// over 128 bytes for each of them: 25.7 GB of traffic for 2.4 GB of algorithmic bytes (profiles/README.md, r1r).  Here every
// random access is kept inside L2 instead, and DRAM only sees streams:
//   P0 part_hist_kernel      q_seg -> per (span, bucket) counts; a bucket is a contiguous range of segments whose compact index
//                            (overlap_index.cuh) is a few MB; a span is the contiguous run of queries one CTA owns
//   P1 part_scan_kernel      exclusive scan of the counts in (bucket, span) order = where every (span, bucket) run goes
//   P2 part_scatter_kernel   queries -> (start, stop, tag) records grouped by bucket; each CTA sorts a sub-tile in shared memory
//                            first, so a bucket receives runs of consecutive records.  tag = position inside the region << sb |
//                            segment inside the bucket.
//   P3 part_lookup_kernel    records in bucket order: the compact forms of ~2 buckets are live in L2 at any time; counts are
//                            written next to the records, again as a stream
//   P4 part_unpermute_kernel region by region (a region = 2^rb consecutive queries, its slice of out[] is <= 16 MB and stays
//                            in L2): the counts of the region's records — a contiguous piece of every bucket, because spans
//                            never straddle regions — are stored to out[region base + position].
// DRAM traffic: 4 + 12 + 12 + 12 + 4 + 8 + 4 B per query + the compact index once = 5.6 GB + 0.4 GB for C5.
// Queries whose segment is outside the index go to bucket 0 with the segment field all ones and count 0 (utils.rs:29-32).
#include "common.cuh"
#include "overlap_index.cuh"

namespace gams {

constexpr int PT_MAX_B = 256;         // buckets
constexpr int PT_THREADS = 512;
constexpr int PT_PER = 8;             // queries per thread and sub-tile
constexpr int PT_SUB = PT_THREADS * PT_PER;   // queries sorted in shared memory at a time
constexpr int PT_MAX_SPANS = 4096;
constexpr size_t PT_SCATTER_SMEM = (size_t)PT_SUB * 13;

// 32 bytes per segment: what one lookup needs, in one load
struct alignas(32) SegLite {
    uint64_t cofs;    // byte offset of the segment's compact form in the compact buffer
    uint64_t off;     // first interval of the segment in starts[] / stops[]
    uint32_t lo, nb, n;
    uint32_t kind;    // l | fmt << 8 (E8) ; PT_KIND_GENERAL: read the two SegCompact records ; PT_KIND_EMPTY
};
constexpr uint32_t PT_KIND_GENERAL = 0xFFFFFFFEu, PT_KIND_EMPTY = 0xFFFFFFFFu;

__global__ void __launch_bounds__(256)
part_lite_kernel(const SegCompact *__restrict__ par, const uint64_t *__restrict__ cofs, const uint64_t *__restrict__ iv_off,
                 int64_t n_seg, SegLite *__restrict__ lite) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const SegCompact A = par[2 * s], B = par[2 * s + 1];
    SegLite t;
    t.cofs = cofs[s];
    t.off = iv_off[s];
    t.lo = A.lo, t.nb = A.nb, t.n = A.n;
    if (A.n == 0) t.kind = PT_KIND_EMPTY;
    else if (A.fmt == 1u && B.l == ST_ALIAS && A.l != ST_PLAIN) t.kind = A.l | (1u << 8);
    else t.kind = PT_KIND_GENERAL;
    lite[s] = t;
}

struct PartGeom {
    int32_t n_bkt;        // buckets
    int32_t sb;           // bits of the segment-inside-bucket field of a tag
    int32_t rb;           // log2 of the region size
    int32_t span_shift;   // log2 of the span size (spans per region = 1 << (rb - span_shift))
    int32_t n_span;       // spans (n_region << (rb - span_shift))
    int32_t n_region;
};

__device__ __forceinline__ uint32_t bucket_of(int32_t seg, int64_t n_seg, const uint8_t *__restrict__ seg2bkt) {
    return (seg >= 0 && seg < n_seg) ? (uint32_t)__ldg(seg2bkt + seg) : 0u;
}

// P0: one CTA per span
__global__ void __launch_bounds__(PT_THREADS)
part_hist_kernel(const int32_t *__restrict__ q_seg, int64_t n_q, int64_t n_seg, const uint8_t *__restrict__ seg2bkt, PartGeom G,
                 uint32_t *__restrict__ cnt /* [n_bkt][n_span] */) {
    __shared__ uint32_t hist[PT_MAX_B];
    for (int b = threadIdx.x; b < PT_MAX_B; b += PT_THREADS) hist[b] = 0;
    __syncthreads();
    const int64_t q0 = (int64_t)blockIdx.x << G.span_shift, q1 = min(q0 + ((int64_t)1 << G.span_shift), n_q);
    for (int64_t q = q0 + threadIdx.x; q < q1; q += (int64_t)PT_THREADS * 4) {
        int32_t s[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) s[k] = (q + k * PT_THREADS < q1) ? __ldg(q_seg + q + k * PT_THREADS) : 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (q + k * PT_THREADS < q1) atomicAdd(&hist[bucket_of(s[k], n_seg, seg2bkt)], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < G.n_bkt; b += PT_THREADS) cnt[(size_t)b * G.n_span + blockIdx.x] = hist[b];
}

// P1: exclusive scan of cnt[0 .. n) in place, total to cnt[n]; one CTA, 4 elements per thread and round
__global__ void __launch_bounds__(1024)
part_scan_kernel(uint32_t *__restrict__ cnt, uint32_t n) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 4096u) {
        const uint32_t i = base + threadIdx.x * 4u;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (i + k < n) ? cnt[i + k] : 0u;
        const uint32_t tot = v[0] + v[1] + v[2] + v[3];
        uint32_t inc = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if ((threadIdx.x & 31) >= d) inc += t;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
        __syncthreads();
        uint32_t before = carry_s;
        {
            const uint32_t w = threadIdx.x >> 5;
            uint32_t x = (threadIdx.x & 31) < w ? wsum[threadIdx.x & 31] : 0u;
#pragma unroll
            for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, d);
            before += x;
        }
        uint32_t at = before + inc - tot;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i + k < n) cnt[i + k] = at;
            at += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = before + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) cnt[n] = carry_s;
}

// P2: one CTA per span; sub-tiles of PT_SUB queries are grouped by bucket in shared memory and leave as runs
__global__ void __launch_bounds__(PT_THREADS)
part_scatter_kernel(const int32_t *__restrict__ q_seg, const uint32_t *__restrict__ q_start, const uint32_t *__restrict__ q_stop,
                    int64_t n_q, int64_t n_seg, const uint8_t *__restrict__ seg2bkt, const int32_t *__restrict__ bkt_first,
                    PartGeom G, const uint32_t *__restrict__ off /* scanned cnt */, uint32_t *__restrict__ r_start,
                    uint32_t *__restrict__ r_stop, uint32_t *__restrict__ r_tag) {
    __shared__ uint32_t hist[PT_MAX_B], lbase[PT_MAX_B], cursor[PT_MAX_B];
    __shared__ int32_t first_s[PT_MAX_B];
    __shared__ uint32_t wtot[PT_MAX_B / 32];
    extern __shared__ __align__(16) uint8_t pt_smem[];   // PT_SCATTER_SMEM bytes
    uint32_t *s_start = reinterpret_cast<uint32_t *>(pt_smem), *s_stop = s_start + PT_SUB, *s_tag = s_stop + PT_SUB;
    uint8_t *s_b = reinterpret_cast<uint8_t *>(s_tag + PT_SUB);
    const int tid = threadIdx.x;
    for (int b = tid; b < PT_MAX_B; b += PT_THREADS) {
        hist[b] = 0;
        cursor[b] = b < G.n_bkt ? off[(size_t)b * G.n_span + blockIdx.x] : 0u;
        first_s[b] = b < G.n_bkt ? bkt_first[b] : 0;
    }
    __syncthreads();
    const int64_t q0 = (int64_t)blockIdx.x << G.span_shift, q1 = min(q0 + ((int64_t)1 << G.span_shift), n_q);
    const uint32_t region_mask = (1u << G.rb) - 1u, sent = (1u << G.sb) - 1u;
    for (int64_t base = q0; base < q1; base += PT_SUB) {
        const uint32_t m = (uint32_t)min((int64_t)PT_SUB, q1 - base);
        int32_t sg[PT_PER];
        uint32_t qs[PT_PER], qe[PT_PER], bk[PT_PER], rk[PT_PER];
#pragma unroll
        for (int k = 0; k < PT_PER; ++k) {
            const uint32_t i = k * PT_THREADS + tid;
            if (i < m) {
                sg[k] = __ldg(q_seg + base + i);
                qs[k] = __ldg(q_start + base + i);
                qe[k] = __ldg(q_stop + base + i);
            }
        }
#pragma unroll
        for (int k = 0; k < PT_PER; ++k) {
            const uint32_t i = k * PT_THREADS + tid;
            if (i < m) {
                bk[k] = bucket_of(sg[k], n_seg, seg2bkt);
                rk[k] = atomicAdd(&hist[bk[k]], 1u);
            }
        }
        __syncthreads();
        // exclusive scan of hist over PT_MAX_B = 256 buckets: the first 8 warps
        if (tid < PT_MAX_B) {
            const uint32_t h = hist[tid];
            uint32_t inc = h;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if ((tid & 31) >= d) inc += t;
            }
            if ((tid & 31) == 31) wtot[tid >> 5] = inc;
            lbase[tid] = inc - h;
        }
        __syncthreads();
        if (tid < PT_MAX_B) {
            uint32_t add = 0;
            for (int w = 0; w < (tid >> 5); ++w) add += wtot[w];
            lbase[tid] += add;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PT_PER; ++k) {
            const uint32_t i = k * PT_THREADS + tid;
            if (i < m) {
                const uint32_t slot = lbase[bk[k]] + rk[k];
                const bool valid = sg[k] >= 0 && sg[k] < n_seg;
                const uint32_t local = valid ? (uint32_t)(sg[k] - first_s[bk[k]]) : sent;
                s_start[slot] = qs[k];
                s_stop[slot] = qe[k];
                s_tag[slot] = (((uint32_t)(base + i) & region_mask) << G.sb) | local;
                s_b[slot] = (uint8_t)bk[k];
            }
        }
        __syncthreads();
        for (uint32_t slot = tid; slot < m; slot += PT_THREADS) {
            const uint32_t b = s_b[slot];
            const uint32_t dst = cursor[b] + (slot - lbase[b]);
            r_start[dst] = s_start[slot];
            r_stop[dst] = s_stop[slot];
            r_tag[dst] = s_tag[slot];
        }
        __syncthreads();
        if (tid < PT_MAX_B) {
            cursor[tid] += hist[tid];
            hist[tid] = 0;
        }
        __syncthreads();
    }
}

// #elements < key of one endpoint array of a general (two-array) segment, compact form in global memory
__device__ __forceinline__ uint32_t general_lower_bound(const SegCompact &P, const uint32_t *__restrict__ a,
                                                        const uint8_t *__restrict__ form, uint32_t key) {
    return staged_lower_bound(P, a, form, key);
}

// P3: records in bucket order
constexpr int PL_THREADS = 256;
constexpr int PL_PER = 4;
__global__ void __launch_bounds__(PL_THREADS)
part_lookup_kernel(const uint32_t *__restrict__ starts, const uint32_t *__restrict__ stops, const SegLite *__restrict__ lite,
                   const SegCompact *__restrict__ par, const uint8_t *__restrict__ cbuf, const int32_t *__restrict__ bkt_first,
                   PartGeom G, const uint32_t *__restrict__ off, uint32_t n_rec, const uint32_t *__restrict__ r_start,
                   const uint32_t *__restrict__ r_stop, const uint32_t *__restrict__ r_tag, uint32_t *__restrict__ r_cnt) {
    __shared__ uint32_t boff[PT_MAX_B + 1];   // first record of every bucket
    __shared__ int32_t first_s[PT_MAX_B];
    for (int b = threadIdx.x; b <= G.n_bkt; b += PL_THREADS) boff[b] = off[(size_t)b * G.n_span];
    for (int b = threadIdx.x; b < G.n_bkt; b += PL_THREADS) first_s[b] = bkt_first[b];
    __syncthreads();
    const uint32_t p0 = blockIdx.x * (uint32_t)(PL_THREADS * PL_PER);
    if (p0 >= n_rec) return;
    // bucket of the CTA's first record
    int lo = 0, hi = G.n_bkt;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (boff[mid] <= p0) lo = mid; else hi = mid;
    }
    uint32_t b = (uint32_t)lo;
    const uint32_t sent = (1u << G.sb) - 1u;
    uint32_t qs[PL_PER], qe[PL_PER], tg[PL_PER];
#pragma unroll
    for (int k = 0; k < PL_PER; ++k) {
        const uint32_t p = p0 + k * PL_THREADS + threadIdx.x;
        if (p < n_rec) {
            qs[k] = __ldg(r_start + p);
            qe[k] = __ldg(r_stop + p);
            tg[k] = __ldg(r_tag + p);
        }
    }
#pragma unroll
    for (int k = 0; k < PL_PER; ++k) {
        const uint32_t p = p0 + k * PL_THREADS + threadIdx.x;
        if (p >= n_rec) break;
        while (p >= boff[b + 1]) ++b;
        const uint32_t local = tg[k] & sent;
        uint32_t r = 0;
        if (local != sent) {
            const int64_t s = (int64_t)first_s[b] + local;
            const uint4 *lp = reinterpret_cast<const uint4 *>(lite + s);
            const uint4 l0 = __ldg(lp), l1 = __ldg(lp + 1);
            const uint64_t cofs = (uint64_t)l0.x | ((uint64_t)l0.y << 32), ivo = (uint64_t)l0.z | ((uint64_t)l0.w << 32);
            const uint32_t kind = l1.w;
            if (kind == PT_KIND_EMPTY) {
                r = 0;
            } else if (kind != PT_KIND_GENERAL) {
                SegCompact A;
                A.lo = l1.x, A.nb = l1.y, A.n = l1.z, A.l = kind & 0xFFu;
                const uint2 *ent = reinterpret_cast<const uint2 *>(cbuf + cofs);
                const uint32_t *a = starts + ivo;
                const uint32_t last = e8_lower_bound(A, a, ent, qe[k]);
                const uint32_t first = (qs[k] + 1u) ? e8_lower_bound(A, a, ent, qs[k]) : 0u;
                r = last - first;
            } else {
                const SegCompact A = par[2 * s];
                SegCompact B = par[2 * s + 1];
                const bool alias = B.l == ST_ALIAS;
                const uint8_t *fA = cbuf + cofs, *fB = fA + st_arr_cap_bytes(A.n);
                const uint32_t last = general_lower_bound(A, starts + ivo, fA, qe[k]);
                const uint32_t kb = qs[k] + 1u;
                uint32_t first;
                if (alias) first = kb ? general_lower_bound(A, starts + ivo, fA, kb - 1u) : 0u;
                else first = general_lower_bound(B, stops + ivo, fB, kb);
                r = last - first;
            }
        }
        r_cnt[p] = r;
    }
}

// P4: CTA (region r, part j of PU_PARTS): its share of the region's records, bucket after bucket
constexpr int PU_THREADS = 256;
__global__ void __launch_bounds__(PU_THREADS)
part_unpermute_kernel(PartGeom G, int32_t parts, const uint32_t *__restrict__ off, const uint32_t *__restrict__ r_tag,
                      const uint32_t *__restrict__ r_cnt, int64_t n_q, uint32_t *__restrict__ out) {
    __shared__ uint32_t sb_begin[PT_MAX_B], vstart[PT_MAX_B + 1];
    __shared__ uint32_t wtot[PT_MAX_B / 32];
    const int r = blockIdx.x / parts, j = blockIdx.x % parts;
    const int spr = 1 << (G.rb - G.span_shift);   // spans per region
    const int tid = threadIdx.x;
    if (tid < PT_MAX_B) {
        uint32_t len = 0;
        if (tid < G.n_bkt) {
            const size_t row = (size_t)tid * G.n_span;
            const uint32_t a = off[row + (size_t)r * spr], e = off[row + (size_t)(r + 1) * spr];   // the row after the last is the next bucket's first
            sb_begin[tid] = a;
            len = e - a;
        }
        uint32_t inc = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if ((tid & 31) >= d) inc += t;
        }
        if ((tid & 31) == 31) wtot[tid >> 5] = inc;
        vstart[tid + 1] = inc;
    }
    __syncthreads();
    if (tid < PT_MAX_B) {
        uint32_t add = 0;
        for (int w = 0; w < (tid >> 5); ++w) add += wtot[w];
        vstart[tid + 1] += add;
    }
    if (tid == 0) vstart[0] = 0;
    __syncthreads();
    const uint32_t total = vstart[G.n_bkt];
    const uint32_t v0 = (uint32_t)((uint64_t)total * j / parts), v1 = (uint32_t)((uint64_t)total * (j + 1) / parts);
    if (v1 <= v0) return;
    int lo = 0, hi = G.n_bkt;   // bucket holding v0
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (vstart[mid] <= v0) lo = mid; else hi = mid;
    }
    uint32_t *__restrict__ o = out + ((int64_t)r << G.rb);
    for (int b = lo; b < G.n_bkt && vstart[b] < v1; ++b) {
        const uint32_t a = max(v0, vstart[b]), e = min(v1, vstart[b + 1]);
        const uint32_t p0 = sb_begin[b] + (a - vstart[b]), cnt = e - a;
        for (uint32_t i = tid; i < cnt; i += PU_THREADS) {
            const uint32_t t = __ldg(r_tag + p0 + i), c = __ldg(r_cnt + p0 + i);
            o[t >> G.sb] = c;
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
static int32_t env_int(const char *name, int32_t dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// Buckets: contiguous segment ranges of about `target` bytes of compact index each (at most PT_MAX_B of them).
static void plan_buckets(gams_cuda_count_plan *p, std::vector<uint8_t> &seg2bkt, std::vector<int32_t> &first, int &sb) {
    const int64_t n_seg = p->n_seg;
    std::vector<uint64_t> w((size_t)n_seg + 1, 0);
    for (int64_t s = 0; s < n_seg; ++s) {
        const uint64_t n = p->h_iv_off[s + 1] - p->h_iv_off[s];
        const uint64_t form = 2 * st_arr_cap_bytes((uint32_t)std::min<uint64_t>(n, 0xFFFFFFFFull));
        w[s + 1] = w[s] + (form ? form : 8 * n) + 64;
    }
    const uint64_t target = (uint64_t)std::max(1, env_int("GAMS_CUDA_PART_MB", 12)) << 20;
    int n_bkt = (int)std::min<uint64_t>(PT_MAX_B, std::max<uint64_t>(1, (w[n_seg] + target - 1) / target));
    n_bkt = (int)std::min<int64_t>(n_bkt, std::max<int64_t>(1, n_seg));
    first.assign((size_t)n_bkt + 1, 0);
    seg2bkt.assign((size_t)std::max<int64_t>(n_seg, 1), 0);
    int64_t s = 0;
    uint64_t max_local = 1;
    for (int b = 0; b < n_bkt; ++b) {
        first[b] = (int32_t)s;
        const uint64_t lim = w[n_seg] * (uint64_t)(b + 1) / (uint64_t)n_bkt;
        const int64_t s_begin = s;
        while (s < n_seg && (b == n_bkt - 1 || w[s + 1] <= lim || s == s_begin)) seg2bkt[s++] = (uint8_t)b;
        max_local = std::max<uint64_t>(max_local, (uint64_t)(s - s_begin));
    }
    first[n_bkt] = (int32_t)n_seg;
    sb = 1;
    while ((1ull << sb) <= max_local) ++sb;
}

bool count_part_eligible(const gams_cuda_count_plan *p, int64_t n_q) {
    const int64_t min_q = (int64_t)env_int("GAMS_CUDA_PART_MIN_Q", 1 << 20);
    return p->mode == 0 && n_q >= min_q && n_q < (1ll << 31) && p->n_seg > 0 && p->total_iv > 0;
}

// tmp[6]: records (start, stop, tag, cnt: 16 B per query); tmp[7]: seg2bkt | bkt_first | SegLite table; tmp[8]: counts / offsets
int32_t launch_count_overlaps_part(gams_cuda_count_plan *p, cudaStream_t st, const uint32_t *d_starts, const uint32_t *d_stops,
                                   int64_t n_q, const int32_t *d_q_seg, const uint32_t *d_q_start, const uint32_t *d_q_stop,
                                   uint32_t *d_out, bool *done) {
    *done = false;
    int32_t rc;
    if (!p->cidx_valid || p->cidx_starts != d_starts || p->cidx_stops != d_stops) {
        if ((rc = build_compact_index(p, st, d_starts, d_stops))) return rc;
        p->part_valid = false;
    }
    const size_t lite_off = ((size_t)std::max<int64_t>(p->n_seg, 1) + (PT_MAX_B + 1) * sizeof(int32_t) + 63) & ~(size_t)31;
    if (!p->part_valid) {
        std::vector<uint8_t> seg2bkt;
        std::vector<int32_t> first;
        int sb = 1;
        plan_buckets(p, seg2bkt, first, sb);
        p->part_nbkt = (int)first.size() - 1;
        p->part_sb = sb;
        if ((rc = p->tmp[7].reserve(lite_off + sizeof(SegLite) * (size_t)std::max<int64_t>(p->n_seg, 1)))) return rc;
        uint8_t *base = p->tmp[7].as<uint8_t>();
        const size_t first_off = ((size_t)std::max<int64_t>(p->n_seg, 1) + 3) & ~(size_t)3;
        CUDA_TRY(cudaMemcpyAsync(base, seg2bkt.data(), seg2bkt.size(), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(base + first_off, first.data(), first.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        const SegCompact *d_par = reinterpret_cast<const SegCompact *>(p->tmp[4].as<uint8_t>() + p->cpar_off);
        part_lite_kernel<<<(unsigned)((p->n_seg + 255) / 256), 256, 0, st>>>(d_par, p->tmp[4].as<uint64_t>(), p->d_iv_off, p->n_seg,
                                                                              reinterpret_cast<SegLite *>(base + lite_off));
        ++p->launches;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(st));   // seg2bkt / first are locals
        p->part_valid = true;
    }
    PartGeom G;
    G.n_bkt = p->part_nbkt;
    G.sb = p->part_sb;
    G.rb = std::min(env_int("GAMS_CUDA_PART_RB", 22), 32 - G.sb);
    if (G.rb < 12) return GAMS_CUDA_OK;   // too many segments per bucket for a 32-bit tag: the caller takes the ranked kernel
    G.n_region = (int32_t)((n_q + (1ll << G.rb) - 1) >> G.rb);
    // spans: a power-of-two fraction of a region, at least two sub-tiles each, about 148 * 8 of them in all
    int span_shift = G.rb;
    while (span_shift > 13 && ((int64_t)G.n_region << (G.rb - span_shift)) < 148 * 8) --span_shift;
    while (((int64_t)G.n_region << (G.rb - span_shift)) > PT_MAX_SPANS) ++span_shift;
    if (span_shift > G.rb) return GAMS_CUDA_OK;
    G.span_shift = span_shift;
    G.n_span = G.n_region << (G.rb - span_shift);
    const size_t n_cnt = (size_t)G.n_bkt * G.n_span;
    if ((rc = p->tmp[6].reserve((size_t)n_q * 16 + 64))) return rc;
    if ((rc = p->tmp[8].reserve((n_cnt + 1) * sizeof(uint32_t)))) return rc;
    uint8_t *base = p->tmp[7].as<uint8_t>();
    const size_t first_off = ((size_t)std::max<int64_t>(p->n_seg, 1) + 3) & ~(size_t)3;
    const uint8_t *d_seg2bkt = base;
    const int32_t *d_first = reinterpret_cast<const int32_t *>(base + first_off);
    const SegLite *d_lite = reinterpret_cast<const SegLite *>(base + lite_off);
    uint32_t *d_cnt = p->tmp[8].as<uint32_t>();
    uint32_t *r_start = p->tmp[6].as<uint32_t>(), *r_stop = r_start + n_q, *r_tag = r_stop + n_q, *r_cnt = r_tag + n_q;
    part_hist_kernel<<<(unsigned)G.n_span, PT_THREADS, 0, st>>>(d_q_seg, n_q, p->n_seg, d_seg2bkt, G, d_cnt);
    part_scan_kernel<<<1, 1024, 0, st>>>(d_cnt, (uint32_t)n_cnt);
    CUDA_TRY(cudaFuncSetAttribute(part_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SCATTER_SMEM));
    part_scatter_kernel<<<(unsigned)G.n_span, PT_THREADS, PT_SCATTER_SMEM, st>>>(d_q_seg, d_q_start, d_q_stop, n_q, p->n_seg, d_seg2bkt, d_first, G,
                                                                  d_cnt, r_start, r_stop, r_tag);
    const SegCompact *d_par = reinterpret_cast<const SegCompact *>(p->tmp[4].as<uint8_t>() + p->cpar_off);
    const unsigned lgrid = (unsigned)((n_q + PL_THREADS * PL_PER - 1) / (PL_THREADS * PL_PER));
    part_lookup_kernel<<<lgrid, PL_THREADS, 0, st>>>(d_starts, d_stops, d_lite, d_par, p->tmp[5].as<uint8_t>(), d_first, G, d_cnt,
                                                    (uint32_t)n_q, r_start, r_stop, r_tag, r_cnt);
    const int parts = std::max(1, env_int("GAMS_CUDA_PART_UPARTS", 592));
    part_unpermute_kernel<<<(unsigned)(G.n_region * parts), PU_THREADS, 0, st>>>(G, parts, d_cnt, r_tag, r_cnt, n_q, d_out);
    p->launches += 5;
    CUDA_TRY(cudaGetLastError());
    *done = true;
    return GAMS_CUDA_OK;
}

}  // namespace gams
