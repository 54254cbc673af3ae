// overlap_part.cu — Kernel C for queries in ARBITRARY order (the order `gams locate --count -f` walks its input file in,
// src/cmd_gams/locate.rs:111-141): sort spans by index bucket -> lookups through L2 -> un-permute spans.
//
// Why.  count = lower_bound(starts_seg, stop) - lower_bound(stops_seg, start + 1) (utils.rs:24-36 -> Lapper::count) touches two
// random places of an index of 10^8 intervals per query.  Answered in input order every touch is a DRAM access, and DRAM hands
// over 128 bytes for each of them: 25.7 GB of traffic for 2.4 GB of algorithmic bytes (profiles/README.md, r1r).  Here every
// random access stays inside L2 or shared memory, and DRAM only sees streams:
//   S1 part_sort_kernel       one CTA per span of PT_SPAN consecutive queries: counting sort by BUCKET in shared memory (a bucket
//                             is a contiguous range of segments whose compact index, overlap_index.cuh, is a few MB), written
//                             back as (start, stop, tag) records in the span's own slot of the record arrays — linear stores —
//                             with the span's bucket counts and bases.  tag = position inside the span << sb | segment inside
//                             the bucket.
//   S2 part_lookup_kernel     run by run in BUCKET-MAJOR order (run = the records of one bucket in one span, contiguous): at any
//                             time the CTAs in flight work on one or two buckets, whose compact forms stay in L2; the count goes
//                             next to its record.
//   S3 part_unpermute_kernel  one CTA per span: counts to their original positions through shared memory, linear stores to out[].
// DRAM traffic: 12 + 12 | 12 + 4 | 8 + 4 = 52 B per query + the compact index once: 5.2 GB + 0.4 GB for C5 (10^8 x 10^8).
// There is no histogram pass and no global scan: a run is found from (bucket, span) alone.
// Queries whose segment is outside the index go to bucket 0 with the segment field all ones and count 0 (utils.rs:29-32).
#include "common.cuh"
#include "overlap_index.cuh"

namespace gams {

constexpr int PT_MAX_B = 256;                 // buckets
#ifndef PT_THREADS_N
#define PT_THREADS_N 512
#endif
#ifndef PT_SPAN_SHIFT_N
#define PT_SPAN_SHIFT_N 13
#endif
constexpr int PT_THREADS = PT_THREADS_N;
constexpr int PT_SPAN_SHIFT = PT_SPAN_SHIFT_N;
constexpr int PT_PER = (1 << PT_SPAN_SHIFT) / PT_THREADS;   // queries per thread
constexpr int PT_SPAN = 1 << PT_SPAN_SHIFT;   // queries per span = PT_THREADS * PT_PER
static_assert(PT_SPAN == PT_THREADS * PT_PER, "span geometry");
constexpr size_t PT_SORT_SMEM = (size_t)PT_SPAN * 12;

// 16 bytes per segment: what one E8 lookup needs, in one load
struct alignas(16) SegLite {
    uint64_t cofs;    // byte offset of the segment's compact form in the compact buffer
    uint32_t lo;      // first element of the segment's starts[]
    uint32_t nbl;     // E8 form: buckets | l << 16 ; PT_GENERAL: read the two SegCompact records ; PT_EMPTY: no intervals
};
constexpr uint32_t PT_GENERAL = 0xFFFFFFFEu, PT_EMPTY = 0xFFFFFFFFu;

__global__ void __launch_bounds__(256)
part_lite_kernel(const SegCompact *__restrict__ par, const uint64_t *__restrict__ cofs, int64_t n_seg, SegLite *__restrict__ lite) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const SegCompact A = par[2 * s], B = par[2 * s + 1];
    SegLite t;
    t.cofs = cofs[s];
    t.lo = A.lo;
    if (A.n == 0) t.nbl = PT_EMPTY;
    else if (A.fmt == 1u && B.l == ST_ALIAS && A.l != ST_PLAIN) t.nbl = A.nb | (A.l << 16);
    else t.nbl = PT_GENERAL;
    lite[s] = t;
}

struct PartGeom {
    int32_t n_bkt;        // buckets: bucket of segment s = s >> bshift
    int32_t bshift;
    int32_t sb;           // bits of the segment-inside-bucket field of a tag = bshift + 1 (all ones: segment outside the index)
    int32_t n_span;       // spans
};

__device__ __forceinline__ uint32_t ld_stream(const uint32_t *p) {   // read once: do not keep in L1
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_stream4(const uint32_t *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ void cp_async4(uint32_t *smem_dst, const uint32_t *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

// S1: one CTA per span
__global__ void __launch_bounds__(PT_THREADS)
part_sort_kernel(const int32_t *__restrict__ q_seg, const uint32_t *__restrict__ q_start, const uint32_t *__restrict__ q_stop,
                 int64_t n_q, int64_t n_seg, PartGeom G, uint16_t *__restrict__ run_cnt /* [n_bkt][n_span] */,
                 uint16_t *__restrict__ run_base,
                 uint32_t *__restrict__ r_start, uint32_t *__restrict__ r_stop, uint32_t *__restrict__ r_tag) {
    __shared__ uint32_t hist[PT_MAX_B], lbase[PT_MAX_B];
    __shared__ uint32_t wtot[PT_MAX_B / 32];
    extern __shared__ __align__(16) uint8_t pt_smem[];   // PT_SORT_SMEM bytes
    uint32_t *s_start = reinterpret_cast<uint32_t *>(pt_smem), *s_stop = s_start + PT_SPAN, *s_tag = s_stop + PT_SPAN;
    const int tid = threadIdx.x;
    for (int b = tid; b < PT_MAX_B; b += PT_THREADS) hist[b] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x << PT_SPAN_SHIFT;
    const uint32_t m = (uint32_t)min((int64_t)PT_SPAN, n_q - base);
    const uint32_t sent = (1u << G.sb) - 1u;
    uint32_t br[PT_PER], tg[PT_PER];   // bucket << 16 | rank inside the bucket ; tag
    {
        int32_t sg[PT_PER];
#pragma unroll
        for (int k = 0; k < PT_PER; ++k) {
            const uint32_t i = k * PT_THREADS + tid;
            sg[k] = i < m ? __ldg(q_seg + base + i) : -1;
        }
#pragma unroll
        for (int k = 0; k < PT_PER; ++k) {
            const uint32_t i = k * PT_THREADS + tid;
            if (i < m) {
                const bool valid = sg[k] >= 0 && sg[k] < n_seg;
                const uint32_t b = valid ? (uint32_t)sg[k] >> G.bshift : 0u;
                br[k] = (b << 16) | atomicAdd(&hist[b], 1u);
                tg[k] = (i << G.sb) | (valid ? ((uint32_t)sg[k] & (sent >> 1)) : sent);
            }
        }
    }
    __syncthreads();
    if (tid < PT_MAX_B) {   // exclusive scan of hist over the buckets: the first 8 warps
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
        const uint32_t lb = lbase[tid] + add;
        lbase[tid] = lb;
        if (tid < G.n_bkt) {
            run_cnt[(size_t)tid * G.n_span + blockIdx.x] = (uint16_t)hist[tid];   // <= PT_SPAN = 8192
            run_base[(size_t)tid * G.n_span + blockIdx.x] = (uint16_t)lb;
        }
    }
    __syncthreads();
    // start / stop go from global memory straight to their sorted slot: 4-byte cp.async (LDGSTS) to a computed shared
    // address — no register staging, all 32 copies of a thread in flight at once
#pragma unroll
    for (int k = 0; k < PT_PER; ++k) {
        const uint32_t i = k * PT_THREADS + tid;
        if (i < m) {
            const uint32_t slot = lbase[br[k] >> 16] + (br[k] & 0xFFFFu);
            cp_async4(s_start + slot, q_start + base + i);
            cp_async4(s_stop + slot, q_stop + base + i);
            s_tag[slot] = tg[k];
        }
    }
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // the record arrays are 16-byte aligned and padded to whole spans: linear 16-byte stores
    uint4 *o_start = reinterpret_cast<uint4 *>(r_start + base), *o_stop = reinterpret_cast<uint4 *>(r_stop + base),
          *o_tag = reinterpret_cast<uint4 *>(r_tag + base);
    const uint4 *i_start = reinterpret_cast<const uint4 *>(s_start), *i_stop = reinterpret_cast<const uint4 *>(s_stop),
                *i_tag = reinterpret_cast<const uint4 *>(s_tag);
    for (uint32_t j = tid; j < (m + 3u) / 4u; j += PT_THREADS) {
        o_start[j] = i_start[j];
        o_stop[j] = i_stop[j];
        o_tag[j] = i_tag[j];
    }
}

// S2: CTA (bucket b, chunk of PL_RUNS spans): each warp takes PL_WRUNS consecutive runs of the bucket as one stream of records
// and handles it 4 records per lane at a time: 12 record loads, then the 8 index loads, then the 4 stores — the latencies
// of a phase overlap inside the warp.  The bucket's SegLite records sit in shared memory (CACHE) when they fit.
constexpr int PL_THREADS = 256;
#ifndef PL_WRUNS_N
#define PL_WRUNS_N 8
#endif
#ifndef PL_U_N
#define PL_U_N 4
#endif
constexpr int PL_WRUNS = PL_WRUNS_N;                     // runs per warp
constexpr int PL_RUNS = PL_WRUNS * PL_THREADS / 32;      // runs (spans) per CTA
constexpr int PL_U = PL_U_N;                             // records per lane and round
#ifndef PL_MIN_CTAS
#define PL_MIN_CTAS 2
#endif
constexpr int PL_MAX_CACHE_SEGS = 3072;                  // 48 KB of SegLite

struct LookupArgs {
    const uint32_t *starts, *stops;
    const uint64_t *iv_off;
    const SegLite *lite;
    const SegCompact *par;
    const uint8_t *cbuf;
};

// general (two-array) segments and everything else that is not the E8 form: straight through the shared helpers
__device__ __noinline__ uint32_t lookup_general(const LookupArgs &A_, int64_t s, uint32_t qs, uint32_t qe) {
    const SegCompact A = A_.par[2 * s], B = A_.par[2 * s + 1];
    if (A.n == 0) return 0u;
    const uint64_t cofs = A_.lite[s].cofs, ivo = A_.iv_off[s];
    const bool alias = B.l == ST_ALIAS;
    const uint8_t *fA = A_.cbuf + cofs, *fB = fA + st_arr_cap_bytes(A.n);
    const uint32_t kb = qs + 1u;
    if (A.fmt == 1u && alias && A.l != ST_PLAIN) {
        const uint2 *ent = reinterpret_cast<const uint2 *>(fA);
        const uint32_t last = e8_lower_bound(A, A_.starts + ivo, ent, qe);
        return last - (kb ? e8_lower_bound(A, A_.starts + ivo, ent, qs) : 0u);
    }
    const uint32_t last = staged_lower_bound(A, A_.starts + ivo, fA, qe);
    uint32_t first;
    if (alias) first = kb ? staged_lower_bound(A, A_.starts + ivo, fA, kb - 1u) : 0u;
    else first = staged_lower_bound(B, A_.stops + ivo, fB, kb);
    return last - first;
}

// one E8 entry.  A plain read-only load: with L1::no_allocate the sector is not kept in L2 either (measured: every lookup
// became a DRAM access, 8.0 GB instead of 1.6 GB), and L2 residency of the bucket's entries is the point of this path
__device__ __forceinline__ uint2 ld_ent(const uint2 *p) { return __ldg(p); }

// E8 lower bound in two halves: which entry, and what the entry says (see e8_lower_bound, overlap_index.cuh)
__device__ __forceinline__ uint32_t e8_entry_index(uint32_t lo, uint32_t l, uint32_t nb, uint32_t key) {
    return min((max(key, lo) - lo) >> l, nb);
}
__device__ __forceinline__ uint32_t e8_finish(uint32_t lo, uint32_t l, uint2 e, uint32_t key) {   // the six slots of the entry only
    const uint32_t d = max(key, lo) - lo, rem = d & ((1u << l) - 1u), rr = rem * 0x01010101u;
    return (e.x & 0x7FFFu) + (uint32_t)__popc(~((e.x | 0x8080FFFFu) - rr) & 0x80808080u) + (uint32_t)__popc(~((e.y | 0x80808080u) - rr) & 0x80808080u);
}

// crowded E8 bucket (flag set, key beyond the entry's sixth element): the seventh and later elements, in the raw array
__device__ __noinline__ uint32_t e8_crowded_extra(const uint32_t *__restrict__ a, const uint2 *ent, uint32_t b, uint32_t cum, uint32_t key) {
    const uint32_t nxt = __ldg(&ent[b + 1].x) & 0x7FFFu;
    uint32_t cnt = 0;
    for (uint32_t i = cum + 6u; i < nxt; ++i) cnt += (__ldg(a + i) < key) ? 1u : 0u;
    return cnt;
}

template <bool CACHE>
__global__ void __launch_bounds__(PL_THREADS, PL_MIN_CTAS)
part_lookup_kernel(LookupArgs A, int64_t n_seg, PartGeom G, uint32_t ctas_per_bkt,
                   const uint16_t *__restrict__ run_cnt, const uint16_t *__restrict__ run_base, const uint32_t *__restrict__ r_start,
                   const uint32_t *__restrict__ r_stop, const uint32_t *__restrict__ r_tag, uint32_t *__restrict__ r_cnt) {
    // [warp][stage][field][k][lane]: the records of the next round arrive by cp.async while the current round is answered
    __shared__ uint32_t stage[PL_THREADS / 32][2][3][PL_U][32];
    extern __shared__ __align__(32) uint8_t pl_smem[];
    const uint32_t b = blockIdx.x / ctas_per_bkt, chunk = blockIdx.x - b * ctas_per_bkt;
    const int64_t s0 = (int64_t)b << G.bshift;
    const SegLite *lite = A.lite + s0;
    if (CACHE) {
        const uint32_t ns = (uint32_t)min((int64_t)1 << G.bshift, n_seg - s0);
        const uint4 *src = reinterpret_cast<const uint4 *>(lite);
        uint4 *dst = reinterpret_cast<uint4 *>(pl_smem);
        for (uint32_t i = threadIdx.x; i < ns; i += PL_THREADS) dst[i] = __ldg(src + i);
        __syncthreads();
        lite = reinterpret_cast<const SegLite *>(pl_smem);
    }
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t g0 = chunk * PL_RUNS + warp * PL_WRUNS;
    if (g0 >= (uint32_t)G.n_span) return;
    // the warp's runs: lane j < PL_WRUNS holds run g0 + j
    uint32_t c = 0, pb = 0;
    if (lane < PL_WRUNS && g0 + lane < (uint32_t)G.n_span) {
        const size_t f = (size_t)b * G.n_span + g0 + lane;
        c = __ldg(run_cnt + f);
        pb = ((g0 + lane) << PT_SPAN_SHIFT) + __ldg(run_base + f);
    }
    uint32_t inc = c;
#pragma unroll
    for (int d = 1; d < PL_WRUNS; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += t;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, PL_WRUNS - 1);
    uint32_t pre[PL_WRUNS], phys[PL_WRUNS];   // first record of run j in the warp's stream; where run j starts in the record arrays
#pragma unroll
    for (int j = 0; j < PL_WRUNS; ++j) {
        pre[j] = __shfl_sync(0xFFFFFFFFu, inc - c, j);
        phys[j] = __shfl_sync(0xFFFFFFFFu, pb, j) - pre[j];   // record v of run j sits at phys[j] + v
    }
    const uint32_t sent = (1u << G.sb) - 1u;
    uint32_t p_next[PL_U], p[PL_U];
    auto prefetch = [&](uint32_t v0, int st) {   // where the records of round v0 are, and their copies into stage st
#pragma unroll
        for (int k = 0; k < PL_U; ++k) {
            const uint32_t v = v0 + k * 32 + lane;
            uint32_t ph = phys[0];
#pragma unroll
            for (int j = 1; j < PL_WRUNS; ++j) ph = v >= pre[j] ? phys[j] : ph;
            p_next[k] = ph + v;
            if (v < total) {
                cp_async4(&stage[warp][st][0][k][lane], r_start + p_next[k]);
                cp_async4(&stage[warp][st][1][k][lane], r_stop + p_next[k]);
                cp_async4(&stage[warp][st][2][k][lane], r_tag + p_next[k]);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(0, 0);
    int st = 0;
    for (uint32_t v0 = 0; v0 < total; v0 += PL_U * 32, st ^= 1) {
#pragma unroll
        for (int k = 0; k < PL_U; ++k) p[k] = p_next[k];
        if (v0 + PL_U * 32 < total) {
            prefetch(v0 + PL_U * 32, st ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        uint32_t qs[PL_U], qe[PL_U], loc[PL_U];
#pragma unroll
        for (int k = 0; k < PL_U; ++k) {   // every lane reads back what it copied itself
            loc[k] = sent;
            if (v0 + k * 32 + lane < total) {
                qs[k] = stage[warp][st][0][k][lane];
                qe[k] = stage[warp][st][1][k][lane];
                loc[k] = stage[warp][st][2][k][lane] & sent;
            }
        }
        uint2 eh[PL_U], el[PL_U];
        uint32_t lo[PL_U], nbl[PL_U];
#pragma unroll
        for (int k = 0; k < PL_U; ++k) {
            nbl[k] = PT_EMPTY;
            if (loc[k] != sent) {
                const uint4 t = *reinterpret_cast<const uint4 *>(lite + loc[k]);   // cofs | lo | nbl
                lo[k] = t.z, nbl[k] = t.w;
                if (nbl[k] < PT_GENERAL) {
                    const uint2 *ent = reinterpret_cast<const uint2 *>(A.cbuf + ((uint64_t)t.x | ((uint64_t)t.y << 32)));
                    eh[k] = ld_ent(ent + e8_entry_index(lo[k], nbl[k] >> 16, nbl[k] & 0xFFFFu, qe[k]));
                    el[k] = ld_ent(ent + e8_entry_index(lo[k], nbl[k] >> 16, nbl[k] & 0xFFFFu, qs[k]));
                }
            }
        }
        uint32_t res[PL_U];
        uint32_t slow = 0;   // bits k / k + 8: the upper / lower endpoint of record k needs the raw array; bit k + 16: not the E8 form
#pragma unroll
        for (int k = 0; k < PL_U; ++k) {
            res[k] = 0;
            if (nbl[k] < PT_GENERAL) {
                const uint32_t l = nbl[k] >> 16;
                // a bucket with more than six elements (flag) keeps its six smallest in the entry: they decide unless the key is
                // beyond the sixth
                const uint32_t mask = (1u << l) - 1u;
                const uint32_t rem_h = (max(qe[k], lo[k]) - lo[k]) & mask, rem_l = (max(qs[k], lo[k]) - lo[k]) & mask;
                const uint32_t last = e8_finish(lo[k], l, eh[k], qe[k]);
                const uint32_t first = (qs[k] + 1u) ? e8_finish(lo[k], l, el[k], qs[k]) : 0u;
                res[k] = last - first;
                if ((eh[k].x & 0x8000u) && rem_h > (eh[k].y >> 24)) slow |= 1u << k;
                if ((qs[k] + 1u) && (el[k].x & 0x8000u) && rem_l > (el[k].y >> 24)) slow |= 1u << (k + 8);
            } else if (nbl[k] == PT_GENERAL) {
                slow |= 1u << (k + 16);
            }
        }
        if (__any_sync(0xFFFFFFFFu, slow != 0)) {   // once per round, not once per record
#pragma unroll
            for (int k = 0; k < PL_U; ++k) {
                if (slow & (1u << (k + 16))) res[k] = lookup_general(A, s0 + loc[k], qs[k], qe[k]);
                if (slow & (0x101u << k)) {
                    const uint4 t = *reinterpret_cast<const uint4 *>(lite + loc[k]);
                    const uint2 *ent = reinterpret_cast<const uint2 *>(A.cbuf + ((uint64_t)t.x | ((uint64_t)t.y << 32)));
                    const uint32_t *a = A.starts + __ldg(A.iv_off + s0 + loc[k]);
                    const uint32_t l = nbl[k] >> 16, nb = nbl[k] & 0xFFFFu;
                    if (slow & (1u << k)) res[k] += e8_crowded_extra(a, ent, e8_entry_index(lo[k], l, nb, qe[k]), eh[k].x & 0x7FFFu, qe[k]);
                    if (slow & (1u << (k + 8))) res[k] -= e8_crowded_extra(a, ent, e8_entry_index(lo[k], l, nb, qs[k]), el[k].x & 0x7FFFu, qs[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < PL_U; ++k)
            if (v0 + k * 32 + lane < total) r_cnt[p[k]] = res[k];
    }
}

// S3: one CTA per span
constexpr int PU_THREADS = 512;
__global__ void __launch_bounds__(PU_THREADS)
part_unpermute_kernel(PartGeom G, const uint32_t *__restrict__ r_tag, const uint32_t *__restrict__ r_cnt, int64_t n_q,
                      uint32_t *__restrict__ out) {
    __shared__ __align__(16) uint32_t s_out[PT_SPAN];
    const int64_t base = (int64_t)blockIdx.x << PT_SPAN_SHIFT;
    const uint32_t m = (uint32_t)min((int64_t)PT_SPAN, n_q - base);
    for (uint32_t j = threadIdx.x; j < (m + 3u) / 4u; j += PU_THREADS) {
        const uint4 t = ld_stream4(r_tag + base + 4 * j), c = ld_stream4(r_cnt + base + 4 * j);
        const uint32_t k = 4u * j;
        s_out[t.x >> G.sb] = c.x;                       // (positions of the padding past m are >= m: never read below, but
        if (k + 1 < m) s_out[t.y >> G.sb] = c.y;        //  their tags are stale, so they are not stored at all)
        if (k + 2 < m) s_out[t.z >> G.sb] = c.z;
        if (k + 3 < m) s_out[t.w >> G.sb] = c.w;
    }
    __syncthreads();
    uint32_t *o = out + base;
    if ((reinterpret_cast<uintptr_t>(o) & 15u) == 0) {
        for (uint32_t j = threadIdx.x; j < m / 4u; j += PU_THREADS) reinterpret_cast<uint4 *>(o)[j] = reinterpret_cast<const uint4 *>(s_out)[j];
        for (uint32_t j = (m & ~3u) + threadIdx.x; j < m; j += PU_THREADS) o[j] = s_out[j];
    } else {
        for (uint32_t j = threadIdx.x; j < m; j += PU_THREADS) o[j] = s_out[j];
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
static int32_t env_int(const char *name, int32_t dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// Buckets: 2^bshift consecutive segments each, at most PT_MAX_B of them, as large as the target of compact index bytes per
// bucket allows (a bucket is what has to stay in L2 while its queries are answered).
static void plan_buckets(const gams_cuda_count_plan *p, int &n_bkt, int &bshift) {
    const int64_t n_seg = p->n_seg;
    std::vector<uint64_t> w((size_t)n_seg + 1, 0);
    for (int64_t s = 0; s < n_seg; ++s) {
        const uint64_t n = p->h_iv_off[s + 1] - p->h_iv_off[s];
        const uint64_t form = 2 * st_arr_cap_bytes((uint32_t)std::min<uint64_t>(n, 0xFFFFFFFFull));
        w[s + 1] = w[s] + (form ? form : 8 * n) + 64;
    }
    const uint64_t target = (uint64_t)std::max(1, env_int("GAMS_CUDA_PART_MB", 12)) << 20;
    auto heaviest = [&](int sh) {
        uint64_t mx = 0;
        for (int64_t s = 0; s < n_seg; s += (int64_t)1 << sh) mx = std::max(mx, w[std::min<int64_t>(n_seg, s + ((int64_t)1 << sh))] - w[s]);
        return mx;
    };
    bshift = 0;
    while ((((int64_t)std::max<int64_t>(n_seg, 1) - 1) >> bshift) + 1 > PT_MAX_B) ++bshift;
    while ((((int64_t)std::max<int64_t>(n_seg, 1) - 1) >> bshift) + 1 > 1 && heaviest(bshift + 1) <= target) ++bshift;
    n_bkt = (int)(((std::max<int64_t>(n_seg, 1) - 1) >> bshift) + 1);
}

bool count_part_eligible(const gams_cuda_count_plan *p, int64_t n_q) {
    const int64_t min_q = (int64_t)env_int("GAMS_CUDA_PART_MIN_Q", 1 << 20);
    return p->mode == 0 && n_q >= min_q && n_q < (1ll << 31) && p->n_seg > 0 && p->total_iv > 0;
}

// tmp[6]: records (start, stop, tag, cnt: 16 B per query, whole spans); tmp[7]: SegLite table;
// tmp[8]: run counts and bases (2 x u16 per (bucket, span))
int32_t launch_count_overlaps_part(gams_cuda_count_plan *p, cudaStream_t st, const uint32_t *d_starts, const uint32_t *d_stops,
                                   int64_t n_q, const int32_t *d_q_seg, const uint32_t *d_q_start, const uint32_t *d_q_stop,
                                   uint32_t *d_out, bool *done) {
    *done = false;
    int32_t rc;
    if (!p->cidx_valid || p->cidx_starts != d_starts || p->cidx_stops != d_stops) {
        if ((rc = build_compact_index(p, st, d_starts, d_stops))) return rc;
        p->part_valid = false;
    }
    if (!p->part_valid) {
        plan_buckets(p, p->part_nbkt, p->part_bshift);
        if ((rc = p->tmp[7].reserve(sizeof(SegLite) * (size_t)std::max<int64_t>(p->n_seg, 1)))) return rc;
        const SegCompact *d_par = reinterpret_cast<const SegCompact *>(p->tmp[4].as<uint8_t>() + p->cpar_off);
        part_lite_kernel<<<(unsigned)((p->n_seg + 255) / 256), 256, 0, st>>>(d_par, p->tmp[4].as<uint64_t>(), p->n_seg, p->tmp[7].as<SegLite>());
        ++p->launches;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaFuncSetAttribute(part_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SORT_SMEM));
        CUDA_TRY(cudaFuncSetAttribute(part_lookup_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(PL_MAX_CACHE_SEGS * sizeof(SegLite))));
        p->part_valid = true;
    }
    PartGeom G;
    G.n_bkt = p->part_nbkt;
    G.bshift = p->part_bshift;
    G.sb = G.bshift + 1;
    if (G.sb + PT_SPAN_SHIFT > 32) return GAMS_CUDA_OK;   // too many segments per bucket for a 32-bit tag: the ranked kernel
    G.n_span = (int32_t)((n_q + PT_SPAN - 1) >> PT_SPAN_SHIFT);
    const size_t n_run = (size_t)G.n_bkt * (size_t)G.n_span, n_pad = (size_t)G.n_span << PT_SPAN_SHIFT;
    if (n_run >= (1ull << 32)) return GAMS_CUDA_OK;
    if ((rc = p->tmp[6].reserve(n_pad * 16 + 64))) return rc;
    if ((rc = p->tmp[8].reserve(n_run * 2 * sizeof(uint16_t) + 64))) return rc;
    uint16_t *d_rcnt = p->tmp[8].as<uint16_t>(), *d_rbase = d_rcnt + n_run;
    uint32_t *r_start = p->tmp[6].as<uint32_t>(), *r_stop = r_start + n_pad, *r_tag = r_stop + n_pad, *r_cnt = r_tag + n_pad;
    part_sort_kernel<<<(unsigned)G.n_span, PT_THREADS, PT_SORT_SMEM, st>>>(d_q_seg, d_q_start, d_q_stop, n_q, p->n_seg, G, d_rcnt,
                                                                           d_rbase, r_start, r_stop, r_tag);
    LookupArgs A;
    A.starts = d_starts, A.stops = d_stops;
    A.iv_off = p->d_iv_off;
    A.lite = p->tmp[7].as<SegLite>();
    A.par = reinterpret_cast<const SegCompact *>(p->tmp[4].as<uint8_t>() + p->cpar_off);
    A.cbuf = p->tmp[5].as<uint8_t>();
    const uint32_t ctas_per_bkt = (uint32_t)((G.n_span + PL_RUNS - 1) / PL_RUNS);
    const size_t lgrid = (size_t)ctas_per_bkt * G.n_bkt;
    if (lgrid >= (1ull << 31)) return GAMS_CUDA_OK;
    if (((int64_t)1 << G.bshift) <= PL_MAX_CACHE_SEGS)
        part_lookup_kernel<true><<<(unsigned)lgrid, PL_THREADS, (sizeof(SegLite) << G.bshift), st>>>(
            A, p->n_seg, G, ctas_per_bkt, d_rcnt, d_rbase, r_start, r_stop, r_tag, r_cnt);
    else
        part_lookup_kernel<false><<<(unsigned)lgrid, PL_THREADS, 0, st>>>(A, p->n_seg, G, ctas_per_bkt, d_rcnt, d_rbase, r_start, r_stop,
                                                                         r_tag, r_cnt);
    part_unpermute_kernel<<<(unsigned)G.n_span, PU_THREADS, 0, st>>>(G, r_tag, r_cnt, n_q, d_out);
    p->launches += 3;
    CUDA_TRY(cudaGetLastError());
    *done = true;
    return GAMS_CUDA_OK;
}

}  // namespace gams
