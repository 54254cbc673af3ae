// overlap_index.cuh — the compact per-segment forms of the sorted endpoint arrays of Kernel C (built by build_compact_kernel,
// overlap.cu) and the lower_bound evaluations on them; shared by the staged kernel (forms copied to shared memory, overlap.cu)
// and the partitioned kernel for queries in arbitrary order (forms read through L2, overlap_part.cu).
#pragma once
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

#ifndef ST_THREADS_N
#define ST_THREADS_N 1024   // two CTAs per SM = 64 warps: the two-table lookups of general intervals are latency-bound (0.673 -> 0.594 ms)
#endif
constexpr int ST_THREADS = ST_THREADS_N;
constexpr int ST_NB = 8192;          // buckets per endpoint array
constexpr int ST_MAX_N = 24576;      // elements per endpoint array
constexpr int ST_CHUNK = 16384;      // queries per work item
constexpr uint32_t ST_PLAIN = 0xFFu;
constexpr uint32_t ST_ALIAS = 0xFEu;   // stops[i] == starts[i] + 1 for every i (point intervals): lower_bound(stops, k) = lower_bound(starts, k - 1)
constexpr size_t ST_ARR_BYTES = 4 * (ST_NB + 4) + ST_MAX_N + 32;
constexpr size_t ST_SMEM = 2 * ST_ARR_BYTES;

struct alignas(16) SegCompact {      // one per (segment, endpoint array)
    uint32_t lo, hi, n, l;           // l == ST_PLAIN: no compact form
    uint32_t nb;                     // buckets
    uint32_t low_off;                // byte offset of low[] behind cum2[] (multiple of 16)
    uint32_t bytes;                  // bytes of the compact form (multiple of 16)
    uint32_t fmt;                    // 0: cum2[] + low[] (two arrays);  1: one 8-byte entry per bucket (E8, point intervals)
};

// bucket budget of an array of n elements (host and device agree on it: the host sizes the buffer with it)
__host__ __device__ __forceinline__ uint32_t st_nb_cap(uint32_t n) {
    return n < 64u ? 64u : (n > (uint32_t)ST_NB ? (uint32_t)ST_NB : n);
}
__host__ __device__ __forceinline__ uint64_t st_arr_cap_bytes(uint32_t n) {   // upper bound of SegCompact::bytes
    if (n == 0 || n > (uint32_t)ST_MAX_N) return 0;
    return ((4ull * (st_nb_cap(n) + 1) + 15) & ~15ull) + (((uint64_t)n + 16 + 15) & ~15ull);
}

// #elements < key of one endpoint array: compact form in shared memory (sm), or the raw array `a` in global memory
__device__ __forceinline__ uint32_t staged_lower_bound(const SegCompact &P, const uint32_t *__restrict__ a, const uint8_t *sm,
                                                       uint32_t key) {
    if (P.l == ST_PLAIN) return lower_bound_u32(a, P.n, key);
    // clamp instead of branching: key <= lo gives d = 0 (bucket 0, rem 0: nothing below), key > hi is patched at the end
    const uint32_t d = max(key, P.lo) - P.lo, b = min(d >> P.l, P.nb - 1u), rem = d & ((1u << P.l) - 1u);
    const uint32_t w = reinterpret_cast<const uint32_t *>(sm)[b];
    const uint32_t c0 = w & 0xFFFFu, m = (w >> 16) - c0;                 // the bucket's elements are low[c0 .. c0 + m)
    const uint8_t *low = sm + P.low_off;
    uint32_t cnt;
    if (m > 8u) {                                                         // rare: a crowded bucket
        cnt = c0;
        for (uint32_t i = 0; i < m; ++i) cnt += (low[c0 + i] < rem) ? 1u : 0u;
    } else {
        // 8-byte window of low[] at c0 out of three aligned words; byte < rem <=> bit 7 of (byte | 0x80) - rem is clear
        const uint32_t *lw = reinterpret_cast<const uint32_t *>(low) + (c0 >> 2);
        const uint32_t x0 = lw[0], x1 = lw[1], x2 = lw[2], sh = (c0 & 3u) * 8u;
        const uint32_t y0 = __funnelshift_r(x0, x1, sh), y1 = __funnelshift_r(x1, x2, sh);
        // byte < rem <=> bit 7 of (byte | 0x80) - rem is clear; byte index < m <=> bit 7 of (0x80 + index) - m is clear
        const uint32_t rr = rem * 0x01010101u, mm = m * 0x01010101u;
        const uint32_t a0 = (y0 | 0x80808080u) - rr, a1 = (y1 | 0x80808080u) - rr;
        const uint32_t i0 = 0x83828180u - mm, i1 = 0x87868584u - mm;
        cnt = c0 + (uint32_t)__popc(~a0 & ~i0 & 0x80808080u) + (uint32_t)__popc(~a1 & ~i1 & 0x80808080u);
    }
    return key > P.hi ? P.n : cnt;
}

// E8 format: #elements < key with one 64-bit load.  Entry b: cum | overflow << 15 | six low bytes (0xFF = empty slot).
// Entry nb is a sentinel (cum = n, no elements): every key past the last bucket lands there, so key > hi needs no
// special case (inside the last bucket the byte compare already counts every element).
__device__ __forceinline__ uint32_t e8_lower_bound(const SegCompact &P, const uint32_t *__restrict__ a, const uint2 *ent, uint32_t key) {
    const uint32_t d = max(key, P.lo) - P.lo, b = min(d >> P.l, P.nb), rem = d & ((1u << P.l) - 1u);
    const uint2 e = ent[b];
    const uint32_t cum = e.x & 0x7FFFu, rr = rem * 0x01010101u;
    // byte < rem <=> bit 7 of (byte | 0x80) - rem is clear; the two cum bytes are forced to 0xFF, empty slots are 0xFF
    uint32_t cnt = cum + (uint32_t)__popc(~((e.x | 0x8080FFFFu) - rr) & 0x80808080u) + (uint32_t)__popc(~((e.y | 0x80808080u) - rr) & 0x80808080u);
    // rare: more than six elements in the bucket; the entry holds the six smallest, so the seventh and later ones (compared in
    // the raw array) matter only for a key beyond the sixth
    if ((e.x & 0x8000u) && rem > (e.y >> 24)) {
        const uint32_t nxt = ent[b + 1].x & 0x7FFFu;
#pragma unroll 1
        for (uint32_t i = cum + 6u; i < nxt; ++i) cnt += (__ldg(a + i) < key) ? 1u : 0u;
    }
    return cnt;
}

}  // namespace gams
