// gc_wave_fused.cu — Kernel AB: sequence bytes -> per-window GC AND peak signals in one pass (influence == 1).
//
// Replaces wave::proc_ctg's compute, src/cmd_gams/wave.rs:138-155: gams::sliding (window.rs:78-94), the
// per-window gc_content loop (wave.rs:143-153) and gams::thresholding_algo (stat.rs:16-56).
//
// Why fuse.  With a certified filter the z-score step does not need the f32 GC values: it needs the exact mean
// and centred sum of squares of each stats window.  For GC data those follow from INTEGER window counts c_k
// (x_k = fl(c_k / size)), which the GC phase already holds in shared memory.  So the tile is extended backwards
// by lag+1 windows, the counts are scanned twice more (sum c mod 2^16, sum c^2 mod 2^32: exact differences),
// and each window is decided from D = n c_i - sum c and V = n sum c^2 - (sum c)^2 (both exact u32):
//     signal certainly      if (|D| - e2)^2  >  K_hi ((1+eps) V + B_hi)        and |D| > e2
//     certainly no signal   if (|D| + e2)^2  <  K_lo max((1-eps) V - B_lo, 0)
// where e2, K_*, B_*, eps are host-computed bounds on everything the reference's f32 evaluation can do to these
// exact quantities, widened for the f32 arithmetic of the test itself (derivation at make_fused_const and in
// DESIGN.md §3 "Kernel AB"); otherwise the window is replayed exactly (zscore_decide on x_k = gc_tab[c_k], the same
// f32 values the reference sums).  No pass re-reads the 1.24 GB of GC values and no f64 is involved.
//
// The kernel is instruction-issue bound, so the per-window phases are laid out for few instructions: eight
// consecutive windows per thread, 128-bit shared-memory accesses, everything unrolled.
//
// Bound: HBM (1 B/base read + 5 B/window written, + the lag+1-window halo re-read per tile, 5 % at 100/10/100).
#include "gc_common.cuh"
#include "zscore_common.cuh"

namespace gams {

#ifndef FW_MIN_CTAS
#define FW_MIN_CTAS 8   // 64 warps per SM at 32 registers (a few spilled bytes): 0.928 -> 0.913 ms; 5 CTAs per SM: 1.109 ms
#endif
constexpr int FW_NCH = GC_THREADS * FW_K;  // 16-byte chunks per tile
constexpr int FW_NBLK = ((FW_MAX_EL + FW_MAX_R + GC_THREADS - 1) / GC_THREADS) * GC_THREADS;  // block sums per tile
static_assert((31 + (FW_NBLK - 1) * 16) / 16 < FW_NBLK + 4, "stale block reads stay inside WW");
static_assert(FW_MAX_EL == GC_THREADS * FW_E && FW_E == 8 && FW_PAD % 8 == 0 && FW_PAD >= FW_MAX_LAG, "fused tile geometry");

// v + (value of lane - d, if that lane exists): one SHFL and one predicated add
__device__ __forceinline__ uint32_t scan_step(uint32_t v, int d) {
    uint32_t out;
    asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 t;\n\tshfl.sync.up.b32 t|p, %1, %2, 0, 0xffffffff;\n\tmov.u32 %0, %1;\n\t@p add.u32 %0, %1, t;\n\t}"
                 : "=r"(out) : "r"(v), "r"(d));
    return out;
}

// Decision for one window from the exact integers D = n c_i - sum c and V = n sum c^2 - (sum c)^2:
// z = D^2 - k V against a margin that covers every rounding of the reference and of this test (make_fused_const).
// Returns 1 / 0xFF (signal up / down), 0 (certainly none) or 2 (inconclusive: replay exactly).
__device__ __forceinline__ uint32_t fused_certify(int D, uint32_t V, const FusedConst &fc) {
    const float fd = fabsf(__int2float_rn(D));      // exact: |D| < 2^16
    const float fv = __uint2float_rn(V);
    const float t = fd * fd;
    const float z = fmaf(-fc.k, fv, t);
    const float m = fmaf(fd, fc.e2x2, fmaf(fv, fc.epsk, fc.beta));
    const uint32_t sgn = (D > 0) ? 1u : 0xFFu;
    const uint32_t cert = (z > 0.0f) ? sgn : 0u;
    return (fabsf(z) > m) ? cert : 2u;
}

// GC bit of each of the 32 bytes in v (bit i <=> byte i is one of C G c g): the four flags of a word are gathered
// into the top nibble by one multiply (gc_gather) and shifted into the mask by one funnel shift.
// `one` is the constant 1 passed as a kernel argument: t * one + 0x7F.. keeps that add on the FMA pipe (IMAD), which
// idles here, instead of the ALU pipe, which is the busiest unit of this kernel.
__device__ __forceinline__ uint32_t gc_gather_fma(uint32_t w, uint32_t one) {
    const uint32_t t = lop3<0x6A>(w, 0x5B5B5B5Bu, 0x43434343u);
    const uint32_t a = t * one + 0x7F7F7F7Fu;
    const uint32_t g7 = lop3<0x02>(a, w, 0x80808080u);
    return g7 * 0x00204081u;
}
// u8 x s8 dot product with accumulator (IDP.4A, FMA pipe)
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// flags in bit 7 of every byte (0x80 / 0x00), not yet gathered
__device__ __forceinline__ uint32_t gc_flags_fma(uint32_t w, uint32_t one) {
    const uint32_t t = lop3<0x6A>(w, 0x5B5B5B5Bu, 0x43434343u);
    const uint32_t a = t * one + 0x7F7F7F7Fu;
    return lop3<0x02>(a, w, 0x80808080u);
}
// Two words per dot-product chain: weights 1 2 4 8 / 16 32 64 128 put the eight flags (each worth 128) of a word pair
// into bits 7..14; the four pairs are merged by two multiply-adds and one shift: 17 ALU + 19 FMA-pipe instructions
// per 32 bytes instead of 24 + 16.
__device__ __forceinline__ uint32_t gc_mask32(const Bytes32 &v, uint32_t one) {
    uint32_t P[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        P[q] = __dp4a(gc_flags_fma(v.w[2 * q], one), 0x08040201u, __dp4a(gc_flags_fma(v.w[2 * q + 1], one), 0x80402010u, 0u));
    const uint32_t lo = P[1] * 256u + P[0], hi = P[3] * 256u + P[2];   // 128 x (16 flags)
    return hi * 512u + (lo >> 7);
}

// bits [P, P + 10) of the bit string x[0] | x[1] << 32 | ..., UNMASKED above the field (P is a compile-time constant)
template <int P>
__device__ __forceinline__ uint32_t bits_from(const uint32_t (&x)[6]) {
    constexpr int w = P >> 5, o = P & 31;
    if (o == 0) return x[w];
    if (o <= 22) return x[w] >> o;
    return __funnelshift_r(x[w], x[w + 1], o);
}
// popc(in field) - popc(out field) + 10 for two 10-bit fields, with ONE population count:
// the inverted out-field is placed above the in-field
template <int PIN, int POUT>
__device__ __forceinline__ uint32_t popc_in_minus_out_plus10(const uint32_t (&x)[6]) {
    const uint32_t a = bits_from<PIN>(x) & 0x3FFu;
    const uint32_t b = ~bits_from<POUT>(x) & 0x3FFu;
    return (uint32_t)__popc(b * 1024u + a);
}

// Tile layout.  A tile holds FW_MAX_EL consecutive windows ("elements") of one ctg; element e is global window
// g0 + e with g0 = 7 (mod 8), so that the eight elements 8t+1 .. 8t+8 a thread decides start at a multiple of 8 in the
// output arrays: 32 bytes of GC fractions go out as two 128-bit stores and the eight signals as one 64-bit store.
// Elements [e_out0, e_out1) are the tile's outputs; those before are the lag+1-window halo (or, in a ctg's first tile,
// the 1-8 alignment elements before the ctg's window 0, which read the neighbouring bytes and are never used).
//
// LAG4: lag % 4 == 0, so the counts lag windows back sit at 8-byte aligned shared-memory addresses.
// R: size / step when known at compile time (10), 0 = read it from `r_rt`.
// S10: size == 100 and step == 10 (the reference's documented geometry, README `gams wave --size 100 --step 10`): the
// window counts come straight out of the bit array, without the per-block sums and their barrier.
template <bool WRITE_CNT, bool LAG4, int R, bool S10>
__global__ void __launch_bounds__(GC_THREADS, FW_MIN_CTAS)
gc_wave_fused_kernel(const uint8_t *__restrict__ seq, const FusedTile *__restrict__ tiles,
                     const float *__restrict__ gc_tab, int size, int step, int r_rt, int lag, float threshold,
                     uint32_t one, const FusedConst fc, uint32_t *__restrict__ cnt_out, float *__restrict__ gc_out,
                     int8_t *__restrict__ sig_out, unsigned long long *__restrict__ stats) {
    // WW[i] = GC bits of bytes [16 i, 16 i + 32) of the tile's span (consecutive words overlap by 16 bits so that a
    // block of <= 16 bases always comes out of one word)
    __shared__ __align__(16) uint32_t WW[S10 ? FW_NCH / 2 + 8 : FW_NBLK + 4];   // S10: plain bit array, word i = bits [32 i, 32 i + 32)
    __shared__ __align__(16) uint8_t R1[S10 ? 16 : FW_NBLK + 8];   // bsum[]: GC bases per step-sized block
    __shared__ __align__(16) uint16_t cwp[FW_PAD + FW_MAX_EL + 8];  // window counts c_e at cwp[FW_PAD + e]
    __shared__ __align__(16) uint2 Pq[(FW_PAD + FW_MAX_EL) / 4];    // (sum c, sum c^2) over elements < 4 m, at Pq[FW_PAD/4 + m]
    __shared__ float stab[FW_STAB];                                 // gc_tab when size < FW_STAB
    __shared__ uint32_t ws1[GC_THREADS / 32], ws2[GC_THREADS / 32];
    static_assert(FW_NBLK + 4 >= FW_NCH + 1, "WW holds the bit array");
    uint16_t *const WW16 = reinterpret_cast<uint16_t *>(WW);
    uint16_t *const cw = cwp + FW_PAD;
    uint8_t *const bsum = R1;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const FusedTile td = tiles[blockIdx.x];
    const uint32_t e_out0 = td.e_out0, e_out1 = td.e_out1;          // e_out1 is also the number of elements the tile holds
    const uint32_t lead = (uint32_t)(td.byte0 & 31);
    const uint32_t r = R ? (uint32_t)R : (uint32_t)r_rt;
    const uint32_t span = (e_out1 - 1u) * (uint32_t)step + (uint32_t)size;
    const uint32_t nch = (lead + span + 15u) >> 4;                  // <= FW_NCH by construction of the tiles
    const bool use_stab = size < FW_STAB;
    if (use_stab && tid <= size) stab[tid] = __ldg(gc_tab + tid);
    // Elements before the tile count as zeros: a thread whose first stats window starts before element 0 then slides
    // into correct sums by the time its window lies inside the tile.
    if (tid < FW_PAD / 8) reinterpret_cast<uint4 *>(cwp)[tid] = make_uint4(0u, 0u, 0u, 0u);
    else if (tid < FW_PAD / 4) reinterpret_cast<uint4 *>(Pq)[tid - FW_PAD / 8] = make_uint4(0u, 0u, 0u, 0u);

    // ---- phase 1: classify, 32 bytes per lane per load (see gc_common.cuh); one bit per base into WW
    {
        const uint8_t *src = seq + (td.byte0 - (int64_t)lead);
        Bytes32 v[FW_K / 2];
#pragma unroll
        for (int j = 0; j < FW_K / 2; ++j) {
            const uint32_t pi = j * GC_THREADS + tid;
            if (2u * pi < nch && pi >= td.pi_min) v[j] = ld_stream32(src + 32u * pi);
        }
#pragma unroll
        for (int j = 0; j < FW_K / 2; ++j) {
            const uint32_t pi = j * GC_THREADS + tid;
            if (2u * (pi - lane) < nch) {                           // warp-uniform
                const uint32_t m = gc_mask32(v[j], one);
                if (S10) {                                          // plain bit array: word pi = bits [32 pi, 32 pi + 32)
                    WW[pi] = m;
                    continue;
                }
                const uint32_t nx = __shfl_down_sync(0xFFFFFFFFu, m, 1);
                if (lane != 31) *reinterpret_cast<uint2 *>(WW + 2 * pi) = make_uint2(m, __funnelshift_r(m, nx, 16));
                else {
                    WW[2 * pi] = m;
                    WW16[4 * pi + 2] = (uint16_t)(m >> 16);         // low half of WW[2 pi + 1] ...
                }
                if (lane == 0 && pi) WW16[4 * pi - 1] = (uint16_t)m;  // ... whose high half comes from the next warp's lane 0
            }
        }
    }
    __syncthreads();

    // ---- GC bases of every step-sized block: bsum[m] = popc(bits [lead + m step, + step)).  Branch-free: blocks past
    //      the tile's last boundary read stale (in-bounds) words and are never used.
    if (!S10) {
        const uint32_t smask = (1u << step) - 1u;                   // step <= 16
        const uint32_t p0 = lead + (uint32_t)tid * (uint32_t)step;
        const uint32_t sh = p0 & 15u;                               // 256 * step is a multiple of 16: the same for every j
        const uint32_t *wp = WW + (p0 >> 4);
        const uint32_t dw = 16u * (uint32_t)step;                   // words between blocks m and m + 256
#pragma unroll
        for (int j = 0; j < FW_NBLK / GC_THREADS; ++j, wp += dw)
            bsum[j * GC_THREADS + tid] = (uint8_t)__popc((*wp >> sh) & smask);
        __syncthreads();
    }

    // ---- window counts by sliding over the block sums (thread t owns windows [8t, 8t+8)), then exact prefix sums of
    //      c and c^2 over the tile, kept for every fourth element
    uint32_t c[FW_E + 1];
    uint32_t q1, q2;                            // sums over the elements before 8t
    {
        const uint8_t *bt = bsum + FW_E * tid;
        if (S10) {
            // the thread's eight windows are bits [p0, p0 + 170) of the tile's bit array: seven words, shifted so that
            // block k sits at bits [10 k, 10 k + 10); window 0 is one 100-bit population count and window j + 1 differs
            // from window j by the block that enters and the block that leaves.  Words past the tile's last chunk are
            // stale (in bounds) and only reach elements the tile does not have.
            const uint32_t p0 = lead + 80u * (uint32_t)tid;
            const uint32_t *wp = WW + (p0 >> 5);
            const uint32_t sh = p0 & 31u;
            uint32_t w[7], x[6];
#pragma unroll
            for (int k = 0; k < 7; ++k) w[k] = wp[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) x[k] = __funnelshift_r(w[k], w[k + 1], sh);
            c[0] = (uint32_t)(__popc(x[0]) + __popc(x[1])) + (uint32_t)(__popc(x[2]) + __popc(x[3] & 0xFu));
            c[1] = c[0] + popc_in_minus_out_plus10<100, 0>(x) - 10u;
            c[2] = c[1] + popc_in_minus_out_plus10<110, 10>(x) - 10u;
            c[3] = c[2] + popc_in_minus_out_plus10<120, 20>(x) - 10u;
            c[4] = c[3] + popc_in_minus_out_plus10<130, 30>(x) - 10u;
            c[5] = c[4] + popc_in_minus_out_plus10<140, 40>(x) - 10u;
            c[6] = c[5] + popc_in_minus_out_plus10<150, 50>(x) - 10u;
            c[7] = c[6] + popc_in_minus_out_plus10<160, 60>(x) - 10u;
        } else if (R == 10) {
            const uint2 a = *reinterpret_cast<const uint2 *>(bt), b = *reinterpret_cast<const uint2 *>(bt + 8);
            const uint32_t d = *reinterpret_cast<const uint32_t *>(bt + 16);
            // byte sums and byte differences as dot products (IDP.4A, FMA pipe) instead of extract + add on the ALU pipe
            c[0] = (uint32_t)dp4a_us(a.x, 0x01010101, dp4a_us(a.y, 0x01010101, dp4a_us(b.x, 0x00000101, 0)));
            c[1] = (uint32_t)dp4a_us(b.x, 0x00010000, dp4a_us(a.x, (int)0x000000FF, (int)c[0]));
            c[2] = (uint32_t)dp4a_us(b.x, 0x01000000, dp4a_us(a.x, (int)0x0000FF00, (int)c[1]));
            c[3] = (uint32_t)dp4a_us(b.y, 0x00000001, dp4a_us(a.x, (int)0x00FF0000, (int)c[2]));
            c[4] = (uint32_t)dp4a_us(b.y, 0x00000100, dp4a_us(a.x, (int)0xFF000000, (int)c[3]));
            c[5] = (uint32_t)dp4a_us(b.y, 0x00010000, dp4a_us(a.y, (int)0x000000FF, (int)c[4]));
            c[6] = (uint32_t)dp4a_us(b.y, 0x01000000, dp4a_us(a.y, (int)0x0000FF00, (int)c[5]));
            c[7] = (uint32_t)dp4a_us(d, 0x00000001, dp4a_us(a.y, (int)0x00FF0000, (int)c[6]));
        } else {
            uint32_t s = 0;
            for (uint32_t m = 0; m < r; ++m) s += bt[m];
            c[0] = s;
#pragma unroll
            for (int j = 1; j < FW_E; ++j) c[j] = c[j - 1] + bt[r + j - 1] - bt[j - 1];
        }
        reinterpret_cast<uint4 *>(cw)[tid] = make_uint4(c[0] | (c[1] << 16), c[2] | (c[3] << 16), c[4] | (c[5] << 16), c[6] | (c[7] << 16));
        if (tid < 4) reinterpret_cast<uint32_t *>(cw + FW_MAX_EL)[tid] = 0u;
        const uint32_t h1 = c[0] + c[1] + c[2] + c[3], h2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3];
        const uint32_t a1 = h1 + c[4] + c[5] + c[6] + c[7], a2 = h2 + c[4] * c[4] + c[5] * c[5] + c[6] * c[6] + c[7] * c[7];
        uint32_t s1 = a1, s2 = a2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            s1 = scan_step(s1, d);
            s2 = scan_step(s2, d);
        }
        if (lane == 31) { ws1[warp] = s1; ws2[warp] = s2; }
        __syncthreads();
        const uint32_t o1 = (lane < warp) ? ws1[lane & 7] : 0u, o2 = (lane < warp) ? ws2[lane & 7] : 0u;
        q1 = s1 - a1 + __reduce_add_sync(0xFFFFFFFFu, o1);
        q2 = s2 - a2 + __reduce_add_sync(0xFFFFFFFFu, o2);
        reinterpret_cast<uint4 *>(Pq + FW_PAD / 4)[tid] = make_uint4(q1, q2, q1 + h1, q2 + h2);
    }
    __syncthreads();

    // ---- GC fractions of elements 8t+1 .. 8t+8: 32 aligned bytes per thread
    const uint32_t e0 = FW_E * tid + 1;
    const bool whole = e0 >= e_out0 && e0 + FW_E <= e_out1;         // all eight elements are outputs of this tile
    c[FW_E] = cw[FW_E * tid + FW_E];
    {
        float g[FW_E];
        if (use_stab) {
#pragma unroll
            for (int j = 0; j < FW_E; ++j) g[j] = stab[c[j + 1] & (FW_STAB - 1)];
        } else {
#pragma unroll
            for (int j = 0; j < FW_E; ++j) g[j] = __ldg(gc_tab + min(c[j + 1], (uint32_t)size));
        }
        float *__restrict__ out = gc_out + (td.g0 + (int64_t)e0);
        uint32_t *__restrict__ outc = WRITE_CNT ? cnt_out + (td.g0 + (int64_t)e0) : nullptr;
        if (whole) {
            reinterpret_cast<float4 *>(out)[0] = make_float4(g[0], g[1], g[2], g[3]);
            reinterpret_cast<float4 *>(out)[1] = make_float4(g[4], g[5], g[6], g[7]);
            if (WRITE_CNT) {
                reinterpret_cast<uint4 *>(outc)[0] = make_uint4(c[1], c[2], c[3], c[4]);
                reinterpret_cast<uint4 *>(outc)[1] = make_uint4(c[5], c[6], c[7], c[8]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < FW_E; ++j)
                if (e0 + j >= e_out0 && e0 + j < e_out1) {
                    out[j] = g[j];
                    if (WRITE_CNT) outc[j] = c[j + 1];
                }
        }
    }

    // ---- certificate: thread t decides elements e = 8t+1 .. 8t+8.  The stats window of e is [e-lag-1, e-1): for
    //      e = 8t+1 its sums are (sums before 8t) - (sums before 8t-lag), and they slide by one count in, one out.
    const int e_sp = td.e_ctg0 + lag;           // the ctg's window `lag` (only in a ctg's first tile; else far negative)
    const uint32_t e_lo = (uint32_t)max((int)e_out0, e_sp + 1);   // first element that can signal by the general formula
    {
        const int i0 = FW_E * tid - lag;        // first element of the first stats window; >= -FW_PAD
        uint32_t cb[FW_E];                      // counts leaving the stats window: c[i0 + j]
        uint32_t s1 = q1, s2 = q2;
        if (LAG4) {
            const uint2 u0 = *reinterpret_cast<const uint2 *>(cw + i0), u1 = *reinterpret_cast<const uint2 *>(cw + i0 + 4);
            cb[0] = u0.x & 0xFFFFu; cb[1] = u0.x >> 16; cb[2] = u0.y & 0xFFFFu; cb[3] = u0.y >> 16;
            cb[4] = u1.x & 0xFFFFu; cb[5] = u1.x >> 16; cb[6] = u1.y & 0xFFFFu; cb[7] = u1.y >> 16;
            const uint2 pb = Pq[(FW_PAD + i0) >> 2];
            s1 -= pb.x;
            s2 -= pb.y;
        } else {
#pragma unroll
            for (int j = 0; j < FW_E; ++j) cb[j] = cw[i0 + j];
            const int m4 = (FW_PAD + i0) >> 2, rho = (FW_PAD + i0) & 3;   // rho is the same for every thread
            uint2 pb = Pq[m4];
            for (int k = 0; k < rho; ++k) {
                const uint32_t x = cw[4 * m4 - FW_PAD + k];
                pb.x += x;
                pb.y += x * x;
            }
            s1 -= pb.x;
            s2 -= pb.y;
        }
        uint32_t pk0 = 0u, pk1 = 0u;
#pragma unroll
        for (int j = 0; j < FW_E; ++j) {
            const int D = (int)(c[j + 1] * (uint32_t)lag) - (int)s1;
            const uint32_t V = s2 * (uint32_t)lag - s1 * s1;
            const uint32_t dec = fused_certify(D, V, fc);
            if (j < 4) pk0 = dec * (1u << (8 * (j & 3))) + pk0;
            else pk1 = dec * (1u << (8 * (j & 3))) + pk1;
            s1 += c[j] - cb[j];
            s2 += c[j] * c[j] - cb[j] * cb[j];
        }
        if (e0 < e_lo || e0 + FW_E > e_out1) {   // edge threads: elements that cannot signal (before `lag`) or do not exist
#pragma unroll
            for (int j = 0; j < FW_E; ++j)
                if (e0 + j < e_lo || e0 + j >= e_out1) {
                    if (j < 4) pk0 &= ~(0xFFu << (8 * (j & 3)));
                    else pk1 &= ~(0xFFu << (8 * (j & 3)));
                }
        }
        // ---- exact replay of inconclusive windows by the lane that owns them (rare: 7.6e-6 of windows): x_k = gc_tab[c_k]
        //      are the f32 values the reference sums.  Decisions never leave the thread, so no barrier and no list.
        // bytes equal to 2 (bit 1 set, bit 0 clear; 0xFF has both): some window is inconclusive
        const bool own_sp = e_sp >= (int)e0 && e_sp < (int)(e0 + FW_E) && (uint32_t)e_sp < e_out1;
        if ((((pk0 & ~(pk0 << 1)) | (pk1 & ~(pk1 << 1))) & 0x02020202u) || own_sp) {
            const float flag = __ull2float_rn((unsigned long long)lag);
            const float flagm1 = __fsub_rn(flag, 1.0f);
            auto ld = [&](int64_t e) { return __ldg(gc_tab + cw[e]); };
            unsigned long long n_rep = 0;
#pragma unroll 1
            for (int j = 0; j < FW_E; ++j) {
                const uint32_t e = e0 + (uint32_t)j;
                uint32_t cur = (((j < 4) ? pk0 : pk1) >> (8 * (j & 3))) & 0xFFu;
                if (own_sp && (int)e == e_sp) {
                    // window `lag` of the ctg takes its statistics from windows [0, lag) (stat.rs:30-31), not [lag-1-lag, lag-1)
                    cur = (uint32_t)zscore_decide(ld, (int64_t)td.e_ctg0, (uint64_t)lag, flag, flagm1, ld(e), threshold) & 0xFFu;
                } else if (cur == 2u) {
                    cur = (uint32_t)zscore_decide(ld, (int64_t)e - 1 - (int64_t)lag, (uint64_t)lag, flag, flagm1, ld(e), threshold) & 0xFFu;
                    ++n_rep;
                } else continue;
                const uint32_t sh = 8u * (uint32_t)(j & 3), keep = ~(0xFFu << sh);
                if (j < 4) pk0 = (pk0 & keep) | (cur << sh);
                else pk1 = (pk1 & keep) | (cur << sh);
            }
            if (stats && n_rep) atomicAdd(&stats[0], n_rep);
        }
        // ---- signals of elements 8t+1 .. 8t+8: one aligned 64-bit store per thread
        int8_t *__restrict__ so = sig_out + (td.g0 + (int64_t)e0);
        if (whole) *reinterpret_cast<uint2 *>(so) = make_uint2(pk0, pk1);
        else {
#pragma unroll
            for (int j = 0; j < FW_E; ++j)
                if (e0 + j >= e_out0 && e0 + j < e_out1) so[j] = (int8_t)(((j < 4) ? pk0 : pk1) >> (8 * (j & 3)));
        }
    }
}

// Host-side bounds of the certificate (derivation: DESIGN.md §3 "Kernel AB").  With u = 2^-24, n = lag,
// x_j = fl(c_j/size) in [0,1], x^c_j = c_j/size exactly, mu_c and S_c the exact mean / centred sum of squares of x^c:
//   |m^ - mu_x|           <= E_m  (recursive f32 sum of n terms + one divide: ~ n u, taken as 1.01 (n+2) u)
//   |fl(x_i - m^) - (x^c_i - mu_c)| <= E_m + 4u =: E2      (x_i vs x^c_i, mu_x vs mu_c, the subtraction's rounding)
//   fl(thr * sd^) = thr sqrt(Q*/(n-1)) (1 +- R),  R = (0.6 n + 8) u,  Q* = sum (x_j - m^)^2 in [S_x, S_x + n E_m^2]
//   sqrt(S_x) = sqrt(S_c) +- sqrt(n) u  =>  S_x in [(1-eps) S_c - (1/eps - 1) n u^2, (1+eps) S_c + (1 + 1/eps) n u^2]
// and everything is scaled by n*size (D = n*size*(x^c_i - mu_c)) resp. n*size^2 (V = n*size^2*S_c).
// The device evaluates the two tests in f32 (|D| exact, V rounded once, one rounding per operation): the
// constants are widened by 8u and rounded away from the uncertain band so that a computed "certain" implies the
// exact-arithmetic one.
static float f_up(double x) {
    float f = (float)x;
    if ((double)f < x) f = nextafterf(f, INFINITY);
    return f;
}

FusedConst make_fused_const(int32_t size, uint64_t lag, float threshold) {
    const double u = 5.9604644775390625e-8;
    const float flag = (float)lag, flagm1 = flag - 1.0f;  // `data.len() as f32`, `len - 1.`
    const double n = (double)lag, nm1 = (double)flagm1, sz = (double)size;
    const double E_m = 1.01 * (n + 2.0) * u;
    const double E2 = E_m + 4.0 * u;
    const double R = (0.6 * n + 8.0) * u;
    const double eps = 2e-6;
    const double b2 = 4.0 * n * u * u;            // 4x the n u^2 needed
    const double k = (double)threshold * (double)threshold * n / nm1;
    const double rho_hi = (1.0 + R) * (1.0 + R) * (1.0 + 3.0 * u) - 1.0, rho_lo = 1.0 - (1.0 - R) * (1.0 - R) * (1.0 - 3.0 * u);
    const double e2 = E2 * n * sz;                                                   // on |D|
    const double B_hi = n * sz * sz * (b2 * (1.0 + 1.0 / eps) + n * E_m * E_m);      // on V
    const double B_lo = n * sz * sz * (b2 * (1.0 / eps - 1.0));
    // exact tests:  signal        <=  (|D| - e2)^2 > k (1 + rho_hi) ((1 + eps) V + B_hi)   and |D| > e2
    //               no signal     <=  (|D| + e2)^2 < k (1 - rho_lo) max((1 - eps) V - B_lo, 0)
    // both follow from |D^2 - k V| > 2 e2 |D| + k eps' V + beta with the sign of D^2 - k V choosing between them:
    const double eps1 = (1.0 + rho_hi) * (1.0 + eps) - 1.0, eps2 = 1.0 - (1.0 - rho_lo) * (1.0 - eps);
    const double beta = std::max(k * (1.0 + rho_hi) * B_hi, k * (1.0 - rho_lo) * B_lo + e2 * e2);
    // f32 evaluation on the device: z = fma(-k_f, fl(V), fl(D^2)) is within 3u D^2 + 4u k V of D^2 - k V, and the margin
    // itself carries three roundings: 3u |D| n size >= 3u D^2 goes into the |D| term, 6u into eps', (1 + 8u) on top.
    const double infl = 1.0 + 8.0 * u;
    FusedConst c;
    c.k = (float)k;
    c.e2x2 = f_up((2.0 * e2 + 3.0 * u * n * sz) * infl);
    c.epsk = f_up(k * (std::max(eps1, eps2) + 6.0 * u) * infl);
    c.beta = f_up(beta * infl);
    return c;
}

// The fused kernel covers the default statistics only: influence exactly 1 (filtered == data, stat.rs:42,47),
// a finite non-negative threshold (the certificate squares it) and an eligible geometry.
bool wave_fused_ok(const gams_cuda_wave_plan *p, float threshold, float influence) {
    // thresholds in [2^-16, 2^30]: k = thr^2 n / (n - 1) and every product with it stay normal f32 numbers, as the error
    // model of the certificate assumes (GC fractions are in [0, 1]: magnitudes need no guard here); outside that range
    // Kernel A + Kernel B replay every window exactly
    return p->fused_wt > 0 && influence == 1.0f && threshold >= 1.52587890625e-5f && threshold <= 1073741824.0f &&
           p->ctx->zscore_mode == 0 && p->win.min_len >= p->lag;
}

// Output windows per fused tile, or 0 when the fused path does not apply to this geometry.
int fused_windows_per_tile(int32_t size, int32_t step, uint64_t lag) {
    if (step > 16 || size % step != 0 || size / step > FW_MAX_R) return 0;   // a block's bits come from one 32-bit word
    if (lag < 2 || lag > FW_MAX_LAG) return 0;
    if ((uint64_t)lag * (uint64_t)size >= 65536) return 0;   // D and V = n sum c^2 - (sum c)^2 exact in 32 bits
    const int64_t span_cap = (int64_t)FW_NCH * 16 - 31;      // worst case lead = 31
    if (size > span_cap) return 0;
    const int64_t h = (int64_t)lag + 1;
    const int64_t n_el = std::min<int64_t>((span_cap - size) / step + 1, FW_MAX_EL);
    const int64_t wt = ((n_el - h - 7) / 8) * 8;             // up to 7 alignment elements before the halo
    if (wt < 4 * h || wt < 64) return 0;                     // halo re-read would exceed 25 %
    return (int)wt;
}

// Tiles of one ctg (see the layout comment at the kernel).  seq_off / win_off: the ctg's first byte / window in the
// device arrays; n_c windows.
void fused_tiles_for_ctg(std::vector<FusedTile> &out, uint64_t seq_off, uint64_t win_off, uint64_t n_c, int32_t step,
                         uint64_t lag, int wt) {
    const int64_t h = (int64_t)lag + 1;
    for (uint64_t w = 0; w < n_c; w += (uint64_t)wt) {
        FusedTile t{};
        const int64_t g_first = (int64_t)(win_off + w);            // first output window, global index
        const int64_t need = w ? h : 1;                            // elements wanted before it
        int64_t g0 = g_first - need;
        g0 -= ((g0 % 8) + 8 + 1) % 8;                              // largest g0' <= g0 with g0' = 7 (mod 8)
        t.g0 = g0;
        t.e_out0 = (uint32_t)(g_first - g0);
        t.e_out1 = t.e_out0 + (uint32_t)std::min<uint64_t>((uint64_t)wt, n_c - w);
        t.e_ctg0 = w ? -(1 << 30) : (int32_t)t.e_out0;
        t.byte0 = (int64_t)seq_off + ((int64_t)w - (int64_t)t.e_out0) * (int64_t)step;
        const int64_t base = t.byte0 - (t.byte0 & 31);             // 32-byte aligned (also for negative byte0)
        t.pi_min = base < 0 ? (uint32_t)((-base + 31) / 32) : 0u;  // pieces before the start of the sequence buffer
        out.push_back(t);
    }
}

int32_t launch_gc_wave_fused(gams_cuda_wave_plan *p, cudaStream_t st, const uint8_t *d_seq, float threshold,
                             uint32_t *d_cnt, float *d_gc, int8_t *d_sig, unsigned long long *d_stats, int64_t c0,
                             int64_t c1) {
    if ((reinterpret_cast<uintptr_t>(d_seq) & 31u) != 0) {
        set_error("fused GC-wave: d_seq must be 32-byte aligned");
        return GAMS_CUDA_EINVAL;
    }
    const uint64_t t0 = p->h_ftile_off[c0], t1 = p->h_ftile_off[c1];
    if (t1 == t0) return GAMS_CUDA_OK;
    const FusedConst fc = make_fused_const(p->size, p->lag, threshold);
    const unsigned grid = (unsigned)(t1 - t0);
    const FusedTile *tiles = p->d_ftiles + t0;
    const int r = p->size / p->step;
#define GAMS_FUSED_LAUNCH(WC, L4, RR, S)                                                                                \
    gc_wave_fused_kernel<WC, L4, RR, S><<<grid, GC_THREADS, 0, st>>>(d_seq, tiles, p->d_gc_tab, p->size, p->step, r,      \
                                                                  (int)p->lag, threshold, 1u, fc, d_cnt, d_gc, d_sig, d_stats)
#define GAMS_FUSED_LAUNCH_R(WC, L4)                                                                                     \
    do {                                                                                                               \
        if (r == 10 && p->step == 10) GAMS_FUSED_LAUNCH(WC, L4, 10, true);                                             \
        else if (r == 10) GAMS_FUSED_LAUNCH(WC, L4, 10, false);                                                        \
        else GAMS_FUSED_LAUNCH(WC, L4, 0, false);                                                                      \
    } while (0)
    if (p->lag % 4 == 0) { if (d_cnt) GAMS_FUSED_LAUNCH_R(true, true); else GAMS_FUSED_LAUNCH_R(false, true); }
    else { if (d_cnt) GAMS_FUSED_LAUNCH_R(true, false); else GAMS_FUSED_LAUNCH_R(false, false); }
#undef GAMS_FUSED_LAUNCH_R
#undef GAMS_FUSED_LAUNCH
    ++p->launches;
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

}  // namespace gams
