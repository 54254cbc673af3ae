// gc_common.cuh — device helpers shared by Kernel A (gc_window.cu) and the fused GC-wave kernel (gc_wave_fused.cu)
#pragma once
#include "common.cuh"

namespace gams {

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}

// Top nibble of the result: bit (28 + j) is set iff byte j of w is one of C G c g.
//   t  = (w & 0x5B) ^ 0x43 per byte: zero exactly for {C,G,c,g} and their bit-7 twins  (one LOP3)
//   a  = t + 0x7F          per byte: bit 7 <=> t != 0, no cross-byte carry (t <= 0x5B)
//   g7 = ~(a | w) & 0x80   per byte: bit 7 <=> t == 0 and w < 0x80                      (one LOP3)
//   g7 * 0x204081 moves bits 7,15,23,31 to 28,29,30,31 (all partial products land on distinct bits)
__device__ __forceinline__ uint32_t gc_gather(uint32_t w) {
    const uint32_t t = lop3<0x6A>(w, 0x5B5B5B5Bu, 0x43434343u);
    const uint32_t a = t + 0x7F7F7F7Fu;
    const uint32_t g7 = lop3<0x02>(a, w, 0x80808080u);
    return g7 * 0x00204081u;
}

// bit (4*i + j) is set iff byte j of word i of the 16-byte chunk is one of C G c g
__device__ __forceinline__ uint32_t gc_mask16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    uint32_t m = gc_gather(w0) >> 28;
    m = lop3<0xEA>(gc_gather(w1) >> 24, 0xF0u, m);     // (a & b) | c
    m = lop3<0xEA>(gc_gather(w2) >> 20, 0xF00u, m);
    m = lop3<0xEA>(gc_gather(w3) >> 16, 0xF000u, m);
    return m;
}

// 32 bytes per lane in one instruction (Blackwell LDG.E.256), streamed once: read-only path, no L1 allocation
struct Bytes32 { uint32_t w[8]; };
__device__ __forceinline__ Bytes32 ld_stream32(const void *p) {
    Bytes32 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}

// ent[c]: low 16 bits = GC mask of chunk c, high 16 bits = #GC bytes in chunks [0, c) of the tile.
__device__ __forceinline__ uint32_t prefix_at(const uint32_t *ent, uint32_t x) {
    uint32_t e = ent[x >> 4];
    return (e >> 16) + __popc(e & ((1u << (x & 15u)) - 1u));
}

// opaque 32-bit subtract: keeps the table index a plain u32 (no 64-bit range reasoning in the address math)
__device__ __forceinline__ uint32_t sub_u32(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("sub.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

constexpr int GC_MAX_BOUNDS = 4096;  // ALIGNED path: window boundaries kept per tile (u16 each)


}  // namespace gams
