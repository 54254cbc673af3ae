// gc_window.cu — Kernel A: sequence bytes -> per-window GC counts and f32 fractions.
//
// Replaces the GC loop of wave::proc_ctg (src/cmd_gams/wave.rs:143-153, bio gc_content per
// window) for every window gams::sliding yields (src/libs/window.rs:78-94).
//
// Bound: HBM.  Algorithmic traffic 1 B/base read + 4 B/window written (+4 with counts).
// Design: one CTA per tile of `wt` consecutive windows of one ctg.  The tile's bytes are read
// exactly once, 32 bytes per lane per instruction (LDG.E.256); each 16-byte chunk is classified with SIMD-in-
// register byte tricks into a 16-bit "is C/G" mask; chunk counts are scanned across the CTA
// (byte-packed local scan + warp-shuffle scan + one smem hop); a window's count is then
// prefix(end) - prefix(start) with prefix(x) = chunk_prefix[x>>4] + popc(mask[x>>4] & low(x&15)).
// So each base is touched once however much windows overlap (the reference rescans every
// window: size/step = 10x re-reads at 100/10).
#include "gc_common.cuh"

namespace gams {

// ALIGNED: size % step == 0, so windows share boundaries: prefix(boundary) is evaluated once per boundary
// and a window is the difference of two table entries.
template <int K, bool WRITE_CNT, bool ALIGNED>
__global__ void __launch_bounds__(GC_THREADS)
gc_window_kernel(const uint8_t *__restrict__ seq, const uint64_t *__restrict__ seq_off,
                 const uint64_t *__restrict__ win_off, const TileDesc *__restrict__ tiles,
                 const float *__restrict__ gc_tab, int size, int step, int wt,
                 uint32_t *__restrict__ cnt_out, float *__restrict__ gc_out) {
    constexpr int NCH = GC_THREADS * K;
    __shared__ uint32_t ent[NCH + 1];
    __shared__ __align__(16) uint8_t c8[NCH];
    __shared__ uint32_t wsum[GC_THREADS / 32];
    __shared__ uint16_t bnd[ALIGNED ? GC_MAX_BOUNDS : 1];

    const int tid = threadIdx.x;
    const TileDesc td = tiles[blockIdx.x];
    const uint64_t wbase = win_off[td.item];
    const uint32_t n_c = (uint32_t)(win_off[td.item + 1] - wbase);
    const uint32_t nw = min((uint32_t)wt, n_c - td.first);
    const uint64_t byte0 = seq_off[td.item] + (uint64_t)td.first * (uint64_t)step;  // first byte needed
    const uint32_t lead = (uint32_t)(byte0 & 31u);
    const uint32_t span = (nw - 1u) * (uint32_t)step + (uint32_t)size;
    const uint32_t nch = (lead + span + 15u) >> 4;  // <= NCH by construction of wt

    // ---- phase 1: classify.  Lane t of group j reads bytes [32 (j*256 + t), +32): 1 KB contiguous per warp instruction
    const uint8_t *src = seq + (byte0 - lead);
    Bytes32 v[K / 2];
#pragma unroll
    for (int j = 0; j < K / 2; ++j) {
        const uint32_t pi = j * GC_THREADS + tid;
        if (2u * pi < nch) v[j] = ld_stream32(src + 32u * pi);
        else {
#pragma unroll
            for (int q = 0; q < 8; ++q) v[j].w[q] = 0u;
        }
    }
    uint16_t *ent16 = reinterpret_cast<uint16_t *>(ent);
#pragma unroll
    for (int j = 0; j < K / 2; ++j) {
        const uint32_t pi = j * GC_THREADS + tid;
        const uint32_t m0 = gc_mask16(v[j].w[0], v[j].w[1], v[j].w[2], v[j].w[3]);
        const uint32_t m1 = gc_mask16(v[j].w[4], v[j].w[5], v[j].w[6], v[j].w[7]);
        ent16[4 * pi] = (uint16_t)m0;
        ent16[4 * pi + 2] = (uint16_t)m1;
        reinterpret_cast<uint16_t *>(c8)[pi] = (uint16_t)(__popc(m0) | (__popc(m1) << 8));
    }
    __syncthreads();

    // ---- scan: thread t owns chunks [t*K, t*K+K); byte-packed inclusive scan of their counts
    uint32_t inc[K / 4];
    uint32_t total = 0;
#pragma unroll
    for (int q = 0; q < K / 4; ++q) {
        const uint32_t c4 = reinterpret_cast<const uint32_t *>(c8)[tid * (K / 4) + q];
        inc[q] = c4 * 0x01010101u + total * 0x01010101u;  // byte i = total + sum_{k<=i} c_k  (<= 16*K <= 128)
        total = inc[q] >> 24;
    }
    uint32_t sc = total;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, sc, d);
        if (lane >= d) sc += n;
    }
    if (lane == 31) wsum[warp] = sc;
    __syncthreads();
    uint32_t excl = sc - total;
#pragma unroll
    for (int w = 0; w < GC_THREADS / 32; ++w)
        if (w < warp) excl += wsum[w];
#pragma unroll
    for (int q = 0; q < K / 4; ++q) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t before = (i == 0) ? (q == 0 ? 0u : (inc[q - 1] >> 24)) : ((inc[q] >> (8 * (i - 1))) & 0xFFu);
            ent16[2 * (tid * K + q * 4 + i) + 1] = (uint16_t)(excl + before);
        }
    }
    if (tid == GC_THREADS - 1) ent[NCH] = (excl + total) << 16;
    __syncthreads();

    // ---- phase 2: windows
    float *__restrict__ out = gc_out + (wbase + td.first);
    uint32_t *__restrict__ outc = WRITE_CNT ? cnt_out + (wbase + td.first) : nullptr;
    if (ALIGNED) {
        const uint32_t r = (uint32_t)size / (uint32_t)step;
        const uint32_t nb = nw + r;  // boundaries lead + m*step, m = 0 .. nw-1+r
#pragma unroll 2
        for (uint32_t m = tid; m < nb; m += GC_THREADS) bnd[m] = (uint16_t)prefix_at(ent, lead + m * (uint32_t)step);
        __syncthreads();
        const uint16_t *b0 = bnd + tid, *b1 = bnd + tid + r;
#pragma unroll 4
        for (uint32_t w = tid; w < nw; w += GC_THREADS, b0 += GC_THREADS, b1 += GC_THREADS) {
            const uint32_t cnt = sub_u32(*b1, *b0);
            out[w] = __ldg(gc_tab + cnt);
            if (WRITE_CNT) outc[w] = cnt;
        }
    } else {
#pragma unroll 2
        for (uint32_t w = tid; w < nw; w += GC_THREADS) {
            const uint32_t x0 = lead + w * (uint32_t)step;
            const uint32_t cnt = sub_u32(prefix_at(ent, x0 + (uint32_t)size), prefix_at(ent, x0));
            out[w] = __ldg(gc_tab + cnt);
            if (WRITE_CNT) outc[w] = cnt;
        }
    }
}

// Any window size (used when one window does not fit a tile): one warp per window, byte scan.
__global__ void __launch_bounds__(256)
gc_window_generic_kernel(const uint8_t *__restrict__ seq, const uint64_t *__restrict__ seq_off,
                         const uint64_t *__restrict__ win_off, int64_t n_ctg, uint64_t g_begin, uint64_t g_end,
                         int size, int step, uint32_t *__restrict__ cnt_out, float *__restrict__ gc_out) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t g = g_begin + (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); g < g_end; g += warps) {
        int64_t lo = 0, hi = n_ctg;  // last ctg with win_off[c] <= g
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (win_off[mid] <= g) lo = mid; else hi = mid;
        }
        const uint8_t *s = seq + seq_off[lo] + (g - win_off[lo]) * (uint64_t)step;
        uint32_t c = 0;
        for (int i = lane; i < size; i += 32) {
            const uint32_t b = s[i];
            c += ((b & 0xDBu) == 0x43u);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
        if (lane == 0) {
            gc_out[g] = __fdiv_rn(__uint2float_rn(c), __int2float_rn(size));
            if (cnt_out) cnt_out[g] = c;
        }
    }
}

GcGeom plan_gc_geometry(int32_t size, int32_t step) {
    GcGeom g;
    for (int k : {4, 8}) {
        const int64_t span_cap = (int64_t)GC_THREADS * k * 16 - 31;  // worst case lead = 31
        if (size <= span_cap / 2 || (k == 8 && size <= span_cap)) {
            g.k = k;
            int64_t wt = (span_cap - size) / step + 1;
            if (wt > (1 << 20)) wt = 1 << 20;
            g.wt = (int)wt;
            return g;
        }
    }
    g.k = 0;
    g.wt = 0;
    return g;
}

template <int K, bool ALIGNED>
static void launch_tiled(const gams_cuda_wave_plan *p, cudaStream_t st, const uint8_t *d_seq, uint32_t *d_cnt,
                         float *d_gc, int64_t c0, int64_t c1) {
    const uint64_t t0 = p->h_gtile_off[c0], t1 = p->h_gtile_off[c1];
    if (t1 == t0) return;
    const unsigned grid = (unsigned)(t1 - t0);
    if (d_cnt)
        gc_window_kernel<K, true, ALIGNED><<<grid, GC_THREADS, 0, st>>>(d_seq, p->d_seq_off, p->win.d_off, p->d_gtiles + t0,
                                                                        p->d_gc_tab, p->size, p->step, p->geom.wt, d_cnt, d_gc);
    else
        gc_window_kernel<K, false, ALIGNED><<<grid, GC_THREADS, 0, st>>>(d_seq, p->d_seq_off, p->win.d_off, p->d_gtiles + t0,
                                                                         p->d_gc_tab, p->size, p->step, p->geom.wt, nullptr, d_gc);
}

int32_t launch_gc_windows(const gams_cuda_wave_plan *p, cudaStream_t st, const uint8_t *d_seq, uint32_t *d_cnt,
                          float *d_gc, uint64_t *launches, int64_t c0, int64_t c1) {
    const uint64_t g0 = p->h_win_off[c0], g1 = p->h_win_off[c1];
    if (g1 == g0) return GAMS_CUDA_OK;
    if ((reinterpret_cast<uintptr_t>(d_seq) & 31u) != 0) {
        set_error("gc windows: d_seq must be 32-byte aligned");
        return GAMS_CUDA_EINVAL;
    }
    const bool aligned = p->size % p->step == 0 && (int64_t)p->geom.wt + p->size / p->step <= GC_MAX_BOUNDS;
    if (p->geom.k == 4) aligned ? launch_tiled<4, true>(p, st, d_seq, d_cnt, d_gc, c0, c1) : launch_tiled<4, false>(p, st, d_seq, d_cnt, d_gc, c0, c1);
    else if (p->geom.k == 8) aligned ? launch_tiled<8, true>(p, st, d_seq, d_cnt, d_gc, c0, c1) : launch_tiled<8, false>(p, st, d_seq, d_cnt, d_gc, c0, c1);
    else {
        uint64_t blocks = (g1 - g0 + 7) / 8;
        if (blocks > 148ull * 64) blocks = 148ull * 64;
        gc_window_generic_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_seq, p->d_seq_off, p->win.d_off, p->n_ctg, g0, g1,
                                                                   p->size, p->step, d_cnt, d_gc);
    }
    if (launches) ++*launches;
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

}  // namespace gams
