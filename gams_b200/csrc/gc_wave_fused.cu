// gc_wave_fused.cu — Kernel AB: sequence bytes -> per-window GC AND peak signals in one pass (influence == 1).
//
// Replaces wave::proc_ctg's compute, src/cmd_gams/wave.rs:138-155: gams::sliding (window.rs:78-94), the
// per-window gc_content loop (wave.rs:143-153) and gams::thresholding_algo (stat.rs:16-56).
//
// Why fuse.  With the certified filter (zscore.cu) the z-score step no longer needs the f32 GC values:
// it needs the exact mean and centred sum of squares of each stats window.  For GC data those follow from
// INTEGER window counts c_k (x_k = fl(c_k / size)), which Kernel A already holds in shared memory.  So the
// tile is extended backwards by lag+1 windows, the counts are scanned twice more (sum c, sum c^2: exact, u32),
// and each window is decided from D = n c_i - sum c and V = n sum c^2 - (sum c)^2:
//     signal certainly      if (|D| - e2)^2            >  K_hi (V (1+eps) + B_hi)        and |D| > e2
//     certainly no signal   if (|D| + e2)^2            <  K_lo max(V (1-eps) - B_lo, 0)
// where e2, K_*, B_*, eps are host-computed bounds on everything the reference's f32 evaluation can do to
// these exact quantities (derivation in DESIGN.md §3 "Kernel AB"); otherwise the window is replayed exactly
// (zscore_decide on x_k = gc_tab[c_k], the same f32 values the reference sums).  No separate pass re-reads the
// 1.24 GB of GC values, and no f64 prefix sums are needed.
//
// Bound: HBM (1 B/base read + 5 B/window written, + the lag+1-window halo re-read per tile, ~6 % at 100/10/100).
#include "gc_common.cuh"
#include "zscore_common.cuh"

namespace gams {

template <int K, bool WRITE_CNT>
__global__ void __launch_bounds__(GC_THREADS)
gc_wave_fused_kernel(const uint8_t *__restrict__ seq, const uint64_t *__restrict__ seq_off,
                     const uint64_t *__restrict__ win_off, const TileDesc *__restrict__ tiles,
                     const float *__restrict__ gc_tab, int size, int step, int wt, int lag, float threshold,
                     const FusedConst fc, uint32_t *__restrict__ cnt_out, float *__restrict__ gc_out,
                     int8_t *__restrict__ sig_out, unsigned long long *__restrict__ stats) {
    constexpr int NCH = GC_THREADS * K;
    __shared__ uint32_t ent[NCH + 1];
    __shared__ __align__(16) uint8_t c8[NCH];
    __shared__ uint32_t wsum[GC_THREADS / 32];
    __shared__ uint2 wsum2[GC_THREADS / 32];
    __shared__ uint16_t bnd[GC_MAX_BOUNDS];
    __shared__ uint16_t cw[FUSED_MAX_EL];
    __shared__ __align__(16) uint2 CC[FUSED_MAX_EL + 1];  // exclusive prefixes (sum c, sum c^2); reused as f32 x[] on replay
    __shared__ int8_t sg[FUSED_MAX_EL];
    __shared__ int list[FUSED_LIST_CAP];
    __shared__ int list_n;

    const int tid = threadIdx.x;
    const TileDesc td = tiles[blockIdx.x];
    const uint64_t wbase = win_off[td.item];
    const uint32_t n_c = (uint32_t)(win_off[td.item + 1] - wbase);
    const uint32_t w0 = td.first;
    const uint32_t nw = min((uint32_t)wt, n_c - w0);
    const uint32_t h = (uint32_t)lag + 1u;
    const uint32_t hh = min(h, w0);       // halo windows actually present before the tile
    const uint32_t n_el = nw + hh;        // windows whose counts this tile holds: [w0 - hh, w0 + nw)
    const uint64_t byte0 = seq_off[td.item] + (uint64_t)(w0 - hh) * (uint64_t)step;
    const uint32_t lead = (uint32_t)(byte0 & 31u);
    const uint32_t span = (n_el - 1u) * (uint32_t)step + (uint32_t)size;
    const uint32_t nch = (lead + span + 15u) >> 4;  // <= NCH by construction of wt
    if (tid == 0) list_n = 0;

    // ---- phase 1: classify (see gc_window.cu)
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

    // ---- scan of chunk counts
    uint32_t inc[K / 4];
    uint32_t total = 0;
#pragma unroll
    for (int q = 0; q < K / 4; ++q) {
        const uint32_t c4 = reinterpret_cast<const uint32_t *>(c8)[tid * (K / 4) + q];
        inc[q] = c4 * 0x01010101u + total * 0x01010101u;
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

    // ---- window boundaries, counts, GC output
    const uint32_t r = (uint32_t)size / (uint32_t)step;
    const uint32_t nb = n_el + r;
#pragma unroll 2
    for (uint32_t m = tid; m < nb; m += GC_THREADS) bnd[m] = (uint16_t)prefix_at(ent, lead + m * (uint32_t)step);
    __syncthreads();
    float *__restrict__ out = gc_out + (wbase + w0);
    uint32_t *__restrict__ outc = WRITE_CNT ? cnt_out + (wbase + w0) : nullptr;
#pragma unroll 4
    for (uint32_t e = tid; e < n_el; e += GC_THREADS) {
        const uint32_t cnt = sub_u32(bnd[e + r], bnd[e]);
        cw[e] = (uint16_t)cnt;
        if (e >= hh) {
            out[e - hh] = __ldg(gc_tab + cnt);
            if (WRITE_CNT) outc[e - hh] = cnt;
        }
    }
    __syncthreads();

    // ---- exact integer prefix sums of c and c^2 over the tile's windows (thread-contiguous chunks)
    {
        const uint32_t per = ((n_el + GC_THREADS - 1) / GC_THREADS) | 1u;
        const uint32_t pb = min((uint32_t)tid * per, n_el), pe = min(pb + per, n_el);
        uint32_t a1 = 0, a2 = 0;
        for (uint32_t p = pb; p < pe; ++p) {
            const uint32_t c = cw[p];
            a1 += c;
            a2 += c * c;
        }
        uint32_t s1 = a1, s2 = a2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t1 = __shfl_up_sync(0xFFFFFFFFu, s1, d), t2 = __shfl_up_sync(0xFFFFFFFFu, s2, d);
            if (lane >= d) { s1 += t1; s2 += t2; }
        }
        if (lane == 31) wsum2[warp] = make_uint2(s1, s2);
        __syncthreads();
        uint32_t b1 = s1 - a1, b2 = s2 - a2;
#pragma unroll
        for (int w = 0; w < GC_THREADS / 32; ++w)
            if (w < warp) { b1 += wsum2[w].x; b2 += wsum2[w].y; }
        for (uint32_t p = pb; p < pe; ++p) {
            CC[p] = make_uint2(b1, b2);
            const uint32_t c = cw[p];
            b1 += c;
            b2 += c * c;
        }
        if (pe == n_el && pb < pe) CC[n_el] = make_uint2(b1, b2);
    }
    __syncthreads();

    // ---- certificate per window of the tile
    const uint32_t k_min = (w0 >= (uint32_t)lag) ? 0u : (uint32_t)lag - w0;  // windows before `lag` never signal
    const int e_shift = (int)hh - (int)h;                                   // stats window of tile window k starts at max(k + e_shift, 0)
    unsigned unc = 0;
#pragma unroll 2
    for (uint32_t k = tid; k < nw; k += GC_THREADS) {
        const int ea = max((int)k + e_shift, 0);
        const uint2 A = CC[ea], B = CC[ea + lag];
        const double s1 = (double)(B.x - A.x), s2 = (double)(B.y - A.y), ci = (double)cw[k + hh];
        const double D = fma(fc.n, ci, -s1);
        const double V = fma(fc.n, s2, -s1 * s1);
        const double aD = fabs(D);
        const double lo = aD - fc.e2, hi = aD + fc.e2;
        const bool yes = (lo > 0.0) & (lo * lo > fc.k_hi * fma(V, fc.one_plus_eps, fc.b_hi));
        const bool no = hi * hi < fc.k_lo * fmax(fma(V, fc.one_minus_eps, -fc.b_lo), 0.0);
        int dec = yes ? (D > 0.0 ? 1 : -1) : (no ? 0 : 2);
        if (k < k_min) dec = 0;
        if (dec == 2) {
            ++unc;
            const int slot = atomicAdd(&list_n, 1);
            if (slot < FUSED_LIST_CAP) list[slot] = (int)k;
        }
        sg[k] = (int8_t)dec;
    }
    if (stats) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) unc += __shfl_xor_sync(0xFFFFFFFFu, unc, d);
        if (lane == 0 && unc) atomicAdd(&stats[0], (unsigned long long)unc);
    }
    __syncthreads();

    // ---- exact replay of inconclusive windows (rare): x_k = gc_tab[c_k] are the f32 values the reference sums
    const int n_list = list_n;  // block-uniform
    if (n_list > 0) {
        float *xs = reinterpret_cast<float *>(CC);  // the prefix sums are no longer needed
        for (uint32_t e = tid; e < n_el; e += GC_THREADS) xs[e] = __ldg(gc_tab + cw[e]);
        __syncthreads();
        const float flag = __ull2float_rn((unsigned long long)lag);
        const float flagm1 = __fsub_rn(flag, 1.0f);
        auto ld = [&](int64_t e) { return xs[e]; };
        const int n_todo = (n_list <= FUSED_LIST_CAP) ? n_list : (int)nw;  // overflow: replay the whole tile
        for (int t = tid; t < n_todo; t += GC_THREADS) {
            const uint32_t k = (n_list <= FUSED_LIST_CAP) ? (uint32_t)list[t] : (uint32_t)t;
            if (k < k_min) continue;
            const int ea = max((int)k + e_shift, 0);
            sg[k] = (int8_t)zscore_decide(ld, (int64_t)ea, (uint64_t)lag, flag, flagm1, xs[k + hh], threshold);
        }
        __syncthreads();
    }
    int8_t *__restrict__ so = sig_out + (wbase + w0);
    for (uint32_t k = tid; k < nw; k += GC_THREADS) so[k] = sg[k];
}

// Host-side bounds of the certificate (derivation: DESIGN.md §3 "Kernel AB").  With u = 2^-24, n = lag,
// x_j = fl(c_j/size) in [0,1], x^c_j = c_j/size exactly, mu_c and S_c the exact mean / centred sum of squares of x^c:
//   |m^ - mu_x|           <= E_m  (recursive f32 sum of n terms + one divide: ~ n u, taken as 1.01 (n+2) u)
//   |fl(x_i - m^) - (x^c_i - mu_c)| <= E_m + 4u =: E2      (x_i vs x^c_i, mu_x vs mu_c, the subtraction's rounding)
//   fl(thr * sd^) = thr sqrt(Q*/(n-1)) (1 +- R),  R = (0.6 n + 8) u,  Q* = sum (x_j - m^)^2 in [S_x, S_x + n E_m^2]
//   sqrt(S_x) = sqrt(S_c) +- sqrt(n) u  =>  S_x in [(1-eps) S_c - (1/eps - 1) n u^2, (1+eps) S_c + (1 + 1/eps) n u^2]
// and everything is scaled by n*size (D = n*size*(x^c_i - mu_c)) resp. n*size^2 (V = n*size^2*S_c).
FusedConst make_fused_const(int32_t size, uint64_t lag, float threshold) {
    const double u = 5.9604644775390625e-8;
    const float flag = (float)lag, flagm1 = flag - 1.0f;  // `data.len() as f32`, `len - 1.`
    const double n = (double)lag, nm1 = (double)flagm1, sz = (double)size;
    const double E_m = 1.01 * (n + 2.0) * u;
    const double E2 = E_m + 4.0 * u;
    const double R = (0.6 * n + 8.0) * u;
    const double eps = 2e-6;
    const double b2 = 4.0 * n * u * u;            // 4x the n u^2 needed
    const double t2 = (double)threshold * (double)threshold;
    FusedConst c;
    c.n = n;
    c.e2 = E2 * n * sz * (1.0 + 1e-9);
    c.k_hi = t2 * (1.0 + R) * (1.0 + R) * (1.0 + 3.0 * u) * n / nm1 * (1.0 + 1e-9);
    c.k_lo = t2 * (1.0 - R) * (1.0 - R) * (1.0 - 3.0 * u) * n / nm1 * (1.0 - 1e-9);
    c.one_plus_eps = 1.0 + eps;
    c.one_minus_eps = 1.0 - eps;
    c.b_hi = n * sz * sz * (b2 * (1.0 + 1.0 / eps) + n * E_m * E_m) * (1.0 + 1e-9);
    c.b_lo = n * sz * sz * (b2 * (1.0 / eps - 1.0)) * (1.0 + 1e-9);
    return c;
}

// The fused kernel covers the default statistics only: influence exactly 1 (filtered == data, stat.rs:42,47),
// a finite non-negative threshold (the certificate squares it) and an eligible geometry.
bool wave_fused_ok(const gams_cuda_wave_plan *p, float threshold, float influence) {
    return p->fused_wt > 0 && influence == 1.0f && threshold >= 0.0f && threshold < 1.0e18f && p->ctx->zscore_mode == 0 &&
           p->win.min_len >= p->lag;
}

// Windows per fused tile, or 0 when the fused path does not apply to this geometry.
int fused_windows_per_tile(int32_t size, int32_t step, uint64_t lag, int k) {
    if (k != 4 && k != 8) return 0;
    if (size % step != 0 || lag < 2 || lag > 4000) return 0;
    const int64_t span_cap = (int64_t)GC_THREADS * k * 16 - 31;
    const int64_t h = (int64_t)lag + 1, r = size / step;
    int64_t wt = (span_cap - size) / step + 1 - h;           // tile windows + halo windows must fit the byte span
    wt = std::min<int64_t>(wt, FUSED_MAX_EL - h);            // counts / prefixes / decisions per tile
    wt = std::min<int64_t>(wt, GC_MAX_BOUNDS - h - r);       // boundary table
    if (wt < 4 * h || wt < 64) return 0;                     // halo re-read would exceed 25 %
    if ((uint64_t)size * (uint64_t)size * (uint64_t)(wt + h) >= (1ull << 32)) return 0;  // sum c^2 must fit u32
    if ((double)lag * (double)size * (double)lag * (double)size >= 4e15) return 0;        // n s2, s1^2 exact in f64
    return (int)wt;
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
    const TileDesc *tiles = p->d_ftiles + t0;
#define GAMS_FUSED_LAUNCH(KK, WC)                                                                                       \
    gc_wave_fused_kernel<KK, WC><<<grid, GC_THREADS, 0, st>>>(d_seq, p->d_seq_off, p->win.d_off, tiles, p->d_gc_tab,   \
                                                              p->size, p->step, p->fused_wt, (int)p->lag, threshold,   \
                                                              fc, d_cnt, d_gc, d_sig, d_stats)
    if (p->geom.k == 4) { if (d_cnt) GAMS_FUSED_LAUNCH(4, true); else GAMS_FUSED_LAUNCH(4, false); }
    else { if (d_cnt) GAMS_FUSED_LAUNCH(8, true); else GAMS_FUSED_LAUNCH(8, false); }
#undef GAMS_FUSED_LAUNCH
    ++p->launches;
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

}  // namespace gams
