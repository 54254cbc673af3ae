// peaks.cu — what `gams wave` prints and what `gams peak` stores, computed on the device.
//
// (1) Peak merge: wave::proc_ctg's emit step, src/cmd_gams/wave.rs:170-212 with merge_ints (wave.rs:217-252).
//     The reference builds a graph over the signal windows of one sign (edge = the two windows intersect and
//     size/intersect >= coverage, an O(p^2) loop with string keys) and prints every connected component once, at
//     its first member.  All windows of a ctg have one size and lie on a line, so when every intersecting pair
//     passes the coverage test (always for coverage <= 1: size/intersect >= 1) a component is a maximal run of
//     same-sign signal windows whose consecutive members are at most (size-1)/step windows apart.  The kernels
//     find the run starts (no same-sign signal among the preceding (size-1)/step windows), order them by a
//     tile count + scan, and walk each run to its last member.  Output: one record per printed line, in the
//     reference's line order within a ctg (window order of the first member).
// (2) Peak linking: peak::execute, src/cmd_gams/peak.rs:66-160: GC content of each peak range
//     (cache_gc_content, utils.rs:141-162: count/len rounded to 4 decimals) and the distances / GC differences to
//     the previous and next peak of the ctg.
//
// Both are tiny next to Kernel AB (0.3 GB of signals read once); they exist so that the 1.55 GB of per-window
// results need not cross PCIe when only the peaks are wanted.
#include "common.cuh"

namespace gams {

constexpr int PK_THREADS = 256;
constexpr int PK_PER = 8;
constexpr int PK_TILE = PK_THREADS * PK_PER;

// last ctg whose first window is <= g
__device__ __forceinline__ int64_t ctg_of_window(const uint64_t *__restrict__ win_off, int64_t n_ctg, uint64_t g) {
    int64_t lo = 0, hi = n_ctg;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (win_off[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// window g (signal s != 0) starts a printed line: no window of the same sign among the `gap` windows before it in
// its ctg
__device__ __forceinline__ bool run_starts_at(const int8_t *__restrict__ sig, uint64_t g, uint64_t ctg_first, int8_t s,
                                              uint32_t gap) {
    const uint64_t lo = (g - ctg_first > gap) ? g - gap : ctg_first;
    for (uint64_t j = g; j-- > lo;)
        if (sig[j] == s) return false;
    return true;
}

// MODE 0: find the run starts (one flag byte per thread: bit j <=> window t0 + j starts a line) and count them per tile;
// MODE 1: write the records (tile_base = exclusive scan of the counts) from the flag bytes.  A thread's eight signals are one
// 8-byte load where the address allows it, and 92 % of the threads see eight zeros and are done.
template <int MODE>
__global__ void __launch_bounds__(PK_THREADS)
peak_runs_kernel(const int8_t *__restrict__ sig, const float *__restrict__ gc, const uint64_t *__restrict__ win_off,
                 int64_t n_ctg, uint64_t g0, uint64_t g1, uint32_t gap, uint64_t *__restrict__ tile_cnt,
                 const uint64_t *__restrict__ tile_base, uint8_t *__restrict__ flag_bytes, uint64_t cap, uint32_t *__restrict__ pk_ctg,
                 uint32_t *__restrict__ pk_first, uint32_t *__restrict__ pk_last, float *__restrict__ pk_gc,
                 int8_t *__restrict__ pk_sig) {
    __shared__ uint32_t wsum[PK_THREADS / 32];
    __shared__ int64_t tile_ctg;      // the ctg of the tile's first window: a thread's ctg is that one or one of the next few
    if (threadIdx.x == 0) tile_ctg = ctg_of_window(win_off, n_ctg, min(g0 + (uint64_t)blockIdx.x * PK_TILE, g1 - 1));
    __syncthreads();
    const uint64_t t0 = g0 + (uint64_t)blockIdx.x * PK_TILE + (uint64_t)threadIdx.x * PK_PER;
    const size_t fidx = (size_t)blockIdx.x * PK_THREADS + threadIdx.x;
    uint32_t flags = 0;   // bit j: window t0 + j starts a line
    if (MODE == 0) {
        uint64_t w8 = 0;    // byte j: sig[t0 + j] (0 past g1)
        if (t0 + PK_PER <= g1 && (reinterpret_cast<uintptr_t>(sig + t0) & 7u) == 0) w8 = *reinterpret_cast<const uint64_t *>(sig + t0);
        else
            for (int j = 0; j < PK_PER; ++j)
                if (t0 + j < g1) w8 |= (uint64_t)(uint8_t)sig[t0 + j] << (8 * j);
        if (w8) {
            int64_t c = tile_ctg;
            for (int j = 0; j < PK_PER; ++j) {
                const int8_t s = (int8_t)(w8 >> (8 * j));
                if (s == 0) continue;
                const uint64_t g = t0 + j;
                while (g >= win_off[c + 1]) ++c;          // (g < g1 <= win_off[n_ctg]: c stays below n_ctg)
                if (run_starts_at(sig, g, win_off[c], s, gap)) flags |= 1u << j;
            }
        }
        flag_bytes[fidx] = (uint8_t)flags;
    } else {
        flags = flag_bytes[fidx];
    }
    const uint32_t mine = __popc(flags);
    uint32_t inc = mine;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t before = inc - mine;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    if (MODE == 0) {
        if (threadIdx.x == PK_THREADS - 1) tile_cnt[blockIdx.x] = before + mine;
        return;
    }
    if (!flags) return;
    uint64_t pos = tile_base[blockIdx.x] + before;
    int64_t c = tile_ctg;
#pragma unroll 1
    for (int j = 0; j < PK_PER; ++j) {
        if (!(flags & (1u << j))) continue;
        const uint64_t g = t0 + j;
        const int8_t s = sig[g];
        while (g >= win_off[c + 1]) ++c;
        const uint64_t c_first = win_off[c], c_end = win_off[c + 1];
        uint64_t last = g;   // walk the run: the next member is the nearest same-sign signal within `gap` windows
        for (;;) {
            uint64_t nx = 0;
            const uint64_t hi = min(last + gap, c_end - 1);
            for (uint64_t k = last + 1; k <= hi; ++k)
                if (sig[k] == s) { nx = k; break; }
            if (!nx) break;
            last = nx;
        }
        if (pos < cap) {
            pk_ctg[pos] = (uint32_t)c;
            pk_first[pos] = (uint32_t)(g - c_first);
            pk_last[pos] = (uint32_t)(last - c_first);
            pk_gc[pos] = gc[g];
            pk_sig[pos] = s;
        }
        ++pos;
    }
}

// exclusive scan of n values by one CTA; total in out[n]
__global__ void __launch_bounds__(1024)
scan_u64_kernel(const uint64_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ out) {
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const uint64_t v = i < n ? in[i] : 0;
        uint64_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if ((threadIdx.x & 31) >= d) inc += t;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
        __syncthreads();
        uint64_t before = carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += wsum[w];
        if (i < n) out[i] = before + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

// scratch of launch_peak_runs for n_win windows: tile counts, tile bases, one flag byte per thread
size_t peak_tmp_bytes(uint64_t n_win) {
    const uint64_t tiles = (n_win + PK_TILE - 1) / PK_TILE;
    return sizeof(uint64_t) * 2 * (tiles + 2) + (size_t)tiles * PK_THREADS + 64;
}

// Every intersecting pair of windows passes merge_ints' coverage test (wave.rs:229-231, f32): then components are runs.
bool peak_merge_is_runs(int32_t size, int32_t step, float coverage) {
    const float fs = (float)size;
    for (int64_t d = 1; d * step < size; ++d) {
        const float isz = (float)(int32_t)(size - d * step);
        if (!(fs / isz >= coverage)) return false;
    }
    return true;
}

// d_tmp: at least peak_tmp_bytes(windows) bytes.  d_n_peaks receives the number of records (may exceed cap: nothing past cap is written).
int32_t launch_peak_runs(cudaStream_t st, const int8_t *d_sig, const float *d_gc, const uint64_t *d_win_off, int64_t n_ctg,
                         uint64_t g0, uint64_t g1, int32_t size, int32_t step, uint64_t *d_tmp, uint64_t cap,
                         uint32_t *pk_ctg, uint32_t *pk_first, uint32_t *pk_last, float *pk_gc, int8_t *pk_sig,
                         uint64_t *d_n_peaks) {
    const uint64_t tiles = (g1 - g0 + PK_TILE - 1) / PK_TILE;
    if (tiles == 0) {
        CUDA_TRY(cudaMemsetAsync(d_n_peaks, 0, sizeof(uint64_t), st));
        return GAMS_CUDA_OK;
    }
    if (tiles >= (1ull << 31)) {
        set_error("wave_peaks: too many windows for one launch");
        return GAMS_CUDA_EINVAL;
    }
    const uint32_t gap = (uint32_t)((size - 1) / step);
    uint64_t *cnt = d_tmp, *base = d_tmp + (tiles + 1);
    uint8_t *flag_bytes = reinterpret_cast<uint8_t *>(d_tmp + 2 * (tiles + 2));
    peak_runs_kernel<0><<<(unsigned)tiles, PK_THREADS, 0, st>>>(d_sig, d_gc, d_win_off, n_ctg, g0, g1, gap, cnt, nullptr, flag_bytes, 0,
                                                                nullptr, nullptr, nullptr, nullptr, nullptr);
    scan_u64_kernel<<<1, 1024, 0, st>>>(cnt, tiles, base);
    peak_runs_kernel<1><<<(unsigned)tiles, PK_THREADS, 0, st>>>(d_sig, d_gc, d_win_off, n_ctg, g0, g1, gap, nullptr, base, flag_bytes, cap,
                                                                pk_ctg, pk_first, pk_last, pk_gc, pk_sig);
    CUDA_TRY(cudaMemcpyAsync(d_n_peaks, base + tiles, sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

// ---- peak linking (peak.rs:66-160) -------------------------------------------------------------------
__device__ __forceinline__ float round4_rn(float x) {  // utils.rs:135-138 with decimals = 4
    return __fdiv_rn(roundf(__fmul_rn(x, 10000.0f)), 10000.0f);
}

// peaks sorted by (ctg, start); cnt[i] = GC bases of peak i (from gc_ranges_kernel)
__global__ void __launch_bounds__(256)
peak_link_kernel(int64_t n_pk, const uint32_t *__restrict__ p_ctg, const int32_t *__restrict__ p_start,
                 const int32_t *__restrict__ p_end, const int8_t *__restrict__ p_sig, const uint32_t *__restrict__ cnt,
                 const int32_t *__restrict__ chr_start, const int32_t *__restrict__ chr_end, float *__restrict__ gc,
                 int8_t *__restrict__ l_sig, int32_t *__restrict__ l_len, float *__restrict__ l_amp,
                 int8_t *__restrict__ r_sig, int32_t *__restrict__ r_len, float *__restrict__ r_amp) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pk; i += stride) {
        auto gc_of = [&](int64_t k) {   // gc_content = count as f32 / len as f32 (bio), then round(.., 4)
            return round4_rn(__fdiv_rn(__uint2float_rn(cnt[k]), __int2float_rn(p_end[k] - p_start[k] + 1)));
        };
        const uint32_t c = p_ctg[i];
        const float g = gc_of(i);
        gc[i] = g;
        const bool has_prev = i > 0 && p_ctg[i - 1] == c, has_next = i + 1 < n_pk && p_ctg[i + 1] == c;
        // left: prev_* start as the first peak's own signal / gc and the ctg's chr_start (peak.rs:112-114)
        l_sig[i] = has_prev ? p_sig[i - 1] : p_sig[i];
        l_len[i] = p_start[i] - (has_prev ? p_end[i - 1] : chr_start[c]) + 1;
        l_amp[i] = fabsf(__fsub_rn(g, has_prev ? gc_of(i - 1) : g));
        // right: next_* start as the last peak's own and chr_end (peak.rs:137-139)
        r_sig[i] = has_next ? p_sig[i + 1] : p_sig[i];
        r_len[i] = (has_next ? p_start[i + 1] : chr_end[c]) - p_end[i] + 1;
        r_amp[i] = fabsf(__fsub_rn(g, has_next ? gc_of(i + 1) : g));
    }
}

int32_t launch_peak_link(cudaStream_t st, int64_t n_pk, const uint32_t *p_ctg, const int32_t *p_start, const int32_t *p_end,
                         const int8_t *p_sig, const uint32_t *cnt, const int32_t *chr_start, const int32_t *chr_end, float *gc,
                         int8_t *l_sig, int32_t *l_len, float *l_amp, int8_t *r_sig, int32_t *r_len, float *r_amp) {
    if (n_pk == 0) return GAMS_CUDA_OK;
    int64_t blocks = (n_pk + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    peak_link_kernel<<<(unsigned)blocks, 256, 0, st>>>(n_pk, p_ctg, p_start, p_end, p_sig, cnt, chr_start, chr_end, gc, l_sig, l_len,
                                                       l_amp, r_sig, r_len, r_amp);
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

}  // namespace gams

// ---- anno coverage (src/cmd_gams/anno.rs:128-139) -------------------------------------------------------
// prop = |set(chr) ∩ ctg ∩ range| / |range| as f32: the set of a chromosome is a sorted list of disjoint inclusive runs;
// with cum[k] = positions covered by the runs before k, the covered positions <= x are one binary search away.
namespace gams {

// positions <= x covered by runs [rs[k], re[k]], k = 0 .. n-1 (sorted, disjoint); cum[k] = sum of the lengths before k
__device__ __forceinline__ int64_t covered_upto(const int32_t *__restrict__ rs, const int32_t *__restrict__ re,
                                                const int64_t *__restrict__ cum, uint32_t n, int64_t x) {
    uint32_t lo = 0, len = n;   // first run with start > x
    while (len > 0) {
        const uint32_t half = len >> 1;
        const bool right = (int64_t)__ldg(rs + lo + half) <= x;
        lo = right ? lo + half + 1 : lo;
        len = right ? len - half - 1 : half;
    }
    if (lo == 0) return 0;
    const uint32_t k = lo - 1;  // last run starting at or before x
    const int64_t e = min((int64_t)__ldg(re + k), x);
    return __ldg(cum + k) + (e - (int64_t)__ldg(rs + k) + 1);
}

__global__ void __launch_bounds__(256)
anno_coverage_kernel(const int32_t *__restrict__ run_start, const int32_t *__restrict__ run_end,
                     const int64_t *__restrict__ run_cum, const uint64_t *__restrict__ run_off, int64_t n_chr, int64_t n_q,
                     const int32_t *__restrict__ q_chr, const int32_t *__restrict__ q_ctg_start,
                     const int32_t *__restrict__ q_ctg_end, const int32_t *__restrict__ q_start,
                     const int32_t *__restrict__ q_end, float *__restrict__ prop) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_q; q += stride) {
        const int32_t c = q_chr[q];
        const int64_t s = q_start[q], e = q_end[q];
        float p = 0.0f;
        if (c >= 0 && c < n_chr && e >= s) {
            const int64_t a = max(s, (int64_t)q_ctg_start[q]), b = min(e, (int64_t)q_ctg_end[q]);
            int64_t cov = 0;
            if (b >= a) {
                const uint64_t o = run_off[c];
                const uint32_t n = (uint32_t)(run_off[c + 1] - o);
                cov = covered_upto(run_start + o, run_end + o, run_cum + o, n, b) -
                      covered_upto(run_start + o, run_end + o, run_cum + o, n, a - 1);
            }
            p = __fdiv_rn(__ll2float_rn(cov), __ll2float_rn(e - s + 1));   // `cardinality() as f32 / cardinality() as f32`
        }
        prop[q] = p;
    }
}

int32_t launch_anno_coverage(cudaStream_t st, const int32_t *run_start, const int32_t *run_end, const int64_t *run_cum,
                             const uint64_t *run_off, int64_t n_chr, int64_t n_q, const int32_t *q_chr, const int32_t *q_ctg_start,
                             const int32_t *q_ctg_end, const int32_t *q_start, const int32_t *q_end, float *prop) {
    if (n_q == 0) return GAMS_CUDA_OK;
    int64_t blocks = (n_q + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    anno_coverage_kernel<<<(unsigned)blocks, 256, 0, st>>>(run_start, run_end, run_cum, run_off, n_chr, n_q, q_chr, q_ctg_start,
                                                           q_ctg_end, q_start, q_end, prop);
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

}  // namespace gams
