// count_windows.cu — rg overlap counts along the sliding windows of a ctg set ("counting of rg intervals against windows").
//
// What the reference computes when it walks gams::sliding (src/libs/window.rs:78-94) and asks gams::count_rg
// (src/libs/utils.rs:24-36 -> rust_lapper::Lapper::count) for every window: for window k of ctg c,
//     ws = chr_start[c] + k step,  we = ws + size - 1,
//     out[win_off[c] + k] = lower_bound(starts_c, we) - lower_bound(stops_c, ws + 1)          (#elements < key)
// with gams' half-open quirk (the inclusive end is passed as Lapper's exclusive stop).
//
// The queries are an arithmetic progression, so nothing is searched per window.  A tile is CW_TILE consecutive windows
// of one ctg; its keys we_k = E0 + j step and ws_k + 1 = S0 + j step (j = window in tile).  For a sorted array a[] and
// keys K0 + j step:   #{a < K0 + j step} = #{a < K0} + #{a in tile range with floor((a - K0) / step) < j}
// i.e. one histogram over step-wide bins and its exclusive prefix sum.  cw_desc_kernel finds #{a < K0} of every tile
// (binary searches, all tiles in parallel) and packs what a tile needs into one 48-byte descriptor; cw_tile_kernel streams each tile's slice of starts[] and
// stops[] once (coalesced), bins it with shared-memory atomics (+1 per start, -1 per stop: one signed histogram),
// scans, and writes 4 bytes per window.
//
// Bound: HBM.  Algorithmic bytes: 8 B per interval read + 4 B per window written; no per-window query arrays exist.
#include "common.cuh"
#include <memory>

namespace gams {

constexpr int CW_THREADS = 256;
constexpr int CW_PER = 16;                         // windows per thread in the scan
constexpr int CW_TILE = CW_THREADS * CW_PER;       // windows per tile

// Tiles are cut on the OUTPUT grid: tile j of ctg c holds the windows whose global index lies in
// [A + j CW_TILE, A + (j+1) CW_TILE), A = win_off[c] rounded down to a multiple of 4, so that every thread's 16 results
// are 16-byte aligned in out[] and leave as four 128-bit stores.  The up to three slots before the ctg's window 0 are
// phantom windows (k < 0): computed like any other, never stored.
struct CwTile {
    uint32_t ctg;
    uint32_t j;
};

__device__ __forceinline__ uint32_t cw_lower_bound(const uint32_t *__restrict__ a, uint32_t n, uint64_t key) {
    if (key > 0xFFFFFFFFull) return n;
    const uint32_t k = (uint32_t)key;
    uint32_t lo = 0, len = n;
    while (len > 0) {
        const uint32_t half = len >> 1;
        const bool right = __ldg(a + lo + half) < k;
        lo = right ? lo + half + 1 : lo;
        len = right ? len - half - 1 : half;
    }
    return lo;
}

// everything a tile needs, assembled once per plan launch by cw_desc_kernel so that the tile kernel starts its
// index loads after ONE dependent global load
struct alignas(16) CwDesc {
    uint64_t o;                // first interval of the ctg's segment
    uint64_t wbase;            // first output window of the tile
    uint32_t f0, f1, g0, g1;   // starts[f0, f1) lie in [E0, E0 + CW_TILE step), stops[g0, g1) in [S0, S0 + CW_TILE step)
    uint32_t E0, S0;           // key of the tile's first window: we_0 and ws_0 + 1
    uint32_t lo, hi;           // windows [lo, hi) of the tile exist (lo > 0 only in a ctg's first tile, see CwTile)
};

// one thread per tile: #{starts_c < E0}, #{stops_c < S0} of this tile and of the next one (binary searches, all tiles in parallel)
__global__ void __launch_bounds__(256)
cw_desc_kernel(const uint32_t *__restrict__ starts, const uint32_t *__restrict__ stops, const uint64_t *__restrict__ iv_off,
               const int32_t *__restrict__ chr_start, const uint64_t *__restrict__ win_off, const CwTile *__restrict__ tiles,
               uint64_t n_tiles, int32_t size, int32_t step, CwDesc *__restrict__ desc) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_tiles; t += stride) {
        const CwTile td = tiles[t];
        CwDesc d;
        d.o = iv_off[td.ctg];
        const uint32_t n = (uint32_t)(iv_off[td.ctg + 1] - d.o);
        const uint64_t wo = win_off[td.ctg], a = wo & 3u, n_c = win_off[td.ctg + 1] - wo;
        const int64_t k0 = (int64_t)((uint64_t)td.j * CW_TILE) - (int64_t)a;                  // window (in the ctg) of tile slot 0
        const int64_t w0 = (int64_t)chr_start[td.ctg] + k0 * step;                            // its start coordinate
        const int64_t w1 = w0 + (int64_t)CW_TILE * step;
        const int64_t kE = w0 + size - 1, kS = w0 + 1;
        d.f0 = kE <= 0 ? 0u : cw_lower_bound(starts + d.o, n, (uint64_t)kE);
        d.g0 = kS <= 0 ? 0u : cw_lower_bound(stops + d.o, n, (uint64_t)kS);
        d.f1 = w1 + size - 1 <= 0 ? 0u : cw_lower_bound(starts + d.o, n, (uint64_t)(w1 + size - 1));
        d.g1 = w1 + 1 <= 0 ? 0u : cw_lower_bound(stops + d.o, n, (uint64_t)(w1 + 1));
        d.E0 = (uint32_t)kE;   // elements f0 .. f1 lie in [E0, E0 + CW_TILE step): their distance to E0 fits 32 bits (mod 2^32 for kE < 0)
        d.S0 = (uint32_t)kS;
        d.wbase = wo - a + (uint64_t)td.j * CW_TILE;                                           // multiple of 4
        d.lo = td.j == 0 ? (uint32_t)a : 0u;
        d.hi = (uint32_t)min((int64_t)CW_TILE, (int64_t)n_c - k0);
        desc[t] = d;
    }
}

// bin b of the tile lives at word b + 4 (b >> 4): each thread's run of 16 bins starts 20 words after its neighbour's,
// which makes the 128-bit accesses of the scan conflict-free (16-word pitch would be 4-way conflicted)
constexpr int CW_WORDS = CW_THREADS * (CW_PER + 4);
__device__ __forceinline__ uint32_t cw_slot(uint32_t bin) { return bin + ((bin >> 4) << 2); }
// floor(d / step) for d < CW_TILE * step: one multiply-high with magic = ceil(2^32 / step) when step < 1024
// (d * (magic * step - 2^32) < CW_TILE * step^2 < 2^32 keeps it exact), else a division (magic == 0)
__device__ __forceinline__ uint32_t cw_bin(uint32_t d, uint32_t step, uint32_t magic) {
    const uint32_t b = magic ? __umulhi(d, magic) : d / step;
    return min(b, (uint32_t)CW_TILE - 1u);
}
__host__ inline uint32_t cw_magic(int32_t step) {
    if (step < 2 || step >= 1024) return 0u;    // step 1: 2^32 does not fit; the division by 1 is free anyway
    return (uint32_t)(((1ull << 32) + (uint64_t)step - 1) / (uint64_t)step);
}

#ifndef CW_MIN_CTAS
#define CW_MIN_CTAS 8
#endif
#ifndef CW_PRE
#define CW_PRE 6
#endif
__global__ void __launch_bounds__(CW_THREADS, CW_MIN_CTAS)
cw_tile_kernel(const uint32_t *__restrict__ starts, const uint32_t *__restrict__ stops, const CwDesc *__restrict__ desc,
               int32_t step, uint32_t magic, uint32_t *__restrict__ out) {
    // ONE signed histogram: +1 for a start in bin floor((a - E0) / step), -1 for a stop in bin floor((a - S0) / step);
    // its exclusive prefix, plus f0 - g0, is lower_bound(starts, we_j) - lower_bound(stops, ws_j + 1)
    __shared__ __align__(16) int32_t h[CW_WORDS];
    __shared__ int32_t wsum[CW_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint4 *dp = reinterpret_cast<const uint4 *>(desc + blockIdx.x);
    const uint4 d0 = __ldg(dp), d1 = __ldg(dp + 1), d2 = __ldg(dp + 2);
    const uint64_t o = (uint64_t)d0.x | ((uint64_t)d0.y << 32), wbase = (uint64_t)d0.z | ((uint64_t)d0.w << 32);
    const uint32_t f0 = d1.x, f1 = d1.y, g0 = d1.z, g1 = d1.w, E0 = d2.x, S0 = d2.y, w_lo = d2.z, w_hi = d2.w, ustep = (uint32_t)step;
    // the tile's slices of the index: loads in flight before the histogram is cleared
    constexpr int PRE = CW_PRE;
    uint32_t pa[PRE], pb[PRE];
#pragma unroll
    for (int k = 0; k < PRE; ++k) {
        const uint32_t i = f0 + k * CW_THREADS + tid, j = g0 + k * CW_THREADS + tid;
        pa[k] = i < f1 ? __ldg(starts + o + i) : 0xFFFFFFFFu;
        pb[k] = j < g1 ? __ldg(stops + o + j) : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int k = 0; k < (CW_PER + 4) / 4; ++k) reinterpret_cast<int4 *>(h)[k * CW_THREADS + tid] = make_int4(0, 0, 0, 0);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PRE; ++k) {
        if (f0 + k * CW_THREADS + tid < f1) atomicAdd(&h[cw_slot(cw_bin(pa[k] - E0, ustep, magic))], 1);
        if (g0 + k * CW_THREADS + tid < g1) atomicAdd(&h[cw_slot(cw_bin(pb[k] - S0, ustep, magic))], -1);
    }
    for (uint32_t i = f0 + PRE * CW_THREADS + tid; i < f1; i += CW_THREADS)
        atomicAdd(&h[cw_slot(cw_bin(__ldg(starts + o + i) - E0, ustep, magic))], 1);
    for (uint32_t i = g0 + PRE * CW_THREADS + tid; i < g1; i += CW_THREADS)
        atomicAdd(&h[cw_slot(cw_bin(__ldg(stops + o + i) - S0, ustep, magic))], -1);
    __syncthreads();
    // exclusive prefix over the tile's bins: thread t owns bins [16 t, 16 t + 16)
    int32_t v[CW_PER];
    const int4 *mine = reinterpret_cast<const int4 *>(h + (CW_PER + 4) * tid);
#pragma unroll
    for (int k = 0; k < CW_PER / 4; ++k) {
        const int4 a = mine[k];
        v[4 * k] = a.x; v[4 * k + 1] = a.y; v[4 * k + 2] = a.z; v[4 * k + 3] = a.w;
    }
    int32_t tot = 0;
#pragma unroll
    for (int k = 0; k < CW_PER; ++k) tot += v[k];
    int32_t inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int32_t a = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += a;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int32_t acc = (int32_t)(f0 - g0) + inc - tot;
    for (int w = 0; w < warp; ++w) acc += wsum[w];
    int4 *mine_w = reinterpret_cast<int4 *>(h + (CW_PER + 4) * tid);
#pragma unroll
    for (int k = 0; k < CW_PER / 4; ++k) {
        int4 r;
        r.x = acc; acc += v[4 * k];
        r.y = acc; acc += v[4 * k + 1];
        r.z = acc; acc += v[4 * k + 2];
        r.w = acc; acc += v[4 * k + 3];
        mine_w[k] = r;
    }
    __syncthreads();
    // out: lane-contiguous 128-bit stores (wbase is a multiple of 4: full sectors), four windows per lane and round
    uint32_t *__restrict__ ob = out + wbase;
#pragma unroll
    for (int i = 0; i < CW_PER / 4; ++i) {
        const uint32_t s0 = 4u * (uint32_t)(i * CW_THREADS + tid);    // tile slot of r.x
        const int4 r = *reinterpret_cast<const int4 *>(h + cw_slot(s0));
        if (s0 >= w_lo && s0 + 4 <= w_hi) *reinterpret_cast<int4 *>(ob + s0) = r;
        else {
            if (s0 >= w_lo && s0 < w_hi) ob[s0] = (uint32_t)r.x;
            if (s0 + 1 >= w_lo && s0 + 1 < w_hi) ob[s0 + 1] = (uint32_t)r.y;
            if (s0 + 2 >= w_lo && s0 + 2 < w_hi) ob[s0 + 2] = (uint32_t)r.z;
            if (s0 + 3 >= w_lo && s0 + 3 < w_hi) ob[s0 + 3] = (uint32_t)r.w;
        }
    }
}

}  // namespace gams

using namespace gams;

struct gams_cuda_cwin_plan {
    gams_cuda_ctx *ctx = nullptr;
    int dev_index = 0, device = 0;
    int64_t n_ctg = 0;
    int32_t size = 0, step = 0;
    uint64_t n_win = 0, n_iv = 0, n_tiles = 0, n_ent = 0;
    DevBuf meta;       // chr_start | win_off | iv_off | boff | tiles
    DevBuf bounds;     // CwDesc per tile
    size_t o_win = 0, o_iv = 0, o_boff = 0, o_tiles = 0;
    uint64_t launches = 0;
    const uint32_t *desc_starts = nullptr, *desc_stops = nullptr;   // index the descriptors were built for
    bool desc_valid = false;
};

GAMS_API int32_t gams_cuda_cwin_plan_create(gams_cuda_ctx *ctx, int32_t dev_index, int64_t n_ctg, const int32_t *chr_start,
                                            const uint64_t *win_off, int32_t size, int32_t step, const uint64_t *iv_off,
                                            gams_cuda_cwin_plan **out) {
    ARG_CHECK(ctx && out && win_off && iv_off && (chr_start || n_ctg == 0), "cwin_plan_create: NULL argument");
    ARG_CHECK(dev_index >= 0 && dev_index < (int32_t)ctx->devs.size(), "cwin_plan_create: dev_index %d out of range", dev_index);
    ARG_CHECK(n_ctg >= 0 && n_ctg < (1ll << 31), "cwin_plan_create: n_ctg out of range");
    ARG_CHECK(size >= 1 && step >= 1, "cwin_plan_create: size and step must be >= 1");
    ARG_CHECK((int64_t)step * CW_TILE < (1ll << 32), "cwin_plan_create: step too large");
    *out = nullptr;
    std::unique_ptr<gams_cuda_cwin_plan, void (*)(gams_cuda_cwin_plan *)> p(new gams_cuda_cwin_plan, gams_cuda_cwin_plan_destroy);
    PerDevice &pd = ctx->devs[dev_index];
    p->ctx = ctx, p->dev_index = dev_index, p->device = pd.device, p->n_ctg = n_ctg, p->size = size, p->step = step;
    std::vector<uint64_t> boff((size_t)n_ctg + 1, 0), w((size_t)n_ctg + 1), iv((size_t)n_ctg + 1);
    std::vector<CwTile> tiles;
    for (int64_t c = 0; c <= n_ctg; ++c) { w[c] = win_off[c] - win_off[0]; iv[c] = iv_off[c] - iv_off[0]; }
    for (int64_t c = 0; c < n_ctg; ++c) {
        ARG_CHECK(win_off[c + 1] >= win_off[c] && iv_off[c + 1] >= iv_off[c], "cwin_plan_create: offsets decrease at %lld", (long long)c);
        ARG_CHECK(iv_off[c + 1] - iv_off[c] < (1ull << 32), "cwin_plan_create: ctg %lld has more than 2^32-1 intervals", (long long)c);
        const uint64_t nt = (w[c + 1] - w[c] + (w[c] & 3u) + CW_TILE - 1) / CW_TILE;   // + the phantom slots before window 0
        ARG_CHECK(nt < (1ull << 32), "cwin_plan_create: ctg %lld has too many windows", (long long)c);
        boff[c + 1] = boff[c] + nt + 1;
        for (uint64_t j = 0; j < nt; ++j) tiles.push_back(CwTile{(uint32_t)c, (uint32_t)j});
    }
    ARG_CHECK(tiles.size() < (1ull << 31), "cwin_plan_create: too many tiles for one launch");
    p->n_win = w[n_ctg], p->n_iv = iv[n_ctg], p->n_tiles = tiles.size(), p->n_ent = boff[n_ctg];
    DeviceGuard g(pd.device);
    p->meta.cache = p->bounds.cache = pd.cache;
    auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t nc = (size_t)n_ctg;
    p->o_win = up16(4 * std::max<size_t>(nc, 1));
    p->o_iv = p->o_win + 8 * (nc + 1);
    p->o_boff = p->o_iv + 8 * (nc + 1);
    p->o_tiles = p->o_boff + 8 * (nc + 1);
    int32_t rc;
    if ((rc = p->meta.reserve(p->o_tiles + sizeof(CwTile) * std::max<size_t>(tiles.size(), 1)))) return rc;
    if ((rc = p->bounds.reserve(sizeof(CwDesc) * std::max<uint64_t>(p->n_tiles, 1)))) return rc;
    uint8_t *m = p->meta.as<uint8_t>();
    cudaStream_t st = pd.stream;
    if (nc) CUDA_TRY(cudaMemcpyAsync(m, chr_start, 4 * nc, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(m + p->o_win, w.data(), 8 * (nc + 1), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(m + p->o_iv, iv.data(), 8 * (nc + 1), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(m + p->o_boff, boff.data(), 8 * (nc + 1), cudaMemcpyHostToDevice, st));
    if (!tiles.empty()) CUDA_TRY(cudaMemcpyAsync(m + p->o_tiles, tiles.data(), sizeof(CwTile) * tiles.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));   // the vectors are locals
    *out = p.release();
    return GAMS_CUDA_OK;
}

GAMS_API void gams_cuda_cwin_plan_destroy(gams_cuda_cwin_plan *p) {
    if (!p) return;
    DeviceGuard g(p->device);
    cudaDeviceSynchronize();
    p->meta.release();
    p->bounds.release();
    delete p;
}

GAMS_API uint64_t gams_cuda_cwin_plan_total_windows(const gams_cuda_cwin_plan *p) { return p ? p->n_win : 0; }
GAMS_API uint64_t gams_cuda_cwin_plan_launches(const gams_cuda_cwin_plan *p) { return p ? p->launches : 0; }
GAMS_API void gams_cuda_cwin_plan_invalidate(gams_cuda_cwin_plan *p) { if (p) p->desc_valid = false; }

GAMS_API int32_t gams_cuda_count_windows_dev(gams_cuda_cwin_plan *p, void *stream, const uint32_t *d_starts, const uint32_t *d_stops,
                                             uint32_t *d_out) {
    ARG_CHECK(p != nullptr, "count_windows_dev: NULL plan");
    if (p->n_tiles == 0) return GAMS_CUDA_OK;
    ARG_CHECK(d_out && ((d_starts && d_stops) || p->n_iv == 0), "count_windows_dev: NULL device pointer");
    ARG_CHECK((reinterpret_cast<uintptr_t>(d_out) & 15u) == 0, "count_windows_dev: d_out must be 16-byte aligned");
    DeviceGuard g(p->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint8_t *m = p->meta.as<uint8_t>();
    const int32_t *d_cs = reinterpret_cast<const int32_t *>(m);
    const uint64_t *d_win = reinterpret_cast<const uint64_t *>(m + p->o_win), *d_iv = reinterpret_cast<const uint64_t *>(m + p->o_iv);
    const CwTile *d_tiles = reinterpret_cast<const CwTile *>(m + p->o_tiles);
    CwDesc *d_desc = p->bounds.as<CwDesc>();
    const unsigned b0 = (unsigned)std::min<uint64_t>((p->n_tiles + 255) / 256, 148ull * 32);
    // descriptors belong to an index: rebuilt when the caller passes other arrays (as the count plan does with its tables);
    // a caller that rewrites the same arrays in place calls gams_cuda_cwin_plan_invalidate first
    if (!p->desc_valid || p->desc_starts != d_starts || p->desc_stops != d_stops) {
        cw_desc_kernel<<<b0, 256, 0, st>>>(d_starts, d_stops, d_iv, d_cs, d_win, d_tiles, p->n_tiles, p->size, p->step, d_desc);
        p->desc_starts = d_starts, p->desc_stops = d_stops, p->desc_valid = true;
        ++p->launches;
    }
    cw_tile_kernel<<<(unsigned)p->n_tiles, CW_THREADS, 0, st>>>(d_starts, d_stops, d_desc, p->step, cw_magic(p->step), d_out);
    ++p->launches;
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

// One device's share of the host call: ctgs [0, n_ctg) of the (shifted) offset arrays; starts / stops / out are the caller's
// whole arrays, addressed through the absolute offsets.
static int32_t count_windows_one(gams_cuda_ctx *ctx, int d, int64_t n_ctg, const int32_t *chr_start, const uint64_t *win_off,
                                 int32_t size, int32_t step, const uint32_t *starts, const uint32_t *stops, const uint64_t *iv_off,
                                 uint32_t *out) {
    const uint64_t n_win = win_off[n_ctg] - win_off[0], n_iv = iv_off[n_ctg] - iv_off[0];
    if (n_ctg == 0 || n_win == 0) return GAMS_CUDA_OK;
    PerDevice &pd = ctx->devs[d];
    DeviceGuard g(pd.device);
    gams_cuda_cwin_plan *plan = nullptr;
    int32_t rc = gams_cuda_cwin_plan_create(ctx, d, n_ctg, chr_start, win_off, size, step, iv_off, &plan);
    if (rc) return rc;
    std::unique_ptr<gams_cuda_cwin_plan, void (*)(gams_cuda_cwin_plan *)> guard(plan, gams_cuda_cwin_plan_destroy);
    if ((rc = pd.scratch[0].reserve(std::max<uint64_t>(n_iv, 1) * 8))) return rc;
    if ((rc = pd.scratch[1].reserve(n_win * 4))) return rc;
    cudaStream_t st = pd.stream;
    uint32_t *d_st = pd.scratch[0].as<uint32_t>(), *d_sp = d_st + std::max<uint64_t>(n_iv, 1), *d_o = pd.scratch[1].as<uint32_t>();
    if (n_iv) {
        CUDA_TRY(cudaMemcpyAsync(d_st, starts + iv_off[0], n_iv * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_sp, stops + iv_off[0], n_iv * 4, cudaMemcpyHostToDevice, st));
    }
    if ((rc = gams_cuda_count_windows_dev(plan, st, d_st, d_sp, d_o))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out + win_off[0], d_o, n_win * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return GAMS_CUDA_OK;
}

// Host entry point: index up, counts down (4 bytes per window); ctgs are split over the context's devices by windows.
GAMS_API int32_t gams_cuda_count_windows(gams_cuda_ctx *ctx, int64_t n_ctg, const int32_t *chr_start, const uint64_t *win_off,
                                         int32_t size, int32_t step, const uint32_t *starts, const uint32_t *stops,
                                         const uint64_t *iv_off, uint32_t *out) {
    ARG_CHECK(ctx && win_off && iv_off && n_ctg >= 0, "count_windows: bad argument");
    if (n_ctg == 0) return GAMS_CUDA_OK;
    const uint64_t n_win = win_off[n_ctg] - win_off[0], n_iv = iv_off[n_ctg] - iv_off[0];
    if (n_win == 0) return GAMS_CUDA_OK;
    ARG_CHECK(out && ((starts && stops) || n_iv == 0), "count_windows: NULL array");
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int n_dev = (int)std::min<int64_t>((int64_t)ctx->devs.size(), std::max<int64_t>(1, (int64_t)(n_win >> 20)));
    std::vector<int64_t> part((size_t)n_dev + 1);
    int32_t rc = gams_cuda_partition(n_ctg, win_off, n_dev, part.data());
    if (rc) return rc;
    return run_on_devices(n_dev, [&](int d) {
        const int64_t c0 = part[d], c1 = part[d + 1];
        return count_windows_one(ctx, d, c1 - c0, chr_start + c0, win_off + c0, size, step, starts, stops, iv_off + c0, out);
    });
}
