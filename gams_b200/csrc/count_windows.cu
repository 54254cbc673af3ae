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
// i.e. one histogram over step-wide bins and its exclusive prefix sum.  cw_bounds_kernel finds #{a < K0} of every tile
// (one binary search per tile and array, all in parallel); cw_tile_kernel streams each tile's slice of starts[] and
// stops[] once (coalesced), bins it with shared-memory atomics, scans, and writes 4 bytes per window.
//
// Bound: HBM.  Algorithmic bytes: 8 B per interval read + 4 B per window written; no per-window query arrays exist.
#include "common.cuh"
#include <memory>

namespace gams {

constexpr int CW_THREADS = 256;
constexpr int CW_PER = 16;                         // windows per thread in the scan
constexpr int CW_TILE = CW_THREADS * CW_PER;       // windows per tile

struct CwTile {
    uint32_t ctg;
    uint32_t j;        // tile index inside the ctg: windows [j CW_TILE, (j+1) CW_TILE)
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

// bound entry e = boff[c] + j  (j = 0 .. tiles_c):  bF[e] = #{starts_c < E0(c, j)},  bG[e] = #{stops_c < S0(c, j)}
__global__ void __launch_bounds__(256)
cw_bounds_kernel(const uint32_t *__restrict__ starts, const uint32_t *__restrict__ stops, const uint64_t *__restrict__ iv_off,
                 const int32_t *__restrict__ chr_start, const uint64_t *__restrict__ boff, int64_t n_ctg, int32_t size,
                 int32_t step, uint32_t *__restrict__ bF, uint32_t *__restrict__ bG) {
    const uint64_t n_ent = boff[n_ctg];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_ent; e += stride) {
        int64_t lo = 0, hi = n_ctg;   // last ctg with boff[c] <= e
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (boff[mid] <= e) lo = mid; else hi = mid;
        }
        const int64_t c = lo;
        const uint64_t j = e - boff[c];
        const uint64_t o = iv_off[c];
        const uint32_t n = (uint32_t)(iv_off[c + 1] - o);
        const int64_t w0 = (int64_t)chr_start[c] + (int64_t)(j * CW_TILE) * step;   // start of the tile's first window
        const int64_t kE = w0 + size - 1, kS = w0 + 1;
        bF[e] = kE <= 0 ? 0u : cw_lower_bound(starts + o, n, (uint64_t)kE);
        bG[e] = kS <= 0 ? 0u : cw_lower_bound(stops + o, n, (uint64_t)kS);
    }
}

__global__ void __launch_bounds__(CW_THREADS)
cw_tile_kernel(const uint32_t *__restrict__ starts, const uint32_t *__restrict__ stops, const uint64_t *__restrict__ iv_off,
               const int32_t *__restrict__ chr_start, const uint64_t *__restrict__ win_off, const uint64_t *__restrict__ boff,
               const CwTile *__restrict__ tiles, const uint32_t *__restrict__ bF, const uint32_t *__restrict__ bG,
               int32_t size, int32_t step, uint32_t *__restrict__ out) {
    __shared__ __align__(16) uint32_t hF[CW_TILE], hG[CW_TILE];
    __shared__ uint32_t wsF[CW_THREADS / 32], wsG[CW_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const CwTile td = tiles[blockIdx.x];
    const uint64_t e = boff[td.ctg] + td.j, o = iv_off[td.ctg];
    const uint32_t f0 = bF[e], f1 = bF[e + 1], g0 = bG[e], g1 = bG[e + 1];
    const int64_t w0 = (int64_t)chr_start[td.ctg] + (int64_t)((uint64_t)td.j * CW_TILE) * step;
    // elements f0 .. f1 of starts[] lie in [E0, E0 + CW_TILE step): their distance to E0 fits 32 bits (same for stops / S0)
    const uint32_t E0 = (uint32_t)(w0 + size - 1), S0 = (uint32_t)(w0 + 1), ustep = (uint32_t)step;
#pragma unroll
    for (int k = 0; k < CW_PER / 4; ++k) {
        reinterpret_cast<uint4 *>(hF)[k * CW_THREADS + tid] = make_uint4(0u, 0u, 0u, 0u);
        reinterpret_cast<uint4 *>(hG)[k * CW_THREADS + tid] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    for (uint32_t i = f0 + tid; i < f1; i += CW_THREADS) atomicAdd(&hF[min((__ldg(starts + o + i) - E0) / ustep, (uint32_t)CW_TILE - 1u)], 1u);
    for (uint32_t i = g0 + tid; i < g1; i += CW_THREADS) atomicAdd(&hG[min((__ldg(stops + o + i) - S0) / ustep, (uint32_t)CW_TILE - 1u)], 1u);
    __syncthreads();
    // exclusive prefix over the tile's bins: thread t owns bins [16 t, 16 t + 16)
    uint32_t vF[CW_PER], vG[CW_PER];
#pragma unroll
    for (int k = 0; k < CW_PER / 4; ++k) {
        const uint4 a = reinterpret_cast<const uint4 *>(hF)[(CW_PER / 4) * tid + k], b = reinterpret_cast<const uint4 *>(hG)[(CW_PER / 4) * tid + k];
        vF[4 * k] = a.x; vF[4 * k + 1] = a.y; vF[4 * k + 2] = a.z; vF[4 * k + 3] = a.w;
        vG[4 * k] = b.x; vG[4 * k + 1] = b.y; vG[4 * k + 2] = b.z; vG[4 * k + 3] = b.w;
    }
    uint32_t tF = 0, tG = 0;
#pragma unroll
    for (int k = 0; k < CW_PER; ++k) { tF += vF[k]; tG += vG[k]; }
    uint32_t sF = tF, sG = tG;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, sF, d), b = __shfl_up_sync(0xFFFFFFFFu, sG, d);
        if (lane >= d) { sF += a; sG += b; }
    }
    if (lane == 31) { wsF[warp] = sF; wsG[warp] = sG; }
    __syncthreads();
    uint32_t accF = f0 + sF - tF, accG = g0 + sG - tG;
    for (int w = 0; w < warp; ++w) { accF += wsF[w]; accG += wsG[w]; }
    uint32_t r[CW_PER];
#pragma unroll
    for (int k = 0; k < CW_PER; ++k) {
        r[k] = accF - accG;        // lower_bound(starts, we) - lower_bound(stops, ws + 1)
        accF += vF[k];
        accG += vG[k];
    }
#pragma unroll
    for (int k = 0; k < CW_PER / 4; ++k)
        reinterpret_cast<uint4 *>(hF)[(CW_PER / 4) * tid + k] = make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
    __syncthreads();
    const uint64_t wbase = win_off[td.ctg] + (uint64_t)td.j * CW_TILE;
    const uint64_t n_here = min((uint64_t)CW_TILE, win_off[td.ctg + 1] - wbase);
    for (uint32_t k = tid; k < n_here; k += CW_THREADS) out[wbase + k] = hF[k];
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
    DevBuf bounds;     // bF | bG
    size_t o_win = 0, o_iv = 0, o_boff = 0, o_tiles = 0;
    uint64_t launches = 0;
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
        const uint64_t nt = (w[c + 1] - w[c] + CW_TILE - 1) / CW_TILE;
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
    if ((rc = p->bounds.reserve(8 * std::max<uint64_t>(p->n_ent, 1)))) return rc;
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

GAMS_API int32_t gams_cuda_count_windows_dev(gams_cuda_cwin_plan *p, void *stream, const uint32_t *d_starts, const uint32_t *d_stops,
                                             uint32_t *d_out) {
    ARG_CHECK(p != nullptr, "count_windows_dev: NULL plan");
    if (p->n_tiles == 0) return GAMS_CUDA_OK;
    ARG_CHECK(d_out && ((d_starts && d_stops) || p->n_iv == 0), "count_windows_dev: NULL device pointer");
    DeviceGuard g(p->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint8_t *m = p->meta.as<uint8_t>();
    const int32_t *d_cs = reinterpret_cast<const int32_t *>(m);
    const uint64_t *d_win = reinterpret_cast<const uint64_t *>(m + p->o_win), *d_iv = reinterpret_cast<const uint64_t *>(m + p->o_iv),
                   *d_boff = reinterpret_cast<const uint64_t *>(m + p->o_boff);
    const CwTile *d_tiles = reinterpret_cast<const CwTile *>(m + p->o_tiles);
    uint32_t *bF = p->bounds.as<uint32_t>(), *bG = bF + p->n_ent;
    const unsigned b0 = (unsigned)std::min<uint64_t>((p->n_ent + 255) / 256, 148ull * 32);
    cw_bounds_kernel<<<b0, 256, 0, st>>>(d_starts, d_stops, d_iv, d_cs, d_boff, p->n_ctg, p->size, p->step, bF, bG);
    cw_tile_kernel<<<(unsigned)p->n_tiles, CW_THREADS, 0, st>>>(d_starts, d_stops, d_iv, d_cs, d_win, d_boff, d_tiles, bF, bG, p->size,
                                                               p->step, d_out);
    p->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

// Host entry point: index up, counts down (4 bytes per window); everything on device 0 of the context.
GAMS_API int32_t gams_cuda_count_windows(gams_cuda_ctx *ctx, int64_t n_ctg, const int32_t *chr_start, const uint64_t *win_off,
                                         int32_t size, int32_t step, const uint32_t *starts, const uint32_t *stops,
                                         const uint64_t *iv_off, uint32_t *out) {
    ARG_CHECK(ctx && win_off && iv_off && n_ctg >= 0, "count_windows: bad argument");
    if (n_ctg == 0) return GAMS_CUDA_OK;
    const uint64_t n_win = win_off[n_ctg] - win_off[0], n_iv = iv_off[n_ctg] - iv_off[0];
    if (n_win == 0) return GAMS_CUDA_OK;
    ARG_CHECK(out && ((starts && stops) || n_iv == 0), "count_windows: NULL array");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PerDevice &pd = ctx->devs[0];
    DeviceGuard g(pd.device);
    gams_cuda_cwin_plan *plan = nullptr;
    int32_t rc = gams_cuda_cwin_plan_create(ctx, 0, n_ctg, chr_start, win_off, size, step, iv_off, &plan);
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
