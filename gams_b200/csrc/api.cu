// api.cu — C ABI of libgams_cuda.so: context, plans, and the host-buffer entry points.
// See include/gams_cuda.h for the contract of every function and the reference call site it replaces.
#include <algorithm>
#include <cstdlib>
#include <memory>
#include <string>
#include <thread>

#include "common.cuh"
#include "inflate_common.cuh"

namespace gams {

// ---- errors ------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int32_t cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    const char *base = strrchr(file, '/');
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), base ? base + 1 : file, line, what);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        return GAMS_CUDA_ENOMEM;
    }
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return GAMS_CUDA_ENODEV;
    return GAMS_CUDA_ECUDA;
}

size_t BlockCache::round_up(size_t bytes) {
    if (bytes <= 256) return 256;
    if (bytes >= (1u << 20)) return ((bytes + (1u << 20) - 1) >> 20) << 20;
    size_t r = 256;
    while (r < bytes) r <<= 1;
    return r;
}

int32_t BlockCache::get(void **p, size_t bytes) {
    const size_t want = round_up(bytes);
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = free_blocks.lower_bound(want);
        if (it != free_blocks.end() && it->first <= want + want / 4) {   // at most 25 % larger than asked
            *p = it->second;
            free_blocks.erase(it);
            return GAMS_CUDA_OK;
        }
    }
    cudaError_t e = cudaMalloc(p, want);
    if (e == cudaErrorMemoryAllocation) {   // give the cached blocks back and try once more
        cudaGetLastError();
        clear();
        e = cudaMalloc(p, want);
    }
    CUDA_TRY(e);
    return GAMS_CUDA_OK;
}

void BlockCache::put(void *p, size_t bytes) {
    std::lock_guard<std::mutex> lk(mu);
    free_blocks.emplace(round_up(bytes), p);
}

void BlockCache::clear() {
    std::lock_guard<std::mutex> lk(mu);
    for (auto &kv : free_blocks) cudaFree(kv.second);
    free_blocks.clear();
}

BlockCache::~BlockCache() {
    if (h_stage) cudaFreeHost(h_stage);
}

__global__ void __launch_bounds__(256)
upload_table_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int32_t upload_table(BlockCache *cache, void *d_dst, const void *h_src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return GAMS_CUDA_OK;
    if (!cache || (reinterpret_cast<uintptr_t>(d_dst) & 15u)) {
        CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        return GAMS_CUDA_OK;
    }
    std::lock_guard<std::mutex> lk(cache->mu);
    const size_t padded = (bytes + 15) & ~(size_t)15;     // (device blocks are rounded up to 256 bytes and more: the padding fits)
    if (padded > cache->stage_cap) {
        if (cache->h_stage) cudaFreeHost(cache->h_stage);
        cache->h_stage = cache->d_stage = nullptr, cache->stage_cap = 0;
        const size_t cap = std::max<size_t>(padded + padded / 2, 1 << 20);
        CUDA_TRY(cudaHostAlloc(&cache->h_stage, cap, cudaHostAllocMapped));
        CUDA_TRY(cudaHostGetDevicePointer(&cache->d_stage, cache->h_stage, 0));
        cache->stage_cap = cap;
    }
    std::memcpy(cache->h_stage, h_src, bytes);
    const size_t n16 = padded / 16;
    const unsigned grid = (unsigned)std::min<size_t>((n16 + 255) / 256, 148 * 4);
    upload_table_kernel<<<grid, 256, 0, st>>>(static_cast<uint4 *>(d_dst), static_cast<const uint4 *>(cache->d_stage), n16);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(st));
    return GAMS_CUDA_OK;
}

int32_t dev_alloc(BlockCache *cache, void **p, size_t bytes) {
    if (cache) return cache->get(p, bytes);
    CUDA_TRY(cudaMalloc(p, bytes ? bytes : 1));
    return GAMS_CUDA_OK;
}

void dev_free(BlockCache *cache, void *p, size_t bytes) {
    if (!p) return;
    if (cache) cache->put(p, bytes);
    else cudaFree(p);
}

int32_t DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return GAMS_CUDA_OK;
    release();
    const size_t want = BlockCache::round_up(bytes);
    int32_t rc = dev_alloc(cache, &p, want);
    if (rc) return rc;
    cap = want;
    return GAMS_CUDA_OK;
}

void DevBuf::release() {
    dev_free(cache, p, cap);
    p = nullptr;
    cap = 0;
}

int32_t SeriesLayout::build(const uint64_t *h_off, int64_t n, cudaStream_t st) {
    n_series = n;
    total = n ? h_off[n] - h_off[0] : 0;
    min_len = ~0ull;
    std::vector<TileDesc> zt;
    std::vector<uint64_t> &rel = this->h_off;
    rel.assign((size_t)n + 1, 0);
    h_ztile_off.assign((size_t)n + 1, 0);
    for (int64_t c = 0; c <= n; ++c) rel[c] = h_off[c] - h_off[0];
    for (int64_t c = 0; c < n; ++c) {
        if (h_off[c + 1] < h_off[c]) {
            set_error("offsets must be non-decreasing (entry %lld)", (long long)c);
            return GAMS_CUDA_EINVAL;
        }
        const uint64_t len = h_off[c + 1] - h_off[c];
        min_len = std::min(min_len, len);
        for (uint64_t w = 0; w < len; w += ZS_TILE) zt.push_back({(uint32_t)c, (uint32_t)w});
        h_ztile_off[c + 1] = zt.size();
    }
    if (n == 0) min_len = 0;
    n_ztiles = zt.size();
    int32_t rc;
    if ((rc = dev_alloc(cache, reinterpret_cast<void **>(&d_off), sizeof(uint64_t) * (size_t)(n + 1)))) return rc;
    if ((rc = upload_table(cache, d_off, rel.data(), sizeof(uint64_t) * (size_t)(n + 1), st))) return rc;
    if ((rc = dev_alloc(cache, reinterpret_cast<void **>(&d_nonfinite), sizeof(int) * (size_t)std::max<int64_t>(n, 1)))) return rc;
    CUDA_TRY(cudaMemsetAsync(d_nonfinite, 0, sizeof(int) * (size_t)std::max<int64_t>(n, 1), st));
    if (n_ztiles) {
        if ((rc = dev_alloc(cache, reinterpret_cast<void **>(&d_ztiles), sizeof(TileDesc) * n_ztiles))) return rc;
        if ((rc = upload_table(cache, d_ztiles, zt.data(), sizeof(TileDesc) * n_ztiles, st))) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return GAMS_CUDA_OK;
}

// segment tables of series s0 .. s1 - 1 for segments of `seg` values (kept until another range or length is asked for)
int32_t SeriesLayout::seg_tables(int64_t s0, int64_t s1, int64_t seg, uint64_t lag, cudaStream_t st) const {
    int32_t rc;
    if (seg_s0 != s0 || seg_s1 != s1 || seg_len != seg) {
        h_seg_off.assign((size_t)(s1 - s0) + 1, 0);
        for (int64_t c = s0; c < s1; ++c) {
            const uint64_t len = h_off[c + 1] - h_off[c];
            const uint64_t k = h_seg_off[c - s0] + std::max<uint64_t>((len + (uint64_t)seg - 1) / (uint64_t)seg, 1);
            if (k > 0x7FFFFFFFull) {
                set_error("thresholding: too many segments");
                return GAMS_CUDA_EINVAL;
            }
            h_seg_off[c - s0 + 1] = (uint32_t)k;
        }
        const size_t need = h_seg_off.size() * sizeof(uint32_t);
        if (need > seg_off_bytes) {
            dev_free(cache, d_seg_off, seg_off_bytes);
            d_seg_off = nullptr, seg_off_bytes = 0;
            if ((rc = dev_alloc(cache, reinterpret_cast<void **>(&d_seg_off), need))) return rc;
            seg_off_bytes = need;
        }
        CUDA_TRY(cudaMemcpyAsync(d_seg_off, h_seg_off.data(), need, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));     // (h_seg_off may be rebuilt by the next call)
        seg_s0 = s0, seg_s1 = s1, seg_len = seg;
    }
    const size_t n_seg = h_seg_off.back();
    const size_t need_last = std::max<size_t>(n_seg, 1) * sizeof(long long), need_ring = std::max<size_t>(n_seg, 1) * (size_t)(lag + 1) * sizeof(float);
    if (need_last > seg_last_bytes) {
        dev_free(cache, d_seg_last, seg_last_bytes);
        d_seg_last = nullptr, seg_last_bytes = 0;
        if ((rc = dev_alloc(cache, reinterpret_cast<void **>(&d_seg_last), need_last))) return rc;
        seg_last_bytes = need_last;
    }
    if (need_ring > seg_ring_bytes) {
        dev_free(cache, d_seg_ring, seg_ring_bytes);
        d_seg_ring = nullptr, seg_ring_bytes = 0;
        if ((rc = dev_alloc(cache, reinterpret_cast<void **>(&d_seg_ring), need_ring))) return rc;
        seg_ring_bytes = need_ring;
    }
    return GAMS_CUDA_OK;
}

void SeriesLayout::release() {
    dev_free(cache, d_seg_off, seg_off_bytes);
    dev_free(cache, d_seg_last, seg_last_bytes);
    dev_free(cache, d_seg_ring, seg_ring_bytes);
    d_seg_off = nullptr, d_seg_last = nullptr, d_seg_ring = nullptr;
    seg_off_bytes = seg_last_bytes = seg_ring_bytes = 0;
    seg_s0 = seg_s1 = -1;

    dev_free(cache, d_off, sizeof(uint64_t) * (size_t)(n_series + 1));
    dev_free(cache, d_ztiles, sizeof(TileDesc) * n_ztiles);
    dev_free(cache, d_nonfinite, sizeof(int) * (size_t)std::max<int64_t>(n_series, 1));
    d_nonfinite = nullptr;
    d_off = nullptr;
    d_ztiles = nullptr;
}

static int64_t nwin(int64_t len, int32_t size, int32_t step) {
    if (size <= 0 || step <= 0) return -1;
    return len >= size ? (len - size) / step + 1 : 0;
}

// Run fn(d) for every device of the ctx, concurrently when there are several.
template <class F>
static int32_t for_each_device(gams_cuda_ctx *ctx, int n_used, F fn) {
    std::vector<int32_t> rc((size_t)n_used, 0);
    std::vector<std::string> msg((size_t)n_used);
    if (n_used == 1) return fn(0);
    std::vector<std::thread> th;
    for (int d = 0; d < n_used; ++d)
        th.emplace_back([&, d] {
            rc[d] = fn(d);
            if (rc[d]) msg[d] = std::string(g_err);
        });
    for (auto &t : th) t.join();
    for (int d = 0; d < n_used; ++d)
        if (rc[d]) {
            set_error("%s", msg[d].c_str());
            return rc[d];
        }
    (void)ctx;
    return GAMS_CUDA_OK;
}

}  // namespace gams

using namespace gams;

// GAMS_CUDA_TRACE=1: host-clock stage timings of the host API on stderr (synchronises; for debugging only)
#include <chrono>
namespace gams {
bool trace_on() {
    static const bool on = getenv("GAMS_CUDA_TRACE") != nullptr;
    return on;
}
void trace_stage(const char *what, cudaStream_t st, bool sync) {
    if (!trace_on()) return;
    static thread_local auto last = std::chrono::steady_clock::now();
    if (sync) cudaStreamSynchronize(st);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[gams_cuda] %-28s +%8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
}
}  // namespace gams

// ================================================================================ lifecycle =====
GAMS_API int32_t gams_cuda_version(void) { return 100; /* 0.1.0 */ }

GAMS_API const char *gams_cuda_last_error(void) { return g_err; }

GAMS_API int32_t gams_cuda_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (e == cudaErrorNoDevice) return 0;
        return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
    }
    return n;
}

GAMS_API int32_t gams_cuda_create(const int32_t *device_ids, int32_t n_dev, gams_cuda_ctx **out) {
    ARG_CHECK(out != nullptr, "gams_cuda_create: out is NULL");
    *out = nullptr;
    const int32_t avail = gams_cuda_device_count();
    if (avail < 0) return avail;
    if (avail == 0) {
        set_error("gams_cuda_create: no CUDA device visible; this library has no CPU fallback");
        return GAMS_CUDA_ENODEV;
    }
    std::vector<int> ids;
    if (!device_ids || n_dev <= 0)
        for (int d = 0; d < avail; ++d) ids.push_back(d);
    else
        for (int i = 0; i < n_dev; ++i) {
            ARG_CHECK(device_ids[i] >= 0 && device_ids[i] < avail, "gams_cuda_create: device %d not visible (%d devices)",
                      device_ids[i], avail);
            ids.push_back(device_ids[i]);
        }
    std::unique_ptr<gams_cuda_ctx> ctx(new gams_cuda_ctx);
    // every device is pushed into the context BEFORE its resources are created, so that a failure half way releases what
    // the earlier devices (and this one) already own through gams_cuda_destroy
    auto init_device = [&](int id) -> int32_t {
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, id));
        if (prop.major < 10) {
            set_error("gams_cuda_create: device %d is sm_%d%d; this library is built for sm_100a only", id, prop.major,
                      prop.minor);
            return GAMS_CUDA_ENODEV;
        }
        DeviceGuard g(id);
        ctx->devs.emplace_back();
        PerDevice &pd = ctx->devs.back();
        pd.device = id;
        pd.cache = new BlockCache;
        for (auto &b : pd.scratch) b.cache = pd.cache;
        CUDA_TRY(cudaStreamCreateWithFlags(&pd.stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&pd.copy_in, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&pd.copy_out, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&pd.aux, cudaStreamNonBlocking));
        CUDA_TRY(cudaMalloc(&pd.d_zstats, 2 * sizeof(unsigned long long)));
        CUDA_TRY(cudaMemset(pd.d_zstats, 0, 2 * sizeof(unsigned long long)));
        return GAMS_CUDA_OK;
    };
    for (int id : ids) {
        const int32_t rc = init_device(id);
        if (rc) {
            const std::string msg(g_err);
            gams_cuda_destroy(ctx.release());
            set_error("%s", msg.c_str());
            return rc;
        }
    }
    if (const char *m = getenv("GAMS_CUDA_ZSCORE_MODE")) ctx->zscore_mode = atoi(m);
    if (const char *m = getenv("GAMS_CUDA_GUNZIP_MODE")) ctx->gunzip_mode = atoi(m);
    if (const char *m = getenv("GAMS_CUDA_GUNZIP_CRC")) ctx->gunzip_crc = atoi(m);
    *out = ctx.release();
    return GAMS_CUDA_OK;
}

GAMS_API int32_t gams_cuda_set_zscore_mode(gams_cuda_ctx *ctx, int32_t mode) {
    ARG_CHECK(ctx && mode >= 0 && mode <= 2, "set_zscore_mode: mode must be 0 (filter), 1 (replay) or 2 (verify)");
    ctx->zscore_mode = mode;
    return GAMS_CUDA_OK;
}

GAMS_API int32_t gams_cuda_set_gunzip_options(gams_cuda_ctx *ctx, int32_t decoder, int32_t check_crc) {
    ARG_CHECK(ctx && (decoder == 0 || decoder == 1) && (check_crc == 0 || check_crc == 1), "set_gunzip_options: decoder and check_crc are 0 or 1");
    ctx->gunzip_mode = decoder;
    ctx->gunzip_crc = check_crc;
    return GAMS_CUDA_OK;
}

GAMS_API int32_t gams_cuda_zscore_stats(gams_cuda_ctx *ctx, uint64_t *out2, int32_t reset) {
    ARG_CHECK(ctx && out2, "zscore_stats: NULL argument");
    out2[0] = out2[1] = 0;
    for (auto &pd : ctx->devs) {
        DeviceGuard g(pd.device);
        unsigned long long h[2];
        CUDA_TRY(cudaDeviceSynchronize());
        CUDA_TRY(cudaMemcpy(h, pd.d_zstats, sizeof(h), cudaMemcpyDeviceToHost));
        out2[0] += h[0];
        out2[1] += h[1];
        if (reset) CUDA_TRY(cudaMemset(pd.d_zstats, 0, sizeof(h)));
    }
    return GAMS_CUDA_OK;
}

GAMS_API void gams_cuda_destroy(gams_cuda_ctx *ctx) {
    if (!ctx) return;
    for (auto &pd : ctx->devs) {
        DeviceGuard g(pd.device);
        if (pd.stream) cudaStreamSynchronize(pd.stream);
        for (auto &b : pd.scratch) b.release();
        if (pd.cache) {
            pd.cache->clear();
            delete pd.cache;
        }
        if (pd.d_zstats) cudaFree(pd.d_zstats);
        if (pd.stream) cudaStreamDestroy(pd.stream);       // (a context whose creation failed half way has NULL handles)
        if (pd.copy_in) cudaStreamDestroy(pd.copy_in);
        if (pd.copy_out) cudaStreamDestroy(pd.copy_out);
        if (pd.aux) cudaStreamDestroy(pd.aux);
    }
    delete ctx;
}

GAMS_API int32_t gams_cuda_ctx_num_devices(const gams_cuda_ctx *ctx) { return ctx ? (int32_t)ctx->devs.size() : 0; }

GAMS_API void *gams_cuda_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        set_error("gams_cuda_host_alloc: cudaMallocHost(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}

GAMS_API int32_t gams_cuda_memcpy_probe(gams_cuda_ctx *ctx, void *host, size_t bytes, int32_t kind) {
    ARG_CHECK(ctx && host && (kind == 1 || kind == 2), "memcpy_probe: bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PerDevice &pd = ctx->devs[0];
    DeviceGuard g(pd.device);
    int32_t rc;
    if ((rc = pd.scratch[7].reserve(bytes ? bytes : 1))) return rc;
    if (kind == 1) CUDA_TRY(cudaMemcpyAsync(pd.scratch[7].p, host, bytes, cudaMemcpyHostToDevice, pd.stream));
    else CUDA_TRY(cudaMemcpyAsync(host, pd.scratch[7].p, bytes, cudaMemcpyDeviceToHost, pd.stream));
    CUDA_TRY(cudaStreamSynchronize(pd.stream));
    return GAMS_CUDA_OK;
}

GAMS_API void gams_cuda_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

// ======================================================================== window arithmetic =====
GAMS_API int64_t gams_cuda_wave_nwin(int64_t ctg_len, int32_t size, int32_t step) { return nwin(ctg_len, size, step); }

GAMS_API int32_t gams_cuda_wave_win_offsets(int64_t n_ctg, const uint64_t *seq_off, int32_t size, int32_t step,
                                            uint64_t *win_off) {
    ARG_CHECK(n_ctg >= 0 && seq_off && win_off, "wave_win_offsets: NULL argument");
    ARG_CHECK(size > 0 && step > 0, "wave_win_offsets: size and step must be positive (got %d, %d)", size, step);
    win_off[0] = 0;
    for (int64_t c = 0; c < n_ctg; ++c) {
        ARG_CHECK(seq_off[c + 1] >= seq_off[c], "wave_win_offsets: seq_off not monotone at %lld", (long long)c);
        win_off[c + 1] = win_off[c] + (uint64_t)nwin((int64_t)(seq_off[c + 1] - seq_off[c]), size, step);
    }
    return GAMS_CUDA_OK;
}

GAMS_API int32_t gams_cuda_partition(int64_t n_item, const uint64_t *item_off, int32_t n_part, int64_t *part_off) {
    ARG_CHECK(n_item >= 0 && item_off && part_off && n_part > 0, "partition: bad argument");
    const uint64_t base = item_off[0], total = item_off[n_item] - base;
    part_off[0] = 0;
    int64_t c = 0;
    for (int32_t p = 1; p < n_part; ++p) {
        const uint64_t target = base + (uint64_t)((__uint128_t)total * (unsigned)p / (unsigned)n_part);
        // first item whose end passes the target goes to the shard whose boundary is nearer
        while (c < n_item && item_off[c + 1] <= target) ++c;
        if (c < n_item && target - item_off[c] > item_off[c + 1] - target) ++c;
        part_off[p] = std::max(c, part_off[p - 1]);
    }
    part_off[n_part] = n_item;
    return GAMS_CUDA_OK;
}

// ================================================================================ wave plan =====
GAMS_API int32_t gams_cuda_wave_plan_create(gams_cuda_ctx *ctx, int32_t dev_index, int64_t n_ctg,
                                            const uint64_t *seq_off, int32_t size, int32_t step, uint64_t lag,
                                            gams_cuda_wave_plan **out) {
    ARG_CHECK(ctx && out && seq_off, "wave_plan_create: NULL argument");
    ARG_CHECK(dev_index >= 0 && dev_index < (int32_t)ctx->devs.size(), "wave_plan_create: dev_index %d out of range",
              dev_index);
    ARG_CHECK(n_ctg >= 0 && n_ctg < (1ll << 32), "wave_plan_create: n_ctg out of range");
    ARG_CHECK(size > 0 && step > 0, "wave_plan_create: size and step must be positive (got %d, %d)", size, step);
    *out = nullptr;
    std::unique_ptr<gams_cuda_wave_plan, void (*)(gams_cuda_wave_plan *)> p(new gams_cuda_wave_plan,
                                                                            gams_cuda_wave_plan_destroy);
    PerDevice &pd = ctx->devs[dev_index];
    p->ctx = ctx;
    p->dev_index = dev_index;
    p->device = pd.device;
    p->n_ctg = n_ctg;
    p->size = size;
    p->step = step;
    p->lag = lag;
    p->h_seq_off.resize((size_t)n_ctg + 1);
    p->h_win_off.resize((size_t)n_ctg + 1);
    for (int64_t c = 0; c <= n_ctg; ++c) p->h_seq_off[c] = seq_off[c] - seq_off[0];
    p->total_bases = p->h_seq_off[n_ctg];
    p->h_win_off[0] = 0;
    for (int64_t c = 0; c < n_ctg; ++c) {
        ARG_CHECK(seq_off[c + 1] >= seq_off[c], "wave_plan_create: seq_off not monotone at %lld", (long long)c);
        const uint64_t len = seq_off[c + 1] - seq_off[c];
        ARG_CHECK(len < (1ull << 31), "wave_plan_create: ctg %lld longer than 2^31-1 bases", (long long)c);
        p->h_win_off[c + 1] = p->h_win_off[c] + (uint64_t)nwin((int64_t)len, size, step);
    }
    p->geom = plan_gc_geometry(size, step);
    std::vector<TileDesc> gt;
    p->h_gtile_off.assign((size_t)n_ctg + 1, 0);
    if (p->geom.k) {
        for (int64_t c = 0; c < n_ctg; ++c) {
            const uint64_t n_c = p->h_win_off[c + 1] - p->h_win_off[c];
            for (uint64_t w = 0; w < n_c; w += (uint64_t)p->geom.wt) gt.push_back({(uint32_t)c, (uint32_t)w});
            p->h_gtile_off[c + 1] = gt.size();
        }
        p->geom.n_tiles = gt.size();
        ARG_CHECK(gt.size() < (1ull << 31), "wave_plan_create: too many tiles");
    }
    // Kernel AB tiles (bytes -> GC + signals in one pass); only when the geometry is eligible
    std::vector<FusedTile> ft;
    p->h_ftile_off.assign((size_t)n_ctg + 1, 0);
    p->fused_wt = getenv("GAMS_CUDA_NO_FUSED") ? 0 : fused_windows_per_tile(size, step, lag);
    if (p->fused_wt) {
        for (int64_t c = 0; c < n_ctg; ++c) {
            const uint64_t n_c = p->h_win_off[c + 1] - p->h_win_off[c];
            fused_tiles_for_ctg(ft, p->h_seq_off[c], p->h_win_off[c], n_c, step, lag, p->fused_wt);
            p->h_ftile_off[c + 1] = ft.size();
        }
        ARG_CHECK(ft.size() < (1ull << 31), "wave_plan_create: too many tiles");
    }
    DeviceGuard g(pd.device);
    BlockCache *cache = pd.cache;
    p->win.cache = cache;
    p->peak_tmp.cache = cache;
    p->n_ftiles = ft.size();
    int32_t arc;
    if (!ft.empty()) {
        if ((arc = dev_alloc(cache, reinterpret_cast<void **>(&p->d_ftiles), sizeof(FusedTile) * ft.size()))) return arc;
        if ((arc = upload_table(cache, p->d_ftiles, ft.data(), sizeof(FusedTile) * ft.size(), pd.stream))) return arc;
    }
    if ((arc = dev_alloc(cache, reinterpret_cast<void **>(&p->d_seq_off), sizeof(uint64_t) * (size_t)(n_ctg + 1)))) return arc;
    if ((arc = upload_table(cache, p->d_seq_off, p->h_seq_off.data(), sizeof(uint64_t) * (size_t)(n_ctg + 1), pd.stream))) return arc;
    if (!gt.empty()) {
        if ((arc = dev_alloc(cache, reinterpret_cast<void **>(&p->d_gtiles), sizeof(TileDesc) * gt.size()))) return arc;
        if ((arc = upload_table(cache, p->d_gtiles, gt.data(), sizeof(TileDesc) * gt.size(), pd.stream))) return arc;
        std::vector<float> tab((size_t)size + 1);
        const float fs = (float)size;
        for (int32_t c = 0; c <= size; ++c) tab[c] = (float)c / fs;  // `count as f32 / l as f32`, one IEEE divide
        if ((arc = dev_alloc(cache, reinterpret_cast<void **>(&p->d_gc_tab), sizeof(float) * tab.size()))) return arc;
        if ((arc = upload_table(cache, p->d_gc_tab, tab.data(), sizeof(float) * tab.size(), pd.stream))) return arc;
    }
    int32_t rc = p->win.build(p->h_win_off.data(), n_ctg, pd.stream);
    if (rc) return rc;
    *out = p.release();
    return GAMS_CUDA_OK;
}

GAMS_API void gams_cuda_wave_plan_destroy(gams_cuda_wave_plan *p) {
    if (!p) return;
    DeviceGuard g(p->device);
    cudaDeviceSynchronize();   // as cudaFree would: nothing in flight may still read the tables that go back to the cache
    BlockCache *cache = p->ctx ? p->ctx->devs[p->dev_index].cache : nullptr;
    dev_free(cache, p->d_seq_off, sizeof(uint64_t) * (size_t)(p->n_ctg + 1));
    dev_free(cache, p->d_gtiles, sizeof(TileDesc) * (size_t)p->geom.n_tiles);
    dev_free(cache, p->d_ftiles, sizeof(FusedTile) * p->n_ftiles);
    dev_free(cache, p->d_gc_tab, sizeof(float) * ((size_t)p->size + 1));
    p->peak_tmp.release();
    p->win.release();
    delete p;
}

GAMS_API uint64_t gams_cuda_wave_plan_total_windows(const gams_cuda_wave_plan *p) { return p ? p->win.total : 0; }
GAMS_API uint64_t gams_cuda_wave_plan_total_bases(const gams_cuda_wave_plan *p) { return p ? p->total_bases : 0; }
GAMS_API uint64_t gams_cuda_wave_plan_seq_capacity(const gams_cuda_wave_plan *p) {
    return p ? ((p->total_bases + 31) & ~31ull) + 32 : 0;
}
GAMS_API uint64_t gams_cuda_wave_plan_launches(const gams_cuda_wave_plan *p) { return p ? p->launches : 0; }

GAMS_API int32_t gams_cuda_wave_gc_dev(gams_cuda_wave_plan *p, void *stream, const uint8_t *d_seq,
                                       uint32_t *d_gc_count, float *d_gc) {
    ARG_CHECK(p && d_gc && (d_seq || p->total_bases == 0), "wave_gc_dev: NULL argument");
    DeviceGuard g(p->device);
    return launch_gc_windows(p, (cudaStream_t)stream, d_seq, d_gc_count, d_gc, &p->launches, 0, p->n_ctg);
}

GAMS_API int32_t gams_cuda_wave_signal_dev(gams_cuda_wave_plan *p, void *stream, const float *d_gc, float threshold,
                                           float influence, float *d_filtered, int8_t *d_signal) {
    ARG_CHECK(p && d_signal && d_gc, "wave_signal_dev: NULL argument");
    DeviceGuard g(p->device);
    return launch_zscore(p->win, (cudaStream_t)stream, d_gc, p->lag, threshold, influence, d_filtered, d_signal,
                         &p->launches, p->ctx->zscore_mode, p->ctx->devs[p->dev_index].d_zstats, 0, p->n_ctg);
}

GAMS_API int32_t gams_cuda_wave_dev(gams_cuda_wave_plan *p, void *stream, const uint8_t *d_seq, float threshold,
                                    float influence, uint32_t *d_gc_count, float *d_gc, float *d_filtered,
                                    int8_t *d_signal) {
    ARG_CHECK(p && d_gc && d_signal && (d_seq || p->total_bases == 0), "wave_dev: NULL argument");
    DeviceGuard g(p->device);
    // the fused kernel writes 32 / 8 aligned bytes per thread
    const bool aligned = (reinterpret_cast<uintptr_t>(d_gc) & 15u) == 0 && (reinterpret_cast<uintptr_t>(d_signal) & 7u) == 0 &&
                         (reinterpret_cast<uintptr_t>(d_gc_count) & 15u) == 0;
    if (aligned && wave_fused_ok(p, threshold, influence))
        return launch_gc_wave_fused(p, (cudaStream_t)stream, d_seq, threshold, d_gc_count, d_gc, d_signal,
                                    p->ctx->devs[p->dev_index].d_zstats, 0, p->n_ctg);
    // influence != 1 on a fused-eligible geometry: the fused kernel gives GC and the influence-1 decisions, the walk replays
    // the stretches behind the signals
    if (aligned && influence != 1.0f && wave_fused_ok(p, threshold, 1.0f) && zscore_skip_ok(p->lag)) {
        int32_t rc = launch_gc_wave_fused(p, (cudaStream_t)stream, d_seq, threshold, d_gc_count, d_gc, d_signal,
                                          p->ctx->devs[p->dev_index].d_zstats, 0, p->n_ctg);
        if (rc) return rc;
        return launch_zscore_skip(p->win, (cudaStream_t)stream, d_gc, p->lag, threshold, influence, d_signal, &p->launches, 0, p->n_ctg, true);
    }
    int32_t rc = launch_gc_windows(p, (cudaStream_t)stream, d_seq, d_gc_count, d_gc, &p->launches, 0, p->n_ctg);
    if (rc) return rc;
    return launch_zscore(p->win, (cudaStream_t)stream, d_gc, p->lag, threshold, influence, d_filtered, d_signal,
                         &p->launches, p->ctx->zscore_mode, p->ctx->devs[p->dev_index].d_zstats, 0, p->n_ctg, true);
}

GAMS_API int32_t gams_cuda_wave_plan_fused(const gams_cuda_wave_plan *p, float threshold, float influence) {
    return (p && wave_fused_ok(p, threshold, influence)) ? 1 : 0;
}

GAMS_API int32_t gams_cuda_wave_peaks_dev(gams_cuda_wave_plan *p, void *stream, const float *d_gc, const int8_t *d_signal,
                                          float coverage, uint64_t cap, uint32_t *d_pk_ctg, uint32_t *d_pk_first,
                                          uint32_t *d_pk_last, float *d_pk_gc, int8_t *d_pk_signal, uint64_t *d_n_peaks) {
    ARG_CHECK(p && d_n_peaks && (p->win.total == 0 || (d_gc && d_signal)), "wave_peaks_dev: NULL argument");
    ARG_CHECK(cap == 0 || (d_pk_ctg && d_pk_first && d_pk_last && d_pk_gc && d_pk_signal), "wave_peaks_dev: NULL output");
    ARG_CHECK(peak_merge_is_runs(p->size, p->step, coverage),
              "wave_peaks_dev: with coverage %g some intersecting windows are not merged (wave.rs:229-231): components are "
              "not runs; use gams_cuda_wave_text on the per-window arrays", (double)coverage);
    DeviceGuard g(p->device);
    int32_t rc = p->peak_tmp.reserve(peak_tmp_bytes(p->win.total));
    if (rc) return rc;
    p->launches += 3;
    return launch_peak_runs((cudaStream_t)stream, d_signal, d_gc, p->win.d_off, p->n_ctg, 0, p->win.total, p->size, p->step,
                            p->peak_tmp.as<uint64_t>(), cap, d_pk_ctg, d_pk_first, d_pk_last, d_pk_gc, d_pk_signal, d_n_peaks);
}

// ================================================================================ wave: host ====
struct PeakSink {   // per-device results of gams_cuda_wave_peaks_batch
    float coverage = 0.0f;
    // single device: records go straight into the caller's arrays (room for `cap`); otherwise into the vectors below
    uint64_t cap = 0, n = 0;
    uint32_t *o_ctg = nullptr, *o_first = nullptr, *o_last = nullptr;
    float *o_gc = nullptr;
    int8_t *o_sig = nullptr;
    std::vector<uint32_t> ctg, first, last;
    std::vector<float> gc;
    std::vector<int8_t> sig;
};

struct GzSource {   // gams_cuda_wave_gz_batch: the ctgs arrive as the gzip blobs Redis holds (seq:{ctg}), not as bases
    const uint8_t *gz_concat = nullptr;
    const uint64_t *gz_off = nullptr;   // n_ctg + 1
    int32_t *status = nullptr;          // n_ctg inflate status codes
};

static int32_t wave_on_device(gams_cuda_ctx *ctx, int d, int64_t c0, int64_t c1, const uint8_t *seq_concat,
                              const uint64_t *seq_off, int32_t size, int32_t step, uint64_t lag, float threshold,
                              float influence, const uint64_t *win_off, uint32_t *gc_count, float *gc, int8_t *signal,
                              PeakSink *sink = nullptr, const GzSource *gzs = nullptr) {
    if (c1 <= c0) return GAMS_CUDA_OK;
    PerDevice &pd = ctx->devs[d];
    DeviceGuard g(pd.device);
    trace_stage("wave: enter");
    int32_t rc;
    // 1. start the upload: slices of ~64 MB of sequence go out on copy_in before anything else is done, so that the
    //    plan (host loops over every tile + table uploads, ~7 ms for a human-size genome) is built while the link is busy
    const int64_t n = c1 - c0;
    const uint64_t base0 = seq_off[c0], nb = seq_off[c1] - base0;
    if ((rc = pd.scratch[0].reserve(((nb + 31) & ~31ull) + 32))) return rc;
    uint8_t *d_seq = pd.scratch[0].as<uint8_t>();
    uint64_t slice_bytes = 64ull << 20;
    if (const char *e = getenv("GAMS_CUDA_WAVE_SLICE")) slice_bytes = std::max<uint64_t>(1, strtoull(e, nullptr, 10));   // tests: many slices on small inputs
    std::vector<int64_t> cut{0};
    for (int64_t c = 0; c < n; ++c) {
        if (seq_off[c0 + c + 1] < seq_off[c0 + c]) {
            set_error("wave_batch: seq_off not monotone at %lld", (long long)(c0 + c));
            return GAMS_CUDA_EINVAL;
        }
        if (seq_off[c0 + c + 1] - seq_off[c0 + cut.back()] >= slice_bytes || c + 1 == n) cut.push_back(c + 1);
    }
    const size_t n_sl = cut.size() - 1;
    // GAMS_CUDA_TIMELINE=1: the events carry times, and the call prints when each slice had arrived, was computed and was back
    // (device clock, ms after the first copy was enqueued); for measurements, nothing is synchronised on the way
    const bool timeline = getenv("GAMS_CUDA_TIMELINE") != nullptr;
    const unsigned ev_flags = timeline ? cudaEventDefault : cudaEventDisableTiming;
    std::vector<cudaEvent_t> ev_in(n_sl), ev_k(n_sl), ev_out(timeline ? n_sl + 1 : 0);
    for (size_t k = 0; k < n_sl; ++k) {
        CUDA_TRY(cudaEventCreateWithFlags(&ev_in[k], ev_flags));
        CUDA_TRY(cudaEventCreateWithFlags(&ev_k[k], ev_flags));
    }
    for (auto &e : ev_out) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDefault));
    if (timeline) CUDA_TRY(cudaEventRecord(ev_out[n_sl], pd.copy_in));   // time zero
    struct EvGuard {
        std::vector<cudaEvent_t> &a, &b, &c;
        cudaStream_t in, aux;
        std::vector<cudaEvent_t> gz_done;   // one per piece of the gz path
        ~EvGuard() {
            cudaStreamSynchronize(in);   // an early error return must not leave copies from the caller's buffer in flight
            cudaStreamSynchronize(aux);  // ... nor the decoder writing the caller's status array
            for (auto e : a) cudaEventDestroy(e);
            for (auto e : b) cudaEventDestroy(e);
            for (auto e : c) cudaEventDestroy(e);
            for (auto e : gz_done) cudaEventDestroy(e);
        }
    } ev_guard{ev_in, ev_k, ev_out, pd.copy_in, pd.aux, {}};
    const uint8_t *h_seq = gzs ? nullptr : seq_concat + base0;
    // gz source: the compressed bytes go up instead (a third of the bases), in the same slices of ctgs, and are inflated on the
    // device.  When only peak records come back, the whole batch is decoded at once (the decoder's time is the serial chain of
    // its longest blob rather than the number of blobs).  When all windows come back (1.55 GB per 3.1 Gbp: longer on the link
    // than the blobs going up), the batch is decoded in pieces, so that the windows of piece q go down while piece q + 1 is
    // being decoded: upload, decode and download overlap instead of following one another.
    uint8_t *d_gz = nullptr;
    uint64_t gz_base = 0, gz_bytes = 0;
    size_t n_piece = 1;
    if (gzs) {
        gz_base = gzs->gz_off[c0], gz_bytes = gzs->gz_off[c1] - gz_base;
        if ((rc = pd.scratch[6].reserve(gz_bytes + 64 + (size_t)(n + 1) * 16 + (size_t)n * 4 + 64))) return rc;
        d_gz = pd.scratch[6].as<uint8_t>();
        if (!sink && gc) {
            // GAMS_CUDA_GZ_PIECES: the number of pieces, for tests and measurements; by default three, and none below ~128 MB
            // of bases per piece, which is all floor (one deflate block per thread, one blob per warp)
            const char *e = getenv("GAMS_CUDA_GZ_PIECES");
            n_piece = e ? (size_t)std::max<long>(1, std::min<long>(atol(e), 16)) : std::min<size_t>(2, (size_t)(nb >> 27));
            n_piece = std::max<size_t>(1, std::min<size_t>(n_piece, n_sl));
        }
    }
    // slices [n_sl q / n_str, n_sl (q + 1) / n_str) form stretch q: one upload each; a piece is a stretch when there are pieces
    const size_t n_str = n_piece > 1 ? n_piece : 4;
    size_t gz_from = 0;      // first slice whose blobs have not been handed to the copy engine yet
    auto enqueue_h2d = [&](size_t k) -> int32_t {
        if (gzs) {
            // the blobs go up in (up to) four copies, one per stretch the decoder's header search waits for (48 copies of 21 MB
            // reached 49 GB/s of the link's 55); ev_in[k] of a slice inside a stretch fires early and is not waited for
            bool stretch_end = false;
            for (size_t q = 1; q <= n_str; ++q) stretch_end |= k + 1 == n_sl * q / n_str;
            if (stretch_end) {
                const uint64_t b0 = gzs->gz_off[c0 + cut[gz_from]] - gz_base, b1 = gzs->gz_off[c0 + cut[k + 1]] - gz_base;
                if (b1 > b0) CUDA_TRY(cudaMemcpyAsync(d_gz + b0, gzs->gz_concat + gz_base + b0, b1 - b0, cudaMemcpyHostToDevice, pd.copy_in));
                gz_from = k + 1;
            }
        } else {
            const uint64_t b0 = seq_off[c0 + cut[k]] - base0, b1 = seq_off[c0 + cut[k + 1]] - base0;
            if (b1 > b0) CUDA_TRY(cudaMemcpyAsync(d_seq + b0, h_seq + b0, b1 - b0, cudaMemcpyHostToDevice, pd.copy_in));
        }
        CUDA_TRY(cudaEventRecord(ev_in[k], pd.copy_in));
        return GAMS_CUDA_OK;
    };
    // ~10 ms of link time up front (the plan's own table uploads must not queue behind the whole genome); compressed blobs
    // are a third of that and go up at once
    const size_t n_pre = gzs ? n_sl : std::min<size_t>(n_sl, 8);
    for (size_t k = 0; k < n_pre; ++k)
        if ((rc = enqueue_h2d(k))) return rc;
    trace_stage("wave: H2D enqueued");
    std::vector<uint64_t> gz_rel, seq_rel;
    if (gzs) {
        // The decoder starts NOW, on its own stream, before the plan is built (8 ms of host work): its header search follows the
        // upload in (up to) four stretches, and the token and byte phases run as soon as the last blob has arrived.
        // Offsets of the blobs / ctgs relative to this device's buffers first.
        gz_rel.resize((size_t)n + 1), seq_rel.resize((size_t)n + 1);
        for (int64_t c = 0; c <= n; ++c) {
            gz_rel[c] = gzs->gz_off[c0 + c] - gz_base;
            seq_rel[c] = seq_off[c0 + c] - base0;
        }
        uint8_t *meta = d_gz + ((gz_bytes + 63) & ~63ull);
        uint64_t *d_goff = reinterpret_cast<uint64_t *>(meta), *d_ooff = d_goff + (n + 1);
        int32_t *d_status = reinterpret_cast<int32_t *>(d_ooff + (n + 1));
        CUDA_TRY(cudaMemcpyAsync(d_goff, gz_rel.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, pd.aux));
        CUDA_TRY(cudaMemcpyAsync(d_ooff, seq_rel.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, pd.aux));
        // gz_done[q] fires when the bases of piece q are final: the GC-wave kernels run beside the CRC-32 pass, which only sets
        // status codes (fetched at the end: a copy into pageable host memory blocks the host until the decoder is through)
        ev_guard.gz_done.assign(n_piece, nullptr);
        for (auto &e : ev_guard.gz_done) CUDA_TRY(cudaEventCreateWithFlags(&e, ev_flags));
        if (n_piece == 1) {
            std::vector<int64_t> arrive_end;
            std::vector<cudaEvent_t> arrive_ev;
            for (size_t q = 1; q <= n_str && n_sl; ++q) {
                const size_t k = n_sl * q / n_str;
                if (k == 0 || (!arrive_end.empty() && cut[k] == arrive_end.back())) continue;
                arrive_end.push_back(cut[k]), arrive_ev.push_back(ev_in[k - 1]);
            }
            if ((rc = launch_gunzip2(pd, pd.aux, d_gz, d_goff, n, d_seq, d_ooff, d_status, nb, ctx->gunzip_mode, ctx->gunzip_crc != 0,
                                     (int)arrive_end.size(), arrive_end.data(), arrive_ev.data(), ev_guard.gz_done[0])))
                return rc;
        } else {
            // The pieces follow one another on pd.aux and share the decoder's scratch, sized here for the whole batch (no piece
            // needs more).  Pieces side by side on streams of their own were measured and are no faster (profiles/README.md, r2h):
            // a later piece's kernels only get on the device in the tail of the earlier one's.
            size_t by = 0;
            if ((rc = gunzip2_scratch_bytes(pd, n, nb, ctx->gunzip_mode, ctx->gunzip_crc != 0, &by))) return rc;
            if ((rc = pd.scratch[7].reserve(by))) return rc;
            for (size_t q = 0; q < n_piece; ++q) {
                const size_t k0 = n_sl * q / n_piece, k1 = n_sl * (q + 1) / n_piece;   // k1 > k0: n_piece <= n_sl
                const int64_t a = cut[k0], b = cut[k1], all = b - a;
                if ((rc = launch_gunzip2(pd, pd.aux, d_gz, d_goff + a, b - a, d_seq, d_ooff + a, d_status + a, seq_rel[b] - seq_rel[a],
                                         ctx->gunzip_mode, ctx->gunzip_crc != 0, 1, &all, &ev_in[k1 - 1], ev_guard.gz_done[q])))
                    return rc;
            }
        }
        trace_stage("wave: gunzip enqueued");
    }
    // 2. the plan
    gams_cuda_wave_plan *plan = nullptr;
    rc = gams_cuda_wave_plan_create(ctx, d, c1 - c0, seq_off + c0, size, step, lag, &plan);
    if (rc) return rc;
    trace_stage("wave: plan_create");
    std::unique_ptr<gams_cuda_wave_plan, void (*)(gams_cuda_wave_plan *)> guard(plan, gams_cuda_wave_plan_destroy);
    const uint64_t nwn = plan->win.total;
    if (nwn != win_off[c1] - win_off[c0]) {
        set_error("wave_batch: win_off does not match gams_cuda_wave_win_offsets for size=%d step=%d", size, step);
        return GAMS_CUDA_EINVAL;
    }
    if ((signal || sink) && lag == 0) {
        set_error("wave_batch: lag must be >= 1");
        return GAMS_CUDA_EINVAL;
    }
    if ((signal || sink) && plan->win.min_len < lag) {
        set_error("wave_batch: a ctg has %llu windows but lag is %llu (the reference panics, stat.rs:30)",
                  (unsigned long long)plan->win.min_len, (unsigned long long)lag);
        return GAMS_CUDA_ELAG;
    }
    const bool want_sig = signal || sink;   // sink: the per-window arrays stay on the device, only peak records come back
    if ((rc = pd.scratch[1].reserve(std::max<uint64_t>(nwn, 1) * sizeof(float)))) return rc;
    if (gc_count && (rc = pd.scratch[2].reserve(std::max<uint64_t>(nwn, 1) * sizeof(uint32_t)))) return rc;
    if (want_sig && (rc = pd.scratch[3].reserve(std::max<uint64_t>(nwn, 1)))) return rc;
    const bool need_filt = want_sig && !(influence == 1.0f && lag <= ZS_MAX_TILED_LAG);
    if (need_filt && (rc = pd.scratch[4].reserve(std::max<uint64_t>(nwn, 1) * sizeof(float)))) return rc;
    float *d_gc = pd.scratch[1].as<float>();
    uint32_t *d_cnt = gc_count ? pd.scratch[2].as<uint32_t>() : nullptr;
    int8_t *d_sig = want_sig ? pd.scratch[3].as<int8_t>() : nullptr;
    float *d_filt = need_filt ? pd.scratch[4].as<float>() : nullptr;
    // 3. kernels of slice k as soon as it has arrived, D2H of its results behind them: three streams, PCIe both ways
    //    (the kernels are far shorter than either copy)
    const uint64_t wo = win_off[c0];
    CUDA_TRY(cudaStreamSynchronize(pd.stream));  // plan tables were uploaded on pd.stream
    for (size_t k = n_pre; k < n_sl; ++k)
        if ((rc = enqueue_h2d(k))) return rc;
    trace_stage("wave: reserve");
    // The sequential z-score kernel (influence != 1: one warp per ctg) needs thousands of ctgs in flight to hide its
    // latency; on a 64 MB slice (~130 ctgs) it takes as long as on a quarter of a genome.  So it runs once per GROUP of
    // slices (four groups: the first three overlap the upload of the next), 321 -> ~90 ms per 3.1 Gbp end to end.
    const bool fused_walk = want_sig && influence != 1.0f && wave_fused_ok(plan, threshold, 1.0f) && zscore_skip_ok(lag);
    const bool fused = want_sig && (wave_fused_ok(plan, threshold, influence) || fused_walk);
    const bool deferred = want_sig && (fused_walk || (!fused && need_filt));
    const size_t n_grp = deferred ? std::min<size_t>(4, n_sl) : 0;
    size_t grp_first = 0;
    for (size_t k = 0; k < n_sl; ++k) {
        const int64_t a = cut[k], b = cut[k + 1];
        const uint64_t w0 = plan->h_win_off[a], w1 = plan->h_win_off[b];
        if (gzs) {   // the bases of slice k are there once the decoder is through with its piece
            size_t q = 0;
            while (q + 1 < n_piece && k >= n_sl * (q + 1) / n_piece) ++q;
            CUDA_TRY(cudaStreamWaitEvent(pd.stream, ev_guard.gz_done[q], 0));
        } else CUDA_TRY(cudaStreamWaitEvent(pd.stream, ev_in[k], 0));
        if (fused) {
            if ((rc = launch_gc_wave_fused(plan, pd.stream, d_seq, threshold, d_cnt, d_gc, d_sig, pd.d_zstats, a, b))) return rc;
        } else {
            if ((rc = launch_gc_windows(plan, pd.stream, d_seq, d_cnt, d_gc, &plan->launches, a, b))) return rc;
            if (want_sig && !deferred && (rc = launch_zscore(plan->win, pd.stream, d_gc, lag, threshold, influence, d_filt, d_sig,
                                                             &plan->launches, ctx->zscore_mode, pd.d_zstats, a, b, true)))
                return rc;
        }
        CUDA_TRY(cudaEventRecord(ev_k[k], pd.stream));
        CUDA_TRY(cudaStreamWaitEvent(pd.copy_out, ev_k[k], 0));
        if (w1 > w0 && gc) {
            CUDA_TRY(cudaMemcpyAsync(gc + wo + w0, d_gc + w0, (w1 - w0) * sizeof(float), cudaMemcpyDeviceToHost, pd.copy_out));
            if (gc_count)
                CUDA_TRY(cudaMemcpyAsync(gc_count + wo + w0, d_cnt + w0, (w1 - w0) * sizeof(uint32_t), cudaMemcpyDeviceToHost, pd.copy_out));
            if (signal && !deferred) CUDA_TRY(cudaMemcpyAsync(signal + wo + w0, d_sig + w0, w1 - w0, cudaMemcpyDeviceToHost, pd.copy_out));
        }
        if (timeline) CUDA_TRY(cudaEventRecord(ev_out[k], pd.copy_out));
        if (deferred && (k + 1 == n_sl || (k + 1) * n_grp / n_sl != k * n_grp / n_sl)) {   // last slice of its group
            const int64_t ga = cut[grp_first];
            const uint64_t gw0 = plan->h_win_off[ga];
            if (fused_walk) rc = launch_zscore_skip(plan->win, pd.stream, d_gc, lag, threshold, influence, d_sig, &plan->launches, ga, b, true);
            else rc = launch_zscore(plan->win, pd.stream, d_gc, lag, threshold, influence, d_filt, d_sig, &plan->launches,
                                    ctx->zscore_mode, pd.d_zstats, ga, b, true);
            if (rc) return rc;
            if (signal && w1 > gw0) {
                CUDA_TRY(cudaEventRecord(ev_k[k], pd.stream));   // re-recorded behind the group's z-score kernel
                CUDA_TRY(cudaStreamWaitEvent(pd.copy_out, ev_k[k], 0));
                CUDA_TRY(cudaMemcpyAsync(signal + wo + gw0, d_sig + gw0, w1 - gw0, cudaMemcpyDeviceToHost, pd.copy_out));
            }
            grp_first = k + 1;
        }
    }
    trace_stage("wave: enqueue slices");
    if (trace_on()) {
        cudaStreamSynchronize(pd.copy_in);
        trace_stage("wave: H2D done");
        cudaStreamSynchronize(pd.stream);
        trace_stage("wave: kernels done");
        cudaStreamSynchronize(pd.copy_out);
        trace_stage("wave: D2H done");
    }
    if (sink) {
        // peak merge on the device; two attempts at most: the first sizes the record arrays
        uint64_t cap = std::max<uint64_t>(nwn / 64, 4096), n_pk = 0;
        for (int attempt = 0; attempt < 2; ++attempt) {
            const uint64_t rec = cap * 17 + 64;
            if ((rc = pd.scratch[5].reserve(rec + 16))) return rc;
            uint8_t *base = pd.scratch[5].as<uint8_t>();
            uint64_t *d_n = reinterpret_cast<uint64_t *>(base);
            uint32_t *k_ctg = reinterpret_cast<uint32_t *>(base + 16), *k_first = k_ctg + cap, *k_last = k_first + cap;
            float *k_gc = reinterpret_cast<float *>(k_last + cap);
            int8_t *k_sig = reinterpret_cast<int8_t *>(k_gc + cap);
            if ((rc = gams_cuda_wave_peaks_dev(plan, pd.stream, d_gc, d_sig, sink->coverage, cap, k_ctg, k_first, k_last, k_gc,
                                               k_sig, d_n)))
                return rc;
            CUDA_TRY(cudaMemcpyAsync(&n_pk, d_n, sizeof(uint64_t), cudaMemcpyDeviceToHost, pd.stream));
            CUDA_TRY(cudaStreamSynchronize(pd.stream));
            sink->n = n_pk;
            if (sink->o_ctg && n_pk > sink->cap) break;   // the caller's arrays are too small: report the count only
            if (n_pk <= cap) {
                uint32_t *h_ctg = sink->o_ctg, *h_first = sink->o_first, *h_last = sink->o_last;
                float *h_gc = sink->o_gc;
                int8_t *h_sig = sink->o_sig;
                if (!h_ctg) {
                    sink->ctg.resize(n_pk), sink->first.resize(n_pk), sink->last.resize(n_pk), sink->gc.resize(n_pk), sink->sig.resize(n_pk);
                    h_ctg = sink->ctg.data(), h_first = sink->first.data(), h_last = sink->last.data(), h_gc = sink->gc.data(), h_sig = sink->sig.data();
                }
                if (n_pk) {
                    CUDA_TRY(cudaMemcpyAsync(h_ctg, k_ctg, n_pk * 4, cudaMemcpyDeviceToHost, pd.stream));
                    CUDA_TRY(cudaMemcpyAsync(h_first, k_first, n_pk * 4, cudaMemcpyDeviceToHost, pd.stream));
                    CUDA_TRY(cudaMemcpyAsync(h_last, k_last, n_pk * 4, cudaMemcpyDeviceToHost, pd.stream));
                    CUDA_TRY(cudaMemcpyAsync(h_gc, k_gc, n_pk * 4, cudaMemcpyDeviceToHost, pd.stream));
                    CUDA_TRY(cudaMemcpyAsync(h_sig, k_sig, n_pk, cudaMemcpyDeviceToHost, pd.stream));
                }
                break;
            }
            cap = n_pk;
        }
    }
    CUDA_TRY(cudaStreamSynchronize(pd.copy_out));
    if (timeline && !deferred) {
        CUDA_TRY(cudaStreamSynchronize(pd.copy_in));
        CUDA_TRY(cudaStreamSynchronize(pd.stream));
        auto at = [&](cudaEvent_t e) { float ms = -1.0f; cudaEventElapsedTime(&ms, ev_out[n_sl], e); return ms; };
        for (size_t q = 0; q < ev_guard.gz_done.size(); ++q) fprintf(stderr, "[gams_cuda] timeline: piece %zu decoded at %.2f ms\n", q, at(ev_guard.gz_done[q]));
        for (size_t k = 0; k < n_sl; ++k)
            if (k % 4 == 3 || k + 1 == n_sl || n_sl <= 8)
                fprintf(stderr, "[gams_cuda] timeline: slice %3zu arrived %7.2f computed %7.2f back %7.2f ms\n", k, at(ev_in[k]), at(ev_k[k]), at(ev_out[k]));
    }
    if (gzs) {
        const int32_t *d_status = reinterpret_cast<const int32_t *>(reinterpret_cast<const uint64_t *>(d_gz + ((gz_bytes + 63) & ~63ull)) + 2 * (n + 1));
        CUDA_TRY(cudaMemcpyAsync(gzs->status + c0, d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, pd.aux));
        CUDA_TRY(cudaStreamSynchronize(pd.aux));
    }
    trace_stage("wave: sync copy_out");
    CUDA_TRY(cudaStreamSynchronize(pd.stream));
    trace_stage("wave: all streams idle");
    guard.reset();
    trace_stage("wave: plan_destroy");
    return GAMS_CUDA_OK;
}

GAMS_API int32_t gams_cuda_wave_batch(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *seq_concat,
                                      const uint64_t *seq_off, int32_t size, int32_t step, uint64_t lag,
                                      float threshold, float influence, const uint64_t *win_off, uint32_t *gc_count,
                                      float *gc, int8_t *signal) {
    ARG_CHECK(ctx, "wave_batch: ctx is NULL");
    ARG_CHECK(n_ctg >= 0 && seq_off && win_off, "wave_batch: NULL offsets");
    ARG_CHECK(size > 0 && step > 0, "wave_batch: size and step must be positive (got %d, %d)", size, step);
    if (n_ctg == 0) return GAMS_CUDA_OK;
    ARG_CHECK(seq_concat || seq_off[n_ctg] == seq_off[0], "wave_batch: seq_concat is NULL");
    ARG_CHECK(gc || win_off[n_ctg] == 0, "wave_batch: gc is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int n_dev = (int)std::min<int64_t>((int64_t)ctx->devs.size(), n_ctg);
    std::vector<int64_t> part((size_t)n_dev + 1);
    int32_t rc = gams_cuda_partition(n_ctg, seq_off, n_dev, part.data());
    if (rc) return rc;
    return for_each_device(ctx, n_dev, [&](int d) {
        return wave_on_device(ctx, d, part[d], part[d + 1], seq_concat, seq_off, size, step, lag, threshold, influence,
                              win_off, gc_count, gc, signal);
    });
}

// Conn::get_seq + wave::proc_ctg in one call (src/libs/redis.rs:142-161, src/cmd_gams/wave.rs:138-155): the ctgs arrive as the
// gzip blobs Redis holds under seq:{ctg}; they are inflated on the device and go straight into the GC-wave kernels.
GAMS_API int32_t gams_cuda_wave_gz_batch(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *gz_concat, const uint64_t *gz_off,
                                         const uint64_t *seq_off, int32_t size, int32_t step, uint64_t lag, float threshold,
                                         float influence, const uint64_t *win_off, uint32_t *gc_count, float *gc, int8_t *signal,
                                         int32_t *status) {
    ARG_CHECK(ctx, "wave_gz_batch: ctx is NULL");
    ARG_CHECK(n_ctg >= 0 && gz_off && seq_off && win_off, "wave_gz_batch: NULL offsets");
    ARG_CHECK(size > 0 && step > 0, "wave_gz_batch: size and step must be positive (got %d, %d)", size, step);
    if (n_ctg == 0) return GAMS_CUDA_OK;
    ARG_CHECK(gz_concat && status, "wave_gz_batch: NULL blobs / status");
    ARG_CHECK(gc || win_off[n_ctg] == 0, "wave_gz_batch: gc is NULL");
    for (int64_t c = 0; c < n_ctg; ++c)
        ARG_CHECK(gz_off[c + 1] >= gz_off[c], "wave_gz_batch: gz_off decreases at %lld", (long long)c);
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int n_dev = (int)std::min<int64_t>((int64_t)ctx->devs.size(), n_ctg);
    std::vector<int64_t> part((size_t)n_dev + 1);
    int32_t rc = gams_cuda_partition(n_ctg, seq_off, n_dev, part.data());
    if (rc) return rc;
    GzSource src;
    src.gz_concat = gz_concat, src.gz_off = gz_off, src.status = status;
    rc = for_each_device(ctx, n_dev, [&](int d) {
        return wave_on_device(ctx, d, part[d], part[d + 1], nullptr, seq_off, size, step, lag, threshold, influence, win_off, gc_count, gc,
                              signal, nullptr, &src);
    });
    if (rc) return rc;
    for (int64_t c = 0; c < n_ctg; ++c)
        if (status[c] != 0) {
            set_error("wave_gz_batch: seq blob %lld failed with inflate status %d (its windows are undefined)", (long long)c, (int)status[c]);
            return GAMS_CUDA_EINVAL;
        }
    return GAMS_CUDA_OK;
}

// gams_cuda_wave_peaks_batch / gams_cuda_wave_peaks_gz_batch: the bases, or the gzip blobs that hold them (gzs)
static int32_t wave_peaks_impl(gams_cuda_ctx *ctx, const char *who, int64_t n_ctg, const uint8_t *seq_concat, GzSource *gzs,
                               const uint64_t *seq_off, int32_t size, int32_t step, uint64_t lag, float threshold, float influence,
                               float coverage, const uint64_t *win_off, uint64_t cap, uint32_t *pk_ctg, uint32_t *pk_first,
                               uint32_t *pk_last, float *pk_gc, int8_t *pk_signal, uint64_t *n_peaks) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int n_dev = (int)std::min<int64_t>((int64_t)ctx->devs.size(), n_ctg);
    std::vector<int64_t> part((size_t)n_dev + 1);
    int32_t rc = gams_cuda_partition(n_ctg, seq_off, n_dev, part.data());
    if (rc) return rc;
    std::vector<PeakSink> sinks((size_t)n_dev);
    for (auto &sk : sinks) sk.coverage = coverage;
    if (n_dev == 1) {
        PeakSink &sk = sinks[0];
        sk.cap = cap, sk.o_ctg = pk_ctg, sk.o_first = pk_first, sk.o_last = pk_last, sk.o_gc = pk_gc, sk.o_sig = pk_signal;
        if (!sk.o_ctg) sk.o_ctg = reinterpret_cast<uint32_t *>(&sk.n), sk.cap = 0;   // cap == 0: count only
    }
    rc = for_each_device(ctx, n_dev, [&](int d) {
        return wave_on_device(ctx, d, part[d], part[d + 1], seq_concat, seq_off, size, step, lag, threshold, influence, win_off,
                              nullptr, nullptr, nullptr, &sinks[d], gzs);
    });
    if (rc) return rc;
    if (gzs)
        for (int64_t c = 0; c < n_ctg; ++c)
            if (gzs->status[c] != 0) {
                set_error("%s: seq blob %lld failed with inflate status %d (its peaks are undefined)", who, (long long)c, (int)gzs->status[c]);
                return GAMS_CUDA_EINVAL;
            }
    uint64_t total = 0;
    for (auto &sk : sinks) total += sk.n;
    *n_peaks = total;
    if (total > cap) {
        set_error("%s: %llu peak records, room for %llu (n_peaks is set: call again with larger arrays)", who,
                  (unsigned long long)total, (unsigned long long)cap);
        return GAMS_CUDA_ECAP;
    }
    if (n_dev == 1) return GAMS_CUDA_OK;   // already in the caller's arrays (ctg offset of the only shard is 0)
    uint64_t o = 0;
    for (int d = 0; d < n_dev; ++d) {
        const PeakSink &sk = sinks[d];
        for (size_t i = 0; i < sk.ctg.size(); ++i, ++o) {
            pk_ctg[o] = sk.ctg[i] + (uint32_t)part[d];   // shards are contiguous ctg ranges: ctg order is kept
            pk_first[o] = sk.first[i];
            pk_last[o] = sk.last[i];
            pk_gc[o] = sk.gc[i];
            pk_signal[o] = sk.sig[i];
        }
    }
    return GAMS_CUDA_OK;
}

GAMS_API int32_t gams_cuda_wave_peaks_batch(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *seq_concat,
                                            const uint64_t *seq_off, int32_t size, int32_t step, uint64_t lag,
                                            float threshold, float influence, float coverage, const uint64_t *win_off,
                                            uint64_t cap, uint32_t *pk_ctg, uint32_t *pk_first, uint32_t *pk_last,
                                            float *pk_gc, int8_t *pk_signal, uint64_t *n_peaks) {
    ARG_CHECK(ctx && n_peaks, "wave_peaks_batch: NULL argument");
    ARG_CHECK(n_ctg >= 0 && seq_off && win_off, "wave_peaks_batch: NULL offsets");
    ARG_CHECK(size > 0 && step > 0, "wave_peaks_batch: size and step must be positive (got %d, %d)", size, step);
    ARG_CHECK(cap == 0 || (pk_ctg && pk_first && pk_last && pk_gc && pk_signal), "wave_peaks_batch: NULL output");
    *n_peaks = 0;
    if (n_ctg == 0) return GAMS_CUDA_OK;
    ARG_CHECK(seq_concat || seq_off[n_ctg] == seq_off[0], "wave_peaks_batch: seq_concat is NULL");
    return wave_peaks_impl(ctx, "wave_peaks_batch", n_ctg, seq_concat, nullptr, seq_off, size, step, lag, threshold, influence, coverage,
                           win_off, cap, pk_ctg, pk_first, pk_last, pk_gc, pk_signal, n_peaks);
}

// `gams wave` without --signal from the seq:{ctg} values as Redis holds them: Conn::get_seq (redis.rs:142-161) + proc_ctg
// (wave.rs:138-212) + merge_ints (wave.rs:217-252); a third of the bytes go up, a few MB of peak records come back.
GAMS_API int32_t gams_cuda_wave_peaks_gz_batch(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *gz_concat, const uint64_t *gz_off,
                                               const uint64_t *seq_off, int32_t size, int32_t step, uint64_t lag, float threshold,
                                               float influence, float coverage, const uint64_t *win_off, uint64_t cap,
                                               uint32_t *pk_ctg, uint32_t *pk_first, uint32_t *pk_last, float *pk_gc,
                                               int8_t *pk_signal, uint64_t *n_peaks, int32_t *status) {
    ARG_CHECK(ctx && n_peaks, "wave_peaks_gz_batch: NULL argument");
    ARG_CHECK(n_ctg >= 0 && gz_off && seq_off && win_off, "wave_peaks_gz_batch: NULL offsets");
    ARG_CHECK(size > 0 && step > 0, "wave_peaks_gz_batch: size and step must be positive (got %d, %d)", size, step);
    ARG_CHECK(cap == 0 || (pk_ctg && pk_first && pk_last && pk_gc && pk_signal), "wave_peaks_gz_batch: NULL output");
    *n_peaks = 0;
    if (n_ctg == 0) return GAMS_CUDA_OK;
    ARG_CHECK(gz_concat && status, "wave_peaks_gz_batch: NULL blobs / status");
    for (int64_t c = 0; c < n_ctg; ++c)
        ARG_CHECK(gz_off[c + 1] >= gz_off[c], "wave_peaks_gz_batch: gz_off decreases at %lld", (long long)c);
    GzSource src;
    src.gz_concat = gz_concat, src.gz_off = gz_off, src.status = status;
    return wave_peaks_impl(ctx, "wave_peaks_gz_batch", n_ctg, nullptr, &src, seq_off, size, step, lag, threshold, influence, coverage,
                           win_off, cap, pk_ctg, pk_first, pk_last, pk_gc, pk_signal, n_peaks);
}

GAMS_API int32_t gams_cuda_thresholding_batch(gams_cuda_ctx *ctx, int64_t n_series, const float *data,
                                              const uint64_t *off, uint64_t lag, float threshold, float influence,
                                              int8_t *signal) {
    ARG_CHECK(ctx && off, "thresholding_batch: NULL argument");
    ARG_CHECK(n_series >= 0 && n_series < (1ll << 32), "thresholding_batch: n_series out of range");
    if (n_series == 0) return GAMS_CUDA_OK;
    ARG_CHECK(lag >= 1, "thresholding_batch: lag must be >= 1 (the reference underflows lag-1, stat.rs:30)");
    const uint64_t total = off[n_series] - off[0];
    ARG_CHECK((data && signal) || total == 0, "thresholding_batch: NULL data/signal");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PerDevice &pd = ctx->devs[0];
    DeviceGuard g(pd.device);
    SeriesLayout L;
    L.cache = pd.cache;
    int32_t rc = L.build(off, n_series, pd.stream);
    struct Rel { SeriesLayout &l; ~Rel() { l.release(); } } rel{L};
    if (rc) return rc;
    if ((rc = pd.scratch[1].reserve(std::max<uint64_t>(total, 1) * sizeof(float)))) return rc;
    if ((rc = pd.scratch[3].reserve(std::max<uint64_t>(total, 1)))) return rc;
    if ((rc = pd.scratch[4].reserve(std::max<uint64_t>(total, 1) * sizeof(float)))) return rc;
    float *d_x = pd.scratch[1].as<float>();
    int8_t *d_sig = pd.scratch[3].as<int8_t>();
    cudaStream_t st = pd.stream;
    CUDA_TRY(cudaMemcpyAsync(d_x, data + off[0], total * sizeof(float), cudaMemcpyHostToDevice, st));
    if ((rc = launch_zscore(L, st, d_x, lag, threshold, influence, pd.scratch[4].as<float>(), d_sig, nullptr,
                            ctx->zscore_mode, pd.d_zstats, 0, n_series)))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(signal + off[0], d_sig, total, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return GAMS_CUDA_OK;
}

// ============================================================================ overlap counting ==
GAMS_API int32_t gams_cuda_count_plan_create(gams_cuda_ctx *ctx, int32_t dev_index, int64_t n_seg,
                                             const uint64_t *iv_off, gams_cuda_count_plan **out) {
    ARG_CHECK(ctx && out && iv_off, "count_plan_create: NULL argument");
    ARG_CHECK(dev_index >= 0 && dev_index < (int32_t)ctx->devs.size(), "count_plan_create: dev_index %d out of range",
              dev_index);
    ARG_CHECK(n_seg >= 0 && n_seg < (1ll << 31), "count_plan_create: n_seg out of range");
    *out = nullptr;
    std::unique_ptr<gams_cuda_count_plan, void (*)(gams_cuda_count_plan *)> p(new gams_cuda_count_plan,
                                                                              gams_cuda_count_plan_destroy);
    PerDevice &pd = ctx->devs[dev_index];
    p->ctx = ctx;
    p->dev_index = dev_index;
    p->device = pd.device;
    p->n_seg = n_seg;
    p->h_iv_off.resize((size_t)n_seg + 1);
    for (int64_t s = 0; s <= n_seg; ++s) p->h_iv_off[s] = iv_off[s] - iv_off[0];
    for (int64_t s = 0; s < n_seg; ++s) {
        ARG_CHECK(iv_off[s + 1] >= iv_off[s], "count_plan_create: iv_off not monotone at %lld", (long long)s);
        const uint64_t len = iv_off[s + 1] - iv_off[s];
        ARG_CHECK(len < (1ull << 32), "count_plan_create: segment %lld has more than 2^32-1 intervals", (long long)s);
        p->max_seg = std::max(p->max_seg, len);
    }
    p->total_iv = p->h_iv_off[n_seg];
    if (const char *m = getenv("GAMS_CUDA_COUNT_MODE")) p->mode = atoi(m);
    DeviceGuard g(pd.device);
    for (auto &b : p->tmp) b.cache = pd.cache;
    int32_t arc = dev_alloc(pd.cache, reinterpret_cast<void **>(&p->d_iv_off), sizeof(uint64_t) * (size_t)(n_seg + 1));
    if (arc) return arc;
    CUDA_TRY(cudaMemcpyAsync(p->d_iv_off, p->h_iv_off.data(), sizeof(uint64_t) * (size_t)(n_seg + 1),
                             cudaMemcpyHostToDevice, pd.stream));
    CUDA_TRY(cudaStreamSynchronize(pd.stream));
    *out = p.release();
    return GAMS_CUDA_OK;
}

GAMS_API void gams_cuda_count_plan_destroy(gams_cuda_count_plan *p) {
    if (!p) return;
    DeviceGuard g(p->device);
    cudaDeviceSynchronize();   // as cudaFree would
    dev_free(p->ctx ? p->ctx->devs[p->dev_index].cache : nullptr, p->d_iv_off, sizeof(uint64_t) * (size_t)(p->n_seg + 1));
    for (auto &b : p->tmp) b.release();
    delete p;
}

GAMS_API uint64_t gams_cuda_count_plan_launches(const gams_cuda_count_plan *p) { return p ? p->launches : 0; }

GAMS_API void gams_cuda_count_plan_invalidate(gams_cuda_count_plan *p) {
    if (p) p->idx_valid = p->cidx_valid = p->part_valid = false;
}

GAMS_API int32_t gams_cuda_count_overlaps_dev(gams_cuda_count_plan *p, void *stream, const uint32_t *d_starts,
                                              const uint32_t *d_stops, int64_t n_q, const int32_t *d_q_seg,
                                              const uint32_t *d_q_start, const uint32_t *d_q_stop, int32_t grouped,
                                              uint32_t *d_out) {
    ARG_CHECK(p && n_q >= 0, "count_overlaps_dev: bad argument");
    if (n_q == 0) return GAMS_CUDA_OK;
    ARG_CHECK(d_q_seg && d_q_start && d_q_stop && d_out, "count_overlaps_dev: NULL query/output pointer");
    ARG_CHECK((d_starts && d_stops) || p->total_iv == 0, "count_overlaps_dev: NULL index pointer");
    DeviceGuard g(p->device);
    return launch_count_overlaps(p, (cudaStream_t)stream, d_starts, d_stops, n_q, d_q_seg, d_q_start, d_q_stop, grouped,
                                 d_out);
}

GAMS_API int32_t gams_cuda_count_overlaps(gams_cuda_ctx *ctx, int64_t n_seg, const uint32_t *starts,
                                          const uint32_t *stops, const uint64_t *iv_off, int64_t n_q,
                                          const int32_t *q_seg, const uint32_t *q_start, const uint32_t *q_stop,
                                          uint32_t *out) {
    ARG_CHECK(ctx && iv_off && n_seg >= 0 && n_q >= 0, "count_overlaps: bad argument");
    if (n_q == 0) return GAMS_CUDA_OK;
    ARG_CHECK(q_seg && q_start && q_stop && out, "count_overlaps: NULL query/output pointer");
    const uint64_t n_iv = iv_off[n_seg] - iv_off[0];
    ARG_CHECK((starts && stops) || n_iv == 0, "count_overlaps: NULL index pointer");
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int n_dev = (int)std::min<int64_t>((int64_t)ctx->devs.size(), std::max<int64_t>(1, n_q / 65536));
    // Several devices.  Queries grouped by ctg (q_seg non-decreasing: the order read_range's BTreeMap yields, utils.rs:39-79)
    // shard BY CTG: device d gets a contiguous range of segments, their slice of the index and exactly the queries that name
    // them — nothing is replicated and nothing is exchanged.  Queries in arbitrary order are split evenly and the
    // index is replicated: it is small against 180 GB of HBM, and routing 10^8 queries to the owners of their ctgs would be
    // an all-to-all the path does not otherwise need.
    bool grouped = false;
    std::vector<int64_t> qb((size_t)n_dev + 1, 0), sb((size_t)n_dev + 1, 0);
    if (n_dev > 1) {
        std::vector<int> ok((size_t)n_dev, 1);
        run_on_devices(n_dev, [&](int d) -> int32_t {   // (host threads only: the order check of one slice each)
            const int64_t q0 = n_q * d / n_dev, q1 = n_q * (d + 1) / n_dev;
            int good = 1;
            for (int64_t q = std::max<int64_t>(q0, 1); q < q1 && good; ++q) good = q_seg[q - 1] <= q_seg[q];
            ok[d] = good;
            return GAMS_CUDA_OK;
        });
        grouped = std::all_of(ok.begin(), ok.end(), [](int v) { return v != 0; });
    }
    if (grouped) {
        auto clamp_seg = [&](int32_t v) { return (int64_t)std::min<int64_t>(std::max<int64_t>(v, 0), n_seg); };
        qb[n_dev] = n_q, sb[n_dev] = n_seg;
        for (int d = 1; d < n_dev; ++d) {
            const int64_t q = n_q * d / n_dev;
            const int32_t sv = q_seg[q];
            qb[d] = std::max<int64_t>(qb[d - 1], std::lower_bound(q_seg, q_seg + n_q, sv) - q_seg);   // first query of that segment
            sb[d] = std::max<int64_t>(sb[d - 1], clamp_seg(sv));
        }
    }
    return for_each_device(ctx, n_dev, [&](int d) -> int32_t {
        const int64_t q0 = grouped ? qb[d] : n_q * d / n_dev, q1 = grouped ? qb[d + 1] : n_q * (d + 1) / n_dev, nq = q1 - q0;
        if (nq == 0) return GAMS_CUDA_OK;
        PerDevice &pd = ctx->devs[d];
        DeviceGuard g(pd.device);
        gams_cuda_count_plan *plan = nullptr;
        trace_stage("count: enter");
        // this device's view of the index: every segment outside [s_lo, s_hi) is empty
        const int64_t s_lo = grouped ? sb[d] : 0, s_hi = grouped ? sb[d + 1] : n_seg;
        const uint64_t iv_lo = iv_off[s_lo], iv_hi = iv_off[s_hi], n_iv_d = iv_hi - iv_lo;
        std::vector<uint64_t> view;
        const uint64_t *my_off = iv_off;
        if (grouped) {
            view.resize((size_t)n_seg + 1);
            for (int64_t s = 0; s <= n_seg; ++s) view[s] = std::min(std::max(iv_off[s], iv_lo), iv_hi) - iv_lo;
            my_off = view.data();
        }
        int32_t rc = gams_cuda_count_plan_create(ctx, d, n_seg, my_off, &plan);
        if (rc) return rc;
        trace_stage("count: plan_create");
        std::unique_ptr<gams_cuda_count_plan, void (*)(gams_cuda_count_plan *)> guard(plan, gams_cuda_count_plan_destroy);
        if ((rc = pd.scratch[0].reserve(std::max<uint64_t>(n_iv_d, 1) * 4))) return rc;
        if ((rc = pd.scratch[1].reserve(std::max<uint64_t>(n_iv_d, 1) * 4))) return rc;
        if ((rc = pd.scratch[2].reserve((uint64_t)nq * 4))) return rc;
        if ((rc = pd.scratch[3].reserve((uint64_t)nq * 4))) return rc;
        if ((rc = pd.scratch[4].reserve((uint64_t)nq * 4))) return rc;
        if ((rc = pd.scratch[5].reserve((uint64_t)nq * 4))) return rc;
        cudaStream_t st = pd.stream;
        uint32_t *d_st = pd.scratch[0].as<uint32_t>(), *d_sp = pd.scratch[1].as<uint32_t>();
        if (n_iv_d) {
            CUDA_TRY(cudaMemcpyAsync(d_st, starts + iv_lo, n_iv_d * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(d_sp, stops + iv_lo, n_iv_d * 4, cudaMemcpyHostToDevice, st));
        }
        trace_stage("count: reserve + index H2D", st, true);
        // (the index tables - compact form or rank nodes - are built by the first launch that needs them)
        cudaEvent_t ev_idx;
        CUDA_TRY(cudaEventCreateWithFlags(&ev_idx, cudaEventDisableTiming));
        CUDA_TRY(cudaEventRecord(ev_idx, st));
        CUDA_TRY(cudaStreamWaitEvent(pd.copy_in, ev_idx, 0));  // keeps the slices behind the index on the link
        // Query slices: H2D of slice k+1, kernel of slice k and D2H of slice k-1 overlap (three streams).
        const int64_t slice = 8 << 20;
        const size_t n_sl = (size_t)((nq + slice - 1) / slice);
        std::vector<cudaEvent_t> ev_in(n_sl), ev_k(n_sl);
        for (size_t k = 0; k < n_sl; ++k) {
            CUDA_TRY(cudaEventCreateWithFlags(&ev_in[k], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&ev_k[k], cudaEventDisableTiming));
        }
        struct EvGuard {
            std::vector<cudaEvent_t> &a, &b;
            cudaEvent_t c;
            ~EvGuard() {
                for (auto e : a) cudaEventDestroy(e);
                for (auto e : b) cudaEventDestroy(e);
                cudaEventDestroy(c);
            }
        } ev_guard{ev_in, ev_k, ev_idx};
        int32_t *d_qseg = pd.scratch[2].as<int32_t>();
        uint32_t *d_qs = pd.scratch[3].as<uint32_t>(), *d_qe = pd.scratch[4].as<uint32_t>(), *d_o = pd.scratch[5].as<uint32_t>();
        for (size_t k = 0; k < n_sl; ++k) {
            const int64_t a = (int64_t)k * slice, m = std::min<int64_t>(slice, nq - a);
            CUDA_TRY(cudaMemcpyAsync(d_qseg + a, q_seg + q0 + a, (uint64_t)m * 4, cudaMemcpyHostToDevice, pd.copy_in));
            CUDA_TRY(cudaMemcpyAsync(d_qs + a, q_start + q0 + a, (uint64_t)m * 4, cudaMemcpyHostToDevice, pd.copy_in));
            CUDA_TRY(cudaMemcpyAsync(d_qe + a, q_stop + q0 + a, (uint64_t)m * 4, cudaMemcpyHostToDevice, pd.copy_in));
            CUDA_TRY(cudaEventRecord(ev_in[k], pd.copy_in));
            CUDA_TRY(cudaStreamWaitEvent(st, ev_in[k], 0));
            if ((rc = launch_count_overlaps(plan, st, d_st, d_sp, m, d_qseg + a, d_qs + a, d_qe + a, grouped ? 1 : 0, d_o + a))) return rc;
            CUDA_TRY(cudaEventRecord(ev_k[k], st));
            CUDA_TRY(cudaStreamWaitEvent(pd.copy_out, ev_k[k], 0));
            CUDA_TRY(cudaMemcpyAsync(out + q0 + a, d_o + a, (uint64_t)m * 4, cudaMemcpyDeviceToHost, pd.copy_out));
        }
        CUDA_TRY(cudaStreamSynchronize(pd.copy_out));
        CUDA_TRY(cudaStreamSynchronize(st));
        trace_stage("count: slices");
        guard.reset();
        trace_stage("count: plan_destroy");
        return GAMS_CUDA_OK;
    });
}

GAMS_API int32_t gams_cuda_locate_ctg(gams_cuda_ctx *ctx, int64_t n_chr, const uint32_t *ctg_start,
                                      const uint32_t *ctg_stop, const uint64_t *chr_off, int64_t n_q,
                                      const int32_t *q_chr, const uint32_t *q_start, const uint32_t *q_stop,
                                      int32_t *out) {
    ARG_CHECK(ctx && chr_off && n_chr >= 0 && n_q >= 0, "locate_ctg: bad argument");
    if (n_q == 0) return GAMS_CUDA_OK;
    ARG_CHECK(q_chr && q_start && q_stop && out, "locate_ctg: NULL query/output pointer");
    ARG_CHECK(chr_off[0] == 0, "locate_ctg: chr_off[0] must be 0");
    const uint64_t n_ctg = chr_off[n_chr];
    ARG_CHECK((ctg_start && ctg_stop) || n_ctg == 0, "locate_ctg: NULL index pointer");
    ARG_CHECK(n_ctg < (1ull << 31), "locate_ctg: too many ctgs");
    std::vector<uint32_t> maxlen((size_t)std::max<int64_t>(n_chr, 1), 0);  // Lapper::new: max(stop - start)
    for (int64_t c = 0; c < n_chr; ++c)
        for (uint64_t k = chr_off[c]; k < chr_off[c + 1]; ++k)
            if (ctg_stop[k] > ctg_start[k]) maxlen[c] = std::max(maxlen[c], ctg_stop[k] - ctg_start[k]);
    std::lock_guard<std::mutex> lk(ctx->mu);
    PerDevice &pd = ctx->devs[0];
    DeviceGuard g(pd.device);
    int32_t rc;
    const uint64_t meta = std::max<uint64_t>(n_ctg, 1) * 8 + (uint64_t)(n_chr + 1) * 8 + (uint64_t)std::max<int64_t>(n_chr, 1) * 4;
    if ((rc = pd.scratch[0].reserve(meta))) return rc;
    if ((rc = pd.scratch[2].reserve((uint64_t)n_q * 4))) return rc;
    if ((rc = pd.scratch[3].reserve((uint64_t)n_q * 4))) return rc;
    if ((rc = pd.scratch[4].reserve((uint64_t)n_q * 4))) return rc;
    if ((rc = pd.scratch[5].reserve((uint64_t)n_q * 4))) return rc;
    cudaStream_t st = pd.stream;
    uint8_t *m = pd.scratch[0].as<uint8_t>();
    uint64_t *d_chr_off = reinterpret_cast<uint64_t *>(m);
    uint32_t *d_cs = reinterpret_cast<uint32_t *>(m + (uint64_t)(n_chr + 1) * 8);
    uint32_t *d_ce = d_cs + std::max<uint64_t>(n_ctg, 1);
    uint32_t *d_ml = d_ce + std::max<uint64_t>(n_ctg, 1);
    CUDA_TRY(cudaMemcpyAsync(d_chr_off, chr_off, (uint64_t)(n_chr + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_ctg) {
        CUDA_TRY(cudaMemcpyAsync(d_cs, ctg_start, n_ctg * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_ce, ctg_stop, n_ctg * 4, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemcpyAsync(d_ml, maxlen.data(), maxlen.size() * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(pd.scratch[2].p, q_chr, (uint64_t)n_q * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(pd.scratch[3].p, q_start, (uint64_t)n_q * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(pd.scratch[4].p, q_stop, (uint64_t)n_q * 4, cudaMemcpyHostToDevice, st));
    if ((rc = launch_locate_ctg(st, n_chr, d_cs, d_ce, d_chr_off, d_ml, n_q, pd.scratch[2].as<int32_t>(),
                                pd.scratch[3].as<uint32_t>(), pd.scratch[4].as<uint32_t>(), pd.scratch[5].as<int32_t>())))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(out, pd.scratch[5].p, (uint64_t)n_q * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return GAMS_CUDA_OK;
}

// ============================================================================ GC of ranges / sw ==
static int32_t upload_seq(PerDevice &pd, int64_t n_ctg, const uint8_t *seq_concat, const uint64_t *seq_off,
                          uint8_t **d_seq, uint64_t **d_off) {
    const uint64_t nb = seq_off[n_ctg] - seq_off[0];
    int32_t rc;
    if ((rc = pd.scratch[0].reserve(((nb + 31) & ~31ull) + 32))) return rc;
    if ((rc = pd.scratch[6].reserve((uint64_t)(n_ctg + 1) * 8))) return rc;
    std::vector<uint64_t> rel((size_t)n_ctg + 1);
    for (int64_t c = 0; c <= n_ctg; ++c) {
        if (c && seq_off[c] < seq_off[c - 1]) {
            set_error("seq_off not monotone at %lld", (long long)c);
            return GAMS_CUDA_EINVAL;
        }
        rel[c] = seq_off[c] - seq_off[0];
    }
    *d_seq = pd.scratch[0].as<uint8_t>();
    *d_off = pd.scratch[6].as<uint64_t>();
    if (nb) CUDA_TRY(cudaMemcpyAsync(*d_seq, seq_concat + seq_off[0], nb, cudaMemcpyHostToDevice, pd.stream));
    CUDA_TRY(cudaMemcpyAsync(*d_off, rel.data(), (uint64_t)(n_ctg + 1) * 8, cudaMemcpyHostToDevice, pd.stream));
    CUDA_TRY(cudaStreamSynchronize(pd.stream));  // rel goes out of scope
    return GAMS_CUDA_OK;
}

// ================================================================================ anno ==========
GAMS_API int32_t gams_cuda_anno_coverage(gams_cuda_ctx *ctx, int64_t n_chr, const int32_t *run_start, const int32_t *run_end,
                                         const uint64_t *run_off, int64_t n_q, const int32_t *q_chr, const int32_t *q_ctg_start,
                                         const int32_t *q_ctg_end, const int32_t *q_start, const int32_t *q_end, float *prop) {
    ARG_CHECK(ctx && run_off && n_chr >= 0 && n_q >= 0, "anno_coverage: bad argument");
    if (n_q == 0) return GAMS_CUDA_OK;
    ARG_CHECK(q_chr && q_ctg_start && q_ctg_end && q_start && q_end && prop, "anno_coverage: NULL query/output pointer");
    ARG_CHECK(run_off[0] == 0, "anno_coverage: run_off[0] must be 0");
    const uint64_t n_run = run_off[n_chr];
    ARG_CHECK((run_start && run_end) || n_run == 0, "anno_coverage: NULL run arrays");
    std::vector<int64_t> cum((size_t)std::max<uint64_t>(n_run, 1), 0);   // covered positions before each run of its chromosome
    for (int64_t c = 0; c < n_chr; ++c) {
        ARG_CHECK(run_off[c + 1] >= run_off[c] && run_off[c + 1] - run_off[c] < (1ull << 32), "anno_coverage: bad run_off at %lld", (long long)c);
        int64_t acc = 0;
        for (uint64_t k = run_off[c]; k < run_off[c + 1]; ++k) {
            ARG_CHECK(run_end[k] >= run_start[k] && (k == run_off[c] || run_start[k] > run_end[k - 1]),
                      "anno_coverage: runs of chromosome %lld must be sorted and disjoint (run %llu)", (long long)c, (unsigned long long)k);
            cum[k] = acc;
            acc += (int64_t)run_end[k] - run_start[k] + 1;
        }
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    PerDevice &pd = ctx->devs[0];
    DeviceGuard g(pd.device);
    int32_t rc;
    const uint64_t nr = std::max<uint64_t>(n_run, 1), nq = (uint64_t)n_q;
    if ((rc = pd.scratch[1].reserve(nr * 16 + (uint64_t)(n_chr + 1) * 8 + 64))) return rc;
    if ((rc = pd.scratch[2].reserve(nq * 24 + 64))) return rc;
    cudaStream_t st = pd.stream;
    uint8_t *ib = pd.scratch[1].as<uint8_t>();
    int64_t *d_cum = reinterpret_cast<int64_t *>(ib);
    uint64_t *d_roff = reinterpret_cast<uint64_t *>(d_cum + nr);
    int32_t *d_rs = reinterpret_cast<int32_t *>(d_roff + (n_chr + 1)), *d_re = d_rs + nr;
    CUDA_TRY(cudaMemcpyAsync(d_cum, cum.data(), nr * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_roff, run_off, (uint64_t)(n_chr + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_run) {
        CUDA_TRY(cudaMemcpyAsync(d_rs, run_start, n_run * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_re, run_end, n_run * 4, cudaMemcpyHostToDevice, st));
    }
    int32_t *d_q = pd.scratch[2].as<int32_t>();
    const int32_t *src[5] = {q_chr, q_ctg_start, q_ctg_end, q_start, q_end};
    for (int k = 0; k < 5; ++k) CUDA_TRY(cudaMemcpyAsync(d_q + (uint64_t)k * nq, src[k], nq * 4, cudaMemcpyHostToDevice, st));
    float *d_p = reinterpret_cast<float *>(d_q + 5 * nq);
    if ((rc = launch_anno_coverage(st, d_rs, d_re, d_cum, d_roff, n_chr, n_q, d_q, d_q + nq, d_q + 2 * nq, d_q + 3 * nq, d_q + 4 * nq, d_p)))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(prop, d_p, nq * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));   // cum is a local
    return GAMS_CUDA_OK;
}

// ================================================================================ gen ===========
GAMS_API int32_t gams_cuda_gen_ctgs(gams_cuda_ctx *ctx, int64_t n_chr, const uint8_t *seq_concat, const uint64_t *chr_off,
                                    int32_t fill, int32_t min_len, int32_t piece, uint64_t cap, uint32_t *ctg_chr,
                                    int32_t *ctg_start, int32_t *ctg_end, uint64_t *n_ctg) {
    ARG_CHECK(ctx && chr_off && n_ctg && n_chr >= 0, "gen_ctgs: bad argument");
    ARG_CHECK(fill >= 1 && min_len >= 0 && piece >= 1, "gen_ctgs: fill and piece must be >= 1, min >= 0");
    ARG_CHECK(cap == 0 || (ctg_chr && ctg_start && ctg_end), "gen_ctgs: NULL output");
    *n_ctg = 0;
    if (n_chr == 0) return GAMS_CUDA_OK;
    const uint64_t total = chr_off[n_chr] - chr_off[0];
    ARG_CHECK(seq_concat || total == 0, "gen_ctgs: seq_concat is NULL");
    for (int64_t c = 0; c < n_chr; ++c)
        ARG_CHECK(chr_off[c + 1] >= chr_off[c] && chr_off[c + 1] - chr_off[c] < (1ull << 31), "gen_ctgs: chromosome %lld: bad length",
                  (long long)c);
    std::vector<uint64_t> rs, re;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        PerDevice &pd = ctx->devs[0];
        DeviceGuard g(pd.device);
        uint8_t *d_seq;
        uint64_t *d_off;
        int32_t rc = upload_seq(pd, n_chr, seq_concat, chr_off, &d_seq, &d_off);
        if (rc) return rc;
        const uint64_t tiles = (total + 8191) / 8192;
        if ((rc = pd.scratch[1].reserve(sizeof(uint64_t) * (4 * (tiles + 1) + 2)))) return rc;
        uint64_t *d_tmp = pd.scratch[1].as<uint64_t>(), *d_n = d_tmp + 4 * (tiles + 1);
        uint64_t rcap = std::max<uint64_t>(total / 4096, 4096), n_runs = 0;
        for (int attempt = 0; attempt < 2; ++attempt) {   // the first attempt sizes the run arrays
            if ((rc = pd.scratch[2].reserve(rcap * 16))) return rc;
            uint64_t *d_rs = pd.scratch[2].as<uint64_t>(), *d_re = d_rs + rcap;
            if ((rc = launch_valid_runs(pd.stream, d_seq, total, d_off, n_chr, d_tmp, rcap, d_rs, d_re, d_n))) return rc;
            CUDA_TRY(cudaMemcpyAsync(&n_runs, d_n, sizeof(uint64_t), cudaMemcpyDeviceToHost, pd.stream));
            CUDA_TRY(cudaStreamSynchronize(pd.stream));
            if (n_runs <= rcap) {
                rs.resize(n_runs), re.resize(n_runs);
                if (n_runs) {
                    CUDA_TRY(cudaMemcpyAsync(rs.data(), d_rs, n_runs * 8, cudaMemcpyDeviceToHost, pd.stream));
                    CUDA_TRY(cudaMemcpyAsync(re.data(), d_re, n_runs * 8, cudaMemcpyDeviceToHost, pd.stream));
                    CUDA_TRY(cudaStreamSynchronize(pd.stream));
                }
                break;
            }
            rcap = n_runs;
        }
    }
    // gen.rs:102-126 on the run list: fill(fill - 1), excise(min), pieces with the tail merged into the last one
    uint64_t n_out = 0;
    auto emit = [&](uint32_t c, int64_t s, int64_t e) {
        if (n_out < cap) {
            ctg_chr[n_out] = c;
            ctg_start[n_out] = (int32_t)s;
            ctg_end[n_out] = (int32_t)e;
        }
        ++n_out;
    };
    auto flush_span = [&](uint32_t c, int64_t s, int64_t e) {
        if (e - s + 1 < min_len) return;                      // excise(min), gen.rs:104
        int64_t pos = s;
        bool any = false;
        while (e - pos + 1 > piece) {                         // gen.rs:111-115
            emit(c, pos, pos + piece - 1);
            pos += piece;
            any = true;
        }
        if (!any) emit(c, pos, e);                            // gen.rs:117-119
        else if (n_out - 1 < cap) ctg_end[n_out - 1] = (int32_t)e;   // gen.rs:120-122: the remainder joins the last piece
    };
    int64_t c = 0, cur_s = 0, cur_e = -1;
    bool open = false;
    uint32_t cur_c = 0;
    for (size_t k = 0; k < rs.size(); ++k) {
        while (chr_off[c + 1] - chr_off[0] <= rs[k]) ++c;
        const int64_t s = (int64_t)(rs[k] - (chr_off[c] - chr_off[0])) + 1, e = (int64_t)(re[k] - (chr_off[c] - chr_off[0])) + 1;
        if (open && cur_c == (uint32_t)c && s - cur_e - 1 <= fill - 1) cur_e = e;   // fill
        else {
            if (open) flush_span(cur_c, cur_s, cur_e);
            open = true, cur_c = (uint32_t)c, cur_s = s, cur_e = e;
        }
    }
    if (open) flush_span(cur_c, cur_s, cur_e);
    *n_ctg = n_out;
    if (n_out > cap) {
        set_error("gen_ctgs: %llu ctgs, room for %llu (n_ctg is set: call again with larger arrays)", (unsigned long long)n_out,
                  (unsigned long long)cap);
        return GAMS_CUDA_ECAP;
    }
    return GAMS_CUDA_OK;
}

GAMS_API int32_t gams_cuda_gc_ranges(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *seq_concat,
                                     const uint64_t *seq_off, int64_t n_q, const uint32_t *q_ctg,
                                     const uint32_t *q_from0, const uint32_t *q_to0_excl, uint32_t *gc_count) {
    ARG_CHECK(ctx && seq_off && n_ctg >= 0 && n_q >= 0, "gc_ranges: bad argument");
    if (n_q == 0) return GAMS_CUDA_OK;
    ARG_CHECK(q_ctg && q_from0 && q_to0_excl && gc_count, "gc_ranges: NULL query/output pointer");
    ARG_CHECK(seq_concat || seq_off[n_ctg] == seq_off[0], "gc_ranges: seq_concat is NULL");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PerDevice &pd = ctx->devs[0];
    DeviceGuard g(pd.device);
    uint8_t *d_seq;
    uint64_t *d_off;
    int32_t rc = upload_seq(pd, n_ctg, seq_concat, seq_off, &d_seq, &d_off);
    if (rc) return rc;
    for (int i = 2; i <= 5; ++i)
        if ((rc = pd.scratch[i].reserve((uint64_t)n_q * 4))) return rc;
    cudaStream_t st = pd.stream;
    CUDA_TRY(cudaMemcpyAsync(pd.scratch[2].p, q_ctg, (uint64_t)n_q * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(pd.scratch[3].p, q_from0, (uint64_t)n_q * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(pd.scratch[4].p, q_to0_excl, (uint64_t)n_q * 4, cudaMemcpyHostToDevice, st));
    if ((rc = launch_gc_ranges(st, d_seq, d_off, n_ctg, n_q, pd.scratch[2].as<uint32_t>(), pd.scratch[3].as<uint32_t>(),
                               pd.scratch[4].as<uint32_t>(), pd.scratch[5].as<uint32_t>())))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(gc_count, pd.scratch[5].p, (uint64_t)n_q * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return GAMS_CUDA_OK;
}

GAMS_API int32_t gams_cuda_peak_link_batch(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *seq_concat,
                                           const uint64_t *seq_off, const int32_t *ctg_chr_start, int64_t n_pk,
                                           const uint32_t *p_ctg, const int32_t *p_start, const int32_t *p_end,
                                           const int8_t *p_signal, float *gc, int8_t *left_signal,
                                           int32_t *left_wave_length, float *left_amplitude, int8_t *right_signal,
                                           int32_t *right_wave_length, float *right_amplitude) {
    ARG_CHECK(ctx && seq_off && ctg_chr_start && n_ctg >= 0 && n_pk >= 0, "peak_link_batch: bad argument");
    if (n_pk == 0) return GAMS_CUDA_OK;
    ARG_CHECK(p_ctg && p_start && p_end && p_signal && gc && left_signal && left_wave_length && left_amplitude && right_signal &&
                  right_wave_length && right_amplitude,
              "peak_link_batch: NULL peak/output pointer");
    std::vector<uint32_t> from((size_t)n_pk), to((size_t)n_pk);
    std::vector<int32_t> chr_end((size_t)std::max<int64_t>(n_ctg, 1));
    for (int64_t c = 0; c < n_ctg; ++c) chr_end[c] = ctg_chr_start[c] + (int32_t)(seq_off[c + 1] - seq_off[c]) - 1;
    for (int64_t i = 0; i < n_pk; ++i) {
        const uint32_t c = p_ctg[i];
        ARG_CHECK((int64_t)c < n_ctg, "peak_link_batch: peak %lld names ctg %u of %lld", (long long)i, c, (long long)n_ctg);
        ARG_CHECK(i == 0 || p_ctg[i - 1] < c || (p_ctg[i - 1] == c && p_start[i - 1] <= p_start[i]),
                  "peak_link_batch: peaks must be sorted by (ctg, start) (peak %lld)", (long long)i);
        ARG_CHECK(p_start[i] >= ctg_chr_start[c] && p_end[i] >= p_start[i] && p_end[i] <= chr_end[c],
                  "peak_link_batch: peak %lld (%d-%d) is not inside its ctg", (long long)i, p_start[i], p_end[i]);
        from[i] = (uint32_t)(p_start[i] - ctg_chr_start[c]);      // parent.index(start) - 1 (utils.rs:151-155)
        to[i] = (uint32_t)(p_end[i] - ctg_chr_start[c] + 1);
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    PerDevice &pd = ctx->devs[0];
    DeviceGuard g(pd.device);
    uint8_t *d_seq;
    uint64_t *d_off;
    int32_t rc = upload_seq(pd, n_ctg, seq_concat, seq_off, &d_seq, &d_off);
    if (rc) return rc;
    const uint64_t np = (uint64_t)n_pk, in_bytes = np * 21 + (uint64_t)n_ctg * 8 + 64, out_bytes = np * 26 + 64;
    if ((rc = pd.scratch[1].reserve(in_bytes))) return rc;
    if ((rc = pd.scratch[2].reserve(out_bytes))) return rc;
    cudaStream_t st = pd.stream;
    uint8_t *ib = pd.scratch[1].as<uint8_t>();
    uint32_t *d_ctg = reinterpret_cast<uint32_t *>(ib), *d_from = d_ctg + np, *d_to = d_from + np, *d_cnt = d_to + np;
    int32_t *d_ps = reinterpret_cast<int32_t *>(d_cnt + np), *d_cs = d_ps + np, *d_ce = d_cs + n_ctg;
    // p_end and the signals go behind the ctg tables
    CUDA_TRY(cudaMemcpyAsync(d_ctg, p_ctg, np * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_from, from.data(), np * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_to, to.data(), np * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_ps, p_start, np * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_cs, ctg_chr_start, (uint64_t)n_ctg * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_ce, chr_end.data(), (uint64_t)n_ctg * 4, cudaMemcpyHostToDevice, st));
    if ((rc = pd.scratch[3].reserve(np * 5 + 16))) return rc;
    int32_t *d_pe = pd.scratch[3].as<int32_t>();
    int8_t *d_psig = reinterpret_cast<int8_t *>(d_pe + np);
    CUDA_TRY(cudaMemcpyAsync(d_pe, p_end, np * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_psig, p_signal, np, cudaMemcpyHostToDevice, st));
    if ((rc = launch_gc_ranges(st, d_seq, d_off, n_ctg, n_pk, d_ctg, d_from, d_to, d_cnt))) return rc;
    uint8_t *ob = pd.scratch[2].as<uint8_t>();
    float *o_gc = reinterpret_cast<float *>(ob), *o_la = o_gc + np, *o_ra = o_la + np;
    int32_t *o_ll = reinterpret_cast<int32_t *>(o_ra + np), *o_rl = o_ll + np;
    int8_t *o_ls = reinterpret_cast<int8_t *>(o_rl + np), *o_rs = o_ls + np;
    if ((rc = launch_peak_link(st, n_pk, d_ctg, d_ps, d_pe, d_psig, d_cnt, d_cs, d_ce, o_gc, o_ls, o_ll, o_la, o_rs, o_rl, o_ra)))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(gc, o_gc, np * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(left_amplitude, o_la, np * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(right_amplitude, o_ra, np * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(left_wave_length, o_ll, np * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(right_wave_length, o_rl, np * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(left_signal, o_ls, np, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(right_signal, o_rs, np, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return GAMS_CUDA_OK;
}

// one thread per feature: the ctg (= index segment) of each of its rows
__global__ void __launch_bounds__(256)
sw_row_seg_kernel(int64_t n_feat, const uint32_t *__restrict__ f_ctg, const uint64_t *__restrict__ row_off, int32_t *__restrict__ row_seg) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_feat) return;
    const int32_t c = (int32_t)f_ctg[f];
    for (uint64_t r = row_off[f]; r < row_off[f + 1]; ++r) row_seg[r] = c;
}

static int32_t sw_batch_impl(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *seq_concat,
                             const uint64_t *seq_off, const int32_t *ctg_chr_start, int64_t n_feat,
                             const uint32_t *f_ctg, const int32_t *f_start, const int32_t *f_end, int32_t size,
                             int32_t max, int32_t resize, uint64_t *row_off, int32_t *w_start, int32_t *w_end,
                             int8_t *w_type, int32_t *w_dist, float *gc, float *gc_mean, float *gc_stddev,
                             float *gc_cv, bool want_count, const uint32_t *starts, const uint32_t *stops,
                             const uint64_t *iv_off, uint32_t *rg_count) {
    ARG_CHECK(ctx && seq_off && ctg_chr_start && row_off && n_ctg >= 0 && n_feat >= 0, "sw_batch: bad argument");
    ARG_CHECK(size > 0 && max >= 0 && resize > 0, "sw_batch: size/resize must be positive and max >= 0");
    trace_stage("sw: enter");
    row_off[0] = 0;
    if (n_feat == 0) return GAMS_CUDA_OK;
    ARG_CHECK(f_ctg && f_start && f_end && w_start && w_end && w_type && w_dist && gc && gc_mean && gc_stddev && gc_cv,
              "sw_batch: NULL feature/output pointer");
    uint64_t n_iv = 0;
    if (want_count) {
        ARG_CHECK(iv_off && rg_count, "sw_count_batch: NULL index offsets / rg_count");
        n_iv = iv_off[n_ctg] - iv_off[0];
        ARG_CHECK((starts && stops) || n_iv == 0, "sw_count_batch: NULL index pointer");
    }
    bool grouped = true;   // features sorted by ctg (the reference walks ctg by ctg, sw.rs:95-99): rows are grouped by segment
    for (int64_t f = 1; f < n_feat && grouped; ++f) grouped = f_ctg[f - 1] <= f_ctg[f];
    // Row counts are pure coordinate arithmetic (window.rs:3-56): 1 + #L + #R.
    for (int64_t f = 0; f < n_feat; ++f) {
        const uint32_t c = f_ctg[f];
        ARG_CHECK((int64_t)c < n_ctg, "sw_batch: feature %lld names ctg %u of %lld", (long long)f, c, (long long)n_ctg);
        const int64_t psize = (int64_t)(seq_off[c + 1] - seq_off[c]);
        const int64_t s = (int64_t)f_start[f] - ctg_chr_start[c] + 1, e = (int64_t)f_end[f] - ctg_chr_start[c] + 1;
        ARG_CHECK(s >= 1 && e >= s && e <= psize, "sw_batch: feature %lld (%d-%d) is not inside its ctg", (long long)f,
                  f_start[f], f_end[f]);
        const int64_t half = (e - s + 1) / 2;
        const int64_t ml = half == 0 ? s : s + half - 1, mr = half == 0 ? s : s + half;
        const int64_t ms = std::max<int64_t>(ml - size / 2 + 1, 1), me = std::min<int64_t>(mr + size / 2 - 1, psize);
        const int64_t nl = std::min<int64_t>(max, (ms - 1) / size), nr = std::min<int64_t>(max, (psize - me) / size);
        row_off[f + 1] = row_off[f] + 1 + (uint64_t)nl + (uint64_t)nr;
    }
    const uint64_t n_rows = row_off[n_feat];
    trace_stage("sw: row counts (host)");
    std::lock_guard<std::mutex> lk(ctx->mu);
    PerDevice &pd = ctx->devs[0];
    DeviceGuard g(pd.device);
    uint8_t *d_seq;
    uint64_t *d_off;
    int32_t rc = upload_seq(pd, n_ctg, seq_concat, seq_off, &d_seq, &d_off);
    if (rc) return rc;
    trace_stage("sw: upload_seq");
    // inputs: chr_start | f_ctg | f_start | f_end | row_off ; outputs: 4 x i32/f32 x 2 + i8
    const uint64_t in_bytes = (uint64_t)n_ctg * 4 + (uint64_t)n_feat * 12 + (uint64_t)(n_feat + 1) * 8 + 64;
    if ((rc = pd.scratch[1].reserve(in_bytes))) return rc;
    if ((rc = pd.scratch[2].reserve(n_rows * 29 + 64))) return rc;
    cudaStream_t st = pd.stream;
    uint8_t *ib = pd.scratch[1].as<uint8_t>();
    uint64_t *d_row_off = reinterpret_cast<uint64_t *>(ib);
    int32_t *d_chr_start = reinterpret_cast<int32_t *>(ib + (uint64_t)(n_feat + 1) * 8);
    uint32_t *d_f_ctg = reinterpret_cast<uint32_t *>(d_chr_start + n_ctg);
    int32_t *d_f_start = reinterpret_cast<int32_t *>(d_f_ctg + n_feat);
    int32_t *d_f_end = d_f_start + n_feat;
    CUDA_TRY(cudaMemcpyAsync(d_row_off, row_off, (uint64_t)(n_feat + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_chr_start, ctg_chr_start, (uint64_t)n_ctg * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_f_ctg, f_ctg, (uint64_t)n_feat * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_f_start, f_start, (uint64_t)n_feat * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_f_end, f_end, (uint64_t)n_feat * 4, cudaMemcpyHostToDevice, st));
    uint8_t *ob = pd.scratch[2].as<uint8_t>();
    int32_t *d_ws = reinterpret_cast<int32_t *>(ob), *d_we = d_ws + n_rows, *d_wd = d_we + n_rows;
    float *d_g = reinterpret_cast<float *>(d_wd + n_rows), *d_m = d_g + n_rows, *d_s = d_m + n_rows, *d_c = d_s + n_rows;
    int8_t *d_wt = reinterpret_cast<int8_t *>(d_c + n_rows);
    if ((rc = launch_sw_rows(st, d_seq, d_off, d_chr_start, n_feat, d_f_ctg, d_f_start, d_f_end, size, max, resize,
                             d_row_off, n_rows, d_ws, d_we, d_wt, d_wd, d_g, d_m, d_s, d_c)))
        return rc;
    trace_stage("sw: kernel", st, true);
    std::unique_ptr<gams_cuda_count_plan, void (*)(gams_cuda_count_plan *)> cplan(nullptr, gams_cuda_count_plan_destroy);
    if (want_count && n_rows) {
        // sw --action count (sw.rs:28-32, data.rs:55): gams::count_rg of every generated window against the rg index of its ctg,
        // count(start, end) with the inclusive end as Lapper's exclusive stop exactly as locate.rs:136 calls it (utils.rs:35)
        gams_cuda_count_plan *cp = nullptr;
        if ((rc = gams_cuda_count_plan_create(ctx, 0, n_ctg, iv_off, &cp))) return rc;
        cplan.reset(cp);
        if ((rc = pd.scratch[3].reserve(std::max<uint64_t>(n_iv, 1) * 4))) return rc;
        if ((rc = pd.scratch[4].reserve(std::max<uint64_t>(n_iv, 1) * 4))) return rc;
        if ((rc = pd.scratch[5].reserve(n_rows * 8 + 64))) return rc;
        uint32_t *d_st = pd.scratch[3].as<uint32_t>(), *d_sp = pd.scratch[4].as<uint32_t>();
        int32_t *d_rseg = pd.scratch[5].as<int32_t>();
        uint32_t *d_rcnt = reinterpret_cast<uint32_t *>(d_rseg + n_rows);
        if (n_iv) {
            CUDA_TRY(cudaMemcpyAsync(d_st, starts + iv_off[0], n_iv * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(d_sp, stops + iv_off[0], n_iv * 4, cudaMemcpyHostToDevice, st));
        }
        sw_row_seg_kernel<<<(unsigned)((n_feat + 255) / 256), 256, 0, st>>>(n_feat, d_f_ctg, d_row_off, d_rseg);
        CUDA_TRY(cudaGetLastError());
        if ((rc = launch_count_overlaps(cp, st, d_st, d_sp, (int64_t)n_rows, d_rseg, reinterpret_cast<const uint32_t *>(d_ws),
                                        reinterpret_cast<const uint32_t *>(d_we), grouped ? 1 : 0, d_rcnt)))
            return rc;
        CUDA_TRY(cudaMemcpyAsync(rg_count, d_rcnt, n_rows * 4, cudaMemcpyDeviceToHost, st));
        trace_stage("sw: rg counts", st, true);
    }
    CUDA_TRY(cudaMemcpyAsync(w_start, d_ws, n_rows * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(w_end, d_we, n_rows * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(w_dist, d_wd, n_rows * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(gc, d_g, n_rows * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(gc_mean, d_m, n_rows * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(gc_stddev, d_s, n_rows * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(gc_cv, d_c, n_rows * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(w_type, d_wt, n_rows, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    trace_stage("sw: D2H");
    return GAMS_CUDA_OK;
}

GAMS_API int32_t gams_cuda_sw_batch(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *seq_concat,
                                    const uint64_t *seq_off, const int32_t *ctg_chr_start, int64_t n_feat,
                                    const uint32_t *f_ctg, const int32_t *f_start, const int32_t *f_end, int32_t size,
                                    int32_t max, int32_t resize, uint64_t *row_off, int32_t *w_start, int32_t *w_end,
                                    int8_t *w_type, int32_t *w_dist, float *gc, float *gc_mean, float *gc_stddev,
                                    float *gc_cv) {
    return sw_batch_impl(ctx, n_ctg, seq_concat, seq_off, ctg_chr_start, n_feat, f_ctg, f_start, f_end, size, max, resize, row_off,
                         w_start, w_end, w_type, w_dist, gc, gc_mean, gc_stddev, gc_cv, false, nullptr, nullptr, nullptr, nullptr);
}

GAMS_API int32_t gams_cuda_sw_count_batch(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *seq_concat,
                                          const uint64_t *seq_off, const int32_t *ctg_chr_start, int64_t n_feat,
                                          const uint32_t *f_ctg, const int32_t *f_start, const int32_t *f_end, int32_t size,
                                          int32_t max, int32_t resize, const uint32_t *starts, const uint32_t *stops,
                                          const uint64_t *iv_off, uint64_t *row_off, int32_t *w_start, int32_t *w_end,
                                          int8_t *w_type, int32_t *w_dist, float *gc, float *gc_mean, float *gc_stddev,
                                          float *gc_cv, uint32_t *rg_count) {
    return sw_batch_impl(ctx, n_ctg, seq_concat, seq_off, ctg_chr_start, n_feat, f_ctg, f_start, f_end, size, max, resize, row_off,
                         w_start, w_end, w_type, w_dist, gc, gc_mean, gc_stddev, gc_cv, true, starts, stops, iv_off, rg_count);
}
