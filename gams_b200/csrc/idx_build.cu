// idx_build.cu — the rg index of a ctg set, built on the device.
//
// Replaces the tail of `gams rg`: gams::build_idx_rg (src/libs/redis.rs:276-305) = per ctg
//   rust_lapper::Lapper::new(intervals)  — sort the intervals by (start, stop); starts[] and stops[] sorted on their own;
//                                          max_len = max(stop - start), 0 for stop < start (rust-lapper 1.1.0 lib.rs `new`)
//   bincode::serialize(&lapper)          — the blob stored under idx:rg:{ctg} (layout: SURVEY.md §8 a14, restated in
//                                          oracle/oracle.py lapper_to_bincode; "parity unpinned": no reference test reads bytes)
// What leaves the device is exactly what gams_cuda_count_overlaps consumes (starts[] / stops[] per segment), so the
// index never has to round-trip through the host when queries follow.
//
// Sorting: idx_sort_kernel, one CTA per segment.  A ctg's intervals (16 K SNPs at C5 scale) fit shared memory: 64-bit keys
// start << 32 | stop, a bitonic network whose every comparison is ascending (partner i ^ (2k - 1) in the first step of a
// merge, i ^ stride in the others: the array may then be padded with +inf virtually, any n works), then the stops alone as
// 32-bit keys in the same shared memory.  Each interval is read once and each output written once (8 B in, 16 B out); the
// O(n log^2 n) compare-exchanges stay on chip.  A segment too large for shared memory is sorted by the same network in global
// memory by its CTA (correct for any size; slow, and rare: a ctg is at most a few hundred kbp).
// Bound: shared-memory bandwidth / barriers (105 network steps for 16 K keys), not HBM.  Not on BASELINE's timed path.
#include "common.cuh"

namespace gams {

constexpr int IS_THREADS = 1024;
constexpr uint32_t IS_CAP = 16384;            // 64-bit keys in shared memory (128 KB: one CTA per SM)

// ascending bitonic network over a[0 .. n) (virtually padded with +inf to a power of two); CTA-wide
template <class T>
__device__ void bitonic_sort_cta(T *a, uint32_t n) {
    uint32_t p = 1;
    while (p < n) p <<= 1;
    for (uint32_t k = 2; k <= p; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            const bool first = j == (k >> 1);
            for (uint32_t t = threadIdx.x; t < (p >> 1); t += blockDim.x) {
                // pair t of this step: lo below its partner
                const uint32_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const uint32_t hi = first ? (lo ^ (k - 1)) : (lo | j);     // mirror inside the block of k for the first step
                if (hi < n) {                                             // (lo < hi always; a partner past n is +inf: no exchange)
                    const T x = a[lo], y = a[hi];
                    if (y < x) { a[lo] = y; a[hi] = x; }
                }
            }
            __syncthreads();
        }
    }
}

// one CTA per segment (grid-stride): Lapper::new's three sorted views of the segment
__global__ void __launch_bounds__(IS_THREADS)
idx_sort_kernel(const uint32_t *__restrict__ start, const uint32_t *__restrict__ stop, const uint64_t *__restrict__ iv_off, int64_t n_seg,
                uint64_t *__restrict__ big_keys /* n u64: workspace of the segments that do not fit shared memory */,
                uint32_t *__restrict__ iv_start, uint32_t *__restrict__ iv_stop, uint32_t *__restrict__ starts, uint32_t *__restrict__ stops) {
    extern __shared__ __align__(16) uint8_t is_smem[];
    uint64_t *k64 = reinterpret_cast<uint64_t *>(is_smem);
    uint32_t *k32 = reinterpret_cast<uint32_t *>(is_smem);
    for (int64_t s = blockIdx.x; s < n_seg; s += gridDim.x) {
        const uint64_t o = iv_off[s], n64 = iv_off[s + 1] - o;
        const uint32_t n = (uint32_t)n64;
        if (n64 <= IS_CAP) {
            for (uint32_t i = threadIdx.x; i < n; i += IS_THREADS) k64[i] = ((uint64_t)start[o + i] << 32) | (uint64_t)stop[o + i];
            __syncthreads();
            bitonic_sort_cta(k64, n);
            for (uint32_t i = threadIdx.x; i < n; i += IS_THREADS) {
                const uint64_t k = k64[i];
                iv_start[o + i] = (uint32_t)(k >> 32);
                iv_stop[o + i] = (uint32_t)k;
                starts[o + i] = (uint32_t)(k >> 32);
            }
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < n; i += IS_THREADS) k32[i] = stop[o + i];
            __syncthreads();
            bitonic_sort_cta(k32, n);
            for (uint32_t i = threadIdx.x; i < n; i += IS_THREADS) stops[o + i] = k32[i];
            __syncthreads();
        } else {   // in global memory (n < 2^31 per call)
            uint64_t *g = big_keys + o;
            for (uint32_t i = threadIdx.x; i < n; i += IS_THREADS) g[i] = ((uint64_t)start[o + i] << 32) | (uint64_t)stop[o + i];
            __syncthreads();
            bitonic_sort_cta(g, n);
            for (uint32_t i = threadIdx.x; i < n; i += IS_THREADS) {
                const uint64_t k = g[i];
                iv_start[o + i] = (uint32_t)(k >> 32);
                iv_stop[o + i] = (uint32_t)k;
                starts[o + i] = (uint32_t)(k >> 32);
                stops[o + i] = stop[o + i];
            }
            __syncthreads();
            bitonic_sort_cta(stops + o, n);
        }
    }
}

// one CTA per segment (grid-stride over segments): max_len, then the segment's bincode blob in 16-bit units
// (blob offset 24 iv_off[s] + 30 s is even, not always a multiple of 4)
__global__ void __launch_bounds__(256)
idx_blob_kernel(const uint32_t *__restrict__ iv_start, const uint32_t *__restrict__ iv_stop, const uint32_t *__restrict__ starts,
                const uint32_t *__restrict__ stops, const uint64_t *__restrict__ iv_off, int64_t n_seg,
                uint32_t *__restrict__ max_len, uint8_t *__restrict__ blob) {
    __shared__ uint32_t red[8];
    __shared__ uint32_t ml_sh;
    for (int64_t s = blockIdx.x; s < n_seg; s += gridDim.x) {
        const uint64_t o = iv_off[s], n = iv_off[s + 1] - o;
        uint32_t ml = 0;
        for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t a = iv_start[o + i], b = iv_stop[o + i];
            ml = max(ml, b >= a ? b - a : 0u);   // checked_sub(..).unwrap_or(0)
        }
        ml = __reduce_max_sync(0xFFFFFFFFu, ml);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ml;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t m = 0;
            for (int w = 0; w < 8; ++w) m = max(m, red[w]);
            ml_sh = m;
            max_len[s] = m;
        }
        __syncthreads();
        if (blob) {
            const uint32_t mlen = ml_sh;
            uint16_t *out = reinterpret_cast<uint16_t *>(blob + 24 * o + 30 * (uint64_t)s);
            const uint64_t units = 12 * n + 15;
            // byte offsets: [0,8) n | [8, 8+16n) {start, stop, u64 len("") = 0} | +8 n | +4n starts | +8 n | +4n stops | max_len | 0 | 0
            const uint64_t b_iv = 8, b_n2 = 8 + 16 * n, b_st = b_n2 + 8, b_n3 = b_st + 4 * n, b_sp = b_n3 + 8, b_ml = b_sp + 4 * n;
            for (uint64_t u = threadIdx.x; u < units; u += blockDim.x) {
                const uint64_t b = 2 * u;
                uint64_t word = 0;      // the 8-byte (or 4-byte) field the unit lies in, and the unit's byte offset inside it
                uint32_t in = 0;
                if (b < b_iv) { word = n; in = (uint32_t)b; }
                else if (b < b_n2) {
                    const uint64_t r = b - b_iv, i = r >> 4;
                    const uint32_t f = (uint32_t)(r & 15);
                    if (f < 4) { word = iv_start[o + i]; in = f; }
                    else if (f < 8) { word = iv_stop[o + i]; in = f - 4; }
                    else { word = 0; in = 0; }
                }
                else if (b < b_st) { word = n; in = (uint32_t)(b - b_n2); }
                else if (b < b_n3) { const uint64_t r = b - b_st; word = starts[o + (r >> 2)]; in = (uint32_t)(r & 3); }
                else if (b < b_sp) { word = n; in = (uint32_t)(b - b_n3); }
                else if (b < b_ml) { const uint64_t r = b - b_sp; word = stops[o + (r >> 2)]; in = (uint32_t)(r & 3); }
                else if (b < b_ml + 4) { word = mlen; in = (uint32_t)(b - b_ml); }
                else { word = 0; in = 0; }   // cov: None (tag 0), overlaps_merged: false
                out[u] = (uint16_t)(word >> (8 * in));
            }
        }
        __syncthreads();
    }
}

// All pointers device memory.  d_key_a: n u64 of workspace (touched only for segments of more than IS_CAP intervals).
int32_t launch_build_idx(cudaStream_t st, int64_t n_seg, uint64_t n, const uint32_t *d_start, const uint32_t *d_stop,
                         const uint64_t *d_iv_off, uint64_t *d_key_a, uint32_t *d_iv_start, uint32_t *d_iv_stop, uint32_t *d_starts,
                         uint32_t *d_stops, uint32_t *d_max_len, uint8_t *d_blob) {
    if (n_seg == 0) return GAMS_CUDA_OK;
    if (n) {
        const size_t smem = (size_t)IS_CAP * 8;
        CUDA_TRY(cudaFuncSetAttribute(idx_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned blocks = (unsigned)std::min<int64_t>(n_seg, 148 * 4);
        idx_sort_kernel<<<blocks, IS_THREADS, smem, st>>>(d_start, d_stop, d_iv_off, n_seg, d_key_a, d_iv_start, d_iv_stop, d_starts, d_stops);
    }
    const unsigned sb = (unsigned)std::min<int64_t>(n_seg, 148 * 8);
    idx_blob_kernel<<<sb, 256, 0, st>>>(d_iv_start, d_iv_stop, d_starts, d_stops, d_iv_off, n_seg, d_max_len, d_blob);
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

}  // namespace gams

using namespace gams;

// One device's share: segments [0, n_seg) of the (shifted) iv_off; the interval arrays are the caller's whole arrays.
static int32_t build_idx_one(gams_cuda_ctx *ctx, int dv, int64_t n_seg, uint64_t seg0, const uint32_t *start, const uint32_t *stop,
                             const uint64_t *iv_off, uint32_t *iv_start, uint32_t *iv_stop, uint32_t *starts, uint32_t *stops,
                             uint32_t *max_len, uint8_t *blob) {
    if (n_seg == 0) return GAMS_CUDA_OK;
    const uint64_t base = iv_off[0], n = iv_off[n_seg] - base;
    std::vector<uint64_t> rel((size_t)n_seg + 1);
    for (int64_t s = 0; s <= n_seg; ++s) rel[s] = iv_off[s] - base;
    PerDevice &pd = ctx->devs[dv];
    DeviceGuard g(pd.device);
    cudaStream_t st = pd.stream;
    const uint64_t np = std::max<uint64_t>(n, 1);
    const uint64_t blob_bytes = blob ? 24 * n + 30 * (uint64_t)n_seg : 0;
    int32_t rc;
    // [1] inputs + offsets, [2] key workspace (only when a segment does not fit shared memory), [3] outputs (4 u32 columns +
    // max_len), [5] blob
    bool big = false;
    for (int64_t s = 0; s < n_seg && !big; ++s) big = rel[s + 1] - rel[s] > 16384;
    if ((rc = pd.scratch[1].reserve(np * 8 + (uint64_t)(n_seg + 1) * 8 + 64))) return rc;
    if ((rc = pd.scratch[2].reserve((big ? np * 8 : 8) + 64))) return rc;
    if ((rc = pd.scratch[3].reserve(np * 16 + (uint64_t)n_seg * 4 + 64))) return rc;
    if ((rc = pd.scratch[5].reserve(blob_bytes + 64))) return rc;
    uint64_t *d_off = pd.scratch[1].as<uint64_t>();
    uint32_t *d_start = reinterpret_cast<uint32_t *>(d_off + (n_seg + 1)), *d_stop = d_start + np;
    uint64_t *d_ka = pd.scratch[2].as<uint64_t>();
    uint32_t *d_ivs = pd.scratch[3].as<uint32_t>(), *d_ive = d_ivs + np, *d_sts = d_ive + np, *d_sps = d_sts + np, *d_ml = d_sps + np;
    uint8_t *d_blob = blob ? pd.scratch[5].as<uint8_t>() : nullptr;
    CUDA_TRY(cudaMemcpyAsync(d_off, rel.data(), (uint64_t)(n_seg + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n) {
        CUDA_TRY(cudaMemcpyAsync(d_start, start + base, n * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_stop, stop + base, n * 4, cudaMemcpyHostToDevice, st));
    }
    if ((rc = launch_build_idx(st, n_seg, n, d_start, d_stop, d_off, d_ka, d_ivs, d_ive, d_sts, d_sps, d_ml, d_blob)))
        return rc;
    if (n) {
        CUDA_TRY(cudaMemcpyAsync(iv_start + base, d_ivs, n * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(iv_stop + base, d_ive, n * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(starts + base, d_sts, n * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(stops + base, d_sps, n * 4, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaMemcpyAsync(max_len, d_ml, (uint64_t)n_seg * 4, cudaMemcpyDeviceToHost, st));
    // the share's blobs are contiguous in the caller's buffer: segment seg0 + s sits at 24 iv_off[s] + 30 (seg0 + s)
    if (blob) CUDA_TRY(cudaMemcpyAsync(blob + 24 * base + 30 * seg0, d_blob, blob_bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));   // rel is a local
    return GAMS_CUDA_OK;
}

// Host entry point (declared in include/gams_cuda.h): segments are split over the context's devices by interval count.
GAMS_API int32_t gams_cuda_build_idx_rg(gams_cuda_ctx *ctx, int64_t n_seg, const uint32_t *start, const uint32_t *stop,
                                        const uint64_t *iv_off, uint32_t *iv_start, uint32_t *iv_stop, uint32_t *starts,
                                        uint32_t *stops, uint32_t *max_len, uint8_t *blob) {
    ARG_CHECK(ctx && iv_off && n_seg >= 0, "build_idx_rg: bad argument");
    if (n_seg == 0) return GAMS_CUDA_OK;
    ARG_CHECK(iv_off[0] == 0, "build_idx_rg: iv_off[0] must be 0");
    for (int64_t s = 0; s < n_seg; ++s) ARG_CHECK(iv_off[s + 1] >= iv_off[s], "build_idx_rg: iv_off decreases at %lld", (long long)s);
    const uint64_t n = iv_off[n_seg];
    ARG_CHECK(n < (1ull << 31) && n_seg < (1ll << 31), "build_idx_rg: more than 2^31 intervals or segments in one call");
    ARG_CHECK(max_len != nullptr, "build_idx_rg: max_len is NULL");
    ARG_CHECK(n == 0 || (start && stop && iv_start && iv_stop && starts && stops), "build_idx_rg: NULL interval array");
    std::lock_guard<std::mutex> lk(ctx->mu);
    const int n_dev = (int)std::min<int64_t>((int64_t)ctx->devs.size(), std::max<int64_t>(1, (int64_t)(n >> 20)));
    std::vector<int64_t> part((size_t)n_dev + 1);
    int32_t rc = gams_cuda_partition(n_seg, iv_off, n_dev, part.data());
    if (rc) return rc;
    return run_on_devices(n_dev, [&](int d) {
        const int64_t s0 = part[d], s1 = part[d + 1];
        return build_idx_one(ctx, d, s1 - s0, (uint64_t)s0, start, stop, iv_off + s0, iv_start, iv_stop, starts, stops, max_len + s0,
                             blob);
    });
}
