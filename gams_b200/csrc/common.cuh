// common.cuh — shared declarations of libgams_cuda.so (internal; the public surface is include/gams_cuda.h)
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gams_cuda.h"

#define GAMS_API extern "C" __attribute__((visibility("default")))

namespace gams {

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const char *fmt, ...);
int32_t cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define CUDA_TRY(expr)                                                       \
    do {                                                                     \
        cudaError_t _e = (expr);                                             \
        if (_e != cudaSuccess) return ::gams::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

#define ARG_CHECK(cond, ...)            \
    do {                                \
        if (!(cond)) {                  \
            ::gams::set_error(__VA_ARGS__); \
            return GAMS_CUDA_EINVAL;    \
        }                               \
    } while (0)

// Run fn(d) for d = 0 .. n_used-1, one host thread per device when there are several; the first failure wins and its
// message becomes the caller's gams_cuda_last_error().
template <class F>
inline int32_t run_on_devices(int n_used, F fn) {
    if (n_used <= 1) return fn(0);
    std::vector<int32_t> rc((size_t)n_used, 0);
    std::vector<std::string> msg((size_t)n_used);
    std::vector<std::thread> th;
    for (int d = 0; d < n_used; ++d)
        th.emplace_back([&, d] {
            rc[d] = fn(d);
            if (rc[d]) msg[d] = std::string(gams_cuda_last_error());
        });
    for (auto &t : th) t.join();
    for (int d = 0; d < n_used; ++d)
        if (rc[d]) {
            set_error("%s", msg[d].c_str());
            return rc[d];
        }
    return GAMS_CUDA_OK;
}

// ---- device memory ----------------------------------------------------------------------------
// Freed blocks of one device, kept for reuse: the host API creates and destroys a plan per call, and cudaMalloc /
// cudaFree of its tables (up to GBs of rank nodes) cost from milliseconds to a second each.  Blocks return to the
// driver when the context is destroyed.
struct BlockCache {
    std::mutex mu;
    std::multimap<size_t, void *> free_blocks;
    static size_t round_up(size_t bytes);
    int32_t get(void **p, size_t bytes);      // *p has round_up(bytes) bytes
    void put(void *p, size_t bytes);
    void clear();
    // pinned, device-mapped staging for upload_table (grow-only; one table at a time)
    void *h_stage = nullptr, *d_stage = nullptr;
    size_t stage_cap = 0;
    ~BlockCache();
};
// Small tables (tile descriptors, offsets) go up by a KERNEL that reads them from pinned host memory, not by the copy engine: a
// cudaMemcpyAsync queues behind the 64 MB slices of sequence the host paths have in flight (the plan build waited 9 ms for them),
// a kernel only shares the link with them.  Blocks until the table is on the device (the staging buffer is reused).
// cache == nullptr: plain cudaMemcpyAsync + synchronise.
int32_t upload_table(BlockCache *cache, void *d_dst, const void *h_src, size_t bytes, cudaStream_t st);
int32_t dev_alloc(BlockCache *cache, void **p, size_t bytes);   // cache == nullptr: plain cudaMalloc
void dev_free(BlockCache *cache, void *p, size_t bytes);

struct DevBuf {  // grow-only device scratch
    void *p = nullptr;
    size_t cap = 0;
    BlockCache *cache = nullptr;
    int32_t reserve(size_t bytes);
    void release();
    template <class T> T *as() const { return static_cast<T *>(p); }
};

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- tile descriptors ---------------------------------------------------------------------------
struct TileDesc {
    uint32_t item;   // ctg (GC / z-score tiles) index
    uint32_t first;  // first window of the tile inside the ctg
};

// Per-segment parameters of the rank tables of Kernel C (overlap.cu); [0] = starts[], [1] = stops[]
struct alignas(16) SegRank {
    uint64_t off;   // first interval of the segment in the concatenated arrays
    uint32_t n;     // intervals in the segment
    uint32_t pad;
    uint32_t lo[2], hi[2], shift[2];
    uint32_t pad2[2];
};

// Geometry of Kernel A (gc_window.cu)
constexpr int GC_THREADS = 256;
struct GcGeom {
    int k = 0;            // 16-byte chunks per thread (4 or 8); 0 = generic kernel (huge windows)
    int wt = 0;           // windows per tile
    uint64_t n_tiles = 0;
};

// Kernel AB (gc_wave_fused.cu): tile geometry and the host-computed bounds of its certificate (make_fused_const)
#ifndef FW_THREADS_N
#define FW_THREADS_N 256                       // threads per Kernel-AB tile (a compile-time knob of gc_wave_fused.cu only; A/B: profiles/README.md)
#endif
constexpr int FW_THREADS = FW_THREADS_N;
constexpr int FW_E = 8;                        // consecutive windows per thread in the statistics phases
constexpr int FW_MAX_EL = FW_THREADS * FW_E;   // windows per tile, lag+1 halo windows included
constexpr int FW_K = 6;                        // 16-byte chunks per thread: 24 576 bytes of sequence per tile
constexpr int FW_PAD = 416;                    // front padding of the prefix arrays (>= FW_MAX_LAG, multiple of 8)
constexpr int FW_MAX_LAG = 400;
constexpr int FW_MAX_R = 248;                  // size / step
constexpr int FW_STAB = 256;                   // GC table kept in shared memory when size < FW_STAB
constexpr int FUSED_LIST_CAP = 64;             // inconclusive windows listed per tile before the tile is replayed wholesale
struct FusedConst {
    float k, e2x2, epsk, beta;
};
struct alignas(32) FusedTile {
    int64_t byte0;     // sequence byte where element 0 starts (negative only for the first tile of the buffer)
    int64_t g0;        // global window index of element 0; g0 = 7 (mod 8)
    uint32_t e_out0;   // outputs are elements [e_out0, e_out1)
    uint32_t e_out1;
    int32_t e_ctg0;    // element of the ctg's window 0 (first tile of a ctg), else far negative
    uint32_t pi_min;   // 32-byte pieces of the tile that lie before the start of the sequence buffer (not loaded)
};

// Geometry of Kernel B1 (zscore.cu)
constexpr int ZS_R = 8;                       // windows per thread
constexpr int ZS_THREADS = 128;
constexpr int ZS_TILE = ZS_R * ZS_THREADS;    // windows per tile
constexpr int ZS_MAX_TILED_LAG = 8000;        // (tile + lag) * 21 B of shared memory must fit 227 KB

// A batch of series (one per ctg) resident on one device: offsets + z-score tiles.
struct SeriesLayout {
    int64_t n_series = 0;
    uint64_t total = 0;
    uint64_t min_len = 0;
    uint64_t *d_off = nullptr;       // n_series+1
    TileDesc *d_ztiles = nullptr;    // z-score tiles
    uint64_t n_ztiles = 0;
    int *d_nonfinite = nullptr;      // n_series flags: the series holds Inf / NaN (set by zscore_indep_kernel, see launch_zscore)
    std::vector<uint64_t> h_off;       // n_series+1 (relative to the first series)
    std::vector<uint64_t> h_ztile_off; // n_series+1: first z-score tile of each series
    BlockCache *cache = nullptr;
    // the segmented walk of influence != 1 (zscore.cu): first segment of each series of the last range asked for, what every
    // segment's walk left behind (window of its last signal; filtered history of its last lag + 1 windows)
    mutable std::vector<uint32_t> h_seg_off;
    mutable uint32_t *d_seg_off = nullptr;
    mutable long long *d_seg_last = nullptr;
    mutable float *d_seg_ring = nullptr;
    mutable size_t seg_off_bytes = 0, seg_last_bytes = 0, seg_ring_bytes = 0;
    mutable int64_t seg_s0 = -1, seg_s1 = -1, seg_len = 0;
    int32_t seg_tables(int64_t s0, int64_t s1, int64_t seg, uint64_t lag, cudaStream_t st) const;
    int32_t build(const uint64_t *h_off, int64_t n, cudaStream_t st);
    void release();
};

struct PerDevice {
    int device = -1;
    cudaStream_t stream = nullptr;   // library-owned stream for the host API
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaStream_t aux = nullptr;      // the decoder of the chained gzip path (runs beside the plan build)
    DevBuf scratch[8];
    unsigned long long *d_zstats = nullptr;  // [0] windows replayed after an inconclusive certificate, [1] verify mismatches
    BlockCache *cache = nullptr;             // owned by the context
};

}  // namespace gams

struct gams_cuda_ctx {
    std::vector<gams::PerDevice> devs;
    std::mutex mu;
    int zscore_mode = 0;  // 0 certified filter + exact replay, 1 replay everything, 2 verify
    int gunzip_mode = 0;  // 0 two-phase decoder (inflate2.cu), 1 warp-per-blob decoder (inflate.cu); GAMS_CUDA_GUNZIP_MODE
    int gunzip_crc = 1;   // verify the CRC-32 of the gzip trailer as flate2's GzDecoder does; GAMS_CUDA_GUNZIP_CRC=0 skips it
};

struct gams_cuda_wave_plan {
    gams_cuda_ctx *ctx = nullptr;
    int dev_index = 0;
    int device = 0;
    int64_t n_ctg = 0;
    int32_t size = 0, step = 0;
    uint64_t lag = 0;
    uint64_t total_bases = 0;
    std::vector<uint64_t> h_seq_off, h_win_off;
    uint64_t *d_seq_off = nullptr;
    gams::SeriesLayout win;          // window offsets + z-score tiles
    gams::GcGeom geom;
    gams::TileDesc *d_gtiles = nullptr;
    std::vector<uint64_t> h_gtile_off;  // n_ctg+1: first Kernel-A tile of each ctg
    float *d_gc_tab = nullptr;       // (float)c / (float)size for c = 0..size (host-computed IEEE divide)
    int fused_wt = 0;                // windows per Kernel-AB tile; 0 = geometry not eligible for the fused kernel
    gams::FusedTile *d_ftiles = nullptr;
    size_t n_ftiles = 0;
    std::vector<uint64_t> h_ftile_off;  // n_ctg+1: first Kernel-AB tile of each ctg
    gams::DevBuf peak_tmp;           // tile counts / bases of the peak-merge kernels
    uint64_t launches = 0;
};

struct gams_cuda_count_plan {
    gams_cuda_ctx *ctx = nullptr;
    int dev_index = 0;
    int device = 0;
    int64_t n_seg = 0;
    uint64_t total_iv = 0;
    uint64_t max_seg = 0;
    std::vector<uint64_t> h_iv_off;
    uint64_t *d_iv_off = nullptr;
    gams::DevBuf tmp[9];             // [0] SegRank per segment, [1]/[2] rank tables of starts[] / stops[]; [6]-[8]: overlap_part.cu
    const uint32_t *idx_starts = nullptr, *idx_stops = nullptr;  // arrays the rank tables were built for
    bool idx_valid = false;
    // compact per-segment index of the staged kernel (queries grouped by segment): tmp[4] = offsets + parameters, tmp[5] = data
    const uint32_t *cidx_starts = nullptr, *cidx_stops = nullptr;
    bool cidx_valid = false;
    size_t cpar_off = 0;
    size_t st_smem = 0;              // dynamic shared memory of the staged kernel: the largest staged segment
    int mode = 0;                    // 0 default (partitioned / ranked / staged), 1 plain binary search, 2 ranked only (GAMS_CUDA_COUNT_MODE)
    // partitioned path for queries in arbitrary order (overlap_part.cu): bucket tables in tmp[7], valid for the compact index
    bool part_valid = false;
    int part_nbkt = 0, part_bshift = 0;
    uint64_t launches = 0;
};

namespace gams {
// kernel launchers (each returns 0 or a GAMS_CUDA_E* code; all asynchronous on `st`)
// [c0, c1): ctg / series sub-range to process (the host API pipelines copy and compute over such slices)
int32_t launch_gc_windows(const gams_cuda_wave_plan *p, cudaStream_t st, const uint8_t *d_seq,
                          uint32_t *d_cnt, float *d_gc, uint64_t *launches, int64_t c0, int64_t c1);
int32_t launch_zscore(const SeriesLayout &L, cudaStream_t st, const float *d_x, uint64_t lag, float threshold,
                      float influence, float *d_filtered, int8_t *d_signal, uint64_t *launches, int mode,
                      unsigned long long *d_stats, int64_t s0, int64_t s1, bool gc_values = false);
// gc_values: d_x was produced by the GC kernels of this library (fractions in [0, 1]: no Inf / NaN fix-up pass needed)
int32_t launch_count_overlaps(gams_cuda_count_plan *p, cudaStream_t st, const uint32_t *d_starts,
                              const uint32_t *d_stops, int64_t n_q, const int32_t *d_q_seg,
                              const uint32_t *d_q_start, const uint32_t *d_q_stop, int32_t grouped,
                              uint32_t *d_out);
int32_t build_count_index(gams_cuda_count_plan *p, cudaStream_t st, const uint32_t *d_starts, const uint32_t *d_stops);
int32_t build_compact_index(gams_cuda_count_plan *p, cudaStream_t st, const uint32_t *d_starts, const uint32_t *d_stops);
// overlap_part.cu: queries in arbitrary order through a coarse partition; *done = false: not eligible, nothing was enqueued
bool count_part_eligible(const gams_cuda_count_plan *p, int64_t n_q);
int32_t launch_count_overlaps_part(gams_cuda_count_plan *p, cudaStream_t st, const uint32_t *d_starts, const uint32_t *d_stops,
                                   int64_t n_q, const int32_t *d_q_seg, const uint32_t *d_q_start, const uint32_t *d_q_stop,
                                   uint32_t *d_out, bool *done);
int32_t launch_locate_ctg(cudaStream_t st, int64_t n_chr, const uint32_t *d_ctg_start, const uint32_t *d_ctg_stop,
                          const uint64_t *d_chr_off, const uint32_t *d_chr_maxlen, int64_t n_q,
                          const int32_t *d_q_chr, const uint32_t *d_q_start, const uint32_t *d_q_stop, int32_t *d_out);
int32_t launch_gc_ranges(cudaStream_t st, const uint8_t *d_seq, const uint64_t *d_seq_off, int64_t n_ctg, int64_t n_q,
                         const uint32_t *d_q_ctg, const uint32_t *d_q_from, const uint32_t *d_q_to, uint32_t *d_cnt);
int32_t launch_sw_rows(cudaStream_t st, const uint8_t *d_seq, const uint64_t *d_seq_off, const int32_t *d_chr_start,
                       int64_t n_feat, const uint32_t *d_f_ctg, const int32_t *d_f_start, const int32_t *d_f_end,
                       int32_t size, int32_t max, int32_t resize, const uint64_t *d_row_off, uint64_t n_rows,
                       int32_t *d_w_start, int32_t *d_w_end, int8_t *d_w_type, int32_t *d_w_dist, float *d_gc,
                       float *d_mean, float *d_sd, float *d_cv);
GcGeom plan_gc_geometry(int32_t size, int32_t step);
int32_t launch_anno_coverage(cudaStream_t st, const int32_t *run_start, const int32_t *run_end, const int64_t *run_cum,
                             const uint64_t *run_off, int64_t n_chr, int64_t n_q, const int32_t *q_chr, const int32_t *q_ctg_start,
                             const int32_t *q_ctg_end, const int32_t *q_start, const int32_t *q_end, float *prop);
// gen.cu
int32_t launch_valid_runs(cudaStream_t st, const uint8_t *d_seq, uint64_t total, const uint64_t *d_chr_off, int64_t n_chr,
                          uint64_t *d_tmp, uint64_t cap, uint64_t *d_run_start, uint64_t *d_run_end, uint64_t *d_n_runs);
// peaks.cu
bool peak_merge_is_runs(int32_t size, int32_t step, float coverage);
size_t peak_tmp_bytes(uint64_t n_win);
int32_t launch_peak_runs(cudaStream_t st, const int8_t *d_sig, const float *d_gc, const uint64_t *d_win_off, int64_t n_ctg,
                         uint64_t g0, uint64_t g1, int32_t size, int32_t step, uint64_t *d_tmp, uint64_t cap,
                         uint32_t *pk_ctg, uint32_t *pk_first, uint32_t *pk_last, float *pk_gc, int8_t *pk_sig,
                         uint64_t *d_n_peaks);
int32_t launch_peak_link(cudaStream_t st, int64_t n_pk, const uint32_t *p_ctg, const int32_t *p_start, const int32_t *p_end,
                         const int8_t *p_sig, const uint32_t *cnt, const int32_t *chr_start, const int32_t *chr_end, float *gc,
                         int8_t *l_sig, int32_t *l_len, float *l_amp, int8_t *r_sig, int32_t *r_len, float *r_amp);
// Kernel AB: bytes -> GC + signals in one pass (influence == 1, threshold >= 0, eligible geometry only)
int fused_windows_per_tile(int32_t size, int32_t step, uint64_t lag);
int32_t launch_gc_wave_fused(gams_cuda_wave_plan *p, cudaStream_t st, const uint8_t *d_seq, float threshold,
                             uint32_t *d_cnt, float *d_gc, int8_t *d_sig, unsigned long long *d_stats, int64_t c0,
                             int64_t c1);
void fused_tiles_for_ctg(std::vector<FusedTile> &out, uint64_t seq_off, uint64_t win_off, uint64_t n_c, int32_t step,
                         uint64_t lag, int wt);
bool wave_fused_ok(const gams_cuda_wave_plan *p, float threshold, float influence);
// GAMS_CUDA_TRACE=1: host-clock stage timings on stderr (sync: after synchronising st; for debugging only)
bool trace_on();
void trace_stage(const char *what, cudaStream_t st = nullptr, bool sync = false);
bool zscore_skip_ok(uint64_t lag);
int32_t launch_zscore_skip(const SeriesLayout &L, cudaStream_t st, const float *d_x, uint64_t lag, float threshold, float influence,
                           int8_t *d_signal, uint64_t *launches, int64_t s0, int64_t s1, bool gc_values);
}  // namespace gams
