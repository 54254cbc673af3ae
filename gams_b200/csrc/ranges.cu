// ranges.cu — GC of arbitrary sub-ranges and the `gams sw` row kernel.
//
// Replaces gams::cache_gc_content / cache_gc_stat / gc_stat (src/libs/utils.rs:141-213),
// gams::center_sw / center_resize (src/libs/window.rs:3-56, 96-124) and the per-feature body of
// sw::proc_ctg (src/cmd_gams/sw.rs:141-191, action "gc").
//
// One warp per query/row: the lanes stride over the aligned 4-byte words of the range, classify
// four bases per word with the same byte trick as Kernel A, and shuffle-reduce.  All f32
// arithmetic after the integer count follows the reference's order (sequential sums, separate
// multiply/add, IEEE divide/sqrt, round half away from zero).
#include "common.cuh"

namespace gams {

__device__ __forceinline__ uint32_t gc_flags(uint32_t w) {  // bit 7 of each byte <=> byte in {C,G,c,g}
    const uint32_t t = (w & 0xDBDBDBDBu) ^ 0x43434343u;
    const uint32_t a = (t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(a | t) & 0x80808080u;
}

// #GC bytes in absolute byte range [b0, b1) of seq; every lane returns the total.
__device__ __forceinline__ uint32_t warp_gc_count(const uint8_t *__restrict__ seq, uint64_t b0, uint64_t b1, int lane) {
    uint32_t c = 0;
    if (b1 > b0) {
        const uint64_t w_first = b0 >> 2, w_last = (b1 - 1) >> 2;
        const uint32_t *words = reinterpret_cast<const uint32_t *>(seq);
        for (uint64_t w = w_first + lane; w <= w_last; w += 32) {
            uint32_t f = gc_flags(__ldg(words + w));
            if (w == w_first) f &= 0xFFFFFFFFu << (8u * (uint32_t)(b0 & 3u));
            if (w == w_last) f &= 0xFFFFFFFFu >> (8u * (3u - (uint32_t)((b1 - 1) & 3u)));
            c += __popc(f);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
    return c;
}

__global__ void __launch_bounds__(256)
gc_ranges_kernel(const uint8_t *__restrict__ seq, const uint64_t *__restrict__ seq_off, int64_t n_ctg, int64_t n_q,
                 const uint32_t *__restrict__ q_ctg, const uint32_t *__restrict__ q_from,
                 const uint32_t *__restrict__ q_to, uint32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n_q; q += warps) {
        const uint32_t c = q_ctg[q];
        uint32_t cnt = 0;
        if ((int64_t)c < n_ctg) {
            const uint64_t o = seq_off[c], len = seq_off[c + 1] - o;
            uint64_t a = q_from[q], b = q_to[q];
            if (b > len) b = len;  // clamp; the reference would panic on an out-of-range slice
            if (a > b) a = b;
            cnt = warp_gc_count(seq, o + a, o + b, lane);
        }
        if (lane == 0) out[q] = cnt;
    }
}

int32_t launch_gc_ranges(cudaStream_t st, const uint8_t *d_seq, const uint64_t *d_seq_off, int64_t n_ctg, int64_t n_q,
                         const uint32_t *d_q_ctg, const uint32_t *d_q_from, const uint32_t *d_q_to, uint32_t *d_cnt) {
    if (n_q == 0) return GAMS_CUDA_OK;
    int64_t blocks = (n_q + 7) / 8;
    if (blocks > 148 * 64) blocks = 148 * 64;
    gc_ranges_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_seq, d_seq_off, n_ctg, n_q, d_q_ctg, d_q_from, d_q_to, d_cnt);
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

// ---- sw rows -------------------------------------------------------------------------------------
__device__ __forceinline__ float round4(float x) {  // utils.rs:135-138 with decimals = 4
    return __fdiv_rn(roundf(__fmul_rn(x, 10000.0f)), 10000.0f);
}

// window.rs:96-124 on a single-span parent [1, psize] in index space; child given as indexes [s, e]
__device__ __forceinline__ void center_resize_idx(int32_t psize, int32_t s, int32_t e, int32_t resize, int32_t &os,
                                                  int32_t &oe) {
    const int32_t half_size = (e - s + 1) / 2;
    const int32_t mid_left = (half_size == 0) ? s : s + half_size - 1;
    const int32_t mid_right = (half_size == 0) ? s : s + half_size;
    const int32_t half_resize = resize / 2;
    int32_t l = mid_left - half_resize + 1;
    if (l < 1) l = 1;
    int32_t r = mid_right + half_resize - 1;
    if (r > psize) r = psize;
    os = l;
    oe = r;
}

// #GC bytes in absolute byte range [b0, b1) of seq, by ONE thread (windows are ~100 bytes: 26 aligned words)
__device__ __forceinline__ uint32_t thread_gc_count(const uint8_t *__restrict__ seq, uint64_t b0, uint64_t b1) {
    if (b1 <= b0) return 0;
    const uint64_t w_first = b0 >> 2, w_last = (b1 - 1) >> 2;
    const uint32_t *words = reinterpret_cast<const uint32_t *>(seq);
    const uint32_t m_first = 0xFFFFFFFFu << (8u * (uint32_t)(b0 & 3u));
    const uint32_t m_last = 0xFFFFFFFFu >> (8u * (3u - (uint32_t)((b1 - 1) & 3u)));
    if (w_first == w_last) return __popc(gc_flags(__ldg(words + w_first)) & m_first & m_last);
    uint32_t c = __popc(gc_flags(__ldg(words + w_first)) & m_first) + __popc(gc_flags(__ldg(words + w_last)) & m_last);
#pragma unroll 4
    for (uint64_t w = w_first + 1; w < w_last; ++w) c += __popc(gc_flags(__ldg(words + w)));
    return c;
}

// One thread per row; a CTA owns 256 consecutive rows and finds their features through a shared-memory copy of the
// row offsets of the (at most 257) features that can own them.  (A warp per row with a 20-step binary search in global
// memory per row, the first version of this kernel, spent 96 ms on the 4.1e7 rows of C3 waiting on that search.)
constexpr int SW_THREADS = 256;
constexpr int SW_MAXSUB = 8;   // flank sub-windows kept in registers (resize / size + 1; more are recomputed)

__global__ void __launch_bounds__(SW_THREADS)
sw_rows_kernel(const uint8_t *__restrict__ seq, const uint64_t *__restrict__ seq_off,
               const int32_t *__restrict__ chr_start, int64_t n_feat, const uint32_t *__restrict__ f_ctg,
               const int32_t *__restrict__ f_start, const int32_t *__restrict__ f_end, int32_t size, int32_t maxd,
               int32_t resize, const uint64_t *__restrict__ row_off, uint64_t n_rows, int32_t *__restrict__ w_start,
               int32_t *__restrict__ w_end, int8_t *__restrict__ w_type, int32_t *__restrict__ w_dist,
               float *__restrict__ gc, float *__restrict__ gmean, float *__restrict__ gsd, float *__restrict__ gcv) {
    __shared__ uint64_t roff[SW_THREADS + 2];
    __shared__ int64_t f0_s;
    const uint64_t r0 = (uint64_t)blockIdx.x * SW_THREADS;
    if (threadIdx.x == 0) {
        int64_t lo = 0, hi = n_feat;  // last feature with row_off[f] <= r0
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (row_off[mid] <= r0) lo = mid; else hi = mid;
        }
        f0_s = lo;
    }
    __syncthreads();
    const int64_t f0 = f0_s;
    // every feature has at least one row, so rows r0 .. r0+255 belong to features f0 .. f0+255
    for (int i = threadIdx.x; i < SW_THREADS + 2; i += SW_THREADS) roff[i] = row_off[min(f0 + i, n_feat)];
    __syncthreads();
    const uint64_t row = r0 + threadIdx.x;
    if (row >= n_rows) return;
    int lo = 0, hi = SW_THREADS + 1;   // last i with roff[i] <= row
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (roff[mid] <= row) lo = mid; else hi = mid;
    }
    const int64_t f = f0 + lo;
    const int32_t k = (int32_t)(row - roff[lo]);
    const uint32_t c = f_ctg[f];
    const uint64_t so = seq_off[c];
    const int32_t psize = (int32_t)(seq_off[c + 1] - so);
    const int32_t p_start = chr_start[c];
    // index space of the parent: idx = chr - p_start + 1
    int32_t ms, me;
    center_resize_idx(psize, f_start[f] - p_start + 1, f_end[f] - p_start + 1, size, ms, me);
    int32_t nl = (ms - 1) / size;  // L windows need sw_start = ms - d*size >= 1
    if (nl > maxd) nl = maxd;
    if (nl < 0) nl = 0;
    int32_t ws, we, type, dist;
    if (k == 0) { ws = ms; we = me; type = 0; dist = 0; }
    else if (k <= nl) { dist = k; type = 1; we = ms - 1 - (k - 1) * size; ws = we - size + 1; }
    else { dist = k - nl; type = 2; ws = me + 1 + (dist - 1) * size; we = ws + size - 1; }
    const uint32_t cnt = thread_gc_count(seq, so + (uint64_t)(ws - 1), so + (uint64_t)we);
    const float g = round4(__fdiv_rn(__uint2float_rn(cnt), __uint2float_rn((uint32_t)(we - ws + 1))));

    // flank statistics: utils.rs:189-213
    int32_t rs, re;
    center_resize_idx(psize, ws, we, resize, rs, re);
    const int32_t span = re - rs + 1;
    const int32_t nsub = (span >= size) ? (span - size) / size + 1 : 0;
    const float fsz = __uint2float_rn((uint32_t)size);
    float gsub[SW_MAXSUB];
    float s = 0.0f;
    for (int32_t j = 0; j < nsub; ++j) {
        const uint64_t b0 = so + (uint64_t)(rs - 1) + (uint64_t)j * size;
        const float gj = round4(__fdiv_rn(__uint2float_rn(thread_gc_count(seq, b0, b0 + size)), fsz));
#pragma unroll
        for (int q = 0; q < SW_MAXSUB; ++q)
            if (q == j) gsub[q] = gj;
        s = __fadd_rn(s, gj);
    }
    const float flen = __uint2float_rn((uint32_t)nsub);
    const float m = __fdiv_rn(s, flen);
    float qq = 0.0f;
    if (nsub <= SW_MAXSUB) {
#pragma unroll
        for (int q = 0; q < SW_MAXSUB; ++q)
            if (q < nsub) {
                const float d = __fsub_rn(gsub[q], m);
                qq = __fadd_rn(qq, __fmul_rn(d, d));
            }
    } else {
        for (int32_t j = 0; j < nsub; ++j) {
            const uint64_t b0 = so + (uint64_t)(rs - 1) + (uint64_t)j * size;
            const float d = __fsub_rn(round4(__fdiv_rn(__uint2float_rn(thread_gc_count(seq, b0, b0 + size)), fsz)), m);
            qq = __fadd_rn(qq, __fmul_rn(d, d));
        }
    }
    const float sd = __fsqrt_rn(__fdiv_rn(qq, __fsub_rn(flen, 1.0f)));
    float cv;
    if (m == 0.0f || m == 1.0f) cv = 0.0f;
    else if (m <= 0.5f) cv = __fdiv_rn(sd, m);
    else cv = __fdiv_rn(sd, __fsub_rn(1.0f, m));

    w_start[row] = p_start + ws - 1;
    w_end[row] = p_start + we - 1;
    w_type[row] = (int8_t)type;
    w_dist[row] = dist;
    gc[row] = g;
    gmean[row] = round4(m);
    gsd[row] = round4(sd);
    gcv[row] = round4(cv);
}

int32_t launch_sw_rows(cudaStream_t st, const uint8_t *d_seq, const uint64_t *d_seq_off, const int32_t *d_chr_start,
                       int64_t n_feat, const uint32_t *d_f_ctg, const int32_t *d_f_start, const int32_t *d_f_end,
                       int32_t size, int32_t max, int32_t resize, const uint64_t *d_row_off, uint64_t n_rows,
                       int32_t *d_w_start, int32_t *d_w_end, int8_t *d_w_type, int32_t *d_w_dist, float *d_gc,
                       float *d_mean, float *d_sd, float *d_cv) {
    if (n_rows == 0) return GAMS_CUDA_OK;
    const uint64_t blocks = (n_rows + SW_THREADS - 1) / SW_THREADS;
    if (blocks >= (1ull << 31)) {
        set_error("sw_batch: too many rows for one launch");
        return GAMS_CUDA_EINVAL;
    }
    sw_rows_kernel<<<(unsigned)blocks, SW_THREADS, 0, st>>>(d_seq, d_seq_off, d_chr_start, n_feat, d_f_ctg, d_f_start, d_f_end,
                                                     size, max, resize, d_row_off, n_rows, d_w_start, d_w_end, d_w_type,
                                                     d_w_dist, d_gc, d_mean, d_sd, d_cv);
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

}  // namespace gams
