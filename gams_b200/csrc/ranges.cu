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

__global__ void __launch_bounds__(256)
sw_rows_kernel(const uint8_t *__restrict__ seq, const uint64_t *__restrict__ seq_off,
               const int32_t *__restrict__ chr_start, int64_t n_feat, const uint32_t *__restrict__ f_ctg,
               const int32_t *__restrict__ f_start, const int32_t *__restrict__ f_end, int32_t size, int32_t maxd,
               int32_t resize, const uint64_t *__restrict__ row_off, uint64_t n_rows, int32_t *__restrict__ w_start,
               int32_t *__restrict__ w_end, int8_t *__restrict__ w_type, int32_t *__restrict__ w_dist,
               float *__restrict__ gc, float *__restrict__ gmean, float *__restrict__ gsd, float *__restrict__ gcv) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t row = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += warps) {
        int64_t lo = 0, hi = n_feat;  // last feature with row_off[f] <= row
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (row_off[mid] <= row) lo = mid; else hi = mid;
        }
        const int64_t f = lo;
        const int32_t k = (int32_t)(row - row_off[f]);
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

        const uint32_t cnt = warp_gc_count(seq, so + (uint64_t)(ws - 1), so + (uint64_t)we, lane);
        const float g = round4(__fdiv_rn(__uint2float_rn(cnt), __uint2float_rn((uint32_t)(we - ws + 1))));

        // flank statistics: utils.rs:189-213
        int32_t rs, re;
        center_resize_idx(psize, ws, we, resize, rs, re);
        const int32_t span = re - rs + 1;
        const int32_t nsub = (span >= size) ? (span - size) / size + 1 : 0;
        const float fsz = __uint2float_rn((uint32_t)size);
        float s = 0.0f;
        for (int32_t j = 0; j < nsub; ++j) {
            const uint64_t b0 = so + (uint64_t)(rs - 1) + (uint64_t)j * size;
            const uint32_t cj = warp_gc_count(seq, b0, b0 + size, lane);
            s = __fadd_rn(s, round4(__fdiv_rn(__uint2float_rn(cj), fsz)));
        }
        const float flen = __uint2float_rn((uint32_t)nsub);
        const float m = __fdiv_rn(s, flen);
        float qq = 0.0f;
        for (int32_t j = 0; j < nsub; ++j) {
            const uint64_t b0 = so + (uint64_t)(rs - 1) + (uint64_t)j * size;
            const uint32_t cj = warp_gc_count(seq, b0, b0 + size, lane);
            const float d = __fsub_rn(round4(__fdiv_rn(__uint2float_rn(cj), fsz)), m);
            qq = __fadd_rn(qq, __fmul_rn(d, d));
        }
        const float sd = __fsqrt_rn(__fdiv_rn(qq, __fsub_rn(flen, 1.0f)));
        float cv;
        if (m == 0.0f || m == 1.0f) cv = 0.0f;
        else if (m <= 0.5f) cv = __fdiv_rn(sd, m);
        else cv = __fdiv_rn(sd, __fsub_rn(1.0f, m));

        if (lane == 0) {
            w_start[row] = p_start + ws - 1;
            w_end[row] = p_start + we - 1;
            w_type[row] = (int8_t)type;
            w_dist[row] = dist;
            gc[row] = g;
            gmean[row] = round4(m);
            gsd[row] = round4(sd);
            gcv[row] = round4(cv);
        }
    }
}

int32_t launch_sw_rows(cudaStream_t st, const uint8_t *d_seq, const uint64_t *d_seq_off, const int32_t *d_chr_start,
                       int64_t n_feat, const uint32_t *d_f_ctg, const int32_t *d_f_start, const int32_t *d_f_end,
                       int32_t size, int32_t max, int32_t resize, const uint64_t *d_row_off, uint64_t n_rows,
                       int32_t *d_w_start, int32_t *d_w_end, int8_t *d_w_type, int32_t *d_w_dist, float *d_gc,
                       float *d_mean, float *d_sd, float *d_cv) {
    if (n_rows == 0) return GAMS_CUDA_OK;
    uint64_t blocks = (n_rows + 7) / 8;
    if (blocks > 148ull * 64) blocks = 148ull * 64;
    sw_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_seq, d_seq_off, d_chr_start, n_feat, d_f_ctg, d_f_start, d_f_end,
                                                     size, max, resize, d_row_off, n_rows, d_w_start, d_w_end, d_w_type,
                                                     d_w_dist, d_gc, d_mean, d_sd, d_cv);
    CUDA_TRY(cudaGetLastError());
    return GAMS_CUDA_OK;
}

}  // namespace gams
