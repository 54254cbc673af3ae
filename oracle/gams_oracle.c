/*
 * gams_oracle.c — CPU restatement of the gams hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity checker and the timed CPU baseline.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product (libgams_cuda.so) never links, loads or calls it.
 *
 * The reference (wang-q/gams, Rust) cannot be compiled in this image (no cargo/rustc,
 * crates not vendored), so this is a "port": plain C that follows the reference's
 * loops one for one — same asymptotics (O(size) byte scan per window, three O(lag)
 * passes per z-score step, two binary searches per overlap query), same f32
 * evaluation order, no FMA contraction (build with -ffp-contract=off, see Makefile).
 *
 * Parity pinned by: tests/S288c/I.peaks.tsv (byte-exact, through oracle/oracle.py),
 * src/libs/stat.rs:58-81, src/libs/window.rs:58-76,126-147, tests/cli.rs:426-454,
 * tests/cli.rs:483-495.  Third-party arithmetic restated from the published crates:
 * bio 1.5.0 seq_analysis::gc::gc_content, rust-lapper 1.1.0 Lapper::{new,count,find},
 * intspan 0.7.7 IntSpan::{index,at,slice} for single-span parents.
 * UNPINNED (no reference test discriminates it): the half-open edge of Lapper::count
 * as gams calls it and the sw numeric columns.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* window.rs:78-94  gams::sliding — number of windows on a span of `len`      */
/* start=1; loop { end=start+size-1; if end>len break; push; start+=step }    */
/* ------------------------------------------------------------------------- */
ORC_API int64_t gorc_nwin(int64_t len, int32_t size, int32_t step) {
    int64_t n = 0;
    if (size <= 0 || step <= 0) return -1; /* reference loops forever / misbehaves */
    /* closed form of the loop above (the loop itself is in gorc_wave_ctg) */
    if (len >= size) n = (len - size) / step + 1;
    return n;
}

/* ------------------------------------------------------------------------- */
/* bio 1.5.0 seq_analysis::gc::gc_content: fold over bytes, count c g C G,    */
/* return count as f32 / len as f32.  Every other byte (N, n, ...) only adds  */
/* to the denominator.  Call sites wave.rs:150-151, utils.rs:155-157.         */
/* ------------------------------------------------------------------------- */
static inline uint64_t gc_count_scan(const uint8_t *s, size_t n) {
    uint64_t c = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t b = s[i];
        c += (b == 'c') | (b == 'g') | (b == 'G') | (b == 'C');
    }
    return c;
}

ORC_API float gorc_gc_content(const uint8_t *s, size_t n, uint32_t *count_out) {
    uint64_t c = gc_count_scan(s, n);
    if (count_out) *count_out = (uint32_t)c;
    return (float)c / (float)n; /* n == 0 -> 0/0 = NaN, as in the reference */
}

/* ------------------------------------------------------------------------- */
/* stat.rs:1-6 mean, stat.rs:8-14 stddev — strictly sequential f32            */
/* ------------------------------------------------------------------------- */
ORC_API float gorc_mean(const float *d, size_t n) {
    float len = (float)n;
    float sum = 0.0f; /* iter().sum::<f32>() starts from 0.0 and adds left to right */
    for (size_t i = 0; i < n; i++) sum = sum + d[i];
    return sum / len;
}

ORC_API float gorc_stddev(const float *d, size_t n) {
    float len = (float)n;
    float m = gorc_mean(d, n);
    float sq = 0.0f;
    for (size_t i = 0; i < n; i++) {
        float t = d[i] - m;
        float p = t * t; /* separate multiply ... */
        sq = sq + p;     /* ... and add: Rust never contracts to FMA */
    }
    return sqrtf(sq / (len - 1.0f));
}

/* ------------------------------------------------------------------------- */
/* stat.rs:16-56 thresholding_algo.  Returns 0, or -1 where the reference     */
/* would panic (data[0..lag] out of range, stat.rs:30; lag == 0 underflows    */
/* lag-1).  signals are i32 in the reference; int8 here (values -1/0/1).      */
/* ------------------------------------------------------------------------- */
ORC_API int gorc_thresholding(const float *data, size_t n, size_t lag, float threshold,
                              float influence, int8_t *signals) {
    if (lag == 0 || n < lag) return -1;
    float *filtered = (float *)malloc(sizeof(float) * (n ? n : 1));
    float *avg = (float *)calloc(n ? n : 1, sizeof(float));
    float *sd = (float *)calloc(n ? n : 1, sizeof(float));
    if (!filtered || !avg || !sd) { free(filtered); free(avg); free(sd); return -2; }
    memcpy(filtered, data, sizeof(float) * n);
    memset(signals, 0, n);

    avg[lag - 1] = gorc_mean(data, lag);
    sd[lag - 1] = gorc_stddev(data, lag);

    for (size_t i = lag; i < n; i++) {
        if (fabsf(data[i] - avg[i - 1]) > threshold * sd[i - 1]) {
            signals[i] = (data[i] > avg[i - 1]) ? 1 : -1;
            filtered[i] = influence * data[i] + (1.0f - influence) * filtered[i - 1];
        } else {
            signals[i] = 0;
            filtered[i] = data[i];
        }
        avg[i] = gorc_mean(filtered + (i - lag), lag);
        sd[i] = gorc_stddev(filtered + (i - lag), lag);
    }
    free(filtered); free(avg); free(sd);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* wave.rs:138-155 for one ctg: sliding windows -> per-window GC -> signals.  */
/* Window k covers 0-based bytes [k*step, k*step+size) of the ctg sequence    */
/* (parent.index(window.min())-1 .. parent.index(window.max()), wave.rs:146-150). */
/* ------------------------------------------------------------------------- */
ORC_API int gorc_wave_ctg(const uint8_t *seq, int64_t len, int32_t size, int32_t step,
                          uint64_t lag, float threshold, float influence,
                          uint32_t *cnt /* may be NULL */, float *gc, int8_t *signal) {
    int64_t n = 0;
    int64_t start = 1; /* 1-based index on the parent, as window.rs:81 */
    for (;;) {
        int64_t end = start + size - 1;
        if (end > len) break;
        uint32_t c;
        gc[n] = gorc_gc_content(seq + (start - 1), (size_t)(end - start + 1), &c);
        if (cnt) cnt[n] = c;
        n++;
        start += step;
    }
    if (signal) return gorc_thresholding(gc, (size_t)n, (size_t)lag, threshold, influence, signal);
    return 0;
}

/* ---- thread pool over ctgs: the reference's --parallel N model (wave.rs:255-313) ---- */
typedef struct {
    int64_t n_ctg;
    const uint8_t *seq;
    const uint64_t *seq_off;
    int32_t size, step;
    uint64_t lag;
    float threshold, influence;
    const uint64_t *win_off;
    uint32_t *cnt;
    float *gc;
    int8_t *signal;
    volatile int64_t next;
    volatile int err;
} wave_job;

static void *wave_worker(void *p) {
    wave_job *j = (wave_job *)p;
    for (;;) {
        int64_t c = __sync_fetch_and_add(&j->next, 1);
        if (c >= j->n_ctg) break;
        uint64_t wo = j->win_off[c];
        int r = gorc_wave_ctg(j->seq + j->seq_off[c], (int64_t)(j->seq_off[c + 1] - j->seq_off[c]),
                              j->size, j->step, j->lag, j->threshold, j->influence,
                              j->cnt ? j->cnt + wo : NULL, j->gc + wo, j->signal ? j->signal + wo : NULL);
        if (r) j->err = r;
    }
    return NULL;
}

/* Batched form with the same argument meaning as gams_cuda_wave_batch (include/gams_cuda.h). */
ORC_API int gorc_wave_batch(int64_t n_ctg, const uint8_t *seq, const uint64_t *seq_off, int32_t size,
                            int32_t step, uint64_t lag, float threshold, float influence,
                            const uint64_t *win_off, uint32_t *cnt, float *gc, int8_t *signal,
                            int32_t n_threads) {
    wave_job j = {n_ctg, seq, seq_off, size, step, lag, threshold, influence, win_off, cnt, gc, signal, 0, 0};
    if (n_threads <= 1) {
        wave_worker(&j);
        return j.err;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, wave_worker, &j);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(th);
    return j.err;
}

/* ------------------------------------------------------------------------- */
/* rust-lapper 1.1.0: Lapper::bsearch_seq (branch-light lower bound on a       */
/* sorted slice) and Lapper::count (BITS: two searches on starts[] / stops[]). */
/* gams call site utils.rs:35: lapper.count(rg.start as u32, rg.end as u32).   */
/* ------------------------------------------------------------------------- */
static inline size_t bsearch_seq(uint32_t key, const uint32_t *e, size_t n) {
    if (n == 0 || e[0] >= key) return 0;
    if (e[n - 1] < key) return n;
    size_t cursor = 0, length = n;
    while (length > 1) {
        size_t half = length >> 1;
        length -= half;
        cursor += (size_t)(e[cursor + half - 1] < key) * half;
    }
    return cursor;
}

ORC_API uint32_t gorc_lapper_count(const uint32_t *starts, const uint32_t *stops, uint64_t n,
                                   uint32_t start, uint32_t stop) {
    size_t len = (size_t)n;
    size_t first = bsearch_seq(start + 1u, stops, len);
    size_t last = bsearch_seq(stop, starts, len);
    size_t num_cant_after = len - last;
    return (uint32_t)(len - first - num_cant_after);
}

/* locate.rs:111-141 inner loop, with the ctg already located: per query a segment     */
/* (ctg) id, then count_rg (utils.rs:24-36).  q_seg < 0 or >= n_seg -> 0 ("not found   */
/* in idx", utils.rs:29-32).                                                            */
typedef struct {
    int64_t n_seg;
    const uint32_t *starts, *stops;
    const uint64_t *iv_off;
    int64_t n_q;
    const int32_t *q_seg;
    const uint32_t *q_start, *q_stop;
    uint32_t *out;
    volatile int64_t next;
} count_job;

#define COUNT_CHUNK 65536
static void *count_worker(void *p) {
    count_job *j = (count_job *)p;
    for (;;) {
        int64_t b = __sync_fetch_and_add(&j->next, COUNT_CHUNK);
        if (b >= j->n_q) break;
        int64_t e = b + COUNT_CHUNK < j->n_q ? b + COUNT_CHUNK : j->n_q;
        for (int64_t q = b; q < e; q++) {
            int32_t s = j->q_seg[q];
            if (s < 0 || s >= j->n_seg) { j->out[q] = 0; continue; }
            uint64_t o = j->iv_off[s];
            j->out[q] = gorc_lapper_count(j->starts + o, j->stops + o, j->iv_off[s + 1] - o,
                                          j->q_start[q], j->q_stop[q]);
        }
    }
    return NULL;
}

ORC_API int gorc_count_batch(int64_t n_seg, const uint32_t *starts, const uint32_t *stops,
                             const uint64_t *iv_off, int64_t n_q, const int32_t *q_seg,
                             const uint32_t *q_start, const uint32_t *q_stop, uint32_t *out,
                             int32_t n_threads) {
    count_job j = {n_seg, starts, stops, iv_off, n_q, q_seg, q_start, q_stop, out, 0};
    if (n_threads <= 1) { count_worker(&j); return 0; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, count_worker, &j);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(th);
    return 0;
}

/* rust-lapper 1.1.0 Lapper::find(start, stop).next() on intervals sorted by (start, stop):   */
/* off = lower_bound(start - max_len (saturating), intervals by .start); scan forward,          */
/* return first iv with iv.start < stop && iv.stop > start; stop when iv.start >= stop.         */
/* Returns the index into the sorted interval array, or -1.  Call site utils.rs:16.             */
ORC_API int64_t gorc_lapper_find_first(const uint32_t *iv_start, const uint32_t *iv_stop, uint64_t n,
                                       uint32_t max_len, uint32_t start, uint32_t stop) {
    uint32_t key = start >= max_len ? start - max_len : 0;
    size_t lo = 0, hi = (size_t)n; /* lower_bound over iv.start */
    while (lo < hi) {
        size_t mid = lo + ((hi - lo) >> 1);
        if (iv_start[mid] < key) lo = mid + 1; else hi = mid;
    }
    for (size_t off = lo; off < n; off++) {
        if (iv_start[off] < stop && iv_stop[off] > start) return (int64_t)off;
        if (iv_start[off] >= stop) break;
    }
    return -1;
}

/* ------------------------------------------------------------------------- */
/* window.rs:96-124 center_resize and window.rs:3-56 center_sw for a single-  */
/* span parent [p_start, p_end] (wave.rs:138 / sw.rs:129 build exactly that). */
/* intspan on one span: size = p_end-p_start+1, index(x) = x-p_start+1,        */
/* at(i) = p_start+i-1 (i>=1), slice(a,b) = [at(a), at(b)].                     */
/* The child `intspan` passed to center_resize is also a single span [s,e].    */
/* ------------------------------------------------------------------------- */
ORC_API void gorc_center_resize(int32_t p_start, int32_t p_end, int32_t s, int32_t e, int32_t resize,
                                int32_t *out_s, int32_t *out_e) {
    int32_t isize = e - s + 1;
    int32_t half_size = isize / 2;
    int32_t mid_left = (half_size == 0) ? s : s + half_size - 1;       /* at(half_size)   */
    int32_t mid_right = (half_size == 0) ? s : s + (half_size + 1) - 1; /* at(half_size+1) */
    int32_t mid_left_idx = mid_left - p_start + 1;
    int32_t mid_right_idx = mid_right - p_start + 1;
    int32_t psize = p_end - p_start + 1;
    int32_t half_resize = resize / 2;
    int32_t left_idx = mid_left_idx - half_resize + 1;
    if (left_idx < 1) left_idx = 1;
    int32_t right_idx = mid_right_idx + half_resize - 1;
    if (right_idx > psize) right_idx = psize;
    *out_s = p_start + left_idx - 1;
    *out_e = p_start + right_idx - 1;
}

/* Emits up to 1+2*max windows as (start, end, type, distance); type 0=M 1=L 2=R.  */
/* Returns the number of windows. Order: M, L1..Lk, R1..Rk (window.rs:13-53).      */
ORC_API int32_t gorc_center_sw(int32_t p_start, int32_t p_end, int32_t start, int32_t end, int32_t size,
                               int32_t max, int32_t *w_start, int32_t *w_end, int32_t *w_type,
                               int32_t *w_dist) {
    int32_t n = 0, m_s, m_e;
    int32_t psize = p_end - p_start + 1;
    gorc_center_resize(p_start, p_end, start, end, size, &m_s, &m_e);
    w_start[n] = m_s; w_end[n] = m_e; w_type[n] = 0; w_dist[n] = 0; n++;
    for (int t = 1; t <= 2; t++) { /* "L" then "R" */
        int32_t sw_start, sw_end;
        if (t == 2) { sw_start = (m_e - p_start + 1) + 1; sw_end = sw_start + size - 1; }
        else        { sw_end = (m_s - p_start + 1) - 1;   sw_start = sw_end - size + 1; }
        for (int32_t d = 1; d <= max; d++) {
            if (sw_start < 1) break;
            if (sw_end > psize) break;
            /* slice(sw_start, sw_end) on one span always has sw_end-sw_start+1 == size */
            w_start[n] = p_start + sw_start - 1; w_end[n] = p_start + sw_end - 1;
            w_type[n] = t; w_dist[n] = d; n++;
            if (t == 2) { sw_start = sw_end + 1; sw_end = sw_start + size - 1; }
            else        { sw_end = sw_start - 1; sw_start = sw_end - size + 1; }
        }
    }
    return n;
}

/* utils.rs:135-138 round: (x * 10^d as f32).round() / 10^d ; f32::round = half away from zero */
ORC_API float gorc_round(float x, uint32_t decimals) {
    int32_t p = 1;
    for (uint32_t i = 0; i < decimals; i++) p *= 10;
    float y = (float)p;
    return roundf(x * y) / y;
}

/* utils.rs:141-162 cache_gc_content without the memo (the memo only caches the same value): */
/* GC of 1-based inclusive chr range [rs, re] on a ctg starting at p_start, rounded to 4 dp.    */
ORC_API float gorc_gc_range4(const uint8_t *seq, int32_t p_start, int32_t rs, int32_t re) {
    size_t from = (size_t)(rs - p_start + 1), to = (size_t)(re - p_start + 1);
    float g = gorc_gc_content(seq + (from - 1), to - (from - 1), NULL);
    return gorc_round(g, 4);
}

/* utils.rs:164-187 gc_stat */
ORC_API void gorc_gc_stat(const float *gcs, size_t n, float *mean, float *sd, float *cv) {
    float m = gorc_mean(gcs, n);
    float s = gorc_stddev(gcs, n);
    float c;
    if (m == 0.0f || m == 1.0f) c = 0.0f;
    else if (m <= 0.5f) c = s / m;
    else c = s / (1.0f - m);
    *mean = gorc_round(m, 4); *sd = gorc_round(s, 4); *cv = gorc_round(c, 4);
}

/* utils.rs:189-213 cache_gc_stat: non-overlapping `size` windows (sliding(size, step)) inside */
/* the chr range [rs, re]; each GC rounded to 4 dp; then gc_stat.                               */
ORC_API void gorc_gc_stat_range(const uint8_t *seq, int32_t p_start, int32_t rs, int32_t re, int32_t size,
                                int32_t step, float *mean, float *sd, float *cv) {
    int32_t span = re - rs + 1;
    int64_t n = gorc_nwin(span, size, step);
    float *g = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    int64_t k = 0;
    for (int32_t st = 1; st + size - 1 <= span; st += step) {
        int32_t ws = rs + st - 1, we = ws + size - 1;
        g[k++] = gorc_gc_range4(seq, p_start, ws, we);
    }
    gorc_gc_stat(g, (size_t)k, mean, sd, cv);
    free(g);
}

/* sw.rs:141-191 for one feature [f_start, f_end] on one ctg, action "gc":                  */
/* rows = center_sw windows; per row gc_content (4 dp) and gc_stat of the resized flank.     */
/* out arrays sized 1+2*max.  Returns row count.                                             */
ORC_API int32_t gorc_sw_feature(const uint8_t *seq, int32_t p_start, int32_t p_end, int32_t f_start,
                                int32_t f_end, int32_t size, int32_t max, int32_t resize,
                                int32_t *w_start, int32_t *w_end, int32_t *w_type, int32_t *w_dist,
                                float *gc, float *gc_mean, float *gc_sd, float *gc_cv) {
    int32_t n = gorc_center_sw(p_start, p_end, f_start, f_end, size, max, w_start, w_end, w_type, w_dist);
    for (int32_t i = 0; i < n; i++) {
        gc[i] = gorc_gc_range4(seq, p_start, w_start[i], w_end[i]);
        int32_t rs, re;
        gorc_center_resize(p_start, p_end, w_start[i], w_end[i], resize, &rs, &re);
        gorc_gc_stat_range(seq, p_start, rs, re, size, size, &gc_mean[i], &gc_sd[i], &gc_cv[i]);
    }
    return n;
}
