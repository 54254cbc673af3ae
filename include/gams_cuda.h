/*
 * gams_cuda.h — C ABI of libgams_cuda.so, the B200 (sm_100a) drop-in for the data-parallel
 * hot path of wang-q/gams.  Plain pointers and sizes only; no C++/torch types.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * reference root).  The Rust binding a gams maintainer would add is in INTEGRATION.md and
 * integration/gams-cuda-sys/.
 *
 * Conventions
 *  - return value: 0 = ok, negative = error (GAMS_CUDA_E*); gams_cuda_last_error() gives text
 *    (thread-local).  The library never aborts or throws across this boundary.
 *  - "host API" functions take HOST pointers owned by the caller; outputs are caller-allocated
 *    and fully overwritten; inputs are never modified.
 *  - "device API" functions (suffix _dev) take DEVICE pointers that live on the plan's device
 *    and a cudaStream_t passed as void*; they only enqueue work (no sync).
 *  - coordinates: a ctg is a byte string; window k of a ctg covers 0-based bytes
 *    [k*step, k*step+size) == 1-based chr positions chr_start+k*step .. chr_start+k*step+size-1
 *    (src/libs/window.rs:78-94, src/cmd_gams/wave.rs:146-150).
 *  - there is NO CPU fallback: every compute entry point fails with GAMS_CUDA_ENODEV/ECUDA
 *    when no sm_100 device is usable.
 */
#ifndef GAMS_CUDA_H
#define GAMS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAMS_CUDA_OK 0
#define GAMS_CUDA_EINVAL (-1) /* bad argument (null pointer, size/step <= 0, lag == 0, ...)        */
#define GAMS_CUDA_ECUDA (-2)  /* a CUDA runtime call or kernel failed; see gams_cuda_last_error()   */
#define GAMS_CUDA_ENOMEM (-3) /* host or device allocation failed                                   */
#define GAMS_CUDA_ECAP (-7)   /* caller-provided output arrays are too small; the needed count is reported        */
#define GAMS_CUDA_ELAG (-4)   /* a ctg has fewer windows than lag: the reference panics here        */
                              /* (src/libs/stat.rs:30 indexes data[0..lag])                          */
#define GAMS_CUDA_ENODEV (-5) /* no usable CUDA device                                              */
#define GAMS_CUDA_EPARSE (-6) /* malformed blob / text passed to a host-side helper                 */

typedef struct gams_cuda_ctx gams_cuda_ctx;
typedef struct gams_cuda_wave_plan gams_cuda_wave_plan;
typedef struct gams_cuda_count_plan gams_cuda_count_plan;
typedef struct gams_cuda_cwin_plan gams_cuda_cwin_plan;

/* ---------------------------------------------------------------- lifecycle ---------------- */
int32_t gams_cuda_version(void); /* 10000*major + 100*minor + patch */
const char *gams_cuda_last_error(void);
int32_t gams_cuda_device_count(void); /* >= 0, or negative error */

/* device_ids == NULL or n_dev <= 0: use every visible device.  One ctx may be shared by threads
 * (calls on the same ctx serialise on an internal mutex), cf. the per-worker Conn of
 * src/cmd_gams/wave.rs:288-299. */
int32_t gams_cuda_create(const int32_t *device_ids, int32_t n_dev, gams_cuda_ctx **out);
void gams_cuda_destroy(gams_cuda_ctx *ctx);
int32_t gams_cuda_ctx_num_devices(const gams_cuda_ctx *ctx);

/* Kernel B strategy for influence == 1 (results are identical in every mode):
 * 0 = certified filter, exact replay only where the certificate is inconclusive (default),
 * 1 = exact replay of every window, 2 = verify (replay everything and count certified decisions that
 * disagree with the replay; must stay 0).  Also settable with env GAMS_CUDA_ZSCORE_MODE at create time. */
int32_t gams_cuda_set_zscore_mode(gams_cuda_ctx *ctx, int32_t mode);
/* out2[0] = windows whose certificate was inconclusive (replayed exactly), out2[1] = verify mismatches,
 * both accumulated since creation / the last reset; synchronises the devices. */
int32_t gams_cuda_zscore_stats(gams_cuda_ctx *ctx, uint64_t *out2, int32_t reset);

/* Page-locked host buffers so the host API's copies run at PCIe speed (the analogue of the
 * Vec<u8> that Conn::get_seq fills, src/libs/redis.rs:142-146).  Plain malloc'd memory works
 * too, only slower. */
void *gams_cuda_host_alloc(size_t bytes);
void gams_cuda_host_free(void *p);
/* One plain copy of `bytes` between `host` and a device scratch buffer on device 0 of ctx, synchronous:
 * kind 1 = host->device, 2 = device->host (overwrites host!).  For measuring what the PCIe link gives. */
int32_t gams_cuda_memcpy_probe(gams_cuda_ctx *ctx, void *host, size_t bytes, int32_t kind);

/* ---------------------------------------------------------------- window arithmetic -------- */
/* gams::sliding, src/libs/window.rs:78-94: number of windows on a ctg of ctg_len bases.
 * Returns -1 for size <= 0 or step <= 0. */
int64_t gams_cuda_wave_nwin(int64_t ctg_len, int32_t size, int32_t step);
/* win_off[0] = 0, win_off[c+1] = win_off[c] + nwin(len_c).  seq_off has n_ctg+1 entries. */
int32_t gams_cuda_wave_win_offsets(int64_t n_ctg, const uint64_t *seq_off, int32_t size, int32_t step,
                                   uint64_t *win_off);
/* Contiguous, byte-balanced partition of ctgs over n_part shards (work shards by ctg,
 * no exchange).  part_off has n_part+1 entries: shard p owns ctgs [part_off[p], part_off[p+1]). */
int32_t gams_cuda_partition(int64_t n_item, const uint64_t *item_off, int32_t n_part, int64_t *part_off);

/* ---------------------------------------------------------------- GC-wave: host API -------- */
/* Replaces the body of wave::proc_ctg up to the signals, src/cmd_gams/wave.rs:138-155:
 *   gams::sliding (window.rs:78-94) + bio gc_content per window (wave.rs:143-153)
 *   + gams::thresholding_algo (stat.rs:16-56), for a whole batch of ctgs at once.
 *  seq_concat : all ctg sequences back to back (ASCII, mixed case, any byte)
 *  seq_off    : n_ctg+1 byte offsets into seq_concat
 *  win_off    : n_ctg+1 window offsets from gams_cuda_wave_win_offsets
 *  gc_count   : optional (may be NULL) #bytes in {c,g,C,G} per window          [win_off[n_ctg]]
 *  gc         : count as f32 / size as f32, one IEEE divide                      [win_off[n_ctg]]
 *  signal     : optional (NULL = GC only) -1/0/+1 per window, 0 for the first lag windows
 * Sharded over all devices of ctx by ctg; no inter-device exchange.
 * GAMS_CUDA_ELAG if signal != NULL and some ctg has fewer than lag windows. */
int32_t gams_cuda_wave_batch(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *seq_concat,
                             const uint64_t *seq_off, int32_t size, int32_t step, uint64_t lag,
                             float threshold, float influence, const uint64_t *win_off,
                             uint32_t *gc_count, float *gc, int8_t *signal);

/* The same with the feed step inside: Conn::get_seq (src/libs/redis.rs:142-146: GET seq:{ctg} -> decode_gz, redis.rs:156-161) +
 * wave::proc_ctg (wave.rs:138-155).  The ctgs arrive as the gzip blobs Redis holds (gz_concat / gz_off as gams_cuda_gunzip_batch
 * takes them), a third of the bytes of the bases; they are inflated on the device (ISIZE and CRC-32 checked) and never visit the
 * host.  seq_off: the ctg lengths as offsets (chr_end - chr_start + 1 of the ctg records, src/libs/data.rs:5-16), win_off from
 * them.  status[c]: inflate status of blob c (see gams_cuda_gunzip_batch); GAMS_CUDA_EINVAL if any blob failed. */
int32_t gams_cuda_wave_gz_batch(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *gz_concat, const uint64_t *gz_off,
                                const uint64_t *seq_off, int32_t size, int32_t step, uint64_t lag, float threshold,
                                float influence, const uint64_t *win_off, uint32_t *gc_count, float *gc, int8_t *signal,
                                int32_t *status);

/* `gams wave` without --signal, src/cmd_gams/wave.rs:138-212: the same computation as gams_cuda_wave_batch followed by
 * merge_ints (wave.rs:217-252) ON THE DEVICE, so that only the peaks cross PCIe (a few MB instead of 5 B per window).
 * One record per line wave prints, ctgs in input order, lines of a ctg in the reference's order:
 *   pk_ctg   : ctg index          pk_first / pk_last : first and last member window of the merged peak (pk_last == pk_first
 *   for an unmerged window; the printed range is chr_start + pk_first*step .. chr_start + pk_last*step + size - 1)
 *   pk_gc / pk_signal : GC fraction and signal of the FIRST member (what the reference prints, wave.rs:196-203)
 * coverage: wave's --coverage.  The device merge needs every intersecting pair of windows to pass the coverage test
 * (true for coverage <= 1, where size/intersect >= 1 always holds); otherwise GAMS_CUDA_EINVAL.
 * cap: room in the pk_* arrays; *n_peaks receives the number of records; GAMS_CUDA_ECAP if it exceeds cap. */
int32_t gams_cuda_wave_peaks_batch(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *seq_concat,
                                   const uint64_t *seq_off, int32_t size, int32_t step, uint64_t lag,
                                   float threshold, float influence, float coverage, const uint64_t *win_off,
                                   uint64_t cap, uint32_t *pk_ctg, uint32_t *pk_first, uint32_t *pk_last,
                                   float *pk_gc, int8_t *pk_signal, uint64_t *n_peaks);

/* The same from the seq:{ctg} values as Redis holds them (gams_cuda_wave_gz_batch's input): Conn::get_seq
 * (src/libs/redis.rs:142-161) + wave::proc_ctg without --signal (wave.rs:138-212) + merge_ints (wave.rs:217-252) in one call —
 * what `gams wave` does per ctg between its GET and its print.  A third of the bytes of the bases go up, the peak records come
 * back.  status[c]: inflate status of blob c; GAMS_CUDA_EINVAL if any blob failed. */
int32_t gams_cuda_wave_peaks_gz_batch(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *gz_concat, const uint64_t *gz_off,
                                      const uint64_t *seq_off, int32_t size, int32_t step, uint64_t lag, float threshold,
                                      float influence, float coverage, const uint64_t *win_off, uint64_t cap,
                                      uint32_t *pk_ctg, uint32_t *pk_first, uint32_t *pk_last, float *pk_gc,
                                      int8_t *pk_signal, uint64_t *n_peaks, int32_t *status);

/* `gams peak`, src/cmd_gams/peak.rs:66-160, for peaks sorted by (ctg, start): GC content of each peak range
 * (cache_gc_content, utils.rs:141-162: rounded to 4 decimals) and its links to the previous / next peak of the ctg:
 *   left_wave_length  = start - prev_end + 1   (prev_end = the ctg's chr_start for the first peak)
 *   left_amplitude    = |gc - prev_gc|         (0 for the first peak), left_signal = prev signal (own for the first)
 *   right_*           : the same towards the next peak, with chr_end for the last one.
 * p_start / p_end are 1-based inclusive chr positions inside the ctg. */
int32_t gams_cuda_peak_link_batch(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *seq_concat,
                                  const uint64_t *seq_off, const int32_t *ctg_chr_start, int64_t n_pk,
                                  const uint32_t *p_ctg, const int32_t *p_start, const int32_t *p_end,
                                  const int8_t *p_signal, float *gc, int8_t *left_signal, int32_t *left_wave_length,
                                  float *left_amplitude, int8_t *right_signal, int32_t *right_wave_length,
                                  float *right_amplitude);

/* gams::thresholding_algo on caller-supplied series (stat.rs:16-56), batched: series s is
 * data[off[s] .. off[s+1]).  Same semantics as the signal part of gams_cuda_wave_batch. */
int32_t gams_cuda_thresholding_batch(gams_cuda_ctx *ctx, int64_t n_series, const float *data,
                                     const uint64_t *off, uint64_t lag, float threshold, float influence,
                                     int8_t *signal);

/* ---------------------------------------------------------------- GC-wave: device API ------ */
/* A plan fixes the ctg layout (seq_off), the window geometry and lag on one device of ctx and
 * owns the small device-side tables; the bulk arrays stay with the caller. */
int32_t gams_cuda_wave_plan_create(gams_cuda_ctx *ctx, int32_t dev_index, int64_t n_ctg,
                                   const uint64_t *seq_off /* host */, int32_t size, int32_t step,
                                   uint64_t lag, gams_cuda_wave_plan **out);
void gams_cuda_wave_plan_destroy(gams_cuda_wave_plan *plan);
uint64_t gams_cuda_wave_plan_total_windows(const gams_cuda_wave_plan *plan);
uint64_t gams_cuda_wave_plan_total_bases(const gams_cuda_wave_plan *plan);
/* bytes the caller must make readable behind d_seq (total bases rounded up for 32-byte loads);
 * d_seq itself must be 32-byte aligned */
uint64_t gams_cuda_wave_plan_seq_capacity(const gams_cuda_wave_plan *plan);
/* Kernel A only: d_seq -> d_gc_count (optional) / d_gc. */
int32_t gams_cuda_wave_gc_dev(gams_cuda_wave_plan *plan, void *stream, const uint8_t *d_seq,
                              uint32_t *d_gc_count, float *d_gc);
/* Kernel B only: d_gc -> d_signal.  d_filtered: scratch of total_windows floats, needed only
 * when influence != 1 (may be NULL otherwise). */
int32_t gams_cuda_wave_signal_dev(gams_cuda_wave_plan *plan, void *stream, const float *d_gc,
                                  float threshold, float influence, float *d_filtered, int8_t *d_signal);
/* Kernels A+B in one call, src/cmd_gams/wave.rs:138-155: d_seq -> d_gc_count (optional) / d_gc / d_signal.
 * With the default statistics (influence == 1, threshold >= 0) and size % step == 0 this is ONE fused kernel
 * (gc_wave_fused.cu) that decides the signals from exact integer window counts while the tile is still in
 * shared memory; otherwise it enqueues gams_cuda_wave_gc_dev + gams_cuda_wave_signal_dev.  Results are
 * identical either way.  d_filtered as for gams_cuda_wave_signal_dev.  GAMS_CUDA_ELAG as for wave_batch.
 * (The fused kernel needs d_gc / d_gc_count 16-byte and d_signal 8-byte aligned; other pointers take the two kernels.) */
int32_t gams_cuda_wave_dev(gams_cuda_wave_plan *plan, void *stream, const uint8_t *d_seq, float threshold,
                           float influence, uint32_t *d_gc_count, float *d_gc, float *d_filtered,
                           int8_t *d_signal);
/* Peak merge (see gams_cuda_wave_peaks_batch) on device arrays: d_gc / d_signal as produced by gams_cuda_wave_dev;
 * d_pk_* have room for cap records; *d_n_peaks (device) receives the record count (records past cap are not written). */
int32_t gams_cuda_wave_peaks_dev(gams_cuda_wave_plan *plan, void *stream, const float *d_gc, const int8_t *d_signal,
                                 float coverage, uint64_t cap, uint32_t *d_pk_ctg, uint32_t *d_pk_first,
                                 uint32_t *d_pk_last, float *d_pk_gc, int8_t *d_pk_signal, uint64_t *d_n_peaks);
/* 1 if gams_cuda_wave_dev would take the fused kernel for these statistics, else 0. */
int32_t gams_cuda_wave_plan_fused(const gams_cuda_wave_plan *plan, float threshold, float influence);
/* Kernel launches issued by the plan's _dev calls since creation (bench.py's gpu_launches). */
uint64_t gams_cuda_wave_plan_launches(const gams_cuda_wave_plan *plan);

/* ---------------------------------------------------------------- anno --------------------- */
/* The proportion `gams anno` appends to each line, src/cmd_gams/anno.rs:128-139:
 *   prop = |set(chr) ∩ [ctg_start, ctg_end] ∩ [start, end]| as f32 / |[start, end]| as f32
 * set(chr): the runlist of the chromosome as sorted, disjoint, inclusive runs run_start/run_end[run_off[c] .. run_off[c+1]).
 * q_chr < 0 or >= n_chr (the chromosome is not in the runlist file) gives 0 (anno.rs:127). */
int32_t gams_cuda_anno_coverage(gams_cuda_ctx *ctx, int64_t n_chr, const int32_t *run_start, const int32_t *run_end,
                                const uint64_t *run_off, int64_t n_q, const int32_t *q_chr, const int32_t *q_ctg_start,
                                const int32_t *q_ctg_end, const int32_t *q_start, const int32_t *q_end, float *prop);

/* ---------------------------------------------------------------- gen ---------------------- */
/* The region finding of `gams gen`, src/cmd_gams/gen.rs:80-127: bases other than A C G T a c g t are ambiguous
 * (gen.rs:86-93); the valid regions are the complement with holes of less than `fill` bases filled (IntSpan::fill(fill-1),
 * gen.rs:103) and regions shorter than `min_len` dropped (excise, gen.rs:104); each region is cut into pieces of `piece`
 * bases with the remainder joined to the last piece (gen.rs:107-126).  The per-base scan runs on the device; the run list
 * it returns (one entry per valid run) is processed on the host.
 *  seq_concat / chr_off : the chromosomes back to back, n_chr+1 byte offsets
 *  ctg_chr / ctg_start / ctg_end : per ctg its chromosome index and 1-based inclusive chr positions, in gen's order
 *  cap : room in the output arrays; *n_ctg receives the number of ctgs; GAMS_CUDA_ECAP if it exceeds cap. */
int32_t gams_cuda_gen_ctgs(gams_cuda_ctx *ctx, int64_t n_chr, const uint8_t *seq_concat, const uint64_t *chr_off,
                           int32_t fill, int32_t min_len, int32_t piece, uint64_t cap, uint32_t *ctg_chr,
                           int32_t *ctg_start, int32_t *ctg_end, uint64_t *n_ctg);

/* ---------------------------------------------------------------- GC of ranges / sw -------- */
/* Replaces gams::cache_gc_content without the rounding, src/libs/utils.rs:141-162 (bio
 * gc_content of an arbitrary sub-range of a ctg): query q covers 0-based bytes
 * [q_from0[q], q_to0_excl[q]) of ctg q_ctg[q].  gc_count[q] = #bytes in {c,g,C,G}. */
int32_t gams_cuda_gc_ranges(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *seq_concat,
                            const uint64_t *seq_off, int64_t n_q, const uint32_t *q_ctg,
                            const uint32_t *q_from0, const uint32_t *q_to0_excl, uint32_t *gc_count);

/* Replaces the per-feature body of sw::proc_ctg, src/cmd_gams/sw.rs:141-191 (action "gc"):
 * gams::center_sw (window.rs:3-56) windows around each feature, per window
 * cache_gc_content (utils.rs:141-162) and cache_gc_stat of the center_resize'd flank
 * (utils.rs:189-213, window.rs:96-124), all rounded to 4 dp as the reference does.
 *  ctg_chr_start : n_ctg 1-based chr_start of each ctg (chr_end = chr_start + len - 1)
 *  f_ctg/f_start/f_end : feature's ctg and 1-based inclusive chr range
 *  row_off : n_feat+1, out; feature f owns rows [row_off[f], row_off[f+1]) (at most 1+2*max each)
 *  out arrays sized n_feat*(1+2*max): w_start/w_end (chr coords), w_type (0=M 1=L 2=R), w_dist,
 *  gc, gc_mean, gc_stddev, gc_cv.  Rows are compacted to the front (row_off indexes them). */
int32_t gams_cuda_sw_batch(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *seq_concat,
                           const uint64_t *seq_off, const int32_t *ctg_chr_start, int64_t n_feat,
                           const uint32_t *f_ctg, const int32_t *f_start, const int32_t *f_end,
                           int32_t size, int32_t max, int32_t resize, uint64_t *row_off,
                           int32_t *w_start, int32_t *w_end, int8_t *w_type, int32_t *w_dist,
                           float *gc, float *gc_mean, float *gc_stddev, float *gc_cv);

/* `gams sw --action gc --action count` (src/cmd_gams/sw.rs:22-35): the rows of gams_cuda_sw_batch plus the rg_count column of
 * gams::Sw (src/libs/data.rs:55, printed last by its Display, data.rs:72-82).  The reference accepts the action and leaves the
 * column empty (sw.rs:155-184 never fills rg_count); what the action names is gams::count_rg (src/libs/utils.rs:24-36) of every
 * generated window against the rg index of its ctg, called as locate.rs:136 calls it:
 *   rg_count[row] = lower_bound(starts_c, w_end) - lower_bound(stops_c, w_start + 1)
 * starts / stops / iv_off: the rg index as gams_cuda_count_overlaps takes it, segment c = ctg c (an empty segment counts 0).
 * Features sorted by ctg (the order sw walks them in, sw.rs:95-99, 141) take the staged kernel; any other order works too. */
int32_t gams_cuda_sw_count_batch(gams_cuda_ctx *ctx, int64_t n_ctg, const uint8_t *seq_concat,
                                 const uint64_t *seq_off, const int32_t *ctg_chr_start, int64_t n_feat,
                                 const uint32_t *f_ctg, const int32_t *f_start, const int32_t *f_end,
                                 int32_t size, int32_t max, int32_t resize, const uint32_t *starts,
                                 const uint32_t *stops, const uint64_t *iv_off, uint64_t *row_off,
                                 int32_t *w_start, int32_t *w_end, int8_t *w_type, int32_t *w_dist,
                                 float *gc, float *gc_mean, float *gc_stddev, float *gc_cv, uint32_t *rg_count);

/* ---------------------------------------------------------------- overlap counting --------- */
/* Replaces the count loop of locate::execute, src/cmd_gams/locate.rs:111-141, i.e.
 * gams::count_rg (src/libs/utils.rs:24-36) -> rust_lapper::Lapper::count(start, stop):
 *   out[q] = lower_bound(starts_seg, q_stop[q]) - lower_bound(stops_seg, q_start[q] + 1)
 * where lower_bound(a, key) = #elements < key.  starts/stops are the sorted endpoint arrays that
 * sit verbatim inside the idx:rg:{ctg} bincode blob (src/libs/redis.rs:276-305); segment s (one
 * ctg's index) is [iv_off[s], iv_off[s+1]).  gams passes (rg.start, rg.end) — the inclusive end
 * as Lapper's exclusive stop (utils.rs:35); callers pass the same values here.
 * q_seg[q] < 0 or >= n_seg -> out[q] = 0 ("not found in idx", utils.rs:29-32).
 * Queries may come in any order; results are in query order. */
int32_t gams_cuda_count_overlaps(gams_cuda_ctx *ctx, int64_t n_seg, const uint32_t *starts,
                                 const uint32_t *stops, const uint64_t *iv_off, int64_t n_q,
                                 const int32_t *q_seg, const uint32_t *q_start, const uint32_t *q_stop,
                                 uint32_t *out);

/* Device API: the index stays resident in HBM across query batches (as the lapper BTreeMap stays
 * resident across the loop of locate.rs:104-141). */
int32_t gams_cuda_count_plan_create(gams_cuda_ctx *ctx, int32_t dev_index, int64_t n_seg,
                                    const uint64_t *iv_off /* host */, gams_cuda_count_plan **out);
void gams_cuda_count_plan_destroy(gams_cuda_count_plan *plan);
/* d_starts/d_stops: device copies of the concatenated endpoint arrays.  grouped != 0 promises
 * q_seg is non-decreasing (queries grouped by ctg, the order read_range's BTreeMap yields,
 * utils.rs:39-79); grouped == 0 accepts any order. */
int32_t gams_cuda_count_overlaps_dev(gams_cuda_count_plan *plan, void *stream, const uint32_t *d_starts,
                                     const uint32_t *d_stops, int64_t n_q, const int32_t *d_q_seg,
                                     const uint32_t *d_q_start, const uint32_t *d_q_stop,
                                     int32_t grouped, uint32_t *d_out);
uint64_t gams_cuda_count_plan_launches(const gams_cuda_count_plan *plan);
/* The plan keeps the tables it derives from an index (rank nodes, compact per-segment tables) per (d_starts, d_stops)
 * pointer pair; call this after rewriting the index in place under the same pointers. */
void gams_cuda_count_plan_invalidate(gams_cuda_count_plan *plan);

/* Conn::get_seq's decode_gz (src/libs/redis.rs:142-161: flate2 GzDecoder over the `seq:{ctg}` value), batched, on the
 * device: one gzip member (RFC 1952, deflate RFC 1951: stored / fixed / dynamic blocks) per blob, blobs back to back in
 * gz_concat at gz_off[n_blob + 1].  Blob i inflates to out_concat[out_off[i] .. out_off[i+1]); the caller knows the sizes
 * (ctg lengths) or reads them from the trailers with gams_cuda_gz_sizes.  Like GzDecoder only the first member of a blob is
 * read.  ISIZE, the expected size and the CRC-32 of the trailer are checked, as GzDecoder checks them (decode_gz(..).unwrap()
 * panics on a mismatch, redis.rs:144); gams_cuda_set_gunzip_options can switch the CRC pass off.
 * status[i]: 0 ok, 1 bad gzip header, 2 invalid code / block, 3 output larger than expected, 4 distance before the start,
 * 5 size mismatch, 6 truncated input, 7 CRC-32 mismatch.  Returns GAMS_CUDA_EINVAL if any blob failed (the good ones are
 * still written). */
int32_t gams_cuda_gz_sizes(int64_t n_blob, const uint8_t *gz_concat, const uint64_t *gz_off, uint64_t *out_len);
int32_t gams_cuda_gunzip_batch(gams_cuda_ctx *ctx, int64_t n_blob, const uint8_t *gz_concat, const uint64_t *gz_off,
                               const uint64_t *out_off, uint8_t *out_concat, int32_t *status);
/* decoder: 0 = two-phase (one thread per blob turns the Huffman stream into tokens, one warp per blob turns tokens into bytes;
 * default), 1 = one warp per blob in a single pass (the round-1 kernel; also what a blob falls back to when it compresses worse
 * than 2 bytes per deflate symbol).  check_crc: 1 = verify the CRC-32 (default), 0 = skip that pass.  Results are identical. */
int32_t gams_cuda_set_gunzip_options(gams_cuda_ctx *ctx, int32_t decoder, int32_t check_crc);
/* Device API: all pointers device memory (offsets relative to d_gz / d_out), enqueued on `stream` (which is synchronised once
 * at entry: the total output size comes back to size the decoder's scratch; one call per context device at a time).  The reader loads
 * aligned 32-bit words: the allocation behind d_gz must be readable from d_gz rounded down to a multiple of 4 up to the
 * end rounded up to a multiple of 4 (any cudaMalloc'ed buffer is). */
int32_t gams_cuda_gunzip_dev(gams_cuda_ctx *ctx, int32_t dev_index, void *stream, int64_t n_blob, const uint8_t *d_gz,
                             const uint64_t *d_gz_off, uint8_t *d_out, const uint64_t *d_out_off, int32_t *d_status);

/* rg counts along sliding windows: what the reference computes when gams::sliding (src/libs/window.rs:78-94) is
 * walked and gams::count_rg (src/libs/utils.rs:24-36) is asked for every window -- the "sliding" / "fsw ... rg counting"
 * workload of BASELINE configs[1].  Window k of ctg c is [ws, we] = [chr_start[c] + k*step, ws + size - 1];
 *   out[win_off[c] + k] = lower_bound(starts_c, we) - lower_bound(stops_c, ws + 1)
 * exactly as gams_cuda_count_overlaps answers the query (seg c, ws, we), but with no query arrays at all: the keys are
 * an arithmetic progression, so each tile of windows is one histogram + prefix sum over its slice of the index.
 * win_off: as gams_cuda_wave_win_offsets returns it (n_ctg + 1); the index of ctg c is segment c of starts/stops/iv_off
 * (an empty segment counts 0 everywhere). */
int32_t gams_cuda_count_windows(gams_cuda_ctx *ctx, int64_t n_ctg, const int32_t *chr_start, const uint64_t *win_off,
                                int32_t size, int32_t step, const uint32_t *starts, const uint32_t *stops,
                                const uint64_t *iv_off, uint32_t *out);
/* Device API: geometry and offsets are bound in a plan (host arrays, copied); the index and the counts stay in HBM. */
int32_t gams_cuda_cwin_plan_create(gams_cuda_ctx *ctx, int32_t dev_index, int64_t n_ctg, const int32_t *chr_start,
                                   const uint64_t *win_off, int32_t size, int32_t step, const uint64_t *iv_off,
                                   gams_cuda_cwin_plan **out);
void gams_cuda_cwin_plan_destroy(gams_cuda_cwin_plan *plan);
uint64_t gams_cuda_cwin_plan_total_windows(const gams_cuda_cwin_plan *plan);
uint64_t gams_cuda_cwin_plan_launches(const gams_cuda_cwin_plan *plan);
/* The per-tile slices of the index are found once per (d_starts, d_stops) pair and kept in the plan; call this after
 * rewriting the index in place under the same pointers. */
void gams_cuda_cwin_plan_invalidate(gams_cuda_cwin_plan *plan);
int32_t gams_cuda_count_windows_dev(gams_cuda_cwin_plan *plan, void *stream, const uint32_t *d_starts,
                                    const uint32_t *d_stops, uint32_t *d_out);

/* gams::build_idx_rg (src/libs/redis.rs:276-305): rust_lapper::Lapper::new + bincode::serialize per ctg, on the
 * device.  Input: the intervals of segment s (one ctg) are start[i], stop[i] for i in [iv_off[s], iv_off[s+1]),
 * in any order, stop = rg.end + 1 as redis.rs:291-295 builds them.  Output, segment by segment at the same offsets:
 *   iv_start/iv_stop  the intervals sorted by (start, stop)            (Lapper.intervals; val is "" for rg)
 *   starts, stops     each endpoint array sorted on its own            (Lapper.starts / Lapper.stops: what
 *                                                                        gams_cuda_count_overlaps consumes)
 *   max_len[s]        max(stop - start), 0 where stop < start          (Lapper.max_len)
 *   blob (optional)   the bincode 1.3 bytes of every Lapper back to back, segment s at byte 24*iv_off[s] + 30*s,
 *                     24*n_s + 30 bytes each (layout: SURVEY.md section 8 a14) -- what `SET idx:rg:{ctg}` stores.
 * At most 2^31 - 1 intervals per call. */
int32_t gams_cuda_build_idx_rg(gams_cuda_ctx *ctx, int64_t n_seg, const uint32_t *start, const uint32_t *stop,
                               const uint64_t *iv_off, uint32_t *iv_start, uint32_t *iv_stop, uint32_t *starts,
                               uint32_t *stops, uint32_t *max_len, uint8_t *blob /* NULL: no blobs */);

/* The front end of `gams locate` (src/cmd_gams/locate.rs:111-117) and read_range (src/libs/utils.rs:39-52): the lines of
 * a range file parsed on the device instead of one intspan::Range::from_str (a regex) per line.  The whole line is what
 * read_range hands to the regex (utils.rs:49); `locate -f` hands it the first TAB-separated field only (locate.rs:87-90) —
 * for lines of the strict form both give the same range (the regex cannot run past a TAB).
 * text: n_bytes, lines end in '\n' (a '\r' before it is dropped, as BufRead::lines does; a last line without '\n' counts).
 * chr names: n_chr names back to back in chr_names, name c = [chr_name_off[c], chr_name_off[c+1]).
 * Per line i (0 .. *n_lines): line_off[i] (byte offset; line_off[*n_lines] = n_bytes), q_start / q_end (1-based inclusive) and
 *   q_chr[i] >= 0  index of the chromosome in the name table
 *            -1   a valid range on a chromosome that is not in the table
 *            -2   not a valid range (Range::from_str(..).is_valid() is false: no ':', start == 0, end < start) - skipped
 *            -3   outside the strict form  [name.]chr[(+)|(-)]:start[-end|_end]  anchored at the start of the line: the
 *                 regex might still match somewhere; the caller passes exactly these lines to Range::from_str.
 * Output arrays hold cap_lines entries (line_off: cap_lines + 1); GAMS_CUDA_ECAP with *n_lines set if that is too few. */
int32_t gams_cuda_parse_ranges(gams_cuda_ctx *ctx, const uint8_t *text, uint64_t n_bytes, int64_t n_chr,
                               const uint8_t *chr_names, const uint32_t *chr_name_off, uint64_t cap_lines,
                               uint64_t *n_lines, uint64_t *line_off, int32_t *q_chr, uint32_t *q_start, uint32_t *q_end);

/* The whole loop of `gams locate --count -f file` (src/cmd_gams/locate.rs:80-141) on the device, text in, text out:
 * rg = the first TAB-separated field of each line (locate.rs:87-90: `line.split('\t').first()`), then Range::from_str(rg) ->
 * find_one_idx -> count_rg -> "{rg}\t{cnt}\n" for every rg that is a valid range on a known ctg, in input order (locate.rs:139).
 * Only the first field is parsed and only the first field is echoed, so multi-column files (gams' own .tsv outputs, header
 * rows such as "#range\tgc\tsignal") give exactly the reference's lines.
 * ctg index: as gams_cuda_locate_ctg takes it (ctgs of chromosome c = [chr_off[c], chr_off[c+1]), sorted by start, stop
 * exclusive); rg index: segment k of starts / stops / iv_off belongs to ctg k of that same order.
 * First fields outside the strict form (see gams_cuda_parse_ranges) are NOT printed; *n_handed_back counts them.  A caller that
 * needs byte parity on such input uses gams_cuda_locate_count_lines: it also returns the input line index of every printed
 * line, so the lines the caller resolves itself (gams_cuda_parse_ranges marks them -3 when given the same first fields) merge
 * back in input order.
 * *out_text: malloc'ed, NUL-terminated, *out_len bytes; free with gams_cuda_free. */
int32_t gams_cuda_locate_count_text(gams_cuda_ctx *ctx, const uint8_t *text, uint64_t n_bytes, int64_t n_chr,
                                    const uint8_t *chr_names, const uint32_t *chr_name_off, const uint32_t *ctg_start,
                                    const uint32_t *ctg_stop, const uint64_t *chr_off, const uint32_t *starts,
                                    const uint32_t *stops, const uint64_t *iv_off, char **out_text, size_t *out_len,
                                    uint64_t *n_lines, uint64_t *n_printed, uint64_t *n_handed_back);
/* The same, plus *out_line: malloc'ed, *n_printed entries, out_line[k] = 0-based input line of the k-th printed line
 * (free with gams_cuda_free). */
int32_t gams_cuda_locate_count_lines(gams_cuda_ctx *ctx, const uint8_t *text, uint64_t n_bytes, int64_t n_chr,
                                     const uint8_t *chr_names, const uint32_t *chr_name_off, const uint32_t *ctg_start,
                                     const uint32_t *ctg_stop, const uint64_t *chr_off, const uint32_t *starts,
                                     const uint32_t *stops, const uint64_t *iv_off, char **out_text, size_t *out_len,
                                     uint64_t **out_line, uint64_t *n_lines, uint64_t *n_printed, uint64_t *n_handed_back);

/* gams::find_one_idx (utils.rs:7-22) -> Lapper::find(start, stop).next() on the idx:ctg:{chr}
 * index: ctg intervals [chr_start, chr_end+1) sorted by start within each chromosome segment.
 * out[q] = global index of the first ctg (in start order) with ctg_start < q_stop && ctg_stop >
 * q_start, or -1 (unknown chr or no hit: the query is skipped, locate.rs:118-122). */
int32_t gams_cuda_locate_ctg(gams_cuda_ctx *ctx, int64_t n_chr, const uint32_t *ctg_start,
                             const uint32_t *ctg_stop, const uint64_t *chr_off, int64_t n_q,
                             const int32_t *q_chr, const uint32_t *q_start, const uint32_t *q_stop,
                             int32_t *out);

/* ---------------------------------------------------------------- host-side format helpers -- */
/* Zero-copy view of the sorted endpoint arrays inside a bincode-serialised
 * rust_lapper::Lapper<u32,String> (src/libs/redis.rs:299-300, 316-317).  starts and stops point
 * into blob (little-endian u32, possibly unaligned for 4-byte access on exotic hosts). */
int32_t gams_cuda_lapper_blob_view(const uint8_t *blob, size_t blob_len, uint64_t *n_iv,
                                   const uint8_t **starts, const uint8_t **stops, uint32_t *max_len);

/* `gams wave` text for one ctg, src/cmd_gams/wave.rs:157-212 incl. merge_ints (wave.rs:217-252):
 * given gc/signal of the ctg's n windows, append the lines wave prints (no header) to a
 * malloc'd, NUL-terminated buffer returned in *text (free with gams_cuda_free).  is_signal != 0
 * selects the --signal form (every window, unmerged). */
int32_t gams_cuda_wave_text(const char *chr_id, int32_t chr_start, int32_t size, int32_t step, int64_t n,
                            const float *gc, const int8_t *signal, float coverage, int32_t is_signal,
                            char **text, size_t *text_len);
/* The same text from the records of gams_cuda_wave_peaks_batch (the records of ONE ctg, in order). */
int32_t gams_cuda_wave_peaks_text(const char *chr_id, int32_t chr_start, int32_t size, int32_t step, int64_t n_pk,
                                  const uint32_t *pk_first, const uint32_t *pk_last, const float *pk_gc,
                                  const int8_t *pk_signal, char **text, size_t *text_len);
void gams_cuda_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* GAMS_CUDA_H */
