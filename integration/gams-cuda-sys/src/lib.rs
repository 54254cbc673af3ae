//! Raw bindings to `include/gams_cuda.h` plus the small safe wrappers the gams subcommands call.
//! One declaration per C entry point; argument meaning is documented in the header.
#![allow(non_camel_case_types)]

use std::ffi::CStr;
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct gams_cuda_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct gams_cuda_wave_plan {
    _private: [u8; 0],
}
#[repr(C)]
pub struct gams_cuda_count_plan {
    _private: [u8; 0],
}
#[repr(C)]
pub struct gams_cuda_cwin_plan {
    _private: [u8; 0],
}

pub const GAMS_CUDA_OK: i32 = 0;
pub const GAMS_CUDA_EINVAL: i32 = -1;
pub const GAMS_CUDA_ECUDA: i32 = -2;
pub const GAMS_CUDA_ENOMEM: i32 = -3;
pub const GAMS_CUDA_ELAG: i32 = -4;
pub const GAMS_CUDA_ENODEV: i32 = -5;
pub const GAMS_CUDA_EPARSE: i32 = -6;
/// Caller-provided output arrays are too small; the needed count is reported (call again with larger arrays).
pub const GAMS_CUDA_ECAP: i32 = -7;

extern "C" {
    pub fn gams_cuda_version() -> i32;
    pub fn gams_cuda_last_error() -> *const c_char;
    pub fn gams_cuda_device_count() -> i32;
    pub fn gams_cuda_create(device_ids: *const i32, n_dev: i32, out: *mut *mut gams_cuda_ctx) -> i32;
    pub fn gams_cuda_destroy(ctx: *mut gams_cuda_ctx);
    pub fn gams_cuda_ctx_num_devices(ctx: *const gams_cuda_ctx) -> i32;
    pub fn gams_cuda_host_alloc(bytes: usize) -> *mut c_void;
    pub fn gams_cuda_host_free(p: *mut c_void);
    pub fn gams_cuda_memcpy_probe(ctx: *mut gams_cuda_ctx, host: *mut c_void, bytes: usize, kind: i32) -> i32;
    pub fn gams_cuda_set_zscore_mode(ctx: *mut gams_cuda_ctx, mode: i32) -> i32;
    pub fn gams_cuda_zscore_stats(ctx: *mut gams_cuda_ctx, out2: *mut u64, reset: i32) -> i32;
    pub fn gams_cuda_set_gunzip_options(ctx: *mut gams_cuda_ctx, decoder: i32, check_crc: i32) -> i32;

    pub fn gams_cuda_wave_nwin(ctg_len: i64, size: i32, step: i32) -> i64;
    pub fn gams_cuda_wave_win_offsets(n_ctg: i64, seq_off: *const u64, size: i32, step: i32, win_off: *mut u64) -> i32;
    pub fn gams_cuda_partition(n_item: i64, item_off: *const u64, n_part: i32, part_off: *mut i64) -> i32;

    pub fn gams_cuda_wave_batch(
        ctx: *mut gams_cuda_ctx, n_ctg: i64, seq_concat: *const u8, seq_off: *const u64, size: i32, step: i32,
        lag: u64, threshold: f32, influence: f32, win_off: *const u64, gc_count: *mut u32, gc: *mut f32,
        signal: *mut i8,
    ) -> i32;
    pub fn gams_cuda_wave_gz_batch(
        ctx: *mut gams_cuda_ctx, n_ctg: i64, gz_concat: *const u8, gz_off: *const u64, seq_off: *const u64, size: i32,
        step: i32, lag: u64, threshold: f32, influence: f32, win_off: *const u64, gc_count: *mut u32, gc: *mut f32,
        signal: *mut i8, status: *mut i32,
    ) -> i32;
    pub fn gams_cuda_thresholding_batch(
        ctx: *mut gams_cuda_ctx, n_series: i64, data: *const f32, off: *const u64, lag: u64, threshold: f32,
        influence: f32, signal: *mut i8,
    ) -> i32;

    pub fn gams_cuda_wave_plan_create(
        ctx: *mut gams_cuda_ctx, dev_index: i32, n_ctg: i64, seq_off: *const u64, size: i32, step: i32, lag: u64,
        out: *mut *mut gams_cuda_wave_plan,
    ) -> i32;
    pub fn gams_cuda_wave_plan_destroy(plan: *mut gams_cuda_wave_plan);
    pub fn gams_cuda_wave_plan_total_windows(plan: *const gams_cuda_wave_plan) -> u64;
    pub fn gams_cuda_wave_plan_total_bases(plan: *const gams_cuda_wave_plan) -> u64;
    pub fn gams_cuda_wave_plan_seq_capacity(plan: *const gams_cuda_wave_plan) -> u64;
    pub fn gams_cuda_wave_gc_dev(
        plan: *mut gams_cuda_wave_plan, stream: *mut c_void, d_seq: *const u8, d_gc_count: *mut u32, d_gc: *mut f32,
    ) -> i32;
    pub fn gams_cuda_wave_signal_dev(
        plan: *mut gams_cuda_wave_plan, stream: *mut c_void, d_gc: *const f32, threshold: f32, influence: f32,
        d_filtered: *mut f32, d_signal: *mut i8,
    ) -> i32;
    pub fn gams_cuda_wave_dev(
        plan: *mut gams_cuda_wave_plan, stream: *mut c_void, d_seq: *const u8, threshold: f32, influence: f32,
        d_gc_count: *mut u32, d_gc: *mut f32, d_filtered: *mut f32, d_signal: *mut i8,
    ) -> i32;
    pub fn gams_cuda_wave_plan_fused(plan: *const gams_cuda_wave_plan, threshold: f32, influence: f32) -> i32;
    pub fn gams_cuda_wave_plan_launches(plan: *const gams_cuda_wave_plan) -> u64;

    pub fn gams_cuda_wave_peaks_batch(
        ctx: *mut gams_cuda_ctx, n_ctg: i64, seq_concat: *const u8, seq_off: *const u64, size: i32, step: i32, lag: u64,
        threshold: f32, influence: f32, coverage: f32, win_off: *const u64, cap: u64, pk_ctg: *mut u32, pk_first: *mut u32,
        pk_last: *mut u32, pk_gc: *mut f32, pk_signal: *mut i8, n_peaks: *mut u64,
    ) -> i32;
    pub fn gams_cuda_wave_peaks_gz_batch(
        ctx: *mut gams_cuda_ctx, n_ctg: i64, gz_concat: *const u8, gz_off: *const u64, seq_off: *const u64, size: i32, step: i32,
        lag: u64, threshold: f32, influence: f32, coverage: f32, win_off: *const u64, cap: u64, pk_ctg: *mut u32,
        pk_first: *mut u32, pk_last: *mut u32, pk_gc: *mut f32, pk_signal: *mut i8, n_peaks: *mut u64, status: *mut i32,
    ) -> i32;
    pub fn gams_cuda_wave_peaks_dev(
        plan: *mut gams_cuda_wave_plan, stream: *mut c_void, d_gc: *const f32, d_signal: *const i8, coverage: f32, cap: u64,
        d_pk_ctg: *mut u32, d_pk_first: *mut u32, d_pk_last: *mut u32, d_pk_gc: *mut f32, d_pk_signal: *mut i8,
        d_n_peaks: *mut u64,
    ) -> i32;
    pub fn gams_cuda_peak_link_batch(
        ctx: *mut gams_cuda_ctx, n_ctg: i64, seq_concat: *const u8, seq_off: *const u64, ctg_chr_start: *const i32, n_pk: i64,
        p_ctg: *const u32, p_start: *const i32, p_end: *const i32, p_signal: *const i8, gc: *mut f32, left_signal: *mut i8,
        left_wave_length: *mut i32, left_amplitude: *mut f32, right_signal: *mut i8, right_wave_length: *mut i32,
        right_amplitude: *mut f32,
    ) -> i32;
    pub fn gams_cuda_anno_coverage(
        ctx: *mut gams_cuda_ctx, n_chr: i64, run_start: *const i32, run_end: *const i32, run_off: *const u64, n_q: i64,
        q_chr: *const i32, q_ctg_start: *const i32, q_ctg_end: *const i32, q_start: *const i32, q_end: *const i32,
        prop: *mut f32,
    ) -> i32;
    pub fn gams_cuda_gen_ctgs(
        ctx: *mut gams_cuda_ctx, n_chr: i64, seq_concat: *const u8, chr_off: *const u64, fill: i32, min_len: i32, piece: i32,
        cap: u64, ctg_chr: *mut u32, ctg_start: *mut i32, ctg_end: *mut i32, n_ctg: *mut u64,
    ) -> i32;

    pub fn gams_cuda_gc_ranges(
        ctx: *mut gams_cuda_ctx, n_ctg: i64, seq_concat: *const u8, seq_off: *const u64, n_q: i64, q_ctg: *const u32,
        q_from0: *const u32, q_to0_excl: *const u32, gc_count: *mut u32,
    ) -> i32;
    pub fn gams_cuda_sw_batch(
        ctx: *mut gams_cuda_ctx, n_ctg: i64, seq_concat: *const u8, seq_off: *const u64, ctg_chr_start: *const i32,
        n_feat: i64, f_ctg: *const u32, f_start: *const i32, f_end: *const i32, size: i32, max: i32, resize: i32,
        row_off: *mut u64, w_start: *mut i32, w_end: *mut i32, w_type: *mut i8, w_dist: *mut i32, gc: *mut f32,
        gc_mean: *mut f32, gc_stddev: *mut f32, gc_cv: *mut f32,
    ) -> i32;

    pub fn gams_cuda_sw_count_batch(
        ctx: *mut gams_cuda_ctx, n_ctg: i64, seq_concat: *const u8, seq_off: *const u64, ctg_chr_start: *const i32,
        n_feat: i64, f_ctg: *const u32, f_start: *const i32, f_end: *const i32, size: i32, max: i32, resize: i32,
        starts: *const u32, stops: *const u32, iv_off: *const u64, row_off: *mut u64, w_start: *mut i32, w_end: *mut i32,
        w_type: *mut i8, w_dist: *mut i32, gc: *mut f32, gc_mean: *mut f32, gc_stddev: *mut f32, gc_cv: *mut f32,
        rg_count: *mut u32,
    ) -> i32;

    pub fn gams_cuda_count_overlaps(
        ctx: *mut gams_cuda_ctx, n_seg: i64, starts: *const u32, stops: *const u32, iv_off: *const u64, n_q: i64,
        q_seg: *const i32, q_start: *const u32, q_stop: *const u32, out: *mut u32,
    ) -> i32;
    pub fn gams_cuda_count_plan_create(
        ctx: *mut gams_cuda_ctx, dev_index: i32, n_seg: i64, iv_off: *const u64, out: *mut *mut gams_cuda_count_plan,
    ) -> i32;
    pub fn gams_cuda_count_plan_destroy(plan: *mut gams_cuda_count_plan);
    pub fn gams_cuda_count_overlaps_dev(
        plan: *mut gams_cuda_count_plan, stream: *mut c_void, d_starts: *const u32, d_stops: *const u32, n_q: i64,
        d_q_seg: *const i32, d_q_start: *const u32, d_q_stop: *const u32, grouped: i32, d_out: *mut u32,
    ) -> i32;
    pub fn gams_cuda_count_plan_launches(plan: *const gams_cuda_count_plan) -> u64;
    pub fn gams_cuda_count_plan_invalidate(plan: *mut gams_cuda_count_plan);
    pub fn gams_cuda_gz_sizes(n_blob: i64, gz_concat: *const u8, gz_off: *const u64, out_len: *mut u64) -> i32;
    pub fn gams_cuda_gunzip_batch(
        ctx: *mut gams_cuda_ctx, n_blob: i64, gz_concat: *const u8, gz_off: *const u64, out_off: *const u64,
        out_concat: *mut u8, status: *mut i32,
    ) -> i32;
    pub fn gams_cuda_gunzip_dev(
        ctx: *mut gams_cuda_ctx, dev_index: i32, stream: *mut c_void, n_blob: i64, d_gz: *const u8, d_gz_off: *const u64,
        d_out: *mut u8, d_out_off: *const u64, d_status: *mut i32,
    ) -> i32;
    pub fn gams_cuda_count_windows(
        ctx: *mut gams_cuda_ctx, n_ctg: i64, chr_start: *const i32, win_off: *const u64, size: i32, step: i32,
        starts: *const u32, stops: *const u32, iv_off: *const u64, out: *mut u32,
    ) -> i32;
    pub fn gams_cuda_cwin_plan_create(
        ctx: *mut gams_cuda_ctx, dev_index: i32, n_ctg: i64, chr_start: *const i32, win_off: *const u64, size: i32,
        step: i32, iv_off: *const u64, out: *mut *mut gams_cuda_cwin_plan,
    ) -> i32;
    pub fn gams_cuda_cwin_plan_destroy(plan: *mut gams_cuda_cwin_plan);
    pub fn gams_cuda_cwin_plan_total_windows(plan: *const gams_cuda_cwin_plan) -> u64;
    pub fn gams_cuda_cwin_plan_launches(plan: *const gams_cuda_cwin_plan) -> u64;
    pub fn gams_cuda_cwin_plan_invalidate(plan: *mut gams_cuda_cwin_plan);
    pub fn gams_cuda_count_windows_dev(
        plan: *mut gams_cuda_cwin_plan, stream: *mut c_void, d_starts: *const u32, d_stops: *const u32, d_out: *mut u32,
    ) -> i32;
    pub fn gams_cuda_parse_ranges(
        ctx: *mut gams_cuda_ctx, text: *const u8, n_bytes: u64, n_chr: i64, chr_names: *const u8, chr_name_off: *const u32,
        cap_lines: u64, n_lines: *mut u64, line_off: *mut u64, q_chr: *mut i32, q_start: *mut u32, q_end: *mut u32,
    ) -> i32;
    pub fn gams_cuda_locate_count_text(
        ctx: *mut gams_cuda_ctx, text: *const u8, n_bytes: u64, n_chr: i64, chr_names: *const u8, chr_name_off: *const u32,
        ctg_start: *const u32, ctg_stop: *const u32, chr_off: *const u64, starts: *const u32, stops: *const u32,
        iv_off: *const u64, out_text: *mut *mut c_char, out_len: *mut usize, n_lines: *mut u64, n_printed: *mut u64,
        n_handed_back: *mut u64,
    ) -> i32;
    pub fn gams_cuda_locate_count_lines(
        ctx: *mut gams_cuda_ctx, text: *const u8, n_bytes: u64, n_chr: i64, chr_names: *const u8, chr_name_off: *const u32,
        ctg_start: *const u32, ctg_stop: *const u32, chr_off: *const u64, starts: *const u32, stops: *const u32,
        iv_off: *const u64, out_text: *mut *mut c_char, out_len: *mut usize, out_line: *mut *mut u64, n_lines: *mut u64,
        n_printed: *mut u64, n_handed_back: *mut u64,
    ) -> i32;
    pub fn gams_cuda_build_idx_rg(
        ctx: *mut gams_cuda_ctx, n_seg: i64, start: *const u32, stop: *const u32, iv_off: *const u64,
        iv_start: *mut u32, iv_stop: *mut u32, starts: *mut u32, stops: *mut u32, max_len: *mut u32, blob: *mut u8,
    ) -> i32;
    pub fn gams_cuda_locate_ctg(
        ctx: *mut gams_cuda_ctx, n_chr: i64, ctg_start: *const u32, ctg_stop: *const u32, chr_off: *const u64,
        n_q: i64, q_chr: *const i32, q_start: *const u32, q_stop: *const u32, out: *mut i32,
    ) -> i32;

    pub fn gams_cuda_lapper_blob_view(
        blob: *const u8, blob_len: usize, n_iv: *mut u64, starts: *mut *const u8, stops: *mut *const u8,
        max_len: *mut u32,
    ) -> i32;
    pub fn gams_cuda_wave_text(
        chr_id: *const c_char, chr_start: i32, size: i32, step: i32, n: i64, gc: *const f32, signal: *const i8,
        coverage: f32, is_signal: i32, text: *mut *mut c_char, text_len: *mut usize,
    ) -> i32;
    pub fn gams_cuda_wave_peaks_text(
        chr_id: *const c_char, chr_start: i32, size: i32, step: i32, n_pk: i64, pk_first: *const u32, pk_last: *const u32,
        pk_gc: *const f32, pk_signal: *const i8, text: *mut *mut c_char, text_len: *mut usize,
    ) -> i32;
    pub fn gams_cuda_free(p: *mut c_void);
}

/// Error text of the last failing call on this thread.
pub fn last_error() -> String {
    unsafe { CStr::from_ptr(gams_cuda_last_error()).to_string_lossy().into_owned() }
}

/// Owning handle; `Drop` destroys the context.
pub struct Ctx(pub *mut gams_cuda_ctx);
unsafe impl Send for Ctx {}

impl Ctx {
    /// All visible devices.
    pub fn new() -> Result<Ctx, String> {
        let mut p = std::ptr::null_mut();
        match unsafe { gams_cuda_create(std::ptr::null(), 0, &mut p) } {
            0 => Ok(Ctx(p)),
            _ => Err(last_error()),
        }
    }

    /// Batched replacement of `gams::sliding` + `gc_content` + `gams::thresholding_algo`
    /// (src/cmd_gams/wave.rs:138-155) → (win_off, gc, signal).
    pub fn wave_batch(
        &self, seq_concat: &[u8], seq_off: &[u64], size: i32, step: i32, lag: usize, threshold: f32, influence: f32,
    ) -> Result<(Vec<u64>, Vec<f32>, Vec<i8>), String> {
        let n_ctg = (seq_off.len() - 1) as i64;
        let mut win_off = vec![0u64; seq_off.len()];
        if unsafe { gams_cuda_wave_win_offsets(n_ctg, seq_off.as_ptr(), size, step, win_off.as_mut_ptr()) } != 0 {
            return Err(last_error());
        }
        let n = *win_off.last().unwrap() as usize;
        let (mut gc, mut signal) = (vec![0f32; n], vec![0i8; n]);
        let rc = unsafe {
            gams_cuda_wave_batch(
                self.0, n_ctg, seq_concat.as_ptr(), seq_off.as_ptr(), size, step, lag as u64, threshold, influence,
                win_off.as_ptr(), std::ptr::null_mut(), gc.as_mut_ptr(), signal.as_mut_ptr(),
            )
        };
        if rc != 0 {
            return Err(last_error());
        }
        Ok((win_off, gc, signal))
    }

    /// Batched replacement of `gams::count_rg` (src/libs/utils.rs:24-36).
    pub fn count_overlaps(
        &self, starts: &[u32], stops: &[u32], iv_off: &[u64], q_seg: &[i32], q_start: &[u32], q_stop: &[u32],
    ) -> Result<Vec<u32>, String> {
        let mut out = vec![0u32; q_seg.len()];
        let rc = unsafe {
            gams_cuda_count_overlaps(
                self.0, (iv_off.len() - 1) as i64, starts.as_ptr(), stops.as_ptr(), iv_off.as_ptr(),
                q_seg.len() as i64, q_seg.as_ptr(), q_start.as_ptr(), q_stop.as_ptr(), out.as_mut_ptr(),
            )
        };
        if rc != 0 {
            return Err(last_error());
        }
        Ok(out)
    }
}

impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { gams_cuda_destroy(self.0) }
    }
}
