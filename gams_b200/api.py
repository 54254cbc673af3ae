"""Host-side mirror of the reference's `gams::` interface for the hot path, over the C ABI.

Names and argument meaning follow the reference (src/libs/window.rs, stat.rs, utils.rs and the
proc_ctg bodies of src/cmd_gams/{wave,sw,locate}.rs); arrays are numpy, coordinates 1-based
inclusive chr positions unless a name ends in 0.  Every function runs on the GPU through
libgams_cuda.so; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from ._lib import GamsCudaError, check, load  # noqa: F401


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Context:
    """gams_cuda_ctx: the devices this process computes on (one per process under torchrun)."""

    def __init__(self, devices=None):
        L = load()
        h = C.c_void_p()
        if devices is None:
            check(L.gams_cuda_create(None, 0, C.byref(h)))
        else:
            ids = _arr(devices, np.int32)
            check(L.gams_cuda_create(_p(ids), ids.size, C.byref(h)))
        self._h = h
        self._L = L
        self._plans = weakref.WeakSet()   # plans keep pointers into the context: they are closed before it goes

    @property
    def handle(self):
        return self._h

    @property
    def num_devices(self) -> int:
        return int(self._L.gams_cuda_ctx_num_devices(self._h))

    def set_zscore_mode(self, mode: int):
        """0 = certified filter + exact replay (default), 1 = replay every window, 2 = verify."""
        check(self._L.gams_cuda_set_zscore_mode(self._h, mode))

    def set_gunzip_options(self, decoder=0, check_crc=True):
        """decoder 0 = two-phase (default), 1 = one warp per blob; check_crc: verify the gzip trailer's CRC-32 as GzDecoder does."""
        check(self._L.gams_cuda_set_gunzip_options(self._h, int(decoder), int(bool(check_crc))))

    def zscore_stats(self, reset=True):
        """(windows replayed because the certificate was inconclusive, verify mismatches)."""
        out = np.zeros(2, dtype=np.uint64)
        check(self._L.gams_cuda_zscore_stats(self._h, _p(out), int(reset)))
        return int(out[0]), int(out[1])

    def close(self):
        if getattr(self, "_h", None):
            for plan in list(self._plans):
                plan.close()
            self._L.gams_cuda_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------ window.rs:78-94
def sliding_count(length: int, size: int, step: int) -> int:
    """Number of windows gams::sliding yields on a span of `length` bases."""
    n = int(load().gams_cuda_wave_nwin(length, size, step))
    if n < 0:
        raise ValueError("size and step must be positive")
    return n


def sliding(chr_start: int, chr_end: int, size: int, step: int):
    """gams::sliding on the single-span parent [chr_start, chr_end] → list of (start, end)."""
    n = sliding_count(chr_end - chr_start + 1, size, step)
    return [(chr_start + k * step, chr_start + k * step + size - 1) for k in range(n)]


def win_offsets(seq_off, size: int, step: int) -> np.ndarray:
    seq_off = _arr(seq_off, np.uint64)
    out = np.zeros(seq_off.size, dtype=np.uint64)
    check(load().gams_cuda_wave_win_offsets(seq_off.size - 1, _p(seq_off), size, step, _p(out)))
    return out


def partition(item_off, n_part: int) -> np.ndarray:
    item_off = _arr(item_off, np.uint64)
    out = np.zeros(n_part + 1, dtype=np.int64)
    check(load().gams_cuda_partition(item_off.size - 1, _p(item_off), n_part, _p(out)))
    return out


# --------------------------------------------------------------- wave.rs:138-155 (batched)
def wave_batch(ctx: Context, seq_concat, seq_off, size=100, step=10, lag=100, threshold=3.0, influence=1.0,
               want_count=True, want_signal=True):
    """sliding + gc_content + thresholding_algo for a batch of ctgs → (win_off, gc_count, gc, signal)."""
    seq_concat = _arr(seq_concat, np.uint8)
    seq_off = _arr(seq_off, np.uint64)
    woff = win_offsets(seq_off, size, step)
    n = int(woff[-1])
    cnt = np.empty(n, dtype=np.uint32) if want_count else None
    gc = np.empty(n, dtype=np.float32)
    sig = np.empty(n, dtype=np.int8) if want_signal else None
    check(load().gams_cuda_wave_batch(ctx.handle, seq_off.size - 1, _p(seq_concat), _p(seq_off), size, step, lag,
                                      threshold, influence, _p(woff), _p(cnt), _p(gc), _p(sig)))
    return woff, cnt, gc, sig


def wave_gz_batch(ctx: Context, gz_concat, gz_off, seq_off, size=100, step=10, lag=100, threshold=3.0, influence=1.0,
                  want_count=True, want_signal=True, raise_on_error=True):
    """get_seq + wave for a batch of ctgs held as seq:{ctg} gzip blobs → (win_off, gc_count, gc, signal, inflate status)."""
    gz_concat, gz_off, seq_off = _arr(gz_concat, np.uint8), _arr(gz_off, np.uint64), _arr(seq_off, np.uint64)
    woff = win_offsets(seq_off, size, step)
    n = int(woff[-1])
    cnt = np.empty(n, dtype=np.uint32) if want_count else None
    gc = np.empty(n, dtype=np.float32)
    sig = np.empty(n, dtype=np.int8) if want_signal else None
    status = np.zeros(seq_off.size - 1, dtype=np.int32)
    rc = load().gams_cuda_wave_gz_batch(ctx.handle, seq_off.size - 1, _p(gz_concat), _p(gz_off), _p(seq_off), size, step, lag,
                                        threshold, influence, _p(woff), _p(cnt), _p(gc), _p(sig), _p(status))
    if rc and (raise_on_error or not status.any()):
        check(rc)
    return woff, cnt, gc, sig, status


# --------------------------------------------------------------- anno.rs:128-139
def anno_coverage(ctx: Context, run_start, run_end, run_off, q_chr, q_ctg_start, q_ctg_end, q_start, q_end) -> np.ndarray:
    """`gams anno`: proportion of each range covered by the runlist of its chromosome inside its ctg (f32)."""
    run_start, run_end, run_off = _arr(run_start, np.int32), _arr(run_end, np.int32), _arr(run_off, np.uint64)
    q = [_arr(a, np.int32) for a in (q_chr, q_ctg_start, q_ctg_end, q_start, q_end)]
    prop = np.empty(q[0].size, dtype=np.float32)
    check(load().gams_cuda_anno_coverage(ctx.handle, run_off.size - 1, _p(run_start), _p(run_end), _p(run_off), q[0].size,
                                         *[_p(a) for a in q], _p(prop)))
    return prop


# --------------------------------------------------------------- gen.rs:80-127
def gen_ctgs(ctx: Context, seq_concat, chr_off, piece=500000, fill=50, min_=5000):
    """`gams gen`'s region finding for a batch of chromosomes → (ctg_chr, ctg_start, ctg_end), 1-based inclusive."""
    seq_concat, chr_off = _arr(seq_concat, np.uint8), _arr(chr_off, np.uint64)
    L = load()
    cap = 4096
    while True:
        c, s, e, n = np.empty(cap, dtype=np.uint32), np.empty(cap, dtype=np.int32), np.empty(cap, dtype=np.int32), C.c_uint64()
        rc = L.gams_cuda_gen_ctgs(ctx.handle, chr_off.size - 1, _p(seq_concat), _p(chr_off), fill, min_, piece, cap, _p(c), _p(s),
                                  _p(e), C.byref(n))
        if rc == -7 and n.value > cap:
            cap = int(n.value)
            continue
        check(rc)
        k = int(n.value)
        return c[:k], s[:k], e[:k]


# --------------------------------------------------------------- wave.rs:138-252 (peaks only)
def wave_peaks_batch(ctx: Context, seq_concat, seq_off, size=100, step=10, lag=100, threshold=3.0, influence=1.0,
                     coverage=0.2, cap=None):
    """`gams wave` (merged peaks) for a batch of ctgs → dict(ctg, first, last, gc, signal): one record per printed line."""
    seq_concat = _arr(seq_concat, np.uint8)
    seq_off = _arr(seq_off, np.uint64)
    woff = win_offsets(seq_off, size, step)
    L = load()
    cap = int(cap if cap is not None else max(4096, int(woff[-1]) // 32))
    while True:
        ctg, first, last = (np.empty(cap, dtype=np.uint32) for _ in range(3))
        gc, sig, n = np.empty(cap, dtype=np.float32), np.empty(cap, dtype=np.int8), C.c_uint64()
        rc = L.gams_cuda_wave_peaks_batch(ctx.handle, seq_off.size - 1, _p(seq_concat), _p(seq_off), size, step, lag, threshold,
                                          influence, coverage, _p(woff), cap, _p(ctg), _p(first), _p(last), _p(gc), _p(sig),
                                          C.byref(n))
        if rc == -7 and n.value > cap:  # ECAP: the call reports how many records there are
            cap = int(n.value)
            continue
        check(rc)
        k = int(n.value)
        return dict(ctg=ctg[:k], first=first[:k], last=last[:k], gc=gc[:k], signal=sig[:k])


def wave_peaks_gz_batch(ctx: Context, gz_concat, gz_off, seq_off, size=100, step=10, lag=100, threshold=3.0, influence=1.0,
                        coverage=0.2, cap=None, raise_on_error=True):
    """get_seq + `gams wave` (merged peaks) for ctgs held as seq:{ctg} gzip blobs → (records as wave_peaks_batch, inflate status)."""
    gz_concat, gz_off, seq_off = _arr(gz_concat, np.uint8), _arr(gz_off, np.uint64), _arr(seq_off, np.uint64)
    woff = win_offsets(seq_off, size, step)
    L = load()
    cap = int(cap if cap is not None else max(4096, int(woff[-1]) // 32))
    status = np.zeros(seq_off.size - 1, dtype=np.int32)
    while True:
        ctg, first, last = (np.empty(cap, dtype=np.uint32) for _ in range(3))
        gc, sig, n = np.empty(cap, dtype=np.float32), np.empty(cap, dtype=np.int8), C.c_uint64()
        rc = L.gams_cuda_wave_peaks_gz_batch(ctx.handle, seq_off.size - 1, _p(gz_concat), _p(gz_off), _p(seq_off), size, step, lag,
                                             threshold, influence, coverage, _p(woff), cap, _p(ctg), _p(first), _p(last), _p(gc),
                                             _p(sig), C.byref(n), _p(status))
        if rc == -7 and n.value > cap:  # ECAP: the call reports how many records there are
            cap = int(n.value)
            continue
        if rc and (raise_on_error or not status.any()):
            check(rc)
        k = 0 if rc else int(n.value)
        return dict(ctg=ctg[:k], first=first[:k], last=last[:k], gc=gc[:k], signal=sig[:k]), status


def wave_peaks_text(chr_id: str, chr_start: int, size: int, step: int, first, last, gc, signal) -> str:
    """The per-ctg stdout block of `gams wave` from the peak records of that ctg."""
    first, last = _arr(first, np.uint32), _arr(last, np.uint32)
    gc, signal = _arr(gc, np.float32), _arr(signal, np.int8)
    txt, ln = C.c_void_p(), C.c_size_t()
    L = load()
    check(L.gams_cuda_wave_peaks_text(chr_id.encode(), chr_start, size, step, first.size, _p(first), _p(last), _p(gc), _p(signal),
                                      C.byref(txt), C.byref(ln)))
    try:
        return C.string_at(txt, ln.value).decode()
    finally:
        L.gams_cuda_free(txt)


# --------------------------------------------------------------- peak.rs:66-160
def peak_link_batch(ctx: Context, seq_concat, seq_off, ctg_chr_start, p_ctg, p_start, p_end, p_signal):
    """`gams peak`: GC of each peak range and its links to the neighbouring peaks of the ctg → dict of arrays."""
    seq_concat, seq_off = _arr(seq_concat, np.uint8), _arr(seq_off, np.uint64)
    ctg_chr_start = _arr(ctg_chr_start, np.int32)
    p_ctg, p_start, p_end, p_signal = _arr(p_ctg, np.uint32), _arr(p_start, np.int32), _arr(p_end, np.int32), _arr(p_signal, np.int8)
    n = p_ctg.size
    gc, la, ra = (np.empty(n, dtype=np.float32) for _ in range(3))
    ll, rl = (np.empty(n, dtype=np.int32) for _ in range(2))
    ls, rs = (np.empty(n, dtype=np.int8) for _ in range(2))
    check(load().gams_cuda_peak_link_batch(ctx.handle, seq_off.size - 1, _p(seq_concat), _p(seq_off), _p(ctg_chr_start), n,
                                           _p(p_ctg), _p(p_start), _p(p_end), _p(p_signal), _p(gc), _p(ls), _p(ll), _p(la),
                                           _p(rs), _p(rl), _p(ra)))
    return dict(gc=gc, left_signal=ls, left_wave_length=ll, left_amplitude=la, right_signal=rs, right_wave_length=rl,
                right_amplitude=ra)


# ------------------------------------------------------------------------------ stat.rs:16-56
def thresholding_algo(ctx: Context, data, lag: int, threshold: float, influence: float) -> np.ndarray:
    """gams::thresholding_algo(data, lag, threshold, influence) → int8 signals (-1/0/1)."""
    d = _arr(data, np.float32)
    off = np.array([0, d.size], dtype=np.uint64)
    return thresholding_batch(ctx, d, off, lag, threshold, influence)


def thresholding_batch(ctx: Context, data, off, lag, threshold, influence) -> np.ndarray:
    d = _arr(data, np.float32)
    off = _arr(off, np.uint64)
    sig = np.empty(d.size, dtype=np.int8)
    check(load().gams_cuda_thresholding_batch(ctx.handle, off.size - 1, _p(d), _p(off), lag, threshold, influence, _p(sig)))
    return sig


# --------------------------------------------------------------- utils.rs:24-36, locate.rs:111-141
def count_overlaps(ctx: Context, starts, stops, iv_off, q_seg, q_start, q_stop) -> np.ndarray:
    """Batched gams::count_rg: per query lower_bound(starts, q_stop) - lower_bound(stops, q_start+1)."""
    starts, stops = _arr(starts, np.uint32), _arr(stops, np.uint32)
    iv_off = _arr(iv_off, np.uint64)
    q_seg, q_start, q_stop = _arr(q_seg, np.int32), _arr(q_start, np.uint32), _arr(q_stop, np.uint32)
    out = np.empty(q_seg.size, dtype=np.uint32)
    check(load().gams_cuda_count_overlaps(ctx.handle, iv_off.size - 1, _p(starts), _p(stops), _p(iv_off), q_seg.size,
                                          _p(q_seg), _p(q_start), _p(q_stop), _p(out)))
    return out


def build_idx_rg(ctx: Context, start, stop, iv_off, want_blob=True):
    """gams::build_idx_rg (redis.rs:276-305): Lapper::new + bincode per ctg segment, on the device.
    start/stop: the intervals [s, e+1) of every segment back to back, any order inside a segment.
    Returns dict(iv_start, iv_stop, starts, stops, max_len, blob, blob_off): blob_off[s] = 24*iv_off[s] + 30*s."""
    start, stop, iv_off = _arr(start, np.uint32), _arr(stop, np.uint32), _arr(iv_off, np.uint64)
    n_seg, n = iv_off.size - 1, int(iv_off[-1])
    out = {k: np.empty(n, dtype=np.uint32) for k in ("iv_start", "iv_stop", "starts", "stops")}
    out["max_len"] = np.empty(n_seg, dtype=np.uint32)
    out["blob_off"] = 24 * iv_off + 30 * np.arange(n_seg + 1, dtype=np.uint64)
    out["blob"] = np.empty(int(out["blob_off"][-1]), dtype=np.uint8) if want_blob else None
    check(load().gams_cuda_build_idx_rg(ctx.handle, n_seg, _p(start), _p(stop), _p(iv_off), _p(out["iv_start"]), _p(out["iv_stop"]),
                                        _p(out["starts"]), _p(out["stops"]), _p(out["max_len"]), _p(out["blob"])))
    return out


def locate_ctg(ctx: Context, ctg_start, ctg_stop, chr_off, q_chr, q_start, q_stop) -> np.ndarray:
    """Batched gams::find_one_idx → global ctg index or -1."""
    ctg_start, ctg_stop = _arr(ctg_start, np.uint32), _arr(ctg_stop, np.uint32)
    chr_off = _arr(chr_off, np.uint64)
    q_chr, q_start, q_stop = _arr(q_chr, np.int32), _arr(q_start, np.uint32), _arr(q_stop, np.uint32)
    out = np.empty(q_chr.size, dtype=np.int32)
    check(load().gams_cuda_locate_ctg(ctx.handle, chr_off.size - 1, _p(ctg_start), _p(ctg_stop), _p(chr_off), q_chr.size,
                                      _p(q_chr), _p(q_start), _p(q_stop), _p(out)))
    return out


# --------------------------------------------------------------- utils.rs:141-213, sw.rs:141-191
def gc_ranges(ctx: Context, seq_concat, seq_off, q_ctg, q_from0, q_to0_excl) -> np.ndarray:
    seq_concat, seq_off = _arr(seq_concat, np.uint8), _arr(seq_off, np.uint64)
    q_ctg, q_from0, q_to0_excl = _arr(q_ctg, np.uint32), _arr(q_from0, np.uint32), _arr(q_to0_excl, np.uint32)
    out = np.empty(q_ctg.size, dtype=np.uint32)
    check(load().gams_cuda_gc_ranges(ctx.handle, seq_off.size - 1, _p(seq_concat), _p(seq_off), q_ctg.size, _p(q_ctg),
                                     _p(q_from0), _p(q_to0_excl), _p(out)))
    return out


def sw_batch(ctx: Context, seq_concat, seq_off, ctg_chr_start, f_ctg, f_start, f_end, size=100, max_=20, resize=500, rg_index=None):
    """Rows of `gams sw --action gc` for a batch of features → dict of arrays + row_off.  rg_index = (starts, stops, iv_off)
    with segment c = the rg index of ctg c adds `--action count` (sw.rs:28-32): the rg_count column (data.rs:55)."""
    seq_concat, seq_off = _arr(seq_concat, np.uint8), _arr(seq_off, np.uint64)
    ctg_chr_start = _arr(ctg_chr_start, np.int32)
    f_ctg, f_start, f_end = _arr(f_ctg, np.uint32), _arr(f_start, np.int32), _arr(f_end, np.int32)
    nf = f_ctg.size
    cap = nf * (1 + 2 * max_)
    row_off = np.zeros(nf + 1, dtype=np.uint64)
    ws, we, wd = (np.empty(cap, dtype=np.int32) for _ in range(3))
    wt = np.empty(cap, dtype=np.int8)
    gc, gm, gs, gv = (np.empty(cap, dtype=np.float32) for _ in range(4))
    if rg_index is None:
        check(load().gams_cuda_sw_batch(ctx.handle, seq_off.size - 1, _p(seq_concat), _p(seq_off), _p(ctg_chr_start), nf,
                                        _p(f_ctg), _p(f_start), _p(f_end), size, max_, resize, _p(row_off), _p(ws), _p(we),
                                        _p(wt), _p(wd), _p(gc), _p(gm), _p(gs), _p(gv)))
        n = int(row_off[-1])
        return dict(row_off=row_off, start=ws[:n], end=we[:n], type=wt[:n], dist=wd[:n], gc=gc[:n], mean=gm[:n],
                    sd=gs[:n], cv=gv[:n])
    starts, stops, iv_off = _arr(rg_index[0], np.uint32), _arr(rg_index[1], np.uint32), _arr(rg_index[2], np.uint64)
    rc = np.empty(cap, dtype=np.uint32)
    check(load().gams_cuda_sw_count_batch(ctx.handle, seq_off.size - 1, _p(seq_concat), _p(seq_off), _p(ctg_chr_start), nf,
                                          _p(f_ctg), _p(f_start), _p(f_end), size, max_, resize, _p(starts), _p(stops), _p(iv_off),
                                          _p(row_off), _p(ws), _p(we), _p(wt), _p(wd), _p(gc), _p(gm), _p(gs), _p(gv), _p(rc)))
    n = int(row_off[-1])
    return dict(row_off=row_off, start=ws[:n], end=we[:n], type=wt[:n], dist=wd[:n], gc=gc[:n], mean=gm[:n],
                sd=gs[:n], cv=gv[:n], rg_count=rc[:n])


# --------------------------------------------------------------- host-side format helpers (no GPU)
def wave_text(chr_id: str, chr_start: int, size: int, step: int, gc, signal, coverage=0.2, is_signal=False) -> str:
    """The per-ctg stdout block of `gams wave` (wave.rs:157-212), produced by the C++ host layer."""
    gc, signal = _arr(gc, np.float32), _arr(signal, np.int8)
    txt, ln = C.c_void_p(), C.c_size_t()
    L = load()
    check(L.gams_cuda_wave_text(chr_id.encode(), chr_start, size, step, gc.size, _p(gc), _p(signal), coverage,
                                int(is_signal), C.byref(txt), C.byref(ln)))
    try:
        return C.string_at(txt, ln.value).decode()
    finally:
        L.gams_cuda_free(txt)


def lapper_blob_view(blob: bytes):
    """(starts, stops, max_len) sliced zero-copy out of a bincode Lapper<u32,String> blob."""
    buf = np.frombuffer(blob, dtype=np.uint8)
    n, st, sp, ml = C.c_uint64(), C.c_void_p(), C.c_void_p(), C.c_uint32()
    check(load().gams_cuda_lapper_blob_view(_p(buf), buf.size, C.byref(n), C.byref(st), C.byref(sp), C.byref(ml)))
    base = buf.ctypes.data
    starts = np.frombuffer(blob, dtype="<u4", count=n.value, offset=(st.value or base) - base)
    stops = np.frombuffer(blob, dtype="<u4", count=n.value, offset=(sp.value or base) - base)
    return starts, stops, int(ml.value)


# --------------------------------------------------------------- device API (inputs resident in HBM)
class WavePlan:
    """gams_cuda_wave_plan on one device; tensors are torch CUDA tensors (used for their data_ptr only)."""

    def __init__(self, ctx: Context, seq_off, size=100, step=10, lag=100, dev_index=0):
        self._L = load()
        seq_off = _arr(seq_off, np.uint64)
        h = C.c_void_p()
        check(self._L.gams_cuda_wave_plan_create(ctx.handle, dev_index, seq_off.size - 1, _p(seq_off), size, step, lag,
                                                 C.byref(h)))
        ctx._plans.add(self)
        self._h, self._ctx = h, ctx
        self.total_windows = int(self._L.gams_cuda_wave_plan_total_windows(h))
        self.total_bases = int(self._L.gams_cuda_wave_plan_total_bases(h))
        self.seq_capacity = int(self._L.gams_cuda_wave_plan_seq_capacity(h))

    def gc(self, stream: int, d_seq, d_gc, d_cnt=None):
        check(self._L.gams_cuda_wave_gc_dev(self._h, stream, d_seq.data_ptr(), d_cnt.data_ptr() if d_cnt is not None else None,
                                            d_gc.data_ptr()))

    def signal(self, stream: int, d_gc, threshold, influence, d_signal, d_filtered=None):
        check(self._L.gams_cuda_wave_signal_dev(self._h, stream, d_gc.data_ptr(), threshold, influence,
                                                d_filtered.data_ptr() if d_filtered is not None else None,
                                                d_signal.data_ptr()))

    def wave(self, stream: int, d_seq, threshold, influence, d_gc, d_signal, d_cnt=None, d_filtered=None):
        """GC + signals in one call: the fused kernel when `fused(threshold, influence)`, else gc() + signal()."""
        check(self._L.gams_cuda_wave_dev(self._h, stream, d_seq.data_ptr(), threshold, influence,
                                         d_cnt.data_ptr() if d_cnt is not None else None, d_gc.data_ptr(),
                                         d_filtered.data_ptr() if d_filtered is not None else None, d_signal.data_ptr()))

    def fused(self, threshold, influence) -> bool:
        return bool(self._L.gams_cuda_wave_plan_fused(self._h, threshold, influence))

    @property
    def launches(self) -> int:
        return int(self._L.gams_cuda_wave_plan_launches(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.gams_cuda_wave_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CountPlan:
    """gams_cuda_count_plan: the rg index of a set of ctgs resident on one device."""

    def __init__(self, ctx: Context, iv_off, dev_index=0):
        self._L = load()
        iv_off = _arr(iv_off, np.uint64)
        h = C.c_void_p()
        check(self._L.gams_cuda_count_plan_create(ctx.handle, dev_index, iv_off.size - 1, _p(iv_off), C.byref(h)))
        ctx._plans.add(self)
        self._h, self._ctx = h, ctx

    def count(self, stream: int, d_starts, d_stops, d_q_seg, d_q_start, d_q_stop, d_out, grouped=False):
        check(self._L.gams_cuda_count_overlaps_dev(self._h, stream, d_starts.data_ptr(), d_stops.data_ptr(),
                                                   d_q_seg.numel(), d_q_seg.data_ptr(), d_q_start.data_ptr(),
                                                   d_q_stop.data_ptr(), int(grouped), d_out.data_ptr()))

    @property
    def launches(self) -> int:
        return int(self._L.gams_cuda_count_plan_launches(self._h))

    def invalidate(self):
        """Forget the tables derived from the index (after rewriting it in place under the same device pointers)."""
        self._L.gams_cuda_count_plan_invalidate(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._L.gams_cuda_count_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------- window.rs:78-94 x utils.rs:24-36
def count_windows(ctx: Context, chr_start, win_off, size, step, starts, stops, iv_off) -> np.ndarray:
    """count_rg of every sliding window of every ctg (window k of ctg c = [chr_start[c] + k*step, + size - 1])."""
    chr_start, win_off = _arr(chr_start, np.int32), _arr(win_off, np.uint64)
    starts, stops, iv_off = _arr(starts, np.uint32), _arr(stops, np.uint32), _arr(iv_off, np.uint64)
    out = np.empty(int(win_off[-1]), dtype=np.uint32)
    check(load().gams_cuda_count_windows(ctx.handle, chr_start.size, _p(chr_start), _p(win_off), size, step, _p(starts), _p(stops),
                                         _p(iv_off), _p(out)))
    return out


class CountWindowsPlan:
    """gams_cuda_cwin_plan: window geometry + index offsets of a ctg set bound to one device; index and counts stay in HBM."""

    def __init__(self, ctx: Context, chr_start, win_off, size, step, iv_off, dev_index=0):
        self._L = load()
        chr_start, win_off, iv_off = _arr(chr_start, np.int32), _arr(win_off, np.uint64), _arr(iv_off, np.uint64)
        h = C.c_void_p()
        check(self._L.gams_cuda_cwin_plan_create(ctx.handle, dev_index, chr_start.size, _p(chr_start), _p(win_off), size, step,
                                                 _p(iv_off), C.byref(h)))
        ctx._plans.add(self)
        self._h, self._ctx = h, ctx

    @property
    def total_windows(self) -> int:
        return int(self._L.gams_cuda_cwin_plan_total_windows(self._h))

    def count(self, stream: int, d_starts, d_stops, d_out):
        check(self._L.gams_cuda_count_windows_dev(self._h, stream, d_starts.data_ptr(), d_stops.data_ptr(), d_out.data_ptr()))

    @property
    def launches(self) -> int:
        return int(self._L.gams_cuda_cwin_plan_launches(self._h))

    def invalidate(self):
        self._L.gams_cuda_cwin_plan_invalidate(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._L.gams_cuda_cwin_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------- redis.rs:142-161 (get_seq -> decode_gz)
def pack_blobs(blobs):
    """bytes objects back to back → (uint8 array, uint64 offsets)."""
    off = np.zeros(len(blobs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(b) for b in blobs])
    cat = np.frombuffer(b"".join(blobs), dtype=np.uint8) if blobs else np.zeros(0, np.uint8)
    return cat, off


def gz_sizes(gz_concat, gz_off) -> np.ndarray:
    gz_concat, gz_off = _arr(gz_concat, np.uint8), _arr(gz_off, np.uint64)
    out = np.empty(gz_off.size - 1, dtype=np.uint64)
    check(load().gams_cuda_gz_sizes(gz_off.size - 1, _p(gz_concat), _p(gz_off), _p(out)))
    return out


def gunzip_batch(ctx: Context, gz_concat, gz_off, out_off=None, raise_on_error=True):
    """decode_gz of a batch of seq:{ctg} blobs → (out_concat, out_off, status)."""
    gz_concat, gz_off = _arr(gz_concat, np.uint8), _arr(gz_off, np.uint64)
    if out_off is None:
        out_off = np.concatenate(([0], np.cumsum(gz_sizes(gz_concat, gz_off)))).astype(np.uint64)
    out_off = _arr(out_off, np.uint64)
    out = np.empty(int(out_off[-1]), dtype=np.uint8)
    status = np.zeros(gz_off.size - 1, dtype=np.int32)
    rc = load().gams_cuda_gunzip_batch(ctx.handle, gz_off.size - 1, _p(gz_concat), _p(gz_off), _p(out_off), _p(out), _p(status))
    if rc != 0 and (raise_on_error or not status.any()):
        check(rc)
    return out, out_off, status


# --------------------------------------------------------------- locate.rs:111-117, utils.rs:39-52 (Range::from_str per line)
def parse_ranges(ctx: Context, text: bytes, chr_names):
    """The lines of a range file parsed on the device → dict(line_off, chr, start, end): chr >= 0 index into chr_names,
    -1 unknown chromosome, -2 invalid range (skipped by the reference), -3 outside the strict form (hand to Range::from_str)."""
    t = np.frombuffer(text, dtype=np.uint8)
    names = [n.encode() if isinstance(n, str) else n for n in chr_names]
    noff = np.zeros(len(names) + 1, dtype=np.uint32)
    noff[1:] = np.cumsum([len(n) for n in names])
    ncat = np.frombuffer(b"".join(names), dtype=np.uint8) if names else np.zeros(0, np.uint8)
    L = load()
    n = C.c_uint64()
    # room for the lines: estimated from the head of the text, retried with the exact count if that was too small
    head = text[:1 << 16]
    cap = int(len(text) / max(1.0, len(head) / (head.count(b"\n") + 1)) * 1.25) + 16
    while True:
        out = dict(line_off=np.zeros(cap + 1, np.uint64), chr=np.empty(cap, np.int32), start=np.empty(cap, np.uint32), end=np.empty(cap, np.uint32))
        rc = L.gams_cuda_parse_ranges(ctx.handle, _p(t), t.size, len(names), _p(ncat), _p(noff), cap, C.byref(n), _p(out["line_off"]),
                                      _p(out["chr"]), _p(out["start"]), _p(out["end"]))
        if rc == -7:
            cap = int(n.value)
            continue
        check(rc)
        break
    k = int(n.value)
    out = {"line_off": out["line_off"][:k + 1], "chr": out["chr"][:k], "start": out["start"][:k], "end": out["end"][:k]}
    return out


def locate_count_text(ctx: Context, text: bytes, chr_names, ctg_start, ctg_stop, chr_off, starts, stops, iv_off, want_lines=False):
    """`gams locate --count -f file`, text in → (output text, n_lines, n_printed, n_handed_back[, input line of every printed line])."""
    t = np.frombuffer(text, dtype=np.uint8)
    names = [n.encode() if isinstance(n, str) else n for n in chr_names]
    noff = np.zeros(len(names) + 1, dtype=np.uint32)
    noff[1:] = np.cumsum([len(n) for n in names])
    ncat = np.frombuffer(b"".join(names), dtype=np.uint8) if names else np.zeros(0, np.uint8)
    ctg_start, ctg_stop, chr_off = _arr(ctg_start, np.uint32), _arr(ctg_stop, np.uint32), _arr(chr_off, np.uint64)
    starts, stops, iv_off = _arr(starts, np.uint32), _arr(stops, np.uint32), _arr(iv_off, np.uint64)
    L = load()
    out, n, nl, npr, nb, ol = C.c_void_p(), C.c_size_t(), C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_void_p()
    common = (ctx.handle, _p(t), t.size, len(names), _p(ncat), _p(noff), _p(ctg_start), _p(ctg_stop), _p(chr_off), _p(starts), _p(stops),
              _p(iv_off), C.byref(out), C.byref(n))
    if want_lines:
        check(L.gams_cuda_locate_count_lines(*common, C.byref(ol), C.byref(nl), C.byref(npr), C.byref(nb)))
    else:
        check(L.gams_cuda_locate_count_text(*common, C.byref(nl), C.byref(npr), C.byref(nb)))
    try:
        s = C.string_at(out, n.value) if out.value else b""
        lines = np.ctypeslib.as_array(C.cast(ol, C.POINTER(C.c_uint64)), shape=(int(npr.value),)).copy() if (want_lines and ol.value and npr.value) \
            else np.zeros(0, np.uint64)
    finally:
        L.gams_cuda_free(out)
        if ol.value:
            L.gams_cuda_free(ol)
    if want_lines:
        return s, int(nl.value), int(npr.value), int(nb.value), lines
    return s, int(nl.value), int(npr.value), int(nb.value)
