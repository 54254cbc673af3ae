"""CPU oracle for the gams hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  The product package (gams_b200) never does.

Numeric kernels live in gams_oracle.c (faithful C restatement, timed as the CPU baseline);
this module adds the host-side glue the reference wraps around them, restated in Python:
text emission of `gams wave` (wave.rs:157-252), Range parsing (intspan 0.7.7),
Lapper::new + bincode blob (rust-lapper 1.1.0 / bincode 1.3.3, redis.rs:236-324),
read_range's first-range-dropped quirk (utils.rs:39-67) and gen's ctg partition
(gen.rs:80-127).  Parity status per function is stated in its docstring.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import struct
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "libgams_oracle.so"
    src = _HERE / "gams_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(str(build()))
    u8p, u32p, u64p, i32p, i8p, f32p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_int32, C.c_int8, C.c_float))
    L.gorc_nwin.restype = C.c_int64
    L.gorc_nwin.argtypes = [C.c_int64, C.c_int32, C.c_int32]
    L.gorc_gc_content.restype = C.c_float
    L.gorc_gc_content.argtypes = [C.c_void_p, C.c_size_t, u32p]
    L.gorc_mean.restype = C.c_float
    L.gorc_mean.argtypes = [C.c_void_p, C.c_size_t]
    L.gorc_stddev.restype = C.c_float
    L.gorc_stddev.argtypes = [C.c_void_p, C.c_size_t]
    L.gorc_thresholding.restype = C.c_int
    L.gorc_thresholding.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_float, C.c_float, C.c_void_p]
    L.gorc_wave_batch.restype = C.c_int
    L.gorc_wave_batch.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_uint64,
                                  C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    L.gorc_lapper_count.restype = C.c_uint32
    L.gorc_lapper_count.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32]
    L.gorc_count_batch.restype = C.c_int
    L.gorc_count_batch.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    L.gorc_lapper_find_first.restype = C.c_int64
    L.gorc_lapper_find_first.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
    L.gorc_center_resize.restype = None
    L.gorc_center_resize.argtypes = [C.c_int32] * 5 + [i32p, i32p]
    L.gorc_center_sw.restype = C.c_int32
    L.gorc_center_sw.argtypes = [C.c_int32] * 6 + [C.c_void_p] * 4
    L.gorc_round.restype = C.c_float
    L.gorc_round.argtypes = [C.c_float, C.c_uint32]
    L.gorc_gc_range4.restype = C.c_float
    L.gorc_gc_range4.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    L.gorc_gc_stat.restype = None
    L.gorc_gc_stat.argtypes = [C.c_void_p, C.c_size_t, f32p, f32p, f32p]
    L.gorc_sw_feature.restype = C.c_int32
    L.gorc_sw_feature.argtypes = [C.c_void_p] + [C.c_int32] * 7 + [C.c_void_p] * 8
    _LIB = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ----------------------------------------------------------------------------------------
# numeric wrappers
# ----------------------------------------------------------------------------------------
def nwin(length: int, size: int, step: int) -> int:
    """window.rs:78-94 — number of windows `sliding` yields on a span of `length`."""
    return int(lib().gorc_nwin(length, size, step))


def mean(a) -> np.float32:
    a = np.ascontiguousarray(a, dtype=np.float32)
    return np.float32(lib().gorc_mean(_ptr(a), a.size))


def stddev(a) -> np.float32:
    a = np.ascontiguousarray(a, dtype=np.float32)
    return np.float32(lib().gorc_stddev(_ptr(a), a.size))


def thresholding_algo(data, lag: int, threshold: float, influence: float) -> np.ndarray:
    """stat.rs:16-56.  Raises IndexError where the reference panics (n < lag)."""
    d = np.ascontiguousarray(data, dtype=np.float32)
    sig = np.zeros(d.size, dtype=np.int8)
    r = lib().gorc_thresholding(_ptr(d), d.size, lag, threshold, influence, _ptr(sig))
    if r:
        raise IndexError("thresholding_algo: data shorter than lag (reference panics, stat.rs:30)")
    return sig


def pack_ctgs(seqs):
    """list of bytes/uint8 arrays -> (concat uint8, seq_off uint64[n+1])"""
    arrs = [np.frombuffer(s, dtype=np.uint8) if isinstance(s, (bytes, bytearray)) else np.asarray(s, dtype=np.uint8) for s in seqs]
    off = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        off[1:] = np.cumsum([a.size for a in arrs], dtype=np.uint64)
    cat = np.concatenate(arrs) if arrs else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(cat), off


def win_offsets(seq_off, size: int, step: int) -> np.ndarray:
    lens = np.diff(np.asarray(seq_off, dtype=np.uint64)).astype(np.int64)
    n = np.where(lens >= size, (lens - size) // step + 1, 0)
    out = np.zeros(lens.size + 1, dtype=np.uint64)
    out[1:] = np.cumsum(n, dtype=np.uint64)
    return out


def wave_batch(seq_concat, seq_off, size, step, lag, threshold, influence, threads=1, want_signal=True):
    """wave.rs:138-155 over a batch of ctgs.  Returns (win_off, gc_count u32, gc f32, signal i8)."""
    seq_concat = np.ascontiguousarray(seq_concat, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
    woff = win_offsets(seq_off, size, step)
    n = int(woff[-1])
    cnt = np.zeros(n, dtype=np.uint32)
    gc = np.zeros(n, dtype=np.float32)
    sig = np.zeros(n, dtype=np.int8) if want_signal else None
    r = lib().gorc_wave_batch(seq_off.size - 1, _ptr(seq_concat), _ptr(seq_off), size, step, lag, threshold,
                              influence, _ptr(woff), _ptr(cnt), _ptr(gc), _ptr(sig), threads)
    if r:
        raise IndexError("wave: a ctg has fewer windows than lag (reference panics, stat.rs:30)")
    return woff, cnt, gc, sig


def lapper_count(starts, stops, qs: int, qe: int) -> int:
    """rust-lapper 1.1.0 Lapper::count(start, stop) on prebuilt sorted starts[]/stops[]."""
    starts = np.ascontiguousarray(starts, dtype=np.uint32)
    stops = np.ascontiguousarray(stops, dtype=np.uint32)
    return int(lib().gorc_lapper_count(_ptr(starts), _ptr(stops), starts.size, qs, qe))


def count_batch(starts, stops, iv_off, q_seg, q_start, q_stop, threads=1) -> np.ndarray:
    """locate.rs:111-141 count loop on located queries; q_stop is the value gams passes as
    Lapper `stop`, i.e. the range's inclusive end (utils.rs:35)."""
    starts = np.ascontiguousarray(starts, dtype=np.uint32)
    stops = np.ascontiguousarray(stops, dtype=np.uint32)
    iv_off = np.ascontiguousarray(iv_off, dtype=np.uint64)
    q_seg = np.ascontiguousarray(q_seg, dtype=np.int32)
    q_start = np.ascontiguousarray(q_start, dtype=np.uint32)
    q_stop = np.ascontiguousarray(q_stop, dtype=np.uint32)
    out = np.zeros(q_seg.size, dtype=np.uint32)
    lib().gorc_count_batch(iv_off.size - 1, _ptr(starts), _ptr(stops), _ptr(iv_off), q_seg.size, _ptr(q_seg),
                           _ptr(q_start), _ptr(q_stop), _ptr(out), threads)
    return out


def center_resize(p_start, p_end, s, e, resize):
    a, b = C.c_int32(), C.c_int32()
    lib().gorc_center_resize(p_start, p_end, s, e, resize, C.byref(a), C.byref(b))
    return a.value, b.value


def center_sw(p_start, p_end, start, end, size, max_):
    """window.rs:3-56 → list of (start, end, 'M'|'L'|'R', distance)."""
    k = 1 + 2 * max(max_, 0)
    ws, we, wt, wd = (np.zeros(k, dtype=np.int32) for _ in range(4))
    n = lib().gorc_center_sw(p_start, p_end, start, end, size, max_, _ptr(ws), _ptr(we), _ptr(wt), _ptr(wd))
    return [(int(ws[i]), int(we[i]), "MLR"[wt[i]], int(wd[i])) for i in range(n)]


def round_(x, decimals):
    return np.float32(lib().gorc_round(np.float32(x), decimals))


def gc_stat(gcs):
    g = np.ascontiguousarray(gcs, dtype=np.float32)
    m, s, c = C.c_float(), C.c_float(), C.c_float()
    lib().gorc_gc_stat(_ptr(g), g.size, C.byref(m), C.byref(s), C.byref(c))
    return np.float32(m.value), np.float32(s.value), np.float32(c.value)


def sw_feature(seq, p_start, p_end, f_start, f_end, size=100, max_=20, resize=500):
    """sw.rs:141-191 rows for one feature: dict of arrays."""
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    k = 1 + 2 * max_
    ws, we, wt, wd = (np.zeros(k, dtype=np.int32) for _ in range(4))
    gc, gm, gs, gv = (np.zeros(k, dtype=np.float32) for _ in range(4))
    n = lib().gorc_sw_feature(_ptr(seq), p_start, p_end, f_start, f_end, size, max_, resize,
                              _ptr(ws), _ptr(we), _ptr(wt), _ptr(wd), _ptr(gc), _ptr(gm), _ptr(gs), _ptr(gv))
    return dict(start=ws[:n], end=we[:n], type=wt[:n], dist=wd[:n], gc=gc[:n], mean=gm[:n], sd=gs[:n], cv=gv[:n])


# ----------------------------------------------------------------------------------------
# text glue (wave.rs:157-252).  Pinned by tests/S288c/I.peaks.tsv.
# ----------------------------------------------------------------------------------------
def fmt_f32(x) -> str:
    """Rust `{}` for f32: shortest digits that round-trip, never scientific notation."""
    x = np.float32(x)
    if np.isnan(x):
        return "NaN"
    if np.isinf(x):
        return "inf" if x > 0 else "-inf"
    s = np.format_float_positional(x, unique=True, trim="-")
    return s


def runlist(a: int, b: int) -> str:
    return str(a) if a == b else f"{a}-{b}"


def merge_ints(peaks, coverage: float):
    """wave.rs:217-252.  peaks: list of (start, end), in window order.  Returns {runlist: merged}.
    Edge i-j iff the spans intersect and size/intersect >= coverage for both (f32, wave.rs:229-231);
    components via union-find (tarjan_scc on an undirected graph = connected components).
    Only nodes that own an edge are in the graph, so singletons are absent from the result."""
    cov = np.float32(coverage)
    n = len(peaks)
    parent = list(range(n))
    in_graph = [False] * n

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for i in range(n):
        si, ei = peaks[i]
        for j in range(i + 1, n):
            sj, ej = peaks[j]
            lo, hi = max(si, sj), min(ei, ej)
            if lo > hi:
                if sj > ei:  # peaks arrive sorted by start: no later j can intersect i
                    break
                continue
            isz = np.float32(hi - lo + 1)
            if np.float32(ei - si + 1) / isz >= cov and np.float32(ej - sj + 1) / isz >= cov:
                in_graph[i] = in_graph[j] = True
                ri, rj = find(i), find(j)
                if ri != rj:
                    parent[rj] = ri
    comp = {}
    for i in range(n):
        if in_graph[i]:
            comp.setdefault(find(i), []).append(i)
    merge_of = {}
    for members in comp.values():
        spans = sorted(peaks[m] for m in members)
        merged = []
        for s, e in spans:  # IntSpan union: overlapping or adjacent runs coalesce
            if merged and s <= merged[-1][1] + 1:
                merged[-1][1] = max(merged[-1][1], e)
            else:
                merged.append([s, e])
        text = ",".join(runlist(s, e) for s, e in merged)
        for m in members:
            merge_of[runlist(*peaks[m])] = text
    return merge_of


def wave_text(chr_id: str, chr_start: int, size: int, step: int, gc, signal, coverage=0.2, is_signal=False) -> str:
    """wave.rs:157-212 — the per-ctg block of `gams wave` stdout (no header line)."""
    n = len(gc)
    out = []
    wins = [(chr_start + k * step, chr_start + k * step + size - 1) for k in range(n)]
    if is_signal:
        for k in range(n):
            out.append(f"{chr_id}:{runlist(*wins[k])}\t{fmt_f32(gc[k])}\t{int(signal[k])}\n")
        return "".join(out)
    merge_of = {}
    merge_of.update(merge_ints([wins[k] for k in range(n) if signal[k] == 1], coverage))
    merge_of.update(merge_ints([wins[k] for k in range(n) if signal[k] == -1], coverage))
    seen = set()
    for k in range(n):
        if signal[k] == 0:
            continue
        rl = runlist(*wins[k])
        if rl in merge_of:
            m = merge_of[rl]
            if m not in seen:
                out.append(f"{chr_id}(+):{m}\t{fmt_f32(gc[k])}\t{int(signal[k])}\n")
                seen.add(m)
        else:
            out.append(f"{chr_id}:{rl}\t{fmt_f32(gc[k])}\t{int(signal[k])}\n")
    return "".join(out)


def wave_peak_records(size: int, step: int, gc, signal, coverage=0.2):
    """The lines wave.rs:170-211 prints for one ctg, as records (first_window, last_window, gc, signal): a merged peak is
    listed once, at its first member, with that member's gc/signal (wave.rs:196-203); an unmerged window has last == first.
    Only defined when every merged component is one contiguous run (always for coverage <= 1)."""
    n = len(gc)
    wins = [(1 + k * step, k * step + size) for k in range(n)]
    recs, seen = [], set()
    merge_of = {}
    merge_of.update(merge_ints([wins[k] for k in range(n) if signal[k] == 1], coverage))
    merge_of.update(merge_ints([wins[k] for k in range(n) if signal[k] == -1], coverage))
    for k in range(n):
        if signal[k] == 0:
            continue
        rl = runlist(*wins[k])
        if rl in merge_of:
            m = merge_of[rl]
            if m in seen:
                continue
            seen.add(m)
            assert "," not in m, "component is not a single run"
            a, b = (int(x) for x in m.split("-"))
            recs.append(((a - 1) // step, (b - size) // step, np.float32(gc[k]), int(signal[k])))
        else:
            recs.append((k, k, np.float32(gc[k]), int(signal[k])))
    return recs


def peak_link(seq, chr_start: int, peaks):
    """peak.rs:66-160 for one ctg.  peaks: [(start, end, signal)] sorted by start (1-based inclusive chr positions);
    seq: the ctg's bytes.  -> list of dict(gc, left_signal, left_wave_length, left_amplitude, right_*)."""
    chr_end = chr_start + len(seq) - 1
    out = []
    for s, e, sg in peaks:
        sub = bytes(seq[s - chr_start:e - chr_start + 1])  # utils.rs:151-155
        cnt = sum(sub.count(b) for b in (b"c", b"g", b"C", b"G"))
        g = round_(np.float32(cnt) / np.float32(len(sub)), 4)  # bio gc_content, then utils.rs:161
        out.append(dict(gc=g, signal=sg, start=s, end=e))
    if not out:
        return out
    prev_sig, prev_gc, prev_end = out[0]["signal"], out[0]["gc"], chr_start  # peak.rs:112-114
    for p in out:
        p["left_signal"], p["left_wave_length"] = prev_sig, p["start"] - prev_end + 1
        p["left_amplitude"] = np.float32(abs(np.float32(p["gc"] - prev_gc)))
        prev_sig, prev_end, prev_gc = p["signal"], p["end"], p["gc"]
    next_sig, next_gc, next_start = out[-1]["signal"], out[-1]["gc"], chr_end  # peak.rs:137-139
    for p in reversed(out):
        p["right_signal"], p["right_wave_length"] = next_sig, next_start - p["end"] + 1
        p["right_amplitude"] = np.float32(abs(np.float32(p["gc"] - next_gc)))
        next_sig, next_start, next_gc = p["signal"], p["start"], p["gc"]
    return out


def runlist_to_runs(text: str):
    """IntSpan runlist "a-b,c,d-e" -> sorted disjoint inclusive runs [(a, b), (c, c), (d, e)] (as json2set yields them)."""
    runs = []
    for part in text.split(","):
        if not part or part == "-":
            continue
        a, _, b = part.partition("-")
        runs.append((int(a), int(b) if b else int(a)))
    runs.sort()
    merged = []
    for a, b in runs:  # IntSpan coalesces overlapping and adjacent runs
        if merged and a <= merged[-1][1] + 1:
            merged[-1][1] = max(merged[-1][1], b)
        else:
            merged.append([a, b])
    return [(a, b) for a, b in merged]


def anno_prop(runs, ctg_start: int, ctg_end: int, start: int, end: int) -> np.float32:
    """anno.rs:128-139: |set ∩ ctg ∩ range| as f32 / |range| as f32; runs = the chromosome's set, or None when the
    chromosome is not in the runlist file (prop stays 0.0)."""
    if runs is None:
        return np.float32(0.0)
    a, b = max(start, ctg_start), min(end, ctg_end)
    cov = sum(max(0, min(b, re) - max(a, rs) + 1) for rs, re in runs) if b >= a else 0
    return np.float32(cov) / np.float32(end - start + 1)


WAVE_HEADER = "#range\tgc_content\tsignal\n"  # wave.rs:263-266


# ----------------------------------------------------------------------------------------
# Range grammar (intspan 0.7.7 Range::from_str / to_string / is_valid).  Pinned only through
# the strings the reference tests use ("I:1000-2000", "I:136", "I(+):1661-1780").
# ----------------------------------------------------------------------------------------
_RANGE_RE = re.compile(r"(?:(?P<name>[\w_]+)\.)?(?P<chr>[\w/\-]+)(?:\((?P<strand>.+)\))?:(?P<start>\d+)[_\-]?(?P<end>\d+)?",
                       re.ASCII)


def range_from_str(s: str):
    """-> (chr, strand, start, end) or None when invalid (no match, start == 0, end < start)."""
    m = _RANGE_RE.search(s)
    if not m:
        return None
    start = int(m.group("start"))
    end = int(m.group("end")) if m.group("end") is not None else start
    if start <= 0 or end < start:
        return None
    return m.group("chr"), m.group("strand") or "", start, end


def range_to_str(chr_id: str, start: int, end: int, strand: str = "") -> str:
    s = chr_id + (f"({strand})" if strand else "") + f":{start}"
    return s if start == end else s + f"-{end}"


# ----------------------------------------------------------------------------------------
# gen.rs:80-127 — ctg partition of one chromosome
# ----------------------------------------------------------------------------------------
def gen_ctgs(seq, piece=500000, fill=50, min_=5000):
    """Returns list of (chr_start, chr_end) 1-based inclusive.  Ambiguous = not in ACGTacgt
    (gen.rs:86-93); valid = complement, gaps <= fill-1 filled, spans < min dropped
    (intspan fill/excise, gen.rs:100-104); pieces per gen.rs:107-126 (tail merged into last)."""
    a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.asarray(seq, dtype=np.uint8)
    ok = np.isin(a, np.frombuffer(b"ACGTacgt", dtype=np.uint8))
    if not ok.any():
        return []
    d = np.diff(np.concatenate(([0], ok.view(np.int8), [0])))
    starts = np.flatnonzero(d == 1) + 1
    ends = np.flatnonzero(d == -1)
    spans = []
    for s, e in zip(starts.tolist(), ends.tolist()):
        if spans and s - spans[-1][1] - 1 <= fill - 1:
            spans[-1][1] = e
        else:
            spans.append([s, e])
    spans = [sp for sp in spans if sp[1] - sp[0] + 1 >= min_]
    out = []
    for pos, mx in spans:
        cur = []
        while mx - pos + 1 > piece:
            cur.append([pos, pos + piece - 1])
            pos += piece
        if not cur:
            cur.append([pos, mx])
        else:
            cur[-1][1] = mx
        out.extend((s, e) for s, e in cur)
    return out


# ----------------------------------------------------------------------------------------
# rust-lapper 1.1.0 Lapper::new + bincode 1.3.3 blob (redis.rs:236-324).  Blob layout is
# "parity unpinned": restated from the published crates, no reference test inspects bytes.
# ----------------------------------------------------------------------------------------
def lapper_new(intervals):
    """intervals: iterable of (start, stop, val:str).  Returns dict(start, stop, val, starts, stops, max_len)."""
    ivs = sorted(intervals, key=lambda t: (t[0], t[1]))  # Interval Ord: start, then stop (stable)
    st = np.array([t[0] for t in ivs], dtype=np.uint32)
    sp = np.array([t[1] for t in ivs], dtype=np.uint32)
    max_len = int((sp.astype(np.int64) - st.astype(np.int64)).clip(min=0).max()) if len(ivs) else 0
    return dict(start=st, stop=sp, val=[t[2] for t in ivs], starts=np.sort(st), stops=np.sort(sp), max_len=max_len)


def lapper_to_bincode(lp) -> bytes:
    n = len(lp["val"])
    parts = [struct.pack("<Q", n)]
    for s, e, v in zip(lp["start"].tolist(), lp["stop"].tolist(), lp["val"]):
        vb = v.encode()
        parts.append(struct.pack("<IIQ", s, e, len(vb)) + vb)
    parts.append(struct.pack("<Q", n) + lp["starts"].astype("<u4").tobytes())
    parts.append(struct.pack("<Q", n) + lp["stops"].astype("<u4").tobytes())
    parts.append(struct.pack("<I", lp["max_len"]))
    parts.append(b"\x00")  # cov: Option<u32> = None
    parts.append(b"\x00")  # overlaps_merged = false
    return b"".join(parts)


def lapper_find_first(lp, qs: int, qe: int):
    """Lapper::find(start, stop).next() → val or None (utils.rs:16)."""
    i = lib().gorc_lapper_find_first(_ptr(lp["start"]), _ptr(lp["stop"]), lp["start"].size, lp["max_len"], qs, qe)
    return None if i < 0 else lp["val"][i]


def build_idx_ctg(ctgs):
    """redis.rs:236-258.  ctgs: list of (ctg_id, chr_id, chr_start, chr_end) → {chr_id: lapper}."""
    by = {}
    for cid, chrom, s, e in ctgs:
        by.setdefault(chrom, []).append((s, e + 1, cid))
    return {k: lapper_new(v) for k, v in by.items()}


def find_one_idx(idx_ctg, rg):
    """utils.rs:7-22.  rg = (chr, strand, start, end) → ctg_id or ''."""
    lp = idx_ctg.get(rg[0])
    if lp is None:
        return ""
    return lapper_find_first(lp, rg[2], rg[3]) or ""


def read_range(lines, idx_ctg):
    """utils.rs:39-67 incl. quirk Q7: `.entry().and_modify(push).or_default()` drops the first
    range seen for each ctg.  Returns {ctg_id: [(chr,strand,start,end), ...]} sorted by ctg_id."""
    ranges_of = {}
    for line in lines:
        rg = range_from_str(line)
        if rg is None:
            continue
        cid = find_one_idx(idx_ctg, rg)
        if not cid:
            continue
        if cid in ranges_of:
            ranges_of[cid].append(rg)
        else:
            ranges_of[cid] = []
    return dict(sorted(ranges_of.items()))


def build_idx_rg(ranges_of):
    """redis.rs:276-305: rg chr:s-e → Iv{s, e+1, ""} per ctg."""
    return {cid: lapper_new([(r[2], r[3] + 1, "") for r in rgs]) for cid, rgs in ranges_of.items()}


def count_rg(idx_rg, ctg_id, rg) -> int:
    """utils.rs:24-36."""
    lp = idx_rg.get(ctg_id)
    if lp is None:
        return 0
    return lapper_count(lp["starts"], lp["stops"], rg[2], rg[3])


# ----------------------------------------------------------------------------------------
# Conn::get_seq -> decode_gz (redis.rs:142-161): flate2 1.0.30 GzDecoder::read_to_end.  The algorithm is RFC 1951/1952,
# implemented by flate2's miniz_oxide backend (not vendored in the reference); the checker here is zlib, an independent
# implementation of the same RFCs.  GzDecoder reads ONE member (MultiGzDecoder would read them all), so does this.
# encode_gz (redis.rs:149-154) is GzEncoder with Compression::fast() = level 1.
# ----------------------------------------------------------------------------------------
def decode_gz(blob: bytes) -> bytes:
    import zlib
    d = zlib.decompressobj(wbits=31)
    out = d.decompress(blob)
    if not d.eof:
        raise ValueError("truncated gzip member")
    return out


def encode_gz(seq: bytes, level: int = 1) -> bytes:
    import zlib
    c = zlib.compressobj(level, zlib.DEFLATED, 31)
    return c.compress(seq) + c.flush()
