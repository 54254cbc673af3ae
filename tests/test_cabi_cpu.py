"""CPU-side checks of the C ABI: the library loads, exports every symbol include/gams_cuda.h
declares, its GPU-free helpers agree with the oracle, and compute entry points fail loudly
(no CPU fallback) when there is no device."""
import ctypes as C

import numpy as np
import pytest

from conftest import cut_ctgs
from gams_b200 import _lib, api


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    hs = _lib.header_symbols()
    assert len(hs) >= 30
    for s in hs:
        assert hasattr(L, s), f"libgams_cuda.so does not export {s}"
    assert set(hs) == set(L._gams_signatures), "ctypes table and header disagree"
    assert L.gams_cuda_version() >= 100


def test_rust_sys_crate_declares_what_the_header_declares():
    """integration/gams-cuda-sys/src/lib.rs (the binding a gams maintainer adds) against include/gams_cuda.h: the same
    functions, the same number of arguments, the same error codes.  (No cargo in this image: the crate is checked as text.)"""
    import re
    from pathlib import Path
    root = Path(_lib.__file__).resolve().parent.parent
    rs = (root / "integration" / "gams-cuda-sys" / "src" / "lib.rs").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", (root / "include" / "gams_cuda.h").read_text(), flags=re.S)
    rs_fns = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (gams_cuda_\w+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", rs, flags=re.S)}
    h_fns = {m.group(1): m.group(2) for m in re.finditer(r"\b(gams_cuda_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S)}
    assert set(h_fns) == set(_lib.header_symbols())
    assert set(rs_fns) == set(h_fns), f"missing in lib.rs: {sorted(set(h_fns) - set(rs_fns))}; extra: {sorted(set(rs_fns) - set(h_fns))}"

    def n_args(a):
        a = a.strip()
        return 0 if a in ("", "void") else a.count(",") + 1 - (1 if a.rstrip().endswith(",") else 0)

    for f in h_fns:
        assert n_args(h_fns[f]) == n_args(rs_fns[f]), f"{f}: {n_args(h_fns[f])} arguments in the header, {n_args(rs_fns[f])} in lib.rs"
    h_codes = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(GAMS_CUDA_[A-Z]+)\s+\(?(-?\d+)\)?", hdr)}
    rs_codes = {m.group(1): int(m.group(2)) for m in re.finditer(r"pub const (GAMS_CUDA_[A-Z]+): i32 = (-?\d+);", rs)}
    assert h_codes == rs_codes and h_codes["GAMS_CUDA_ECAP"] == -7


def test_window_arithmetic_matches_oracle(orc):
    rng = np.random.default_rng(1)
    for _ in range(300):
        ln, size, step = int(rng.integers(0, 5000)), int(rng.integers(1, 600)), int(rng.integers(1, 700))
        assert api.sliding_count(ln, size, step) == orc.nwin(ln, size, step)
    off = np.array([0, 99, 100, 1000, 1000, 316000], dtype=np.uint64)
    assert np.array_equal(api.win_offsets(off, 100, 10), orc.win_offsets(off, 100, 10))
    assert api.sliding(1, 230218, 100, 10)[:2] == [(1, 100), (11, 110)] and len(api.sliding(1, 230218, 100, 10)) == 23012
    with pytest.raises(ValueError):
        api.sliding_count(10, 0, 1)


def test_partition_is_contiguous_and_balanced():
    rng = np.random.default_rng(2)
    lens = rng.integers(500_000, 1_000_000, 6200)
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.uint64)
    for n in (1, 2, 4, 8):
        p = api.partition(off, n)
        assert p[0] == 0 and p[-1] == 6200 and np.all(np.diff(p) >= 0)
        loads = np.diff(off[p].astype(np.int64))
        assert loads.max() - loads.min() <= 2_000_000
    assert api.partition(np.array([0, 5], dtype=np.uint64), 4).tolist() in ([0, 0, 0, 1, 1], [0, 0, 1, 1, 1], [0, 1, 1, 1, 1], [0, 0, 0, 0, 1])


def test_wave_text_host_layer_matches_oracle_and_golden(orc, s288c):
    """C++ emit + merge_ints (wave.rs:157-252) against the Python restatement and I.peaks.tsv."""
    ctgs = [c for c in cut_ctgs(orc, s288c, 500000) if c[1] == "I"]
    cat, off = orc.pack_ctgs([c[4] for c in ctgs])
    woff, cnt, gc, sig = orc.wave_batch(cat, off, 100, 10, 100, 3.0, 1.0)
    text = orc.WAVE_HEADER + api.wave_text("I", 1, 100, 10, gc, sig, 0.2)
    assert text.encode() == bytes(s288c["peaks_tsv"])
    for cov in (0.2, 1.0, 1.5, 5.0, 11.0):
        assert api.wave_text("I", 1, 100, 10, gc, sig, cov) == orc.wave_text("I", 1, 100, 10, gc, sig, cov)
    assert api.wave_text("I", 1, 100, 10, gc[:500], sig[:500], 0.2, True) == orc.wave_text("I", 1, 100, 10, gc[:500], sig[:500], 0.2, True)
    # other geometries, random signals
    rng = np.random.default_rng(3)
    for size, step in ((1, 1), (7, 3), (50, 100)):
        g = rng.random(400).astype(np.float32)
        s = rng.choice(np.array([-1, 0, 1], dtype=np.int8), 400, p=[0.3, 0.4, 0.3])
        assert api.wave_text("chrX", 1001, size, step, g, s, 0.2) == orc.wave_text("chrX", 1001, size, step, g, s, 0.2)


def test_f32_display_matches_rust_style(orc):
    vals = np.array([0.18, 0.5, 1.0, 0.0, 0.4321, 1e-7, 123456.79, 0.30000001, 16777216.0, np.nan, np.inf], dtype=np.float32)
    rng = np.random.default_rng(4)
    vals = np.concatenate([vals, (rng.integers(0, 101, 200) / np.float32(100)).astype(np.float32), rng.random(200).astype(np.float32)])
    txt = api.wave_text("c", 1, 1, 1, vals, np.ones(vals.size, dtype=np.int8), 0.2, True)
    got = [l.split("\t")[1] for l in txt.splitlines()]
    assert got == [orc.fmt_f32(v) for v in vals]


def test_lapper_blob_view(orc):
    rng = np.random.default_rng(5)
    ivs = [(int(s), int(s) + int(l), "") for s, l in zip(rng.integers(1, 10**6, 1000), rng.integers(1, 50, 1000))]
    lp = orc.lapper_new(ivs)
    blob = orc.lapper_to_bincode(lp)
    st, sp, ml = api.lapper_blob_view(blob)
    assert np.array_equal(st, lp["starts"]) and np.array_equal(sp, lp["stops"]) and ml == lp["max_len"]
    lp2 = orc.lapper_new([(1, 100001, "ctg:I:1"), (100001, 230219, "ctg:I:2")])  # idx:ctg blob with names
    st, sp, ml = api.lapper_blob_view(orc.lapper_to_bincode(lp2))
    assert st.tolist() == [1, 100001] and sp.tolist() == [100001, 230219] and ml == 130218
    st, sp, ml = api.lapper_blob_view(orc.lapper_to_bincode(orc.lapper_new([])))
    assert st.size == 0 and sp.size == 0
    with pytest.raises(api.GamsCudaError) as e:
        api.lapper_blob_view(blob[:-3])
    assert e.value.code == -6


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.GamsCudaError) as e:
        api.Context()
    assert e.value.code == -5
    L = _lib.load()
    out = np.zeros(1, dtype=np.float32)
    rc = L.gams_cuda_wave_batch(None, 0, None, None, 100, 10, 100, 3.0, 1.0, None, None, out.ctypes.data_as(C.c_void_p), None)
    assert rc == -1  # NULL ctx is rejected, not silently computed on the host


def test_gz_sizes_and_decode_gz_oracle(orc, s288c):
    """gams_cuda_gz_sizes is host arithmetic (the ISIZE trailer); the gunzip oracle (zlib) against the reference's own
    gzip fixture: tests/S288c/genome.fa.gz inflates to the FASTA whose sequences the other goldens are built from."""
    blob = bytes(s288c["genome_fa_gz"])
    fa = orc.decode_gz(blob)
    names, seqs, cur = [], [], []
    for line in fa.splitlines():
        if line.startswith(b">"):
            if names:
                seqs.append(b"".join(cur))
            names.append(line[1:].split()[0].decode()); cur = []
        else:
            cur.append(line)
    seqs.append(b"".join(cur))
    assert names == s288c["chr_names"].tolist()
    assert seqs[0] == s288c["seq_I"].tobytes() and seqs[1] == s288c["seq_Mito"].tobytes()
    parts = [orc.encode_gz(s288c["seq_I"].tobytes()), orc.encode_gz(b""), blob]
    cat, off = api.pack_blobs(parts)
    assert api.gz_sizes(cat, off).tolist() == [230218, 0, len(fa)]
    assert orc.decode_gz(parts[0]) == s288c["seq_I"].tobytes()
    with pytest.raises(api.GamsCudaError) as e:
        api.gz_sizes(np.frombuffer(b"not a gzip member at all", dtype=np.uint8), np.array([0, 24], np.uint64))
    assert e.value.code == -1


def test_new_entry_points_reject_bad_arguments_without_a_device():
    """NULL context / NULL arrays are EINVAL on every entry point added for the index build, window counts and gunzip:
    nothing is computed on the host."""
    L = _lib.load()
    z = np.zeros(4, dtype=np.uint64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.gams_cuda_build_idx_rg(None, 1, None, None, p(z), None, None, None, None, None, None) == -1
    assert L.gams_cuda_count_windows(None, 1, None, p(z), 100, 10, None, None, p(z), None) == -1
    assert L.gams_cuda_gunzip_batch(None, 1, None, p(z), p(z), None, None) == -1
    n = C.c_uint64(7)
    assert L.gams_cuda_parse_ranges(None, None, 10, 0, None, None, 0, C.byref(n), None, None, None, None) == -1
    assert L.gams_cuda_gunzip_dev(None, 0, None, 1, None, None, None, None, None) == -1
    h = C.c_void_p()
    assert L.gams_cuda_cwin_plan_create(None, 0, 1, None, p(z), 100, 10, p(z), C.byref(h)) == -1 and not h.value
    assert L.gams_cuda_count_windows_dev(None, None, None, None, None) == -1
    assert L.gams_cuda_cwin_plan_total_windows(None) == 0 and L.gams_cuda_cwin_plan_launches(None) == 0
    L.gams_cuda_cwin_plan_destroy(None)
    L.gams_cuda_cwin_plan_invalidate(None)
    assert b"NULL" in L.gams_cuda_last_error() or b"bad" in L.gams_cuda_last_error()


def test_window_count_binning_multiply_high_is_exact():
    """count_windows.cu bins an index element with floor(d / step) computed as (d * ceil(2^32 / step)) >> 32 for step < 1024 and
    d < 4096 * step (a tile's span).  The claim in the kernel's comment, checked here for every step on the values around every
    multiple of step (where a wrong quotient would first show) and on random ones."""
    rng = np.random.default_rng(0)
    for step in range(2, 1024):
        magic = ((1 << 32) + step - 1) // step
        k = np.arange(0, 4096, dtype=np.uint64)
        d = np.concatenate([k * step, k * step + (step - 1), np.maximum(k * step, 1) - 1,
                            rng.integers(0, 4096 * step, 2000).astype(np.uint64)])
        d = d[d < 4096 * step]
        assert np.array_equal((d * np.uint64(magic)) >> np.uint64(32), d // np.uint64(step)), step
