"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Needs a B200: -m gpu.

Bar: bit-exact for GC counts, GC fractions (identical f32), signals, overlap counts, window
coordinates.  (north_star's "peak signals identical except |z - threshold| < 1e-9" is met with
zero exceptions because the kernels replay the reference's f32 evaluation order.)
"""
import numpy as np
import pytest

from conftest import cut_ctgs
from gams_b200 import api, synth
from test_oracle_golden import KAT_INPUT, KAT_SIGNALS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context([0])
    yield c
    c.close()


def _eq(name, got, ref):
    if not np.array_equal(got, ref, equal_nan=True):
        bad = np.flatnonzero(~((got == ref) | (np.isnan(got.astype(np.float64)) & np.isnan(ref.astype(np.float64)))))
        raise AssertionError(f"{name}: {bad.size} of {ref.size} differ; first at {bad[:8].tolist()} got {got[bad[:8]].tolist()} ref {ref[bad[:8]].tolist()}")


def _check_wave(ctx, orc, seqs, size, step, lag, thr, infl, signal=True):
    cat, off = orc.pack_ctgs(seqs)
    woff, cnt, gc, sig = api.wave_batch(ctx, cat, off, size, step, lag, thr, infl, want_signal=signal)
    rwoff, rcnt, rgc, rsig = orc.wave_batch(cat, off, size, step, lag, thr, infl, threads=8, want_signal=signal)
    _eq("win_off", woff, rwoff)
    _eq(f"gc_count[{size}/{step}]", cnt, rcnt)
    _eq(f"gc[{size}/{step}]", gc.view(np.uint32), rgc.view(np.uint32))
    if signal:
        _eq(f"signal[{size}/{step} lag {lag} thr {thr} infl {infl}]", sig, rsig)
    return woff, cnt, gc, sig


# ------------------------------------------------------------------------------------------------ wave
def test_wave_fixture_is_byte_identical_to_I_peaks_tsv(ctx, orc, s288c):
    """BASELINE config 1 on the GPU: README.md:155-170 → tests/S288c/I.peaks.tsv."""
    ctgs = cut_ctgs(orc, s288c, 500000)
    woff, cnt, gc, sig = _check_wave(ctx, orc, [c[4] for c in ctgs], 100, 10, 100, 3.0, 1.0)
    i = [k for k, c in enumerate(ctgs) if c[1] == "I"][0]
    a, b = int(woff[i]), int(woff[i + 1])
    text = orc.WAVE_HEADER + api.wave_text("I", 1, 100, 10, gc[a:b], sig[a:b], 0.2)
    assert text.encode() == bytes(s288c["peaks_tsv"])


def test_wave_fixture_piece_100000(ctx, orc, s288c):
    """tests/cli.rs:332-364: 116 stdout lines with the two pinned substrings."""
    ctgs = [c for c in cut_ctgs(orc, s288c, 100000) if c[1] == "I"]
    woff, cnt, gc, sig = _check_wave(ctx, orc, [c[4] for c in ctgs], 100, 10, 100, 3.0, 1.0)
    text = orc.WAVE_HEADER + "".join(
        api.wave_text(c[1], c[2], 100, 10, gc[int(woff[k]):int(woff[k + 1])], sig[int(woff[k]):int(woff[k + 1])], 0.2)
        for k, c in enumerate(ctgs))
    assert len(text.splitlines()) == 116 and "I:7551-7650\t" in text and "I(+):11551-11740\t" in text


@pytest.mark.parametrize("size,step,lag", [
    (100, 10, 100), (100, 20, 50), (100, 5, 200), (97, 13, 31), (1, 1, 100), (16, 16, 8), (15, 17, 7),
    (500, 500, 10), (33, 64, 3), (1000, 7, 100), (9000, 1000, 5), (20000, 5000, 2), (40000, 10000, 2),
])
def test_wave_synthetic_geometries(ctx, orc, size, step, lag):
    """Ragged ctgs (lengths not multiples of 16, one exactly `size` long... ), lowercase, N runs."""
    g = synth.genome(1_300_000, seed=size * 131 + step)
    need = size + (lag + 2) * step
    cuts = [0, need + 17, 2 * need + 17 + 40_001, 700_003, 700_003 + need + lag * step, 1_300_000]
    cuts = sorted(set(min(c, 1_300_000) for c in cuts))
    seqs = [g[a:b] for a, b in zip(cuts[:-1], cuts[1:]) if (b - a - size) // step + 1 >= lag]
    assert len(seqs) >= 2
    for infl in (1.0, 0.3):
        _check_wave(ctx, orc, seqs, size, step, lag, 2.5, infl)


def test_wave_degenerate_constant_gc(ctx, orc):
    """Constant-GC stretches: mean of 100 copies of 0.18f is 0.18000013f and sd is pure rounding noise
     — the outcome is decided by summation order, so this pins the exact order."""
    unit = np.frombuffer(b"GC" + b"A" * 48 + b"G" * 7 + b"T" * 43, dtype=np.uint8)  # 9 GC per 100 -> varying per phase
    flat = np.tile(np.frombuffer(b"ACATTTTTAT" * 5 + b"GCGCGCGCGATATATATATATATATATATATATATATATATATATATATA", dtype=np.uint8), 3000)
    per = np.tile(unit, 3000)
    mix = np.concatenate([flat[:120_000], synth.genome(50_000, seed=5), per[:130_000], np.full(30_000, ord("N"), np.uint8),
                          np.tile(np.frombuffer(b"AC", dtype=np.uint8), 40_000)])
    for thr in (3.0, 1.0, 0.0):
        for infl in (1.0, 0.5, 0.0):
            _check_wave(ctx, orc, [mix, flat[:77_777], per[:50_001]], 100, 10, 100, thr, infl)


def test_wave_gc_only_and_empty_ctgs(ctx, orc):
    """ctgs shorter than one window yield zero windows; with signals requested that is ELAG
    (the reference panics at stat.rs:30), without them it is fine."""
    g = synth.genome(50_000, seed=9)
    seqs = [g[:99], g[99:30_000], g[30_000:30_000], g[30_000:]]
    _check_wave(ctx, orc, seqs, 100, 10, 100, 3.0, 1.0, signal=False)
    cat, off = orc.pack_ctgs(seqs)
    with pytest.raises(api.GamsCudaError) as e:
        api.wave_batch(ctx, cat, off, 100, 10, 100, 3.0, 1.0)
    assert e.value.code == -4
    with pytest.raises(api.GamsCudaError) as e:
        api.wave_batch(ctx, cat, off, 100, 10, 0, 3.0, 1.0)
    assert e.value.code == -1
    woff, cnt, gc, sig = api.wave_batch(ctx, np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert gc.size == 0


def test_wave_exactly_lag_windows(ctx, orc):
    g = synth.genome(5000, seed=10)
    _check_wave(ctx, orc, [g[:100 + 99 * 10], g[:100 + 100 * 10], g[:100 + 101 * 10]], 100, 10, 100, 3.0, 1.0)
    _check_wave(ctx, orc, [g[:100 + 99 * 10], g[:100 + 100 * 10], g[:100 + 101 * 10]], 100, 10, 100, 3.0, 0.0)


def test_wave_non_ascii_bytes(ctx, orc):
    """Any byte value is legal input; only c g C G count (bio gc_content)."""
    rng = np.random.default_rng(12)
    g = rng.integers(0, 256, 400_000, dtype=np.uint8)
    _check_wave(ctx, orc, [g[:123_457], g[123_457:]], 100, 10, 100, 3.0, 1.0)


# ------------------------------------------------------------------------------------ thresholding
def test_thresholding_kat(ctx):
    """src/libs/stat.rs:58-81."""
    assert api.thresholding_algo(ctx, KAT_INPUT, 30, 5.0, 0.0).tolist() == KAT_SIGNALS


@pytest.mark.parametrize("lag", [1, 2, 3, 7, 8, 9, 30, 100, 257])
@pytest.mark.parametrize("infl", [1.0, 0.0, 0.5])
def test_thresholding_random_series(ctx, orc, lag, infl):
    rng = np.random.default_rng(lag * 7 + int(infl * 10))
    series = [rng.normal(1.0, 0.1, n).astype(np.float32) for n in (lag, lag + 1, lag + 2, 1000, 4097, 2048 + lag)]
    for s in series:
        s[rng.integers(0, s.size, max(1, s.size // 50))] += rng.normal(0, 2, max(1, s.size // 50)).astype(np.float32)
    data = np.concatenate(series)
    off = np.concatenate(([0], np.cumsum([s.size for s in series]))).astype(np.uint64)
    for thr in (3.0, 1.5):
        got = api.thresholding_batch(ctx, data, off, lag, thr, infl)
        ref = np.concatenate([orc.thresholding_algo(s, lag, thr, infl) for s in series])
        _eq(f"thresholding lag {lag} infl {infl} thr {thr}", got, ref)


def test_thresholding_short_series_is_an_error(ctx):
    with pytest.raises(api.GamsCudaError) as e:
        api.thresholding_algo(ctx, [0.1] * 10, 30, 3.0, 1.0)
    assert e.value.code == -4


# ------------------------------------------------------------------------------------ overlap counts
def _fixture_index(orc, s288c):
    ctgs = cut_ctgs(orc, s288c, 100000)
    idx_ctg = orc.build_idx_ctg([c[:4] for c in ctgs])
    ranges_of = orc.read_range(bytes(s288c["snp_rg"]).decode().splitlines(), idx_ctg)
    idx_rg = orc.build_idx_rg(ranges_of)
    ids = [c[0] for c in ctgs]
    lps = [idx_rg.get(i, orc.lapper_new([])) for i in ids]
    blobs = [orc.lapper_to_bincode(lp) for lp in lps]
    views = [api.lapper_blob_view(b) for b in blobs]  # the path a Rust caller takes: slice the idx:rg blob
    starts = np.concatenate([v[0] for v in views]).astype(np.uint32)
    stops = np.concatenate([v[1] for v in views]).astype(np.uint32)
    iv_off = np.concatenate(([0], np.cumsum([v[0].size for v in views]))).astype(np.uint64)
    return ctgs, ids, idx_ctg, idx_rg, starts, stops, iv_off


def test_locate_count_fixture(ctx, orc, s288c):
    """BASELINE config 2 / tests/cli.rs:426-454: I:1000-2000 → 12, Mito:1000-2000 → 0, II dropped."""
    ctgs, ids, idx_ctg, idx_rg, starts, stops, iv_off = _fixture_index(orc, s288c)
    chrs = sorted({c[1] for c in ctgs})
    by_chr = {ch: sorted((c for c in ctgs if c[1] == ch), key=lambda c: c[2]) for ch in chrs}
    order = [c for ch in chrs for c in by_chr[ch]]
    chr_off = np.concatenate(([0], np.cumsum([len(by_chr[ch]) for ch in chrs]))).astype(np.uint64)
    cs = np.array([c[2] for c in order], dtype=np.uint32)
    ce = np.array([c[3] + 1 for c in order], dtype=np.uint32)
    qs = ["I:1000-2000", "II:1001-2000", "Mito:1000-2000", "I:100000-100001", "I:100001", "I:1", "I:230218-230300"]
    rgs = [orc.range_from_str(q) for q in qs]
    q_chr = np.array([chrs.index(r[0]) if r[0] in chrs else -1 for r in rgs], dtype=np.int32)
    q_s = np.array([r[2] for r in rgs], dtype=np.uint32)
    q_e = np.array([r[3] for r in rgs], dtype=np.uint32)
    loc = api.locate_ctg(ctx, cs, ce, chr_off, q_chr, q_s, q_e)
    ref_ids = [orc.find_one_idx(idx_ctg, r) for r in rgs]
    assert [order[i][0] if i >= 0 else "" for i in loc] == ref_ids
    seg = np.array([ids.index(order[i][0]) if i >= 0 else -1 for i in loc], dtype=np.int32)
    cnt = api.count_overlaps(ctx, starts, stops, iv_off, seg, q_s, q_e)
    out = [f"{q}\t{c}" for q, c, i in zip(qs, cnt, loc) if i >= 0]
    assert out[:2] == ["I:1000-2000\t12", "Mito:1000-2000\t0"]
    ref = [orc.count_rg(idx_rg, rid, r) if rid else 0 for rid, r in zip(ref_ids, rgs)]
    assert cnt.tolist() == ref


@pytest.mark.parametrize("grouped", [True, False])
def test_count_overlaps_random(ctx, orc, grouped):
    lens = [500_000, 999_999, 5_000, 700_001, 1, 640_000]
    starts, stops, iv_off = synth.rg_index(lens, [16_000, 40_000, 0, 3, 1, 20_000], seed=3, max_len=1)
    s2, e2, o2 = synth.rg_index(lens, [1000, 5000, 10, 0, 0, 800], seed=4, max_len=5000)  # real intervals, duplicates
    st = np.concatenate([np.sort(np.concatenate([starts[iv_off[i]:iv_off[i + 1]], s2[o2[i]:o2[i + 1]]])) for i in range(6)])
    sp = np.concatenate([np.sort(np.concatenate([stops[iv_off[i]:iv_off[i + 1]], e2[o2[i]:o2[i + 1]]])) for i in range(6)])
    off = (iv_off + o2).astype(np.uint64)
    seg, qs, qe = synth.queries(lens, 300_000, seed=5, grouped=grouped)
    seg[::1000] = -1
    seg[7::1000] = 6
    # edge probes: query ending exactly at an interval start, point queries on a SNP
    k = 5000
    qs[:k] = st[off[1]:off[1] + k]
    qe[:k] = qs[:k]
    seg[:k] = 1
    qe[k:2 * k] = st[off[1]:off[1] + k]
    qs[k:2 * k] = np.maximum(qe[k:2 * k].astype(np.int64) - 50, 1).astype(np.uint32)
    seg[k:2 * k] = 1
    if grouped:
        o = np.argsort(seg, kind="stable")
        seg, qs, qe = seg[o], qs[o], qe[o]
    got = api.count_overlaps(ctx, st, sp, off, seg, qs, qe)
    ref = orc.count_batch(st, sp, off, seg, qs, qe, threads=8)
    _eq("overlap counts", got, ref)
    assert got.sum() > 0


# ------------------------------------------------------------------------------------ gc ranges / sw
def test_gc_ranges(ctx, orc):
    g = synth.genome(600_000, seed=21)
    seqs = [g[:250_003], g[250_003:250_003], g[250_003:]]
    cat, off = orc.pack_ctgs(seqs)
    rng = np.random.default_rng(22)
    n = 20_000
    qc = rng.choice([0, 2], n).astype(np.uint32)
    ln = np.diff(off.astype(np.int64))[qc]
    a = (rng.random(n) * ln).astype(np.int64)
    b = np.minimum(a + rng.integers(0, 3000, n), ln)
    a[:10], b[:10] = 0, ln[:10]
    got = api.gc_ranges(ctx, cat, off, qc, a.astype(np.uint32), b.astype(np.uint32))
    isgc = np.isin(cat, np.frombuffer(b"cgCG", dtype=np.uint8))
    cs = np.concatenate(([0], np.cumsum(isgc)))
    base = off.astype(np.int64)[qc]
    _eq("gc_ranges", got, (cs[base + b] - cs[base + a]).astype(np.uint32))


def test_sw_batch_fixture_features(ctx, orc, s288c):
    """`gams sw` rows for the spo11 features (tests/cli.rs:306-330) against the oracle, all columns."""
    ctgs = cut_ctgs(orc, s288c, 100000)
    idx_ctg = orc.build_idx_ctg([c[:4] for c in ctgs])
    feats = orc.read_range(bytes(s288c["spo11_rg"]).decode().splitlines(), idx_ctg)
    cat, off = orc.pack_ctgs([c[4] for c in ctgs])
    f_ctg, f_s, f_e = [], [], []
    for k, c in enumerate(ctgs):
        for rg in feats.get(c[0], []):
            f_ctg.append(k), f_s.append(rg[2]), f_e.append(rg[3])
    # plus features at ctg edges and tiny / huge ones
    for k, c in enumerate(ctgs):
        for s, e in ((c[2], c[2]), (c[3], c[3]), (c[2], c[2] + 49), (c[3] - 20, c[3]), (c[2] + 1000, c[2] + 1001), (c[2], c[3])):
            f_ctg.append(k), f_s.append(s), f_e.append(e)
    r = api.sw_batch(ctx, cat, off, [c[2] for c in ctgs], f_ctg, f_s, f_e)
    assert int(r["row_off"][-1]) > 2000
    for f in range(len(f_ctg)):
        c = ctgs[f_ctg[f]]
        ref = orc.sw_feature(c[4], c[2], c[3], f_s[f], f_e[f])
        a, b = int(r["row_off"][f]), int(r["row_off"][f + 1])
        assert b - a == len(ref["gc"]), (f, b - a, len(ref["gc"]))
        for key in ("start", "end", "type", "dist"):
            _eq(f"sw {key} feature {f}", r[key][a:b].astype(np.int64), ref[key].astype(np.int64))
        for key in ("gc", "mean", "sd", "cv"):
            _eq(f"sw {key} feature {f}", r[key][a:b].view(np.uint32), ref[key].view(np.uint32))


def test_sw_batch_other_geometry(ctx, orc):
    g = synth.genome(300_000, seed=31)
    cat, off = orc.pack_ctgs([g[:120_000], g[120_000:]])
    rng = np.random.default_rng(32)
    f_ctg = rng.integers(0, 2, 300)
    starts = [1, 500_001]
    lens = [120_000, 180_000]
    f_s = np.array([starts[c] + rng.integers(0, lens[c] - 3000) for c in f_ctg])
    f_e = f_s + rng.integers(0, 3000, 300)
    for size, mx, rz in ((100, 20, 500), (50, 3, 1000), (101, 5, 333)):
        r = api.sw_batch(ctx, cat, off, starts, f_ctg, f_s, f_e, size, mx, rz)
        for f in range(300):
            c = int(f_ctg[f])
            seq = g[:120_000] if c == 0 else g[120_000:]
            ref = orc.sw_feature(seq, starts[c], starts[c] + lens[c] - 1, int(f_s[f]), int(f_e[f]), size, mx, rz)
            a, b = int(r["row_off"][f]), int(r["row_off"][f + 1])
            assert b - a == len(ref["gc"])
            for key in ("gc", "mean", "sd", "cv"):
                _eq(f"sw {key} ({size},{mx},{rz}) feature {f}", r[key][a:b].view(np.uint32), ref[key].view(np.uint32))
            _eq("sw start", r["start"][a:b], ref["start"])


# ------------------------------------------------------------------- certified z-score filter (Kernel B)
def _gc_series(orc, seq, size=100, step=10):
    cat, off = orc.pack_ctgs([seq])
    _, _, gc, _ = orc.wave_batch(cat, off, size, step, 100, 3.0, 1.0, want_signal=False)
    return gc


def test_zscore_filter_certificates_are_sound(ctx, orc, s288c):
    """Mode 2 replays every window and counts certified decisions that disagree: must be zero, on the
    fixture, on synthetic genomes, on degenerate constant-GC data and on non-GC series."""
    rng = np.random.default_rng(41)
    flat = np.tile(np.frombuffer(b"ACATTTTTAT" * 5 + b"GCGCGCGCGATATATATATATATATATATATATATATATATATATATATA", dtype=np.uint8), 1500)
    series = [_gc_series(orc, s288c["seq_I"]), _gc_series(orc, s288c["seq_Mito"]), _gc_series(orc, synth.genome(400_000, seed=42)),
              _gc_series(orc, flat), _gc_series(orc, np.full(30_000, ord("N"), np.uint8)),
              rng.normal(1.0, 0.1, 20_000).astype(np.float32), rng.normal(-500.0, 1e-3, 20_000).astype(np.float32),
              (rng.integers(0, 3, 20_000) * np.float32(1e-30)).astype(np.float32), rng.normal(0, 1e4, 20_000).astype(np.float32)]
    data = np.concatenate(series)
    off = np.concatenate(([0], np.cumsum([s.size for s in series]))).astype(np.uint64)
    try:
        for lag in (100, 30, 8, 2):
            for thr in (3.0, 0.0, 1.0, 2.9999, 7.5):
                ref = np.concatenate([orc.thresholding_algo(s, lag, thr, 1.0) for s in series])
                ctx.zscore_stats(reset=True)
                for mode in (2, 1, 0):
                    ctx.set_zscore_mode(mode)
                    got = api.thresholding_batch(ctx, data, off, lag, thr, 1.0)
                    _eq(f"mode {mode} lag {lag} thr {thr}", got, ref)
                    unc, bad = ctx.zscore_stats(reset=True)
                    assert bad == 0, f"{bad} certified decisions disagree with the exact replay (lag {lag}, thr {thr})"
    finally:
        ctx.set_zscore_mode(0)


def test_zscore_filter_borderline_thresholds(ctx, orc, s288c):
    """Thresholds placed exactly on (and one ulp around) the z value of real windows, so some decision is
    as close to a tie as f32 allows; the filter must hand those windows to the exact replay."""
    gc = _gc_series(orc, s288c["seq_I"])
    lag = 100
    rng = np.random.default_rng(43)
    picks = rng.integers(lag + 1, gc.size, 60)
    thrs = []
    for i in picks:
        w = gc[i - 1 - lag:i - 1]
        m, sd = orc.mean(w), orc.stddev(w)
        if sd > 0:
            z = np.float32(abs(np.float32(gc[i] - m)) / sd)
            thrs += [z, np.nextafter(z, np.float32(0)), np.nextafter(z, np.float32(10))]
    off = np.array([0, gc.size], dtype=np.uint64)
    total_unc = 0
    ctx.zscore_stats(reset=True)
    for thr in thrs:
        ref = orc.thresholding_algo(gc, lag, float(thr), 1.0)
        got = api.thresholding_batch(ctx, gc, off, lag, float(thr), 1.0)
        _eq(f"borderline thr {thr!r}", got, ref)
        total_unc += ctx.zscore_stats(reset=True)[0]
    assert total_unc > 0  # the borderline windows did go through the replay
    assert total_unc < 0.01 * gc.size * len(thrs)  # ...and only a sliver of all windows did


def test_zscore_filter_replay_fraction_on_genome(ctx, orc):
    gc = _gc_series(orc, synth.genome(2_000_000, seed=44))
    ctx.zscore_stats(reset=True)
    got = api.thresholding_batch(ctx, gc, np.array([0, gc.size], dtype=np.uint64), 100, 3.0, 1.0)
    _eq("genome", got, orc.thresholding_algo(gc, 100, 3.0, 1.0))
    unc, _ = ctx.zscore_stats(reset=True)
    assert unc < 0.005 * gc.size, f"{unc} of {gc.size} windows needed the exact replay"


# ------------------------------------------------------------------- fused GC-wave kernel (Kernel AB)
def _wave_all_modes(ctx, orc, seqs, size, step, lag, thr, expect_fused=None):
    """Fused kernel (mode 0) vs Kernel A + exact replay of every window (mode 1) vs the oracle."""
    cat, off = orc.pack_ctgs(seqs)
    plan = api.WavePlan(ctx, off, size, step, lag)
    if expect_fused is not None:   # the certificate's error model covers thresholds in [2^-16, 2^30] (wave_fused_ok)
        assert plan.fused(thr, 1.0) == (expect_fused and 2.0 ** -16 <= thr <= 2.0 ** 30), (size, step, lag, thr)
    plan.close()
    _, rcnt, rgc, rsig = orc.wave_batch(cat, off, size, step, lag, thr, 1.0, threads=8)
    out = {}
    try:
        for mode in (1, 0):
            ctx.set_zscore_mode(mode)
            ctx.zscore_stats(reset=True)
            _, cnt, gc, sig = api.wave_batch(ctx, cat, off, size, step, lag, thr, 1.0)
            _eq(f"mode {mode} count", cnt, rcnt)
            _eq(f"mode {mode} gc", gc.view(np.uint32), rgc.view(np.uint32))
            _eq(f"mode {mode} signal [{size}/{step} lag {lag} thr {thr}]", sig, rsig)
            out[mode] = ctx.zscore_stats(reset=True)[0]
    finally:
        ctx.set_zscore_mode(0)
    return out[0], rgc, rsig


@pytest.mark.parametrize("size,step,lag", [(100, 10, 100), (100, 10, 2), (100, 50, 30), (100, 100, 10), (1, 1, 100),
                                           (16, 16, 8), (1000, 100, 20), (5000, 1000, 5), (30, 10, 400), (64, 8, 1000),
                                           (64, 8, 100), (100, 5, 200), (160, 16, 101), (120, 12, 30), (12, 12, 33)])
def test_fused_wave_matches_replay_and_oracle(ctx, orc, size, step, lag):
    g = synth.genome(1_200_000, seed=size * 7 + lag)
    flat = np.tile(np.frombuffer(b"ACATTTTTAT" * 5 + b"GCGCGCGCGATATATATATATATATATATATATATATATATATATATATA", dtype=np.uint8), 1500)
    need = size + (lag + 2) * step
    seqs = [g[:400_003], g[400_003:400_003 + need], g[500_000:], flat, np.full(max(need, 20_000), ord("N"), np.uint8),
            np.concatenate([g[:need], flat[:need], g[need:3 * need]])]
    # eligibility of the fused kernel (fused_windows_per_tile, gc_wave_fused.cu): a block's bits come from one 32-bit
    # word, D and V exact in 32 bits, the lag+1-window halo at most a quarter of a tile
    n_el = min((256 * 6 * 16 - 31 - size) // step + 1, 2048)
    fused = (step <= 16 and size % step == 0 and size // step <= 248 and 2 <= lag <= 400 and lag * size < 65536
             and n_el - (lag + 1) >= max(4 * (lag + 1), 64))
    assert fused == ((size, step, lag) in {(100, 10, 100), (100, 10, 2), (1, 1, 100), (16, 16, 8), (30, 10, 400), (64, 8, 100),
                                           (100, 5, 200), (160, 16, 101), (120, 12, 30), (12, 12, 33)})
    for thr in (3.0, 0.0, 1.0, 7.5):
        _wave_all_modes(ctx, orc, seqs, size, step, lag, thr, expect_fused=fused)


def test_fused_wave_borderline_thresholds(ctx, orc, s288c):
    """Thresholds one ulp around the z value of real windows: the integer certificate must send those
    windows to the exact replay (and only a sliver of all windows)."""
    seq = np.asarray(s288c["seq_I"])
    gc = _gc_series(orc, seq)
    lag = 100
    rng = np.random.default_rng(47)
    thrs = []
    for i in rng.integers(lag + 1, gc.size, 40):
        w = gc[i - 1 - lag:i - 1]
        m, sd = orc.mean(w), orc.stddev(w)
        if sd > 0:
            z = np.float32(abs(np.float32(gc[i] - m)) / sd)
            thrs += [z, np.nextafter(z, np.float32(0)), np.nextafter(z, np.float32(10))]
    total = 0
    for thr in thrs:
        unc, _, _ = _wave_all_modes(ctx, orc, [seq], 100, 10, lag, float(thr), expect_fused=True)
        total += unc
    assert total > 0
    assert total < 0.01 * gc.size * len(thrs)


def test_fused_wave_replay_fraction_and_fallbacks(ctx, orc):
    g = synth.genome(2_000_000, seed=48)
    unc, rgc, _ = _wave_all_modes(ctx, orc, [g], 100, 10, 100, 3.0, expect_fused=True)
    assert unc < 0.005 * rgc.size, f"{unc} of {rgc.size} windows needed the exact replay"
    # geometries / statistics the fused kernel does not cover take Kernel A + Kernel B
    _wave_all_modes(ctx, orc, [g[:300_000]], 97, 13, 31, 3.0, expect_fused=False)      # size % step != 0
    _wave_all_modes(ctx, orc, [g[:300_000]], 100, 10, 100, -1.0, expect_fused=False)    # negative threshold
    _wave_all_modes(ctx, orc, [g[:300_000]], 20000, 5000, 2, 3.0, expect_fused=False)   # step > 16
    _wave_all_modes(ctx, orc, [g[:300_000]], 700, 10, 100, 3.0, expect_fused=False)     # lag * size >= 2^16
    plan = api.WavePlan(ctx, np.array([0, 300_000], dtype=np.uint64), 100, 10, 100)
    assert plan.fused(3.0, 1.0) and not plan.fused(3.0, 0.5) and not plan.fused(float("nan"), 1.0)
    plan.close()


def test_count_overlaps_adversarial_distributions(ctx, orc):
    """Rank tables must stay exact on skewed endpoint distributions: duplicates, far-apart clusters,
    coordinates near 2^32, single-interval and empty segments."""
    rng = np.random.default_rng(51)
    segs = [
        (np.full(5000, 777, np.uint32), np.full(5000, 778, np.uint32)),                                    # all identical
        (np.concatenate([rng.integers(1, 50, 3000), rng.integers(4_000_000_000, 4_000_000_100, 3000)]).astype(np.uint32), None),
        (rng.integers(1, 4_294_967_000, 20000).astype(np.uint32), None),                                   # huge coordinates
        (np.array([123456], np.uint32), np.array([123999], np.uint32)),                                    # single interval
        (np.zeros(0, np.uint32), np.zeros(0, np.uint32)),                                                  # empty
        (np.repeat(rng.integers(1, 10_000, 50), 400).astype(np.uint32), None),                             # heavy duplicates
        (np.arange(1, 30001, dtype=np.uint32) * 3, None),                                                  # perfectly regular
    ]
    st, sp, off = [], [], [0]
    for s, e in segs:
        s = np.sort(s)
        e = np.sort(e if e is not None else (s + rng.integers(1, 200, s.size).astype(np.uint32)))
        st.append(s), sp.append(e), off.append(off[-1] + s.size)
    st, sp, off = np.concatenate(st), np.concatenate(sp), np.asarray(off, dtype=np.uint64)
    n_q = 200_000
    seg = rng.integers(0, len(segs), n_q).astype(np.int32)
    pool = np.concatenate([st, sp, np.array([0, 1, 2, 4_294_967_294, 4_294_967_295], np.uint32)])
    qs = pool[rng.integers(0, pool.size, n_q)].astype(np.int64) + rng.integers(-2, 3, n_q)
    ln = np.where(rng.random(n_q) < 0.3, 0, rng.integers(0, 5000, n_q))
    qs = np.clip(qs, 0, 4_294_967_295)
    qe = np.clip(qs + ln, 0, 4_294_967_295)
    got = api.count_overlaps(ctx, st, sp, off, seg, qs.astype(np.uint32), qe.astype(np.uint32))
    ref = orc.count_batch(st, sp, off, seg, qs.astype(np.uint32), qe.astype(np.uint32), threads=8)
    _eq("adversarial overlap counts", got, ref)


def test_thresholding_lag_longer_than_a_tile(ctx, orc):
    """lag > 1024 (one z-score tile): the first tiles hold only windows before `lag`."""
    rng = np.random.default_rng(61)
    s = rng.normal(0.45, 0.05, 9000).astype(np.float32)
    s[rng.integers(0, s.size, 60)] += 0.5
    for lag in (1500, 2500):
        got = api.thresholding_algo(ctx, s, lag, 3.0, 1.0)
        _eq(f"lag {lag}", got, orc.thresholding_algo(s, lag, 3.0, 1.0))


# ------------------------------------------------------------------------- several devices in one ctx
def test_multi_device_context_shards_by_ctg(orc):
    """gams_cuda_create over every visible device: the host API shards ctgs / queries over them and the
    gathered result is identical to the oracle (and so to the single-device result)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    g = synth.genome(3_000_000, seed=71)
    cuts = np.sort(np.random.default_rng(72).choice(np.arange(20_000, 2_980_000), 11, replace=False))
    seqs = [g[a:b] for a, b in zip(np.concatenate(([0], cuts)), np.concatenate((cuts, [g.size])))]
    with api.Context() as mctx:
        assert mctx.num_devices == torch.cuda.device_count()
        for infl in (1.0, 0.5):
            _check_wave(mctx, orc, seqs, 100, 10, 100, 3.0, infl)
        lens = [s.size for s in seqs]
        st, sp, off = synth.rg_index(lens, [5000] * len(lens), seed=73)
        seg, qs, qe = synth.queries(lens, 400_000, seed=74, grouped=False)
        _eq("multi-device counts", api.count_overlaps(mctx, st, sp, off, seg, qs, qe), orc.count_batch(st, sp, off, seg, qs, qe, threads=8))


def test_count_overlaps_grouped_device_path(ctx, orc):
    """The grouped (q_seg non-decreasing) variant of Kernel C: one contiguous share of queries per SM."""
    import torch
    lens = [500_000 + 37 * k for k in range(300)]
    st, sp, off = synth.rg_index(lens, [8000 + (k % 7) * 1000 for k in range(300)], seed=81)
    seg, qs, qe = synth.queries(lens, 2_500_000, seed=82, grouped=True)
    dev = torch.device("cuda", 0)
    t = lambda a, dt: torch.as_tensor(a.view(dt) if a.dtype != dt else a, device=dev)
    d_st, d_sp = t(st, np.int32), t(sp, np.int32)
    d_seg, d_qs, d_qe = t(seg, np.int32), t(qs, np.int32), t(qe, np.int32)
    plan = api.CountPlan(ctx, off)
    out = torch.empty(seg.size, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ref = orc.count_batch(st, sp, off, seg, qs, qe, threads=8)
    for grouped in (True, False):
        out.zero_()
        plan.count(stream, d_st, d_sp, d_seg, d_qs, d_qe, out, grouped=grouped)
        torch.cuda.synchronize()
        _eq(f"grouped={grouped}", out.cpu().numpy().view(np.uint32), ref)
    plan.close()


def test_count_overlaps_staged_adversarial(ctx, orc):
    """The staged (grouped) kernel on every shape its compact index has to special-case: identical endpoints, clusters
    2^32 apart (no compact form), more elements than a staged array holds, empty and single-interval segments, a
    crowded bucket, segments with no or very few queries, query segment ids outside the index, a q_seg pointer that is
    not 16-byte aligned, and query chunks that end inside a segment."""
    import torch
    rng = np.random.default_rng(91)
    segs = [
        (np.full(5000, 777, np.uint32), np.full(5000, 778, np.uint32)),                                    # all identical
        (np.concatenate([rng.integers(1, 50, 3000), rng.integers(4_000_000_000, 4_000_000_100, 3000)]).astype(np.uint32), None),
        (rng.integers(1, 4_294_967_000, 20000).astype(np.uint32), None),                                   # huge coordinates
        (np.array([123456], np.uint32), np.array([123999], np.uint32)),                                    # single interval
        (np.zeros(0, np.uint32), np.zeros(0, np.uint32)),                                                  # empty
        (np.repeat(rng.integers(1, 10_000, 50), 400).astype(np.uint32), None),                             # crowded buckets
        (np.arange(1, 30001, dtype=np.uint32) * 3, None),                                                  # n > ST_MAX_N
        (rng.integers(1, 500_000, 16000).astype(np.uint32), None),                                         # the C5 shape
        (rng.integers(1, 900_000, 24000).astype(np.uint32), None),                                         # l = 7
        (rng.integers(1, 2_000_000, 24000).astype(np.uint32), None),                                       # l = 8: plain
        (rng.integers(1, 500_000, 16000).astype(np.uint32), None),                                         # gets no queries
        (rng.integers(1, 500_000, 16000).astype(np.uint32), None),                                         # gets 3 queries
    ]
    st, sp, off = [], [], [0]
    for a, e in segs:
        a = np.sort(a)
        e = np.sort(e if e is not None else (a + rng.integers(1, 200, a.size).astype(np.uint32)))
        st.append(a), sp.append(e), off.append(off[-1] + a.size)
    st, sp, off = np.concatenate(st), np.concatenate(sp), np.asarray(off, dtype=np.uint64)
    n_q = 400_000
    w = np.array([3, 3, 3, 1, 1, 3, 3, 40, 10, 10, 0, 0], dtype=np.float64)
    seg = rng.choice(len(segs), size=n_q, p=w / w.sum()).astype(np.int32)
    seg[:7] = [-3, -1, -1, 11, 11, 11, 12]            # outside the index below and above; 3 queries for segment 11
    seg[7:40] = 57
    seg.sort()
    pool = np.concatenate([st, sp, np.array([0, 1, 2, 4_294_967_294, 4_294_967_295], np.uint32)])
    qs = pool[rng.integers(0, pool.size, n_q)].astype(np.int64) + rng.integers(-2, 3, n_q)
    ln = np.where(rng.random(n_q) < 0.3, 0, rng.integers(0, 5000, n_q))
    qs = np.clip(qs, 0, 4_294_967_295)
    qe = np.clip(qs + ln, 0, 4_294_967_295)
    qs, qe = qs.astype(np.uint32), qe.astype(np.uint32)
    ref = orc.count_batch(st, sp, off, seg, qs, qe, threads=8)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.as_tensor(a.view(np.int32) if a.dtype != np.int32 else a, device=dev)
    d_st, d_sp, d_qs, d_qe = t(st), t(sp), t(qs), t(qe)
    plan = api.CountPlan(ctx, off)
    stream = torch.cuda.current_stream().cuda_stream
    for shift in (0, 1, 3):                           # shift > 0: q_seg (and the rest) start off a 16-byte boundary
        d_seg = torch.empty(n_q + 4, dtype=torch.int32, device=dev)[shift:shift + n_q - shift]
        d_seg.copy_(t(seg[shift:]))
        out = torch.full((n_q - shift,), -1, dtype=torch.int32, device=dev)
        plan.count(stream, d_st, d_sp, d_seg, d_qs[shift:], d_qe[shift:], out, grouped=True)
        torch.cuda.synchronize()
        _eq(f"staged adversarial (shift {shift})", out.cpu().numpy().view(np.uint32), ref[shift:])
    # every query outside the index; a single query
    d_seg = torch.full((1000,), 99, dtype=torch.int32, device=dev)
    out = torch.full((1000,), -1, dtype=torch.int32, device=dev)
    plan.count(stream, d_st, d_sp, d_seg, d_qs[:1000], d_qe[:1000], out, grouped=True)
    torch.cuda.synchronize()
    assert int(out.abs().sum()) == 0
    plan.count(stream, d_st, d_sp, t(seg[100_000:100_001]), d_qs[100_000:100_001], d_qe[100_000:100_001], out[:1], grouped=True)
    torch.cuda.synchronize()
    assert int(out[0]) == int(ref[100_000])
    plan.close()


# ------------------------------------------------------------------- peaks: device merge + linking
def _records_of(pk, c):
    m = pk["ctg"] == c
    return list(zip(pk["first"][m].tolist(), pk["last"][m].tolist(), pk["gc"][m].view(np.uint32).tolist(), pk["signal"][m].tolist()))


def test_wave_peaks_fixture_is_byte_identical_to_I_peaks_tsv(ctx, orc, s288c):
    """wave.rs:170-252 on the device: the merged peak records of chr I format to tests/S288c/I.peaks.tsv byte for byte."""
    ctgs = cut_ctgs(orc, s288c, 500000)
    cat, off = orc.pack_ctgs([c[4] for c in ctgs])
    pk = api.wave_peaks_batch(ctx, cat, off, 100, 10, 100, 3.0, 1.0, 0.2)
    i = [k for k, c in enumerate(ctgs) if c[1] == "I"][0]
    m = pk["ctg"] == i
    text = orc.WAVE_HEADER + api.wave_peaks_text("I", 1, 100, 10, pk["first"][m], pk["last"][m], pk["gc"][m], pk["signal"][m])
    assert text.encode() == bytes(s288c["peaks_tsv"])
    # and the same text as the host-side merge of the per-window arrays, for every ctg
    woff, _, gc, sig = api.wave_batch(ctx, cat, off, 100, 10, 100, 3.0, 1.0)
    for k, c in enumerate(ctgs):
        a, b = int(woff[k]), int(woff[k + 1])
        mk = pk["ctg"] == k
        assert api.wave_peaks_text(c[1], c[2], 100, 10, pk["first"][mk], pk["last"][mk], pk["gc"][mk], pk["signal"][mk]) == \
            api.wave_text(c[1], c[2], 100, 10, gc[a:b], sig[a:b], 0.2)


@pytest.mark.parametrize("size,step,lag,thr", [(100, 10, 100, 3.0), (100, 10, 100, 1.5), (100, 50, 30, 2.0), (100, 100, 10, 2.0),
                                               (16, 1, 50, 2.5), (97, 13, 31, 2.0), (1000, 7, 100, 2.5)])
def test_wave_peaks_match_oracle_merge(ctx, orc, size, step, lag, thr):
    """Device run detection against merge_ints (union-find over intersecting same-sign windows) of the oracle, on ragged
    ctgs with constant-GC stretches (long runs), for the fused and the two-kernel paths, several caps."""
    g = synth.genome(900_000, seed=size + step)
    flat = np.tile(np.frombuffer(b"ACATTTTTAT" * 5 + b"GCGCGCGCGATATATATATATATATATATATATATATATATATATATATA", dtype=np.uint8), 1200)
    need = size + (lag + 2) * step
    seqs = [g[:300_003], np.concatenate([flat[:60_000], g[300_003:360_000], flat[:30_000]]), g[400_000:400_000 + need], g[500_000:]]
    cat, off = orc.pack_ctgs(seqs)
    _, _, rgc, rsig = orc.wave_batch(cat, off, size, step, lag, thr, 1.0, threads=8)
    rwoff = orc.win_offsets(off, size, step)
    for cap in (None, 1):
        pk = api.wave_peaks_batch(ctx, cat, off, size, step, lag, thr, 1.0, 0.2, cap=cap)
        assert np.all(np.diff(pk["ctg"].astype(np.int64)) >= 0)
        total = 0
        for c in range(len(seqs)):
            a, b = int(rwoff[c]), int(rwoff[c + 1])
            ref = [(f, l, int(np.float32(x).view(np.uint32)), s) for f, l, x, s in orc.wave_peak_records(size, step, rgc[a:b], rsig[a:b], 0.2)]
            got = _records_of(pk, c)
            assert got == ref, f"ctg {c}: {len(got)} records vs {len(ref)}; first difference at {next((i for i, (x, y) in enumerate(zip(got, ref)) if x != y), None)}"
            total += len(ref)
        assert total == pk["ctg"].size and total > 0
    # a coverage that rejects some intersecting pairs is not a run merge: the device path refuses it
    if (size - 1) // step >= 1:
        with pytest.raises(api.GamsCudaError) as e:
            api.wave_peaks_batch(ctx, cat, off, size, step, lag, thr, 1.0, coverage=float(size))
        assert e.value.code == -1


def test_peak_link_matches_oracle(ctx, orc, s288c):
    """`gams peak` (peak.rs:66-160): the peaks of I.peaks.tsv on the piece-100000 ctgs, plus synthetic corner cases."""
    ctgs = [c for c in cut_ctgs(orc, s288c, 100000) if c[1] == "I"]
    cat, off = orc.pack_ctgs([c[4] for c in ctgs])
    chr_start = np.array([c[2] for c in ctgs], dtype=np.int32)
    peaks = []
    for line in bytes(s288c["peaks_tsv"]).decode().splitlines()[1:]:
        rg, _, sg = line.split("\t")
        _, _, a, b = orc.range_from_str(rg)
        k = [i for i, c in enumerate(ctgs) if c[2] <= a <= c[3]][0]
        if b <= ctgs[k][3]:
            peaks.append((k, a, b, int(sg)))
    peaks += [(0, 1, 1, 1), (0, 5, 5, -1), (1, ctgs[1][3], ctgs[1][3], 1)]   # single-base peaks at the very ends
    peaks.sort()
    p_ctg, p_s, p_e, p_sig = (np.array(x) for x in zip(*peaks))
    got = api.peak_link_batch(ctx, cat, off, chr_start, p_ctg, p_s, p_e, p_sig)
    for k, c in enumerate(ctgs):
        m = np.flatnonzero(p_ctg == k)
        ref = orc.peak_link(c[4], c[2], [(int(p_s[i]), int(p_e[i]), int(p_sig[i])) for i in m])
        for j, i in enumerate(m):
            for key in ("gc", "left_amplitude", "right_amplitude"):
                assert np.float32(got[key][i]).view(np.uint32) == np.float32(ref[j][key]).view(np.uint32), (k, j, key, got[key][i], ref[j][key])
            for key in ("left_signal", "right_signal", "left_wave_length", "right_wave_length"):
                assert int(got[key][i]) == int(ref[j][key]), (k, j, key)
    with pytest.raises(api.GamsCudaError):
        api.peak_link_batch(ctx, cat, off, chr_start, p_ctg[::-1], p_s[::-1], p_e[::-1], p_sig[::-1])   # not sorted


# ------------------------------------------------------------------- gen: ambiguity scan
def test_gen_ctgs_fixture_matches_ctg_tsv(ctx, orc, s288c):
    """tests/S288c/ctg.tsv: `gams gen --piece 100000` on the fixture → I:1-100000, I:100001-230218, Mito:1-85779."""
    names = s288c["chr_names"].tolist()
    cat, off = orc.pack_ctgs([s288c["seq_" + n] for n in names])
    for piece in (100000, 500000, 30000):
        c, s, e = api.gen_ctgs(ctx, cat, off, piece=piece)
        ref = [(k, a, b) for k, n in enumerate(names) for a, b in orc.gen_ctgs(s288c["seq_" + n], piece=piece)]
        assert list(zip(c.tolist(), s.tolist(), e.tolist())) == ref
    c, s, e = api.gen_ctgs(ctx, cat, off, piece=100000)
    assert [(names[k], a, b) for k, a, b in zip(c.tolist(), s.tolist(), e.tolist())] == [("I", 1, 100000), ("I", 100001, 230218), ("Mito", 1, 85779)]


def test_gen_ctgs_synthetic_gaps(ctx, orc):
    """N runs of every length around --fill, non-ACGT bytes of every value, runs touching chromosome ends and 32-byte
    piece boundaries, chromosomes shorter than one load, an all-N and an empty chromosome."""
    rng = np.random.default_rng(71)
    g = synth.genome(700_000, seed=72).copy()
    pos = np.sort(rng.choice(np.arange(1000, 690_000), 400, replace=False))
    for k, p0 in enumerate(pos):
        ln = [1, 2, 31, 32, 33, 48, 49, 50, 51, 64, 100, 4999, 5000][k % 13]
        g[p0:p0 + ln] = ord("N") if k % 3 else rng.integers(0, 256, ln).astype(np.uint8)
    g[:7] = ord("n")
    g[-40:] = ord("-")
    chrs = [g[:250_000], g[250_000:250_031], g[250_031:250_033], np.zeros(0, np.uint8), np.full(9000, ord("N"), np.uint8),
            g[250_033:600_001], g[600_001:], synth.genome(6000, seed=73)]
    cat, off = orc.pack_ctgs(chrs)
    for piece, fill, mn in ((50_000, 50, 5000), (7_777, 1, 1), (500_000, 64, 0), (10_000, 33, 31)):
        c, s, e = api.gen_ctgs(ctx, cat, off, piece=piece, fill=fill, min_=mn)
        ref = [(k, a, b) for k, q in enumerate(chrs) for a, b in orc.gen_ctgs(q, piece=piece, fill=fill, min_=mn)]
        assert list(zip(c.tolist(), s.tolist(), e.tolist())) == ref, (piece, fill, mn)


# ------------------------------------------------------------------- anno coverage
def test_anno_coverage(ctx, orc, s288c):
    """anno.rs:128-139 against the oracle: the fixture's goldens (tests/cli.rs:477-478), then random ranges that start
    before / end after their ctg, touch run ends exactly, lie in chromosomes without a set, and single positions."""
    import json
    js = json.loads(bytes(s288c["intergenic_json"]).decode())
    names = sorted(js)
    sets = [orc.runlist_to_runs(js[n]) for n in names]
    rs = np.array([a for s in sets for a, _ in s], dtype=np.int32)
    re_ = np.array([b for s in sets for _, b in s], dtype=np.int32)
    roff = np.concatenate(([0], np.cumsum([len(s) for s in sets]))).astype(np.uint64)
    ci = names.index("I")
    mito = names.index("Mito") if "Mito" in names else -1
    got = api.anno_coverage(ctx, rs, re_, roff, [mito, ci, ci], [1, 1, 100001], [85779, 100000, 230218], [1, 1, 100001], [85779, 100000, 230218])
    assert f"{got[0]:.4f}" == "0.0000" and f"{got[2]:.4f}" == "0.1072"
    rng = np.random.default_rng(61)
    n_q = 200_000
    q_chr = rng.integers(-1, len(names) + 1, n_q).astype(np.int32)
    ctg_s = rng.integers(1, 400_000, n_q).astype(np.int32)
    ctg_e = (ctg_s + rng.integers(0, 300_000, n_q)).astype(np.int32)
    ends = np.concatenate([rs, re_])
    q_s = np.where(rng.random(n_q) < 0.5, ends[rng.integers(0, ends.size, n_q)] + rng.integers(-2, 3, n_q), rng.integers(1, 700_000, n_q))
    q_s = np.maximum(q_s, 1).astype(np.int32)
    q_e = (q_s + np.where(rng.random(n_q) < 0.2, 0, rng.integers(0, 20_000, n_q))).astype(np.int32)
    got = api.anno_coverage(ctx, rs, re_, roff, q_chr, ctg_s, ctg_e, q_s, q_e)
    for i in rng.integers(0, n_q, 4000):
        c = int(q_chr[i])
        ref = orc.anno_prop(sets[c] if 0 <= c < len(sets) else None, int(ctg_s[i]), int(ctg_e[i]), int(q_s[i]), int(q_e[i]))
        assert np.float32(got[i]).view(np.uint32) == np.float32(ref).view(np.uint32), (i, got[i], ref)
    assert float(got.max()) <= 1.0 and float(got.min()) >= 0.0


# ------------------------------------------------------------------------------------------------ index build
def _check_idx(ctx, orc, per_seg):
    """per_seg: list of lists of (start, stop) in input order → device index == Lapper::new + bincode of the oracle."""
    iv_off = np.concatenate(([0], np.cumsum([len(v) for v in per_seg]))).astype(np.uint64)
    flat = [iv for v in per_seg for iv in v]
    start = np.array([a for a, _ in flat], dtype=np.uint32)
    stop = np.array([b for _, b in flat], dtype=np.uint32)
    got = api.build_idx_rg(ctx, start, stop, iv_off)
    for s, v in enumerate(per_seg):
        lp = orc.lapper_new([(a, b, "") for a, b in v])
        a, b = int(iv_off[s]), int(iv_off[s + 1])
        _eq(f"seg {s} iv_start", got["iv_start"][a:b], lp["start"])
        _eq(f"seg {s} iv_stop", got["iv_stop"][a:b], lp["stop"])
        _eq(f"seg {s} starts", got["starts"][a:b], lp["starts"])
        _eq(f"seg {s} stops", got["stops"][a:b], lp["stops"])
        assert int(got["max_len"][s]) == lp["max_len"], (s, int(got["max_len"][s]), lp["max_len"])
        blob = got["blob"][int(got["blob_off"][s]):int(got["blob_off"][s + 1])].tobytes()
        assert blob == orc.lapper_to_bincode(lp), f"seg {s}: bincode blob differs"
        st, sp, ml = api.lapper_blob_view(blob)   # and the blob parses back through the library's own reader
        assert np.array_equal(st, lp["starts"]) and np.array_equal(sp, lp["stops"]) and ml == lp["max_len"]
    return got, iv_off


def test_build_idx_rg_fixture(ctx, orc, s288c):
    """`gams rg tests/S288c/SK1.snp.rg` tail: the idx:rg:* blobs (redis.rs:276-305), then cli.rs:449's `I:1000-2000 -> 12`
    counted on the device-built index."""
    ctgs, ids, idx_ctg, idx_rg, starts, stops, iv_off = _fixture_index(orc, s288c)
    rng = np.random.default_rng(3)
    per_seg = []
    for cid in ids:
        lp = idx_rg.get(cid, orc.lapper_new([]))
        order = rng.permutation(lp["start"].size)           # the loader sees them in file order, not sorted
        per_seg.append([(int(lp["start"][i]), int(lp["stop"][i])) for i in order])
    got, off = _check_idx(ctx, orc, per_seg)
    _eq("starts vs fixture index", got["starts"], starts)
    _eq("stops vs fixture index", got["stops"], stops)
    seg = np.array([ids.index("ctg:I:1")], dtype=np.int32)
    assert api.count_overlaps(ctx, got["starts"], got["stops"], off, seg, np.array([1000], np.uint32), np.array([2000], np.uint32)).tolist() == [12]


def test_build_idx_rg_synthetic(ctx, orc):
    """Ties on start (order decided by stop), duplicates, stop < start (max_len clamps to 0), empty and single-interval
    segments, a segment larger than any shared-memory sort would hold, coordinates above 2^31."""
    rng = np.random.default_rng(17)
    def rand_seg(n, lo, hi, max_len):
        s = rng.integers(lo, hi, n)
        return [(int(a), int(a) + int(l)) for a, l in zip(s, rng.integers(1, max_len + 1, n))]
    per_seg = [
        rand_seg(5000, 1, 2000, 3),                      # many equal starts, few distinct stops
        [],
        [(7, 8)],
        [(10, 5), (10, 12), (3, 3), (10, 11), (3, 2)],   # stop < start, ties
        rand_seg(70_000, 1, 1_000_000, 500),
        rand_seg(1000, 2**31 - 100, 2**32 - 600, 500),
        [(5, 6)] * 300,
        [],
    ]
    _check_idx(ctx, orc, per_seg)
    # no blob wanted, and an entirely empty index
    got = api.build_idx_rg(ctx, np.array([4, 2], np.uint32), np.array([9, 3], np.uint32), np.array([0, 2], np.uint64), want_blob=False)
    assert got["blob"] is None and got["iv_start"].tolist() == [2, 4] and got["stops"].tolist() == [3, 9] and got["max_len"].tolist() == [5]
    got = api.build_idx_rg(ctx, np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.array([0, 0, 0], np.uint64))
    assert got["blob"].tobytes() == orc.lapper_to_bincode(orc.lapper_new([])) * 2


# ------------------------------------------------------------------------------------------------ counts along windows
def _ref_count_windows(orc, chr_start, lens, size, step, st, sp, off):
    seg, ws, we, woff = [], [], [], [0]
    for c, (cs, ln) in enumerate(zip(chr_start, lens)):
        n = orc.nwin(ln, size, step)
        k = np.arange(n, dtype=np.int64)
        seg.append(np.full(n, c, np.int32)); ws.append(cs + k * step); we.append(cs + k * step + size - 1)
        woff.append(woff[-1] + n)
    seg, ws, we = np.concatenate(seg), np.concatenate(ws).astype(np.uint32), np.concatenate(we).astype(np.uint32)
    return np.asarray(woff, np.uint64), orc.count_batch(st, sp, off, seg, ws, we, threads=8)


@pytest.mark.parametrize("size,step,max_len", [(100, 10, 1), (100, 10, 300), (100, 100, 1), (50, 7, 40), (1, 1, 1), (1000, 3, 5),
                                                (3000, 1500, 50), (200, 1023, 2), (5, 2, 9)])
def test_count_windows_matches_count_rg(ctx, orc, size, step, max_len):
    """Sliding windows x rg index == count_rg per window (utils.rs:24-36 on window.rs:78-94), incl. ctgs shorter than a
    window, ctgs without intervals, intervals before / after the ctg's windows, many tiles per ctg."""
    lens = [230_218, 85_779, 60, 5_000, 1_000_003, 99, 100]
    chr_start = [1, 100_001, 7, 2_000_000, 1, 1, 1]
    npc = [9000, 0, 5, 700, 60_000, 3, 1]
    rng = np.random.default_rng(size * 31 + step)
    starts, stops, off = [], [], [0]
    for cs, ln, k in zip(chr_start, lens, npc):
        s = rng.integers(max(1, cs - 500), cs + ln + 500, k)       # some intervals outside the ctg's range
        e = s + rng.integers(1, max_len + 1, k)
        starts.append(np.sort(s).astype(np.uint32)); stops.append(np.sort(e).astype(np.uint32)); off.append(off[-1] + k)
    st, sp, off = np.concatenate(starts), np.concatenate(stops), np.asarray(off, np.uint64)
    woff, ref = _ref_count_windows(orc, chr_start, lens, size, step, st, sp, off)
    got = api.count_windows(ctx, chr_start, woff, size, step, st, sp, off)
    _eq(f"count_windows[{size}/{step}]", got, ref)


def test_count_windows_fixture(ctx, orc, s288c):
    """BASELINE configs[1]: S288c sliding 100/10 with SK1.snp.rg counting; the window starting at I:1000 of size 1001
    reproduces cli.rs:449's `I:1000-2000 -> 12`."""
    ctgs, ids, idx_ctg, idx_rg, starts, stops, iv_off = _fixture_index(orc, s288c)
    chr_start = [c[2] for c in ctgs]
    lens = [c[3] - c[2] + 1 for c in ctgs]
    woff, ref = _ref_count_windows(orc, chr_start, lens, 100, 10, starts, stops, iv_off)
    _eq("fixture 100/10", api.count_windows(ctx, chr_start, woff, 100, 10, starts, stops, iv_off), ref)
    woff, ref = _ref_count_windows(orc, chr_start, lens, 1001, 1, starts, stops, iv_off)
    got = api.count_windows(ctx, chr_start, woff, 1001, 1, starts, stops, iv_off)
    _eq("fixture 1001/1", got, ref)
    assert ids[0] == "ctg:I:1" and int(got[999]) == 12      # window 999 of ctg:I:1 is I:1000-2000


# ------------------------------------------------------------------------------------------------ gunzip of seq:* blobs
def _gz(data: bytes, level=1, strategy=0, mem=8):
    import zlib
    c = zlib.compressobj(level, zlib.DEFLATED, 31, mem, strategy)
    return c.compress(data) + c.flush()


def test_gunzip_reference_fixture(ctx, orc, s288c):
    """tests/S288c/genome.fa.gz (the reference's own gzip fixture) and the seq:{ctg} blobs `gams gen` would store."""
    blob = bytes(s288c["genome_fa_gz"])
    out, off, st = api.gunzip_batch(ctx, *api.pack_blobs([blob]))
    ref = orc.decode_gz(blob)
    assert st.tolist() == [0] and out.tobytes() == ref
    assert bytes(s288c["seq_I"][:60]) in ref.replace(b"\n", b"")[:100]
    ctgs = cut_ctgs(orc, s288c, 100000)
    blobs = [orc.encode_gz(c[4].tobytes()) for c in ctgs]                    # encode_gz: Compression::fast()
    cat, goff = api.pack_blobs(blobs)
    assert api.gz_sizes(cat, goff).tolist() == [len(c[4]) for c in ctgs]
    out, off, st = api.gunzip_batch(ctx, cat, goff)
    assert not st.any()
    for i, c in enumerate(ctgs):
        assert out[int(off[i]):int(off[i + 1])].tobytes() == c[4].tobytes() == orc.decode_gz(blobs[i])
    # the decoded ctgs feed the wave path unchanged: byte-identical input, so identical output (spot check through gc counts)
    woff, cnt, gc, sig = api.wave_batch(ctx, out, off, 100, 10, 100, 3.0, 1.0)
    rw, rc, rg, rs = orc.wave_batch(out, off, 100, 10, 100, 3.0, 1.0, threads=4)
    _eq("gc after gunzip", cnt, rc); _eq("signal after gunzip", sig, rs)


def test_gunzip_block_types_and_edges(ctx, orc):
    """Every deflate feature: stored blocks (level 0, incompressible data), fixed Huffman (Z_FIXED, tiny inputs), dynamic
    blocks at levels 1/6/9, literals only (Z_HUFFMAN_ONLY), run-length matches (Z_RLE, N-runs: distance 1 and short periods),
    distances up to 32 KB, 258-byte matches, empty input, gzip headers with FNAME / FEXTRA / FCOMMENT / FHCRC, many blobs."""
    import gzip, io, zlib
    rng = np.random.default_rng(11)
    dna = synth.genome(400_000, seed=3)
    dna[5_000:9_000] = ord("N")
    dna[20_000:20_700] = np.frombuffer(b"ACGTTGA" * 100, dtype=np.uint8)
    dna = dna.tobytes()
    rnd = rng.integers(0, 256, 150_000, dtype=np.uint8).tobytes()
    rep = (rnd[:30_000] + b"x" * 5 + rnd[:30_000] + dna[:40_000] + rnd[1000:31_000])       # matches ~30 000 back
    blobs, plain = [], []
    def add(data, **kw):
        blobs.append(_gz(data, **kw)); plain.append(data)
    for lvl in (0, 1, 6, 9):
        add(dna, level=lvl)
    add(rnd, level=6); add(rep, level=9); add(dna[:70_000], level=6, strategy=zlib.Z_FIXED)
    add(dna[:90_000], level=6, strategy=zlib.Z_HUFFMAN_ONLY); add(dna[:90_000], level=6, strategy=zlib.Z_RLE)
    add(b"", level=1); add(b"A", level=1); add(b"ACGT" * 3, level=1); add(b"N" * 100_000, level=1); add(bytes(range(256)) * 40, level=1, mem=1)
    for k in range(40):
        a = int(rng.integers(0, 300_000)); add(dna[a:a + int(rng.integers(1, 60_000))], level=1)
    bio = io.BytesIO()
    with gzip.GzipFile(filename="seq.fa", mode="wb", fileobj=bio, compresslevel=1, mtime=0) as fh:
        fh.write(dna[:12_345])
    blobs.append(bio.getvalue()); plain.append(dna[:12_345])
    body = _gz(dna[:5_000])[10:]                                                             # hand-made header: all optional fields
    hdr = bytes([0x1F, 0x8B, 8, 2 | 4 | 8 | 16, 0, 0, 0, 0, 0, 255]) + bytes([3, 0]) + b"xyz" + b"name\0" + b"comment\0"
    hdr += (zlib.crc32(hdr) & 0xFFFF).to_bytes(2, "little")                                  # FHCRC
    blobs.append(hdr + body); plain.append(dna[:5_000])
    blobs.append(_gz(dna[:3_000]) + _gz(b"second member is ignored")); plain.append(dna[:3_000])
    assert all(orc.decode_gz(b) == p for b, p in zip(blobs, plain))
    cat, goff = api.pack_blobs(blobs)
    ooff = np.concatenate(([0], np.cumsum([len(p) for p in plain]))).astype(np.uint64)
    out, off, st = api.gunzip_batch(ctx, cat, goff, ooff)
    assert not st.any(), st.tolist()
    for i, p in enumerate(plain):
        got = out[int(off[i]):int(off[i + 1])].tobytes()
        assert got == p, f"blob {i}: first difference at {next((k for k in range(min(len(got), len(p))) if got[k] != p[k]), None)}"


def test_gunzip_corrupt_input_is_reported(ctx, orc):
    """Truncated blobs, flipped bits, a wrong expected size and non-gzip bytes end in a status code, not in a crash, and do
    not disturb the good blobs of the same batch (GzDecoder returns an error for all of these too, CRC aside)."""
    dna = synth.genome(50_000, seed=8).tobytes()
    good = _gz(dna)
    bad_trunc = good[:len(good) // 2]
    flipped = bytearray(good); flipped[len(good) // 3] ^= 0x55
    not_gz = b"PK\x03\x04" + good[4:]
    blobs = [good, bad_trunc, bytes(flipped), not_gz, good]
    cat, goff = api.pack_blobs(blobs)
    ooff = (np.arange(6) * len(dna)).astype(np.uint64)
    out, off, st = api.gunzip_batch(ctx, cat, goff, ooff, raise_on_error=False)
    assert st[0] == 0 and st[4] == 0 and st[1] != 0 and st[3] == 1
    assert out[:len(dna)].tobytes() == dna and out[4 * len(dna):].tobytes() == dna
    with pytest.raises(api.GamsCudaError):
        api.gunzip_batch(ctx, cat, goff, ooff)
    _, _, st = api.gunzip_batch(ctx, *api.pack_blobs([good]), np.array([0, len(dna) - 1], np.uint64), raise_on_error=False)
    assert st.tolist() == [3] or st.tolist() == [5]


def test_multi_device_new_host_paths(orc):
    """Window counts, gunzip and the index build through a context over every visible device: ctgs / blobs / segments are
    split over the devices and the gathered result equals the single-device one (and the oracle on samples)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    rng = np.random.default_rng(91)
    lens = [int(x) for x in rng.integers(150_000, 400_000, 140)]          # 140 ctgs, ~3.8 M windows at 100/10
    chr_start = [1 + 7 * k for k in range(len(lens))]
    st, sp, off = synth.rg_index(lens, [int(x) for x in rng.integers(10_000, 30_000, len(lens))], seed=92, max_len=40)
    st = (st.astype(np.int64) + np.repeat(np.asarray(chr_start) - 1, np.diff(off.astype(np.int64)))).astype(np.uint32)
    sp = (sp.astype(np.int64) + np.repeat(np.asarray(chr_start) - 1, np.diff(off.astype(np.int64)))).astype(np.uint32)
    woff, ref = _ref_count_windows(orc, chr_start, lens, 100, 10, st, sp, off)
    g = synth.genome(sum(lens), seed=93)
    cuts = np.concatenate(([0], np.cumsum(lens)))
    blobs = [orc.encode_gz(g[a:b].tobytes()) for a, b in zip(cuts[:-1], cuts[1:])]
    cat, goff = api.pack_blobs(blobs)
    with api.Context() as mctx, api.Context([0]) as one:
        assert mctx.num_devices >= 2
        _eq("multi-device window counts", api.count_windows(mctx, chr_start, woff, 100, 10, st, sp, off), ref)
        out, ooff, status = api.gunzip_batch(mctx, cat, goff)
        assert not status.any() and out.tobytes() == g.tobytes()
        perm = np.concatenate([a + rng.permutation(b - a) for a, b in zip(off[:-1].astype(np.int64), off[1:].astype(np.int64))])
        m, s1 = api.build_idx_rg(mctx, st[perm], sp[perm], off), api.build_idx_rg(one, st[perm], sp[perm], off)
        for k in ("iv_start", "iv_stop", "starts", "stops", "max_len", "blob"):
            _eq(f"multi-device index {k}", m[k], s1[k])
        _eq("index starts", m["starts"], st)
        # the chained paths (gzip blobs -> windows / peaks): every device inflates and walks its own ctgs
        soff = cuts.astype(np.uint64)
        for infl in (1.0, 0.5):
            a = api.wave_gz_batch(mctx, cat, goff, soff, 100, 10, 100, 3.0, infl)
            b = api.wave_gz_batch(one, cat, goff, soff, 100, 10, 100, 3.0, infl)
            assert not a[4].any() and not b[4].any()
            for name, x, y in zip(("win_off", "gc_count", "gc", "signal"), a[:4], b[:4]):
                _eq(f"multi-device wave_gz {name} (influence {infl})", x.view(np.uint32) if x.dtype == np.float32 else x,
                    y.view(np.uint32) if y.dtype == np.float32 else y)
        pm, stm = api.wave_peaks_gz_batch(mctx, cat, goff, soff, 100, 10, 100, 3.0, 1.0, 0.2)
        p1 = api.wave_peaks_batch(one, g, soff, 100, 10, 100, 3.0, 1.0, 0.2)
        assert not stm.any() and p1["ctg"].size > 0
        for k in ("ctg", "first", "last", "signal"):
            _eq(f"multi-device peaks from gz: {k}", pm[k], p1[k])
        _eq("multi-device peaks from gz: gc", pm["gc"].view(np.uint32), p1["gc"].view(np.uint32))


def test_plans_follow_an_index_rewritten_in_place(ctx, orc):
    """Count plans keep tables derived from the index per device-pointer pair; after the index is rewritten in place,
    invalidate() makes the next call rebuild them (grouped and any-order overlap counts, window counts)."""
    import torch
    lens = [400_000 + 11 * k for k in range(40)]
    st, sp, off = synth.rg_index(lens, [9000] * 40, seed=5, max_len=30)
    seg, qs, qe = synth.queries(lens, 300_000, seed=6, grouped=True)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.as_tensor(a.view(np.int32) if a.dtype != np.int32 else a, device=dev)
    d_st, d_sp, d_seg, d_qs, d_qe = t(st), t(sp), t(seg), t(qs), t(qe)
    out = torch.empty(seg.size, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    woff = np.concatenate(([0], np.cumsum([orc.nwin(n, 100, 10) for n in lens]))).astype(np.uint64)
    wout = torch.empty(int(woff[-1]), dtype=torch.int32, device=dev)
    plan, wplan = api.CountPlan(ctx, off), api.CountWindowsPlan(ctx, np.ones(40, np.int32), woff, 100, 10, off)
    for rnd in range(2):
        if rnd == 1:                                  # a different index under the same pointers
            st2, sp2, _ = synth.rg_index(lens, [9000] * 40, seed=77, max_len=30)
            d_st.copy_(t(st2)); d_sp.copy_(t(sp2))
            st, sp = st2, sp2
            plan.invalidate(); wplan.invalidate()
        ref = orc.count_batch(st, sp, off, seg, qs, qe, threads=8)
        for grouped in (True, False):
            out.zero_()
            plan.count(stream, d_st, d_sp, d_seg, d_qs, d_qe, out, grouped=grouped)
            torch.cuda.synchronize()
            _eq(f"round {rnd} grouped={grouped}", out.cpu().numpy().view(np.uint32), ref)
        wplan.count(stream, d_st, d_sp, wout)
        torch.cuda.synchronize()
        _, wref = _ref_count_windows(orc, [1] * 40, lens, 100, 10, st, sp, off)
        _eq(f"round {rnd} window counts", wout.cpu().numpy().view(np.uint32), wref)
    plan.close(); wplan.close()


# ------------------------------------------------------------------------------------------------ range-file parsing
def _check_parse(ctx, orc, lines, names, text=None, allow_complex=True):
    text = ("\n".join(lines) + "\n").encode() if text is None else text
    got = api.parse_ranges(ctx, text, names)
    assert got["chr"].size == len(lines), (got["chr"].size, len(lines))
    offs = got["line_off"].tolist()
    n_complex = 0
    for i, ln in enumerate(lines):
        raw = text[offs[i]:offs[i + 1]].rstrip(b"\n").rstrip(b"\r").decode(errors="replace")
        assert raw == ln, (i, raw, ln)
        ref = orc.range_from_str(ln)
        c = int(got["chr"][i])
        if c == -3:
            n_complex += 1
            assert allow_complex, f"line {i} {ln!r} was handed back although it has the strict form"
            continue
        if ref is None:
            assert c == -2, (i, ln, c)
        else:
            assert c != -2, (i, ln, ref)
            assert (int(got["start"][i]), int(got["end"][i])) == (ref[2], ref[3]), (i, ln, ref)
            assert c == (names.index(ref[0]) if ref[0] in names else -1), (i, ln, c, ref)
    return n_complex


def test_parse_ranges_fixture_and_canonical_forms(ctx, orc, s288c):
    """tests/S288c/SK1.snp.rg and spo11_hot.rg (what `gams locate -f` / `gams rg` read) and every canonical spelling:
    all parsed on the device, none handed back, values equal to Range::from_str's."""
    names = ["I", "II", "Mito", "chr1", "scaffold_12", "2L", "HLA-A", "a/b"]
    for key in ("snp_rg", "spo11_rg"):
        lines = bytes(s288c[key]).decode().splitlines()
        assert _check_parse(ctx, orc, lines, names, allow_complex=False) == 0
    lines = ["I:1000-2000", "I:1000", "S288c.I(+):1000-2000", "S288c.II(-):5", "Mito:1_85779", "chr1:7-", "HLA-A:5-9", "a/b:3-4",
             "2L:12-12", "unknown:1-2", "I:0-5", "I:5-3", "I:2147483647", "I:1000-2000\textra\tcolumns", "I:10 20", "II:1-2 # c",
             "I(+):3-4", "scaffold_12:100-200", "no colon here", "", "I:00012-00013"]
    assert _check_parse(ctx, orc, lines, names, allow_complex=False) == 0
    # no trailing newline, CRLF, a single line, an empty file
    assert _check_parse(ctx, orc, ["I:1-2", "II:3-4"], names, text=b"I:1-2\nII:3-4", allow_complex=False) == 0
    assert _check_parse(ctx, orc, ["I:1-2", "II:3-4"], names, text=b"I:1-2\r\nII:3-4\r\n", allow_complex=False) == 0
    assert api.parse_ranges(ctx, b"", names)["chr"].size == 0
    assert api.parse_ranges(ctx, b"\n\n", names)["chr"].tolist() == [-2, -2]


def test_parse_ranges_random_and_hostile_lines(ctx, orc):
    """Random well-formed lines (many tiles of text, long and short lines) must all be parsed; hostile ones (second colon,
    greedy strand, dots, leading blanks, non-ASCII, huge numbers) are either parsed exactly as the regex does or handed back."""
    rng = np.random.default_rng(23)
    names = [f"chr{k}" for k in range(1, 25)] + ["X_1", "s/7"]
    lines = []
    for _ in range(60_000):
        c = names[int(rng.integers(0, len(names)))] if rng.random() < 0.9 else "other"
        s = int(rng.integers(1, 250_000_000))
        e = s + int(rng.integers(0, 10_000))
        form = int(rng.integers(0, 6))
        lines.append([f"{c}:{s}-{e}", f"{c}:{s}", f"Hs.{c}(+):{s}-{e}", f"{c}(-):{s}_{e}", f"{c}:{s}-{e}\t{int(rng.integers(0, 99))}",
                      f"{c}:{e + 1}-{s}"][form])
    assert _check_parse(ctx, orc, lines, names, allow_complex=False) == 0
    hostile = ["I(+):1-2(x):3-4", "a.b.c:1-2", " I:1-2", "x y:5", "I:abc", "I::5", "a-b.c:1", "é:1-2", "I:99999999999", "I(+-):1", "I(:1",
               ":5", "I:", "I(+)1-2:3", "..:1", "I:1-2:3-4", "I(+):1-2)", "1:1-1", "_:_1", "I:5-99999999999"]
    names2 = ["I", "c", "b", "y", "1", "_"]
    n_complex = _check_parse(ctx, orc, hostile, names2)
    assert n_complex >= 8


def _locate_count_reference(orc, lines, idx_ctg, idx_rg):
    out = []
    for ln in lines:                                     # locate.rs:87-90, 111-141
        ln = ln.split("\t")[0]                           # `line.split('\t').first()`: only the first field is parsed and echoed
        rg = orc.range_from_str(ln)
        if rg is None:
            continue
        cid = orc.find_one_idx(idx_ctg, rg)
        if not cid:
            continue
        out.append(f"{ln}\t{orc.count_rg(idx_rg, cid, rg)}\n")
    return "".join(out)


def test_locate_count_text_fixture(ctx, orc, s288c):
    """`gams locate --count` text in, text out: the README / cli.rs:426-454 queries (`I:1000-2000\\t12`, `Mito:1000-2000\\t0`, II
    dropped) and the whole SK1.snp.rg file as queries, byte for byte against the oracle's locate loop."""
    ctgs, ids, idx_ctg, idx_rg, starts, stops, iv_off = _fixture_index(orc, s288c)
    chrs = sorted({c[1] for c in ctgs})
    order = [c for ch in chrs for c in sorted((c for c in ctgs if c[1] == ch), key=lambda c: c[2])]
    chr_off = np.concatenate(([0], np.cumsum([sum(1 for c in ctgs if c[1] == ch) for ch in chrs]))).astype(np.uint64)
    cs = np.array([c[2] for c in order], dtype=np.uint32)
    ce = np.array([c[3] + 1 for c in order], dtype=np.uint32)
    # the rg index in the ctg order of the locate arrays
    lps = [idx_rg.get(c[0], orc.lapper_new([])) for c in order]
    st = np.concatenate([lp["starts"] for lp in lps]).astype(np.uint32)
    sp = np.concatenate([lp["stops"] for lp in lps]).astype(np.uint32)
    off = np.concatenate(([0], np.cumsum([lp["starts"].size for lp in lps]))).astype(np.uint64)
    qs = ["I:1000-2000", "II:1001-2000", "Mito:1000-2000", "I:100000-100001", "I:100001", "I:1", "I:230218-230300", "bad line", "I:5-1",
          "S288c.I(+):1000-2000", "I:1000-2000\tcomment", "#range\tgc_content\tsignal", "I:7551-7650\t0.18\t-1", "I(+):11551-11740\t0.66\t1",
          "I:1-100\t0.5\t1", "no colon here\tI:1-100", "Mito:5-50\t\t", "\tI:1-100"]
    text = ("\n".join(qs) + "\n").encode()
    got, nl, npr, nb = api.locate_count_text(ctx, text, chrs, cs, ce, chr_off, st, sp, off)
    ref = _locate_count_reference(orc, qs, idx_ctg, idx_rg)
    assert got.decode() == ref and nl == len(qs) and nb == 0 and npr == ref.count("\n")
    assert got.decode().splitlines()[:2] == ["I:1000-2000\t12", "Mito:1000-2000\t0"]
    assert "I:1-100\t0\n" in got.decode() or "I:1-100\t" in got.decode()     # the first field alone is echoed (locate.rs:87-90)
    assert "0.5" not in got.decode() and "comment" not in got.decode()
    # the input line of every printed line
    got2, _, npr2, _, which = api.locate_count_text(ctx, text, chrs, cs, ce, chr_off, st, sp, off, want_lines=True)
    assert got2 == got and npr2 == npr
    assert [qs[int(k)].split("\t")[0] for k in which] == [ln.split("\t")[0] for ln in got.decode().splitlines()]
    lines = bytes(s288c["snp_rg"]).decode().splitlines()
    got, nl, npr, nb = api.locate_count_text(ctx, bytes(s288c["snp_rg"]), chrs, cs, ce, chr_off, st, sp, off)
    assert got.decode() == _locate_count_reference(orc, lines, idx_ctg, idx_rg) and nb == 0 and nl == len(lines)
    # a hostile line is counted, not printed
    got, nl, npr, nb = api.locate_count_text(ctx, b"I:1000-2000\n I:1-2\nI:5\n", chrs, cs, ce, chr_off, st, sp, off)
    assert nb == 1 and nl == 3 and got.decode() == _locate_count_reference(orc, ["I:1000-2000", "I:5"], idx_ctg, idx_rg)
    assert api.locate_count_text(ctx, b"", chrs, cs, ce, chr_off, st, sp, off)[0] == b""


@pytest.mark.parametrize("infl", [0.0, 0.3, 0.9, 2.0])
def test_thresholding_influence_clean_dirty_transitions(ctx, orc, infl):
    """The influence != 1 walk jumps between influence-1 signals and replays only behind them: series built to sit on its
    edges - isolated spikes exactly lag, lag + 1 and lag + 2 apart, bursts, a step (every window after it signals until
    the filtered history catches up - or never, for influence 0), a spike in the last window, GC-like quantised data."""
    rng = np.random.default_rng(int(infl * 100) + 5)
    for lag in (5, 30, 100):
        series = []
        for gap in (lag - 1, lag, lag + 1, lag + 2, lag + 3, 2 * lag + 5):
            s = rng.normal(0.4, 0.01, 40 * gap + 3 * lag + 7).astype(np.float32)
            s[lag + 3::gap] += 1.0
            series.append(s)
        s = rng.normal(0.4, 0.01, 5000).astype(np.float32); s[2500:] += 0.5; series.append(s)                 # a step
        s = rng.normal(0.4, 0.01, 3000).astype(np.float32); s[1000:1040] += rng.normal(0.5, 0.3, 40).astype(np.float32); s[-1] += 2; series.append(s)
        s = (rng.integers(30, 60, 6000) / np.float32(100)).astype(np.float32); s[rng.integers(0, 6000, 40)] = np.float32(0.9); series.append(s)
        s = np.full(2000, 0.25, np.float32); s[700] = 0.5; s[1500:1510] = 0.1; series.append(s)               # constant: std == 0
        data = np.concatenate(series)
        off = np.concatenate(([0], np.cumsum([x.size for x in series]))).astype(np.uint64)
        for thr in (3.0, 1.0):
            got = api.thresholding_batch(ctx, data, off, lag, thr, infl)
            ref = np.concatenate([orc.thresholding_algo(x, lag, thr, infl) for x in series])
            _eq(f"clean/dirty lag {lag} infl {infl} thr {thr}", got, ref)


def test_gunzip_fuzzed_blobs_are_contained(ctx, orc):
    """Corrupt streams (bit flips anywhere, truncations, a valid header followed by noise) must end in a status - no hang, no
    crash - and must not write outside their own output slot: the intact blobs interleaved with them still come out exactly."""
    rng = np.random.default_rng(99)
    dna = synth.genome(60_000, seed=12).tobytes()
    good = [_gz(dna[a:a + 20_000], level=lvl) for a, lvl in ((0, 1), (20_000, 6), (40_000, 9))]
    blobs, plain = [], []
    for k in range(240):
        g = good[k % 3]
        blobs.append(g); plain.append(orc.decode_gz(g))
        b = bytearray(g)
        kind = k % 4
        if kind == 0:
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(10, len(b)))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:
            b = b[:int(rng.integers(11, len(b) - 1))]
        elif kind == 2:
            b = b[:10] + bytes(rng.integers(0, 256, int(rng.integers(8, 3000)), dtype=np.uint8))
        else:
            i = int(rng.integers(10, len(b) - 40)); b[i:i + 32] = bytes(rng.integers(0, 256, 32, dtype=np.uint8))
        blobs.append(bytes(b)); plain.append(None)
    cat, goff = api.pack_blobs(blobs)
    sizes = [20_000] * len(blobs)
    ooff = np.concatenate(([0], np.cumsum(sizes))).astype(np.uint64)
    out, off, st = api.gunzip_batch(ctx, cat, goff, ooff, raise_on_error=False)
    n_bad = 0
    for i, p in enumerate(plain):
        got = out[int(off[i]):int(off[i + 1])].tobytes()
        if p is not None:
            assert st[i] == 0 and got == p, f"intact blob {i} damaged (status {st[i]})"
        elif st[i] == 0:
            # a flip the format does not notice (e.g. inside a stored block, or one the CRC alone would catch): the size still matched
            assert len(got) == 20_000
        else:
            n_bad += 1
    assert n_bad > 150


def test_parse_ranges_fuzzed_text(ctx, orc):
    """Random text over the characters that matter to the grammar: whatever the device does not hand back (-3) must agree with the
    reference expression, line by line; nothing crashes on binary junk either."""
    rng = np.random.default_rng(31)
    alphabet = list("I:-_().+ 0123456789chrX/ab\t") + ["II", "10", "000", "chr1", "(+)", "(-)", ":", ":", "-"]
    lines = ["".join(rng.choice(alphabet, int(rng.integers(0, 14)))) for _ in range(30_000)]
    lines = [ln.replace("\n", "").replace("\r", "") for ln in lines]
    names = ["I", "II", "chr1", "X", "a", "b", "chrX", "c", "h", "r"]
    n_complex = _check_parse(ctx, orc, lines, names)
    assert n_complex < len(lines)
    junk = bytes(rng.integers(0, 256, 200_000, dtype=np.uint8))
    got = api.parse_ranges(ctx, junk, names)
    assert got["chr"].size == junk.count(b"\n") + (0 if junk.endswith(b"\n") else 1)
    assert set(np.unique(got["chr"]).tolist()) <= set(range(-3, len(names)))
