"""Pin the CPU oracle against every golden vector / KAT the reference holds for the hot path
(SURVEY §8c).  CPU only."""
import numpy as np
import pytest

from conftest import cut_ctgs

# src/libs/stat.rs:58-81 — thresholding_sample KAT (lag 30, threshold 5, influence 0)
KAT_INPUT = [
    1.0, 1.0, 1.1, 1.0, 0.9, 1.0, 1.0, 1.1, 1.0, 0.9, 1.0, 1.1, 1.0, 1.0, 0.9, 1.0, 1.0, 1.1, 1.0, 1.0,
    1.0, 1.0, 1.1, 0.9, 1.0, 1.1, 1.0, 1.0, 0.9, 1.0, 1.1, 1.0, 1.0, 1.1, 1.0, 0.8, 0.9, 1.0, 1.2, 0.9,
    1.0, 1.0, 1.1, 1.2, 1.0, 1.5, 1.0, 3.0, 2.0, 5.0, 3.0, 2.0, 1.0, 1.0, 1.0, 0.9, 1.0, 1.0, 3.0, 2.6,
    4.0, 3.0, 3.2, 2.0, 1.0, 1.0, 0.8, 4.0, 4.0, 2.0, 2.5, 1.0, 1.0, 1.0,
]
KAT_SIGNALS = [0] * 45 + [1, 0, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0]


def test_thresholding_kat(orc):
    sig = orc.thresholding_algo(KAT_INPUT, 30, 5.0, 0.0)
    assert sig.tolist() == KAT_SIGNALS


def test_thresholding_panics_when_short(orc):
    with pytest.raises(IndexError):
        orc.thresholding_algo([0.1] * 10, 30, 3.0, 1.0)


# src/libs/window.rs:126-147
@pytest.mark.parametrize("parent,child,resize,exp", [
    ((1, 500), (201, 201), 100, (152, 250)),
    ((1, 500), (200, 200), 100, (151, 249)),
    ((1, 500), (200, 201), 100, (151, 250)),
    ((1, 500), (199, 201), 100, (150, 249)),
    ((1, 500), (199, 202), 100, (151, 250)),
    ((1, 500), (100, 301), 100, (151, 250)),
    ((1, 500), (1, 1), 100, (1, 50)),
    ((1, 500), (500, 500), 100, (451, 500)),
    ((1001, 1500), (1200, 1201), 100, (1151, 1250)),
])
def test_center_resize_kat(orc, parent, child, resize, exp):
    assert orc.center_resize(*parent, *child, resize) == exp


# src/libs/window.rs:58-76
@pytest.mark.parametrize("parent,start,end,exp", [
    ((1, 9999), 500, 500, ((451, 549), "M", 0, 3)),
    ((1, 9999), 500, 800, ((600, 699), "M", 0, 3)),
    ((1, 9999), 101, 101, ((52, 150), "M", 0, 2)),
    ((10001, 19999), 10101, 10101, ((10052, 10150), "M", 0, 2)),
])
def test_center_sw_kat(orc, parent, start, end, exp):
    w = orc.center_sw(*parent, start, end, 100, 1)
    assert (w[0][0], w[0][1]) == exp[0] and w[0][2] == exp[1] and w[0][3] == exp[2]
    assert len(w) == exp[3]


# tests/cli.rs:483-495 and utils.rs:131-134
def test_gc_stat_and_round_kat(orc):
    m, s, c = orc.gc_stat([0.5, 0.5])
    assert (m, s, c) == (np.float32(0.5), np.float32(0.0), np.float32(0.0))
    m, s, c = orc.gc_stat([0.4, 0.5, 0.5, 0.6])
    assert m == np.float32(0.5) and s == pytest.approx(0.0816, rel=1e-6) and c == pytest.approx(0.1633, rel=1e-6)
    assert orc.round_(4.364, 2) == np.float32(4.36)
    assert orc.round_(4.368, 2) == np.float32(4.37)


def test_gen_partition(orc, s288c):
    """tests/S288c/ctg.tsv (piece 100000) and tests/cli.rs:112-146 (3 ctgs)."""
    ctgs = cut_ctgs(orc, s288c, 100000)
    got = {(c[0], c[2], c[3]) for c in ctgs}
    assert got == {("ctg:I:1", 1, 100000), ("ctg:I:2", 100001, 230218), ("ctg:Mito:1", 1, 85779)}
    rows = bytes(s288c["ctg_tsv"]).decode().splitlines()[1:]
    assert {(r.split("\t")[0], int(r.split("\t")[3]), int(r.split("\t")[4])) for r in rows} == got
    assert [(c[2], c[3]) for c in cut_ctgs(orc, s288c, 500000)] == [(1, 230218), (1, 85779)]


def _wave_stdout(orc, ctgs, **kw):
    cat, off = orc.pack_ctgs([c[4] for c in ctgs])
    woff, cnt, gc, sig = orc.wave_batch(cat, off, 100, 10, 100, 3.0, 1.0)
    out = [orc.WAVE_HEADER]
    for i, c in enumerate(ctgs):
        a, b = int(woff[i]), int(woff[i + 1])
        out.append(orc.wave_text(c[1], c[2], 100, 10, gc[a:b], sig[a:b], coverage=0.2, **kw))
    return "".join(out), (woff, cnt, gc, sig)


def test_wave_reproduces_I_peaks_tsv_byte_for_byte(orc, s288c):
    """Golden #1: README.md:155-170 / tests/S288c/I.peaks.tsv (gen --piece 500000, wave --ctg 'ctg:I:*')."""
    ctgs = [c for c in cut_ctgs(orc, s288c, 500000) if c[1] == "I"]
    text, (woff, cnt, gc, sig) = _wave_stdout(orc, ctgs)
    assert text.encode() == bytes(s288c["peaks_tsv"])
    assert int(woff[-1]) == 23012
    assert int((sig != 0).sum()) == 382
    lines = text.splitlines()[1:]
    assert sum(l.endswith("\t1") for l in lines) == 43 and sum(l.endswith("\t-1") for l in lines) == 73


def test_wave_uppercase_only_would_fail(orc, s288c):
    """Quirk Q1: case-insensitive c/g is pinned — counting uppercase only changes the peaks."""
    seq = s288c["seq_I"]
    up = np.where(np.isin(seq, np.frombuffer(b"cg", dtype=np.uint8)), ord("n"), seq).astype(np.uint8)
    cat, off = orc.pack_ctgs([up])
    woff, cnt, gc, sig = orc.wave_batch(cat, off, 100, 10, 100, 3.0, 1.0)
    text = orc.WAVE_HEADER + orc.wave_text("I", 1, 100, 10, gc, sig)
    assert text.encode() != bytes(s288c["peaks_tsv"])


def test_wave_cli_piece_100000(orc, s288c):
    """Golden #2: tests/cli.rs:332-364 — 116 stdout lines, two pinned substrings."""
    ctgs = [c for c in cut_ctgs(orc, s288c, 100000) if c[1] == "I"]
    text, _ = _wave_stdout(orc, ctgs)
    assert len(text.splitlines()) == 116
    assert "I:7551-7650\t" in text and "I(+):11551-11740\t" in text


def test_gc_counts_vs_numpy_cumsum(orc, s288c):
    """Independent restatement of a3: cumulative-sum GC counts must equal the oracle's byte scans."""
    for size, step in [(100, 10), (100, 20), (97, 13), (1, 1), (500, 500), (33, 64)]:
        for name in ("I", "Mito"):
            seq = s288c["seq_" + name]
            cat, off = orc.pack_ctgs([seq])
            woff, cnt, gc, _ = orc.wave_batch(cat, off, size, step, 100, 3.0, 1.0, want_signal=False)
            isgc = np.isin(seq, np.frombuffer(b"cgCG", dtype=np.uint8))
            cs = np.concatenate(([0], np.cumsum(isgc)))
            k = np.arange(int(woff[-1]))
            ref = (cs[k * step + size] - cs[k * step]).astype(np.uint32)
            assert np.array_equal(cnt, ref)
            assert np.array_equal(gc, ref.astype(np.float32) / np.float32(size))


def test_locate_count_golden(orc, s288c):
    """Golden #5: tests/cli.rs:426-454 — I:1000-2000 → 12, Mito:1000-2000 → 0, II dropped."""
    ctgs = cut_ctgs(orc, s288c, 100000)
    idx_ctg = orc.build_idx_ctg([c[:4] for c in ctgs])
    ranges_of = orc.read_range(bytes(s288c["snp_rg"]).decode().splitlines(), idx_ctg)
    idx_rg = orc.build_idx_rg(ranges_of)
    for c in ctgs:  # build_idx_rg covers every ctg, also those without rgs (redis.rs:279-303)
        idx_rg.setdefault(c[0], orc.lapper_new([]))
    out = []
    for q in ["I:1000-2000", "II:1001-2000", "Mito:1000-2000"]:
        rg = orc.range_from_str(q)
        cid = orc.find_one_idx(idx_ctg, rg)
        if not cid:
            continue
        out.append(f"{q}\t{orc.count_rg(idx_rg, cid, rg)}")
    assert out == ["I:1000-2000\t12", "Mito:1000-2000\t0"]


def test_rg_first_range_dropped(orc, s288c):
    """Golden #6 / quirk Q7: tests/cli.rs:235-253 — 79 lines, 71 on chr I, 69 survive (2 ctgs)."""
    ctgs = cut_ctgs(orc, s288c, 100000)
    idx_ctg = orc.build_idx_ctg([c[:4] for c in ctgs])
    lines = bytes(s288c["spo11_rg"]).decode().splitlines()
    located = [orc.find_one_idx(idx_ctg, orc.range_from_str(l)) for l in lines]
    assert len(lines) == 79 and sum(bool(x) for x in located) == 71
    ranges_of = orc.read_range(lines, idx_ctg)
    assert sum(len(v) for v in ranges_of.values()) == 69


def test_sw_feature_rows(orc, s288c):
    """tests/cli.rs:306-330: sw over spo11 features emits > 2000 rows; ids like sw:feature:ctg:I:2:32:1."""
    ctgs = cut_ctgs(orc, s288c, 100000)
    idx_ctg = orc.build_idx_ctg([c[:4] for c in ctgs])
    feats = orc.read_range(bytes(s288c["spo11_rg"]).decode().splitlines(), idx_ctg)
    rows, ids = 0, []
    for c in ctgs:
        for n, rg in enumerate(feats.get(c[0], []), 1):
            r = orc.sw_feature(c[4], c[2], c[3], rg[2], rg[3])
            assert r["type"][0] == 0 and len(r["gc"]) <= 41
            ids += [f"sw:feature:{c[0]}:{n}:{k + 1}" for k in range(len(r["gc"]))]
            rows += len(r["gc"])
    assert rows > 2000 and "sw:feature:ctg:I:2:32:1" in ids


def test_lapper_bincode_layout(orc):
    """a14: 24*N + 30 bytes for val == ""; starts[] at 16N+16, stops[] at 20N+24."""
    lp = orc.lapper_new([(5, 6, ""), (1, 2, ""), (3, 9, "")])
    blob = orc.lapper_to_bincode(lp)
    n = 3
    assert len(blob) == 24 * n + 30
    assert np.frombuffer(blob, dtype="<u4", count=n, offset=16 * n + 16).tolist() == [1, 3, 5]
    assert np.frombuffer(blob, dtype="<u4", count=n, offset=20 * n + 24).tolist() == [2, 6, 9]
    assert lp["max_len"] == 6


def test_fmt_f32(orc):
    assert [orc.fmt_f32(x) for x in (0.18, 0.5, 1.0, 0.0, 0.4321, float("nan"))] == ["0.18", "0.5", "1", "0", "0.4321", "NaN"]


def test_wave_peak_records_reproduce_I_peaks_tsv(orc, s288c):
    """The record form of wave's output (what the device merge returns) formats back to tests/S288c/I.peaks.tsv."""
    seq = np.asarray(s288c["seq_I"])
    cat, off = orc.pack_ctgs([seq])
    _, _, gc, sig = orc.wave_batch(cat, off, 100, 10, 100, 3.0, 1.0)
    recs = orc.wave_peak_records(100, 10, gc, sig, 0.2)
    lines = []
    for first, last, g, s in recs:
        a, b = 1 + first * 10, 1 + last * 10 + 99
        lines.append(f"I{'(+)' if last != first else ''}:{orc.runlist(a, b)}\t{orc.fmt_f32(g)}\t{s}\n")
    assert (orc.WAVE_HEADER + "".join(lines)).encode() == bytes(s288c["peaks_tsv"])
    assert len(recs) == 116


def test_peak_link_oracle_small_case(orc):
    """peak.rs:108-157 on a hand-checked case: first / last peak link to the ctg's ends with amplitude 0."""
    seq = np.frombuffer(b"GGGGGGGGGGAAAAAAAAAAGCGCGCGCGCATATATATATCCCCCAAAAA", dtype=np.uint8)  # chr 101..150
    out = orc.peak_link(seq, 101, [(101, 110, 1), (121, 130, -1), (141, 150, 1)])
    assert [float(p["gc"]) for p in out] == [1.0, 1.0, 0.5]
    assert [p["left_wave_length"] for p in out] == [1, 12, 12]
    assert [p["right_wave_length"] for p in out] == [12, 12, 1]
    assert [float(p["left_amplitude"]) for p in out] == [0.0, 0.0, 0.5]
    assert [float(p["right_amplitude"]) for p in out] == [0.0, 0.5, 0.0]
    assert [p["left_signal"] for p in out] == [1, 1, -1] and [p["right_signal"] for p in out] == [-1, 1, 1]


def test_anno_prop_golden(orc, s288c):
    """tests/cli.rs:457-481: `gams anno intergenic.json ctg.range.tsv -H` prints "85779\t0.0000" and "130218\t0.1072"."""
    import json
    sets = {k: orc.runlist_to_runs(v) for k, v in json.loads(bytes(s288c["intergenic_json"]).decode()).items()}
    assert f"{orc.anno_prop(sets.get('Mito'), 1, 85779, 1, 85779):.4f}" == "0.0000"
    assert f"{orc.anno_prop(sets.get('I'), 100001, 230218, 100001, 230218):.4f}" == "0.1072"
