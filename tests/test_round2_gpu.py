"""Round-2 parity tests of the CUDA path (through the C ABI) against the CPU oracle.  Needs a B200: -m gpu."""
import os

import numpy as np
import pytest

from gams_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context([0])
    yield c
    c.close()


def _eq(name, got, ref):
    if not np.array_equal(got, ref):
        bad = np.flatnonzero(got != ref)
        raise AssertionError(f"{name}: {bad.size} of {ref.size} differ; first at {bad[:8].tolist()} got {got[bad[:8]].tolist()} ref {ref[bad[:8]].tolist()}")


class _Env:
    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _count_dev(ctx, st, sp, off, seg, qs, qe, grouped=False):
    import torch
    dev = torch.device("cuda", 0)
    t = lambda a: torch.as_tensor(a.view(np.int32) if a.dtype != np.int32 else a, device=dev)
    plan = api.CountPlan(ctx, off)
    out = torch.full((seg.size,), -1, dtype=torch.int32, device=dev)
    plan.count(torch.cuda.current_stream().cuda_stream, t(st), t(sp), t(seg), t(qs), t(qe), out, grouped=grouped)
    torch.cuda.synchronize()
    n = plan.launches
    plan.close()
    return out.cpu().numpy().view(np.uint32), n


# --------------------------------------------------------------- Kernel C, arbitrary order: the partitioned path (overlap_part.cu)
def _adversarial_index(rng):
    segs = [
        (np.full(5000, 777, np.uint32), np.full(5000, 778, np.uint32)),                                    # all identical
        (np.concatenate([rng.integers(1, 50, 3000), rng.integers(4_000_000_000, 4_000_000_100, 3000)]).astype(np.uint32), None),
        (rng.integers(1, 4_294_967_000, 20000).astype(np.uint32), None),                                   # huge coordinates
        (np.array([123456], np.uint32), np.array([123999], np.uint32)),                                    # single interval
        (np.zeros(0, np.uint32), np.zeros(0, np.uint32)),                                                  # empty
        (np.repeat(rng.integers(1, 10_000, 50), 400).astype(np.uint32), None),                             # crowded buckets
        (np.arange(1, 30001, dtype=np.uint32) * 3, None),                                                  # n > ST_MAX_N
        (rng.integers(1, 500_000, 16000).astype(np.uint32), None),                                         # the C5 shape, ranges
        (rng.integers(1, 900_000, 24000).astype(np.uint32), None),                                         # l = 7
        (rng.integers(1, 2_000_000, 24000).astype(np.uint32), None),                                       # l = 8: plain
        (rng.integers(1, 500_000, 16000).astype(np.uint32), "points"),                                     # SNP points: E8 form
        (np.repeat(rng.integers(1, 500_000, 2000), 9).astype(np.uint32), "points"),                        # E8 with overflowing buckets
        (rng.integers(1, 5_000_000, 3000).astype(np.uint32), "points"),                                    # points too sparse for E8
    ]
    st, sp, off = [], [], [0]
    for a, e in segs:
        a = np.sort(a)
        if isinstance(e, str):
            e = a + np.uint32(1)
        e = np.sort(e if e is not None else (a + rng.integers(1, 200, a.size).astype(np.uint32)))
        st.append(a), sp.append(e), off.append(off[-1] + a.size)
    return np.concatenate(st), np.concatenate(sp), np.asarray(off, dtype=np.uint64), len(segs)


@pytest.mark.parametrize("mb", [12, 1])
def test_count_overlaps_partitioned_adversarial(ctx, orc, mb):
    """Queries in arbitrary order through the partitioned path (forced on by GAMS_CUDA_PART_MIN_Q) on every shape the
    compact index special-cases, with segment ids outside the index, a last span that is not full, one bucket for the
    whole index and one bucket per segment."""
    rng = np.random.default_rng(191)
    st, sp, off, n_seg = _adversarial_index(rng)
    n_q = 300_007
    seg = rng.integers(-2, n_seg + 2, n_q).astype(np.int32)
    pool = np.concatenate([st, sp, np.array([0, 1, 2, 4_294_967_294, 4_294_967_295], np.uint32)])
    qs = pool[rng.integers(0, pool.size, n_q)].astype(np.int64) + rng.integers(-2, 3, n_q)
    ln = np.where(rng.random(n_q) < 0.3, 0, rng.integers(0, 5000, n_q))
    qs = np.clip(qs, 0, 4_294_967_295)
    qe = np.clip(qs + ln, 0, 4_294_967_295)
    qs, qe = qs.astype(np.uint32), qe.astype(np.uint32)
    ref = orc.count_batch(st, sp, off, seg, qs, qe, threads=8)
    with _Env(GAMS_CUDA_PART_MIN_Q=1, GAMS_CUDA_PART_MB=mb):
        got, launches = _count_dev(ctx, st, sp, off, seg, qs, qe)
        for n_small in (1, 3, 8191, 8192, 8193):   # spans: one partial, exactly one, one and a bit
            g2, _ = _count_dev(ctx, st, sp, off, seg[:n_small], qs[:n_small], qe[:n_small])
            _eq(f"partitioned counts, {n_small} queries", g2, ref[:n_small])
    assert launches >= 5, "the partitioned path did not run"
    _eq(f"partitioned counts ({mb} MB buckets)", got, ref)
    with _Env(GAMS_CUDA_PART_MIN_Q=1 << 30):
        got, launches = _count_dev(ctx, st, sp, off, seg, qs, qe)
    assert launches <= 3
    _eq("ranked counts", got, ref)


def test_count_overlaps_partitioned_many_segments(ctx, orc):
    """5000 ctgs of SNP points and of range intervals, 3 M queries in arbitrary order: default thresholds, several buckets."""
    lens = [200_000 + 13 * k for k in range(5000)]
    for max_len in (1, 300):
        st, sp, off = synth.rg_index(lens, [2000 + (k % 5) * 500 for k in range(5000)], seed=7, max_len=max_len)
        seg, qs, qe = synth.queries(lens, 3_000_000, seed=8, grouped=False)
        seg[::1001] = -1
        seg[5::1001] = 5000
        ref = orc.count_batch(st, sp, off, seg, qs, qe, threads=8)
        with _Env(GAMS_CUDA_PART_MB=1):
            got, launches = _count_dev(ctx, st, sp, off, seg, qs, qe)
        assert launches >= 5
        _eq(f"partitioned counts, max_len {max_len}", got, ref)


def test_count_overlaps_host_api_uses_partitioned_path(ctx, orc):
    """gams_cuda_count_overlaps (host pointers) slices the queries by 8 M: each slice goes through the partitioned path."""
    lens = [500_000] * 64
    st, sp, off = synth.rg_index(lens, [16000] * 64, seed=17)
    seg, qs, qe = synth.queries(lens, 9_000_000, seed=18, grouped=False)
    got = api.count_overlaps(ctx, st, sp, off, seg, qs, qe)
    _eq("host api counts", got, orc.count_batch(st, sp, off, seg, qs, qe, threads=8))


# --------------------------------------------------------------- thresholding_algo outside the certificate's error model
@pytest.mark.parametrize("infl", [1.0, 0.5])
def test_thresholding_extreme_magnitudes(ctx, orc, infl):
    """stat.rs:1-56 on series where the reference's own f32 arithmetic overflows or underflows: (x - mean)^2 = inf gives
    sd = inf and no signal at all (|x| >~ 1.8e19), NaN / Inf poison the statistics of every window that sees them, squares
    of denormal deviations flush to zero.  The f64 certificate must not decide any of those windows: they are replayed."""
    rng = np.random.default_rng(301)
    n, lag = 6000, 50
    base = rng.normal(0, 1, n).astype(np.float32)
    base[rng.integers(0, n, 40)] += 9.0
    series = {
        "N(0,1e25)": base * np.float32(1e25),
        "N(0,1e19)": base * np.float32(1e19),                      # (x - m)^2 on the edge of f32 overflow
        "N(0,3e18)": base * np.float32(3e18),
        "N(0,1e-30)": base * np.float32(1e-30),
        "denormal": base * np.float32(1e-41),
        "N(0,1e-12)": base * np.float32(1e-12),
    }
    mixed = base.copy()
    mixed[1000:1200] *= np.float32(1e22)                           # one tile has huge values, its neighbours do not
    mixed[3000] = np.inf
    mixed[3500] = -np.inf
    mixed[4200] = np.nan
    mixed[5000:5100] = np.float32(1e-44)
    series["mixed Inf/NaN/huge/denormal"] = mixed
    allnan = base.copy()
    allnan[::7] = np.nan
    series["NaN every 7th"] = allnan
    for thr in (3.0, 1e-6, 1e12):
        for name, s in series.items():
            with np.errstate(all="ignore"):
                ref = orc.thresholding_algo(s, lag, thr, infl)
            got = api.thresholding_algo(ctx, s, lag, thr, infl)
            _eq(f"{name}, threshold {thr}, influence {infl}", got, ref)
    # the reference really gives no signal at all once the squares overflow
    with np.errstate(all="ignore"):
        assert not orc.thresholding_algo(series["N(0,1e25)"], lag, 3.0, 1.0).any()


# --------------------------------------------------------------- sw --action count: rg_count per sw row
def _sw_count_reference(orc, ctgs, idx_rg, f_ctg, f_s, f_e, size=100, max_=20, resize=500):
    """sw.rs:141-191 per feature with the rg_count column filled by gams::count_rg (utils.rs:24-36) as locate.rs:136 calls it."""
    rows = []
    for f in range(len(f_ctg)):
        c = ctgs[f_ctg[f]]
        ref = orc.sw_feature(c[4], c[2], c[3], f_s[f], f_e[f], size, max_, resize)
        cnt = [orc.count_rg(idx_rg, c[0], (c[1], "", int(s), int(e))) for s, e in zip(ref["start"], ref["end"])]
        rows.append((ref, np.asarray(cnt, dtype=np.uint32)))
    return rows


@pytest.mark.parametrize("order", ["by_ctg", "shuffled"])
def test_sw_count_batch_fixture(ctx, orc, s288c, order):
    """`gams sw --action gc --action count` for the spo11 features against the SK1 SNP index of tests/S288c: every gc column as
    gams_cuda_sw_batch gives it, rg_count = count_rg of the row's window (oracle: sw_feature composed with count_rg)."""
    from conftest import cut_ctgs
    ctgs = cut_ctgs(orc, s288c, 100000)
    idx_ctg = orc.build_idx_ctg([c[:4] for c in ctgs])
    idx_rg = orc.build_idx_rg(orc.read_range(bytes(s288c["snp_rg"]).decode().splitlines(), idx_ctg))
    lps = [idx_rg.get(c[0], orc.lapper_new([])) for c in ctgs]
    starts = np.concatenate([lp["starts"] for lp in lps]).astype(np.uint32)
    stops = np.concatenate([lp["stops"] for lp in lps]).astype(np.uint32)
    iv_off = np.concatenate(([0], np.cumsum([lp["starts"].size for lp in lps]))).astype(np.uint64)
    feats = orc.read_range(bytes(s288c["spo11_rg"]).decode().splitlines(), idx_ctg)
    cat, off = orc.pack_ctgs([c[4] for c in ctgs])
    f_ctg, f_s, f_e = [], [], []
    for k, c in enumerate(ctgs):
        for rg in feats.get(c[0], []):
            f_ctg.append(k), f_s.append(rg[2]), f_e.append(rg[3])
        for s, e in ((c[2], c[2]), (c[3], c[3]), (c[2], c[2] + 49), (c[3] - 20, c[3]), (c[2] + 1000, c[2] + 1001), (c[2], c[3])):
            f_ctg.append(k), f_s.append(s), f_e.append(e)
    if order == "shuffled":
        perm = np.random.default_rng(5).permutation(len(f_ctg))
        f_ctg, f_s, f_e = [f_ctg[i] for i in perm], [f_s[i] for i in perm], [f_e[i] for i in perm]
    r = api.sw_batch(ctx, cat, off, [c[2] for c in ctgs], f_ctg, f_s, f_e, rg_index=(starts, stops, iv_off))
    plain = api.sw_batch(ctx, cat, off, [c[2] for c in ctgs], f_ctg, f_s, f_e)
    for key in ("start", "end", "type", "dist"):
        _eq(f"sw_count {key} vs sw_batch", r[key], plain[key])
    for key in ("gc", "mean", "sd", "cv"):
        _eq(f"sw_count {key} vs sw_batch", r[key].view(np.uint32), plain[key].view(np.uint32))
    assert int(r["row_off"][-1]) > 2000 and int(r["rg_count"].sum()) > 0
    for f, (ref, cnt) in enumerate(_sw_count_reference(orc, ctgs, idx_rg, f_ctg, f_s, f_e)):
        a, b = int(r["row_off"][f]), int(r["row_off"][f + 1])
        assert b - a == len(ref["gc"])
        _eq(f"rg_count feature {f}", r["rg_count"][a:b], cnt)
        _eq(f"gc feature {f}", r["gc"][a:b].view(np.uint32), ref["gc"].view(np.uint32))


def test_sw_count_batch_synthetic_c3_shape(ctx, orc):
    """The configs[2] shape at small scale: point features on ragged ctgs, SNP and range rgs, ctgs without any rg."""
    rng = np.random.default_rng(41)
    g = synth.genome(1_500_000, seed=42)
    cuts = [0, 300_000, 300_050, 900_001, 1_500_000]
    seqs = [g[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    chr_start = [1, 300_001, 1, 600_000]
    lens = [s.size for s in seqs]
    cat, off = orc.pack_ctgs(seqs)
    for max_len in (1, 200):
        st, sp, ioff = synth.rg_index(lens, [9000, 0, 20000, 15000], seed=43, max_len=max_len)
        # the index holds chr coordinates: shift every segment to its ctg's chr_start
        for c in range(4):
            a, b = int(ioff[c]), int(ioff[c + 1])
            st[a:b] += np.uint32(chr_start[c] - 1)
            sp[a:b] += np.uint32(chr_start[c] - 1)
        n_feat = 3000
        f_ctg = np.sort(rng.integers(0, 4, n_feat)).astype(np.uint32)
        f_s = np.array([chr_start[c] + int(rng.integers(0, lens[c])) for c in f_ctg], dtype=np.int32)
        f_e = np.minimum(f_s + rng.integers(0, 3, n_feat), [chr_start[c] + lens[c] - 1 for c in f_ctg]).astype(np.int32)
        r = api.sw_batch(ctx, cat, off, chr_start, f_ctg, f_s, f_e, rg_index=(st, sp, ioff))
        seg = np.repeat(f_ctg.astype(np.int32), np.diff(r["row_off"]).astype(np.int64))
        ref = orc.count_batch(st, sp, ioff, seg, r["start"].astype(np.uint32), r["end"].astype(np.uint32), threads=8)
        _eq(f"rg_count, max_len {max_len}", r["rg_count"], ref)
        assert not r["rg_count"][seg == 1].any()


# --------------------------------------------------------------- influence != 1 on GC fractions: one thread per segment, O(1) certified steps
@pytest.mark.parametrize("zseg", [0, 137, 700])
@pytest.mark.parametrize("infl", [0.0, 9.5367431640625e-07, 0.001, 0.5, 0.999])
def test_wave_influence_thread_walk(ctx, orc, zseg, infl):
    """zscore_tseg_walk_kernel / zscore_tseg_fix_kernel (stat.rs:40-48 with influence in [0, 1], data = GC fractions) against the
    oracle and against the warp-per-segment walk: segment boundaries every 137 / 700 / 4096 windows, thresholds that make the
    process dirty almost everywhere (1.0: replays run through whole segments, the series is walked again by the sequential
    kernel) or rarely (4.0), constant-GC runs (every certificate inconclusive), a GC shift every 3 kb, lags 2 .. 200."""
    rng = np.random.default_rng(17)
    g = synth.genome(900_000, seed=55)
    shifty = g[:300_000].copy()
    for a in range(0, 300_000, 6000):                          # alternate GC-rich / AT-rich blocks: many long runs of signals
        blk = shifty[a:a + 3000]
        blk[rng.random(blk.size) < 0.5] = ord("G")
    flat = np.tile(np.frombuffer(b"ACATTTTTAT" * 5 + b"GCGCGCGCGATATATATATATATATATATATATATATATATATATATATA", dtype=np.uint8), 1500)
    seqs = [g[300_000:700_000], shifty, np.concatenate([flat, g[700_000:780_000], flat[:40_000]]), g[780_000:782_511]]
    cat, off = orc.pack_ctgs(seqs)
    env = {"GAMS_CUDA_ZSEG": zseg} if zseg else {}
    for size, step, lag, thr in ((100, 10, 100, 3.0), (100, 10, 100, 1.0), (100, 10, 50, 4.0), (100, 20, 2, 2.0), (50, 10, 200, 2.5), (100, 10, 7, 1.5)):
        _, _, _, ref = orc.wave_batch(cat, off, size, step, lag, thr, infl, threads=8)
        with _Env(**env):
            _, _, _, got = api.wave_batch(ctx, cat, off, size, step, lag, thr, infl)
        with _Env(GAMS_CUDA_ZWALK_WARPS=1, **env):
            _, _, _, old = api.wave_batch(ctx, cat, off, size, step, lag, thr, infl)
        _eq(f"thread walk {size}/{step} lag {lag} thr {thr} infl {infl} zseg {zseg}", got, ref)
        _eq(f"warp walk {size}/{step} lag {lag} thr {thr} infl {infl} zseg {zseg}", old, ref)


# --------------------------------------------------------------- gunzip: two-phase decoder (inflate2.cu) and the CRC-32 check
def _gzb(data, level=1, strategy=0, mem=8):
    import zlib
    o = zlib.compressobj(level, zlib.DEFLATED, 31, mem, strategy)
    return o.compress(data) + o.flush()


def test_gunzip_crc32_mismatch_is_status_7(ctx, orc):
    """flate2's GzDecoder verifies the CRC-32 of the trailer (redis.rs:156-161 unwraps the error): a blob whose payload was
    corrupted with ISIZE preserved, or whose trailer CRC was altered, ends in status 7; its neighbours are untouched."""
    import zlib
    dna = synth.genome(120_000, seed=21).tobytes()
    good = _gzb(dna)
    # (a) stored block: flipping a payload byte keeps the stream valid and the size right
    stored = bytearray(_gzb(dna[:30_000], level=0))
    stored[200] ^= 0x04
    # (b) the trailer's CRC itself
    bad_crc = bytearray(good)
    bad_crc[-8] ^= 0x01
    # (c) a literal changed inside a fixed-Huffman block: swap two equal-length codes by re-encoding other data with the trailer of the first
    other = dna[:5000].replace(b"A", b"C", 1)
    swapped = bytearray(_gzb(other, strategy=zlib.Z_FIXED))
    swapped[-8:] = _gzb(dna[:5000], strategy=zlib.Z_FIXED)[-8:]
    blobs = [good, bytes(stored), bytes(bad_crc), bytes(swapped), good]
    sizes = [len(dna), 30_000, len(dna), 5000, len(dna)]
    ooff = np.concatenate(([0], np.cumsum(sizes))).astype(np.uint64)
    for decoder in (0, 1):
        ctx.set_gunzip_options(decoder, True)
        try:
            out, off, st = api.gunzip_batch(ctx, *api.pack_blobs(blobs), ooff, raise_on_error=False)
            assert st.tolist() == [0, 7, 7, 7, 0], (decoder, st.tolist())
            assert out[:len(dna)].tobytes() == dna and out[int(ooff[4]):].tobytes() == dna
            ctx.set_gunzip_options(decoder, False)
            _, _, st = api.gunzip_batch(ctx, *api.pack_blobs(blobs), ooff, raise_on_error=False)
            assert st.tolist() == [0, 0, 0, 0, 0]
        finally:
            ctx.set_gunzip_options(0, True)


def test_gunzip_two_decoders_agree(ctx, orc):
    """The two-phase decoder against the warp-per-blob decoder and zlib on the shapes that stress token resolution: runs of
    one base (distance 1, length 258), tandem repeats with periods 2..40, matches that start inside the batch being resolved,
    literals only, stored data (token budget exceeded: falls back), > 32 blobs (several thread blocks), ragged sizes."""
    import zlib
    rng = np.random.default_rng(31)
    dna = synth.genome(700_000, seed=33).tobytes()
    plain = [dna[:300_001], b"N" * 70_000, (b"AC" * 9 + b"G") * 3000, b"".join(bytes([65 + (i % p)]) for p in (2, 3, 5, 7, 40) for i in range(9000)),
             dna[1000:1040] * 2000, rng.integers(0, 256, 40_000, dtype=np.uint8).tobytes(), b"", b"A", dna[:257], dna[:258] * 3]
    for k in range(60):
        a = int(rng.integers(0, 600_000))
        plain.append(dna[a:a + int(rng.integers(1, 90_000))])
    # N gaps inside sequence, soft-masked stretches
    mixed = bytearray(dna[400_000:600_000])
    for a in range(5000, 190_000, 17_000):
        mixed[a:a + int(rng.integers(1, 3000))] = b"N" * len(mixed[a:a + int(rng.integers(1, 3000))])
    plain.append(bytes(mixed))
    blobs = [_gzb(p, level=(1, 6, 9)[i % 3]) for i, p in enumerate(plain)]
    blobs[3] = _gzb(plain[3], strategy=zlib.Z_RLE)
    blobs[5] = _gzb(plain[5], level=0)
    cat, goff = api.pack_blobs(blobs)
    ooff = np.concatenate(([0], np.cumsum([len(p) for p in plain]))).astype(np.uint64)
    res = {}
    for decoder in (0, 1):
        ctx.set_gunzip_options(decoder, True)
        try:
            out, off, st = api.gunzip_batch(ctx, cat, goff, ooff)
        finally:
            ctx.set_gunzip_options(0, True)
        assert not st.any(), (decoder, st.tolist())
        res[decoder] = out
        for i, p in enumerate(plain):
            got = out[int(off[i]):int(off[i + 1])].tobytes()
            assert got == p, f"decoder {decoder}, blob {i}: first difference at {next((k for k in range(min(len(got), len(p))) if got[k] != p[k]), None)}"
    assert np.array_equal(res[0], res[1])


@pytest.mark.parametrize("pieces", [2, 5, 16])
@pytest.mark.parametrize("decoder", [0, 1])
def test_gunzip_batch_in_pieces(ctx, orc, pieces, decoder):
    """gams_cuda_gunzip_batch cuts large batches into pieces (upload, decode and download of different pieces overlap): same bytes
    and statuses as one piece — blobs of very different sizes, an empty one, a blob with a wrong CRC-32 in a later piece."""
    import zlib
    g = synth.genome(1_200_000, seed=77)
    cuts = [0, 300_000, 300_000, 310_000, 700_001, 700_100, 1_000_000, 1_200_000]
    plain = [g[a:b].tobytes() for a, b in zip(cuts[:-1], cuts[1:])]
    blobs = [bytearray(_gzb(p, level=(1, 6)[i % 2])) for i, p in enumerate(plain)]
    blobs[5][-8] ^= 0x01                                       # CRC-32 of the trailer
    cat, goff = api.pack_blobs([bytes(b) for b in blobs])
    ooff = np.array(cuts, dtype=np.uint64)
    ctx.set_gunzip_options(decoder, True)
    try:
        with _Env(GAMS_CUDA_GZ_PIECES=1):
            ref_out, _, ref_st = api.gunzip_batch(ctx, cat, goff, ooff, raise_on_error=False)
        with _Env(GAMS_CUDA_GZ_PIECES=pieces):
            out, off, st = api.gunzip_batch(ctx, cat, goff, ooff, raise_on_error=False)
    finally:
        ctx.set_gunzip_options(0, True)
    assert st.tolist() == [0, 0, 0, 0, 0, 7, 0] and ref_st.tolist() == st.tolist()
    assert np.array_equal(out, ref_out)
    for i, p in enumerate(plain):
        assert out[int(off[i]):int(off[i + 1])].tobytes() == p, f"blob {i}"


# --------------------------------------------------------------- chained feed: seq:{ctg} gzip blobs -> gunzip -> GC-wave
@pytest.mark.parametrize("infl", [1.0, 0.5])
def test_wave_gz_batch_equals_wave_batch(ctx, orc, infl):
    """gams_cuda_wave_gz_batch (Conn::get_seq + wave::proc_ctg, redis.rs:142-161 + wave.rs:138-155) on the blobs encode_gz writes
    (Compression::fast) gives bit for bit what gams_cuda_wave_batch gives on the bases, and what the oracle gives."""
    g = synth.genome(2_400_000, seed=91)
    cuts = [0, 500_000, 1_000_003, 1_000_903, 1_900_000, 2_400_000]
    seqs = [g[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    cat, off = orc.pack_ctgs(seqs)
    blobs = [orc.encode_gz(s.tobytes()) for s in seqs]
    gcat, goff = api.pack_blobs(blobs)
    woff, cnt, gc, sig, st = api.wave_gz_batch(ctx, gcat, goff, off, 100, 10, 50, 3.0, infl)
    assert not st.any()
    rw, rc, rg, rs = orc.wave_batch(cat, off, 100, 10, 50, 3.0, infl, threads=8)
    _eq("win_off", woff, rw)
    _eq("gc_count", cnt, rc)
    _eq("gc", gc.view(np.uint32), rg.view(np.uint32))
    _eq("signal", sig, rs)
    w2, c2, g2, s2 = api.wave_batch(ctx, cat, off, 100, 10, 50, 3.0, infl)
    assert np.array_equal(c2, cnt) and np.array_equal(s2, sig)


@pytest.mark.parametrize("pieces,infl", [(2, 1.0), (3, 1.0), (5, 1.0), (16, 1.0), (3, 0.5)])
def test_wave_gz_batch_in_pieces(ctx, orc, monkeypatch, pieces, infl):
    """With all windows coming back the chained path decodes the batch in pieces (the windows of one go down while the next is
    decoded).  Small slices force several pieces on a small input: same bits as one piece and as the oracle, a bad blob in the
    later piece keeps its status; a ctg of 61 windows and pieces of a single ctg included."""
    g = synth.genome(1_500_000, seed=93)
    cuts = [0, 130_000, 260_017, 390_000, 520_000, 650_000, 780_001, 910_000, 1_040_000, 1_170_000, 1_170_700, 1_300_000, 1_500_000]
    seqs = [g[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    cat, off = orc.pack_ctgs(seqs)
    blobs = [orc.encode_gz(s.tobytes()) for s in seqs]
    gcat, goff = api.pack_blobs(blobs)
    monkeypatch.setenv("GAMS_CUDA_WAVE_SLICE", "100000")      # one ctg per slice
    monkeypatch.setenv("GAMS_CUDA_GZ_PIECES", "1")
    w1, c1, g1, s1, st1 = api.wave_gz_batch(ctx, gcat, goff, off, 100, 10, 50, 3.0, infl)
    monkeypatch.setenv("GAMS_CUDA_GZ_PIECES", str(pieces))
    woff, cnt, gc, sig, st = api.wave_gz_batch(ctx, gcat, goff, off, 100, 10, 50, 3.0, infl)
    assert not st.any() and not st1.any()
    rw, rc, rg, rs = orc.wave_batch(cat, off, 100, 10, 50, 3.0, infl, threads=8)
    for name, a, b, r in (("win_off", woff, w1, rw), ("gc_count", cnt, c1, rc), ("gc", gc.view(np.uint32), g1.view(np.uint32), rg.view(np.uint32)),
                          ("signal", sig, s1, rs)):
        _eq(name, a, r)
        _eq(name + " (one piece)", b, r)
    bad = [bytearray(b) for b in blobs]
    bad[7][-8] ^= 0x10                                        # CRC of a blob in a later piece
    gcat2, goff2 = api.pack_blobs([bytes(b) for b in bad])
    _, c2, _, s2, st2 = api.wave_gz_batch(ctx, gcat2, goff2, off, 100, 10, 50, 3.0, infl, raise_on_error=False)
    assert st2.tolist() == [0] * 7 + [7] + [0] * (len(blobs) - 8)
    assert np.array_equal(c2, rc) and np.array_equal(s2, rs)  # the payload itself was intact


def test_wave_gz_batch_reports_a_bad_blob(ctx, orc):
    """A corrupted blob (payload changed, CRC-32 no longer matches) and a blob of the wrong length fail the call with their
    status; the other ctgs' statuses stay 0."""
    g = synth.genome(600_000, seed=92)
    seqs = [g[:200_000], g[200_000:400_000], g[400_000:]]
    cat, off = orc.pack_ctgs(seqs)
    blobs = [bytearray(orc.encode_gz(s.tobytes())) for s in seqs]
    blobs[1][-8] ^= 0x10                                    # the trailer's CRC
    gcat, goff = api.pack_blobs([bytes(b) for b in blobs])
    with pytest.raises(api.GamsCudaError):
        api.wave_gz_batch(ctx, gcat, goff, off)
    _, _, _, _, st = api.wave_gz_batch(ctx, gcat, goff, off, raise_on_error=False)
    assert st.tolist() == [0, 7, 0]
    off2 = off.copy()
    off2[1:] += 5                                           # the caller expects 5 bases more in the first ctg
    _, _, _, _, st = api.wave_gz_batch(ctx, *api.pack_blobs([orc.encode_gz(s.tobytes()) for s in seqs]), off2, raise_on_error=False)
    assert st[0] != 0


def _multi_block_gz(rng, dna, pieces, header_extra=b""):
    """One gzip member whose deflate stream is cut into blocks of the given (length, mode): dynamic blocks of level 1 / 6 / 9,
    fixed-Huffman blocks (Z_FIXED), stored blocks (level 0), empty stored blocks (Z_SYNC_FLUSH) and byte-aligned restarts
    (Z_FULL_FLUSH) — block boundaries at irregular distances, headers of every type, some of them only a few bytes apart."""
    import struct, zlib
    raw, body = bytearray(), bytearray()
    comp = None
    for n, mode in pieces:
        a = int(rng.integers(0, len(dna) - n))
        chunk = dna[a:a + n]
        raw += chunk
        level, strategy = {"l1": (1, 0), "l6": (6, 0), "l9": (9, 0), "fixed": (1, zlib.Z_FIXED), "stored": (0, 0)}[mode]
        if comp is None:
            comp = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
            cur = (level, strategy)
        elif cur != (level, strategy):
            # a new compressor for the new parameters: the stream so far ends at a byte boundary (Z_FULL_FLUSH), the next
            # compressor continues it — its first match may not reach back, which zlib guarantees for a fresh object
            body += comp.flush(zlib.Z_FULL_FLUSH)
            comp = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
            cur = (level, strategy)
        body += comp.compress(chunk)
        body += comp.flush(zlib.Z_SYNC_FLUSH if rng.random() < 0.5 else zlib.Z_FULL_FLUSH)
    body += comp.flush(zlib.Z_FINISH)
    flg = 8 if header_extra else 0
    head = b"\x1f\x8b\x08" + bytes([flg]) + b"\0\0\0\0\0\xff" + (header_extra + b"\0" if header_extra else b"")
    return bytes(raw), head + bytes(body) + struct.pack("<II", zlib.crc32(bytes(raw)) & 0xFFFFFFFF, len(raw) & 0xFFFFFFFF)


def test_gunzip_parts_at_irregular_block_boundaries(ctx, orc):
    """The part-parallel decoder (a part per deflate block found by the header search) on streams it cannot predict: blocks
    of 40 bytes to 400 KB in one member, dynamic / fixed / stored / empty blocks in any order, a 30 KB FNAME in front of the
    first block, text and random bytes next to sequence — byte for byte what zlib gives, and the same from both decoders."""
    import zlib
    rng = np.random.default_rng(77)
    dna = synth.genome(1_500_000, seed=78).tobytes()
    text = (b"The quick brown fox jumps over the lazy dog. " * 64 + bytes(rng.integers(32, 127, 4096, dtype=np.uint8))) * 40
    modes = ["l1", "l1", "l1", "l6", "l9", "fixed", "stored"]
    plain, blobs = [], []
    for k in range(24):
        pieces = [(int(rng.choice([40, 300, 5_000, 30_000, 70_000, 200_000, 400_000], p=[.1, .1, .15, .2, .25, .15, .05])), str(rng.choice(modes)))
                  for _ in range(int(rng.integers(2, 14)))]
        raw, gz = _multi_block_gz(rng, dna if k % 4 else text, pieces, header_extra=b"n" * 30_000 if k == 5 else b"")
        assert zlib.decompress(gz, 31) == raw
        plain.append(raw), blobs.append(gz)
    big = dna[:1_400_000]                        # one long member of plain level-1 blocks among them (~21 blocks)
    plain.append(big), blobs.append(_gzb(big))
    gcat, goff = api.pack_blobs(blobs)
    for decoder in (0, 1):
        ctx.set_gunzip_options(decoder, True)
        out, off, st = api.gunzip_batch(ctx, gcat, goff)
        assert not st.any(), st
        for i, p in enumerate(plain):
            assert out[int(off[i]):int(off[i + 1])].tobytes() == p, f"decoder {decoder}, blob {i}"
    ctx.set_gunzip_options(0, True)


@pytest.mark.parametrize("mem", [1, 3, 6, 9])
def test_gunzip_block_sizes_of_other_encoders(ctx, orc, mem):
    """Deflate blocks of 128 symbols (memLevel 1: a header every ~170 compressed bytes, dozens per search range) up to 32 768
    symbols (memLevel 9: twice zlib's default, a header every ~44 KB): the header search learns the block size from the stream,
    whatever the encoder's buffer was.  flate2's miniz_oxide cuts at ~20 000 symbols; zlib's default (memLevel 8) is 16 383."""
    rng = np.random.default_rng(100 + mem)
    dna = synth.genome(2_000_000, seed=101).tobytes()
    plain = [dna[:1_000_000], dna[1_000_000:1_300_000], dna[1_300_000:1_300_700], dna[1_400_000:2_000_000]]
    blobs = [_gzb(p, level=1, mem=mem) for p in plain]
    out, off, st = api.gunzip_batch(ctx, *api.pack_blobs(blobs))
    assert not st.any(), st
    assert out.tobytes() == b"".join(plain)


def test_gunzip_one_huge_blob_and_many_tiny(ctx, orc):
    """Batches at the two ends of the part table: one 12 Mbp member alone (hundreds of ranges, several groups of the header
    search), and 3000 members of 0..300 bases (one range each)."""
    rng = np.random.default_rng(79)
    dna = synth.genome(12_000_000, seed=80).tobytes()
    out, off, st = api.gunzip_batch(ctx, *api.pack_blobs([_gzb(dna)]))
    assert not st.any() and out.tobytes() == dna
    cuts = np.concatenate(([0], np.cumsum(rng.integers(0, 301, 3000))))
    small = [dna[int(a):int(b)] for a, b in zip(cuts[:-1], cuts[1:])]
    out, off, st = api.gunzip_batch(ctx, *api.pack_blobs([_gzb(s) for s in small]))
    assert not st.any() and out.tobytes() == b"".join(small)


def test_wave_peaks_gz_batch_equals_wave_peaks_batch(ctx, orc):
    """gams_cuda_wave_peaks_gz_batch (get_seq + wave without --signal + merge_ints) gives the records gams_cuda_wave_peaks_batch
    gives on the bases; a corrupted blob fails the call with its status."""
    g = synth.genome(2_400_000, seed=93)
    cuts = [0, 500_000, 1_000_003, 1_000_903, 1_900_000, 2_400_000]
    seqs = [g[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    cat, off = orc.pack_ctgs(seqs)
    blobs = [orc.encode_gz(s.tobytes()) for s in seqs]
    gcat, goff = api.pack_blobs(blobs)
    ref = api.wave_peaks_batch(ctx, cat, off, 100, 10, 50, 3.0, 1.0, 0.2)
    got, st = api.wave_peaks_gz_batch(ctx, gcat, goff, off, 100, 10, 50, 3.0, 1.0, 0.2)
    assert not st.any() and ref["ctg"].size > 0
    for k in ("ctg", "first", "last", "signal"):
        _eq(k, got[k], ref[k])
    _eq("gc", got["gc"].view(np.uint32), ref["gc"].view(np.uint32))
    got2, _ = api.wave_peaks_gz_batch(ctx, gcat, goff, off, 100, 10, 50, 3.0, 1.0, 0.2, cap=8)     # ECAP, then the exact size
    _eq("first (small cap)", got2["first"], ref["first"])
    bad = [bytearray(b) for b in blobs]
    bad[3][len(bad[3]) // 2] ^= 0x55
    with pytest.raises(api.GamsCudaError):
        api.wave_peaks_gz_batch(ctx, *api.pack_blobs([bytes(b) for b in bad]), off, 100, 10, 50)
    _, st = api.wave_peaks_gz_batch(ctx, *api.pack_blobs([bytes(b) for b in bad]), off, 100, 10, 50, raise_on_error=False)
    assert st[3] != 0 and not st[:3].any() and st[4] == 0


# --------------------------------------------------------------- rg index build: the hand-written segmented sort (idx_build.cu)
def test_build_idx_rg_segment_sizes(ctx, orc):
    """Lapper::new on the device for segments of every size class: empty, 1, 2, 3, powers of two and their neighbours, the
    shared-memory capacity (16384) and its neighbours, segments that are sorted in global memory (40 000, 70 001), heavy
    duplicates, already sorted and reversed input, stops that sort differently from the starts."""
    rng = np.random.default_rng(71)
    sizes = [0, 1, 2, 3, 31, 32, 33, 1000, 4095, 4096, 4097, 16383, 16384, 16385, 40_000, 0, 70_001, 5]
    st, sp, off = [], [], [0]
    for k, n in enumerate(sizes):
        a = rng.integers(1, 5_000_000, n).astype(np.uint32)
        if k % 4 == 1:
            a = np.resize(np.repeat(a[: max(1, n // 50)], 50), n) if n else a   # heavy duplicates
        if k % 5 == 2:
            a = np.sort(a)[::-1].copy()                                        # reversed
        e = a + rng.integers(1, 3000, n).astype(np.uint32)
        st.append(a), sp.append(e), off.append(off[-1] + n)
    st, sp, off = np.concatenate(st), np.concatenate(sp), np.asarray(off, dtype=np.uint64)
    ix = api.build_idx_rg(ctx, st, sp, off)
    for s in range(len(sizes)):
        a, b = int(off[s]), int(off[s + 1])
        order = np.lexsort((sp[a:b], st[a:b]))
        _eq(f"segment {s} (n = {sizes[s]}) iv_start", ix["iv_start"][a:b], st[a:b][order])
        _eq(f"segment {s} iv_stop", ix["iv_stop"][a:b], sp[a:b][order])
        _eq(f"segment {s} starts", ix["starts"][a:b], np.sort(st[a:b]))
        _eq(f"segment {s} stops", ix["stops"][a:b], np.sort(sp[a:b]))
        assert int(ix["max_len"][s]) == (int((sp[a:b] - st[a:b]).max()) if b > a else 0)
    lp = orc.lapper_new([(int(x), int(y), "") for x, y in zip(st[int(off[7]):int(off[8])], sp[int(off[7]):int(off[8])])])
    assert ix["blob"][int(ix["blob_off"][7]):int(ix["blob_off"][8])].tobytes() == orc.lapper_to_bincode(lp)
