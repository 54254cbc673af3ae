"""BASELINE.json configs at FULL size on the GPU, checked through size-independent properties
(the oracle cannot finish these sizes in seconds) plus oracle spot checks on random ctgs.

  C3  Arabidopsis-scale genome (120 Mbp): GC-wave 100/10 + `sw` around 10^6 features
  C4  human-scale genome (3.1 Gbp, 24 chromosomes): GC-wave 100/10
  C5  10^8 query ranges against 10^8 SNP intervals
"""
import ctypes as C

import numpy as np
import pytest

from gams_b200 import _lib, api, synth

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

PIECE = 500_000


@pytest.fixture(scope="module")
def ctx():
    c = api.Context([0])
    yield c
    c.close()


def _wave_on_device(ctx, d_seq, seq_off, size, step, lag=100, want_signal=True, want_count=True):
    plan = api.WavePlan(ctx, seq_off, size, step, lag)
    assert d_seq.numel() >= plan.seq_capacity
    n = plan.total_windows
    d_gc = torch.empty(n, dtype=torch.float32, device=d_seq.device)
    d_cnt = torch.empty(n, dtype=torch.int32, device=d_seq.device) if want_count else None
    d_sig = torch.empty(n, dtype=torch.int8, device=d_seq.device) if want_signal else None
    st = torch.cuda.current_stream().cuda_stream
    if want_signal:
        # the fused kernel (what `wave` runs by default) ...
        assert plan.fused(3.0, 1.0)
        plan.wave(st, d_seq, 3.0, 1.0, d_gc, d_sig, d_cnt)
        torch.cuda.synchronize()
        # ... against Kernel A + the exact replay of EVERY window (mode 1), over the whole genome
        d_gc2 = torch.empty_like(d_gc)
        d_sig2 = torch.empty_like(d_sig)
        try:
            ctx.set_zscore_mode(1)
            plan.gc(st, d_seq, d_gc2, None)
            plan.signal(st, d_gc2, 3.0, 1.0, d_sig2)
            torch.cuda.synchronize()
        finally:
            ctx.set_zscore_mode(0)
        assert torch.equal(d_gc.view(torch.int32), d_gc2.view(torch.int32)), "fused GC differs from Kernel A"
        nbad = int((d_sig != d_sig2).sum())
        assert nbad == 0, f"{nbad} of {n} fused signals differ from the exact replay"
        del d_gc2, d_sig2
    else:
        plan.gc(st, d_seq, d_gc, d_cnt)
    torch.cuda.synchronize()
    plan.close()
    return d_gc, d_cnt, d_sig


def _wave_properties(ctx, orc, n_bases, chr_lengths, seed, n_spot):
    dev = torch.device("cuda", 0)
    seq_off, ctgs = synth.ctg_layout(chr_lengths, PIECE)
    n_ctg = seq_off.size - 1
    cap = ((n_bases + 31) // 32) * 32 + 64
    d_seq = synth.genome_torch(n_bases, seed=seed, device=dev, capacity=cap)
    lens = np.diff(seq_off.astype(np.int64))

    gc100, cnt100, sig = _wave_on_device(ctx, d_seq, seq_off, 100, 10)
    _, cnt10, _ = _wave_on_device(ctx, d_seq, seq_off, 10, 10, want_signal=False)
    w100 = api.win_offsets(seq_off, 100, 10).astype(np.int64)
    w10 = api.win_offsets(seq_off, 10, 10).astype(np.int64)
    assert int(w100[-1]) == cnt100.numel() == int(((lens - 100) // 10 + 1).sum())

    # (P1) checksum of checksums: all 10/10 blocks together hold exactly the GC bases of the covered prefix of every ctg
    lut = torch.zeros(256, dtype=torch.uint8, device=dev)
    lut[torch.tensor(list(b"cgCG"), device=dev, dtype=torch.long)] = 1
    total_gc = 0
    for s in range(0, n_bases, 1 << 28):
        total_gc += int(lut[d_seq[s:min(n_bases, s + (1 << 28))].long()].sum(dtype=torch.int64))
    tail = 0
    h_off = seq_off.astype(np.int64)
    for c in range(n_ctg):  # bytes past the last whole 10-block of each ctg (< 10 per ctg)
        a = h_off[c] + (lens[c] // 10) * 10
        if a < h_off[c + 1]:
            tail += int(lut[d_seq[a:h_off[c + 1]].long()].sum())
    assert int(cnt10.sum(dtype=torch.int64)) == total_gc - tail

    # (P2) linearity: a 100/10 window is the sum of ten consecutive 10/10 blocks, for every window of every ctg
    cs = torch.zeros(cnt10.numel() + 1, dtype=torch.int64, device=dev)
    cs[1:] = torch.cumsum(cnt10, 0, dtype=torch.int64)
    shift = torch.repeat_interleave(torch.as_tensor(w10[:-1] - w100[:-1], device=dev), torch.as_tensor(np.diff(w100), device=dev))
    idx = torch.arange(cnt100.numel(), device=dev) + shift
    assert torch.equal(cs[idx + 10] - cs[idx], cnt100.to(torch.int64))
    del cs, idx, shift

    # (P3) the f32 fraction is exactly count / size
    # (tensor / tensor: torch turns tensor / python-scalar into a multiply by the reciprocal, which is not IEEE division)
    assert torch.equal(gc100, cnt100.to(torch.float32) / torch.full_like(gc100, 100.0))
    assert int(cnt100.max()) <= 100 and int(cnt100.min()) >= 0

    # (P4) signals: zero in the first `lag` windows of every ctg, values in {-1,0,1}, oracle-exact on random ctgs
    first = torch.as_tensor(w100[:-1], device=dev)
    for k in (0, 50, 99):
        assert int(sig[first + k].abs().max()) == 0
    assert set(torch.unique(sig).tolist()) <= {-1, 0, 1}
    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(n_ctg, min(n_spot, n_ctg), replace=False))
    cat, off = orc.pack_ctgs([d_seq[int(seq_off[c]):int(seq_off[c + 1])].cpu().numpy() for c in pick])
    _, rc, rg, rs = orc.wave_batch(cat, off, 100, 10, 100, 3.0, 1.0, threads=16)
    rw = orc.win_offsets(off, 100, 10).astype(np.int64)
    for k, c in enumerate(pick):
        a, b = int(w100[c]), int(w100[c + 1])
        ra, rb = int(rw[k]), int(rw[k + 1])
        assert np.array_equal(cnt100[a:b].cpu().numpy().view(np.uint32), rc[ra:rb]), f"ctg {c} counts"
        assert np.array_equal(gc100[a:b].cpu().numpy().view(np.uint32), rg[ra:rb].view(np.uint32)), f"ctg {c} gc"
        assert np.array_equal(sig[a:b].cpu().numpy(), rs[ra:rb]), f"ctg {c} signals"
    frac = float((sig != 0).float().mean())
    assert 1e-4 < frac < 0.2, frac
    return d_seq, seq_off, ctgs


def _peaks_properties(ctx, orc, d_seq, seq_off, n_bases):
    """Full-size checks of the device peak merge against a torch restatement of the run rule and the host-side merge."""
    dev = d_seq.device
    st = torch.cuda.current_stream().cuda_stream
    L = _lib.load()
    plan = api.WavePlan(ctx, seq_off, 100, 10, 100)
    n = plan.total_windows
    d_gc = torch.empty(n, dtype=torch.float32, device=dev)
    d_sig = torch.empty(n, dtype=torch.int8, device=dev)
    plan.wave(st, d_seq, 3.0, 1.0, d_gc, d_sig)
    cap = n // 16
    pk = [torch.empty(cap, dtype=dt, device=dev) for dt in (torch.int32, torch.int32, torch.int32, torch.float32, torch.int8)]
    d_n = torch.zeros(1, dtype=torch.int64, device=dev)
    _lib.check(L.gams_cuda_wave_peaks_dev(plan._h, st, d_gc.data_ptr(), d_sig.data_ptr(), 0.2, cap, *[t.data_ptr() for t in pk],
                                          d_n.data_ptr()))
    torch.cuda.synchronize()
    n_pk = int(d_n.item())
    assert 0 < n_pk <= cap
    ctg, first, last, gcf, sgn = (t[:n_pk] for t in pk)
    woff = torch.as_tensor(api.win_offsets(seq_off, 100, 10).astype(np.int64), device=dev)
    # (P5) records are the run starts: a signal window with no same-sign signal among the 9 windows before it in its ctg
    idx = torch.arange(n, device=dev)
    cid = torch.searchsorted(woff, idx, right=True) - 1
    is_start = d_sig != 0
    for d in range(1, 10):
        same = torch.zeros(n, dtype=torch.bool, device=dev)
        same[d:] = (d_sig[d:] == d_sig[:-d]) & (cid[d:] == cid[:-d])
        is_start &= ~same
    starts = torch.nonzero(is_start).flatten()
    assert starts.numel() == n_pk
    g_first = woff[ctg.long()] + first.long()
    assert torch.equal(g_first, starts), "records are not the run starts in window order"
    assert torch.equal(sgn, d_sig[starts]) and torch.equal(gcf.view(torch.int32), d_gc[starts].view(torch.int32))
    # (P6) every run ends on a window of its sign, stays inside its ctg, and is followed by 9 windows without that sign
    g_last = woff[ctg.long()] + last.long()
    assert bool((last >= first).all()) and torch.equal(d_sig[g_last], sgn) and bool((g_last < woff[ctg.long() + 1]).all())
    for d in range(1, 10):
        nxt = torch.clamp(g_last + d, max=n - 1)
        inside = (g_last + d) < woff[ctg.long() + 1]
        assert not bool(((d_sig[nxt] == sgn) & inside).any())
    # ... and the text of a few ctgs equals the host-side merge_ints of the per-window arrays
    rng = np.random.default_rng(3)
    h_ctg = ctg.cpu().numpy()
    for c in rng.choice(seq_off.size - 1, 6, replace=False):
        a, b = int(woff[c]), int(woff[c + 1])
        m = h_ctg == c
        got = api.wave_peaks_text("X", 1, 100, 10, first.cpu().numpy()[m], last.cpu().numpy()[m], gcf.cpu().numpy()[m], sgn.cpu().numpy()[m])
        assert got == api.wave_text("X", 1, 100, 10, d_gc[a:b].cpu().numpy(), d_sig[a:b].cpu().numpy(), 0.2)
    plan.close()
    del d_gc, d_sig, pk, idx, cid, is_start
    torch.cuda.empty_cache()


def _influence_properties(ctx, orc, d_seq, seq_off, influence, n_spot):
    """`wave --influence x` (x != 1) at full size: the segmented walk (zscore_seg_walk_kernel, zscore_seg_fix_kernel) behind the fused kernel.
    (P7) filtered differs from data only AT signals (stat.rs:40-48), so a window whose statistics window [i-1-lag, i-1) holds no
    signal of the result decides exactly as the influence-1 process does: wherever the influence-x result has no signal among
    the lag+1 windows before i, its decision at i equals the influence-1 decision at i — for every window of the genome.
    (P8) oracle-exact on n_spot random ctgs."""
    dev = d_seq.device
    st = torch.cuda.current_stream().cuda_stream
    lag = 100
    plan = api.WavePlan(ctx, seq_off, 100, 10, lag)
    n = plan.total_windows
    d_gc = torch.empty(n, dtype=torch.float32, device=dev)
    d1 = torch.empty(n, dtype=torch.int8, device=dev)
    dx = torch.empty(n, dtype=torch.int8, device=dev)
    d_filt = torch.empty(n, dtype=torch.float32, device=dev)
    plan.wave(st, d_seq, 3.0, 1.0, d_gc, d1)
    plan.wave(st, d_seq, 3.0, influence, d_gc, dx, None, d_filt)
    torch.cuda.synchronize()
    del d_filt
    w = api.win_offsets(seq_off, 100, 10).astype(np.int64)
    woff = torch.as_tensor(w, device=dev)
    # windows since the last signal of dx inside the same ctg (a running maximum of the signal positions)
    idx = torch.arange(n, device=dev)
    last_sig = torch.where(dx != 0, idx, torch.full_like(idx, -1))
    last_sig = torch.cummax(last_sig, 0).values
    prev = torch.full_like(idx, -1)
    prev[1:] = last_sig[:-1]                                     # the last signal strictly before i
    cid = torch.searchsorted(woff, idx, right=True) - 1
    prev = torch.where(prev >= woff[cid], prev, torch.full_like(prev, -(1 << 40)))   # signals of earlier ctgs do not count
    clean = (idx - prev) > (lag + 1)
    nbad = int(((dx != d1) & clean).sum())
    assert nbad == 0, f"{nbad} windows with a clean statistics window differ from the influence-1 decision"
    assert int((dx != d1).sum()) > 1000                          # not vacuous: the influence does change decisions behind signals
    first = woff[:-1]
    for k in (0, 50, 99):
        assert int(dx[first + k].abs().max()) == 0
    rng = np.random.default_rng(77)
    n_ctg = seq_off.size - 1
    pick = np.sort(rng.choice(n_ctg, min(n_spot, n_ctg), replace=False))
    cat, off = orc.pack_ctgs([d_seq[int(seq_off[c]):int(seq_off[c + 1])].cpu().numpy() for c in pick])
    _, _, _, rs = orc.wave_batch(cat, off, 100, 10, lag, 3.0, influence, threads=16)
    rw = orc.win_offsets(off, 100, 10).astype(np.int64)
    for k, c in enumerate(pick):
        a, b = int(w[c]), int(w[c + 1])
        assert np.array_equal(dx[a:b].cpu().numpy(), rs[int(rw[k]):int(rw[k + 1])]), f"ctg {c}: influence {influence} signals"
    plan.close()
    del d_gc, d1, dx, idx, last_sig, prev, cid, clean
    torch.cuda.empty_cache()


def test_c4_human_scale_wave(ctx, orc):
    n = 3_100_000_000
    d_seq, seq_off, _ = _wave_properties(ctx, orc, n, synth.human_like_lengths(n), seed=20240613, n_spot=240)
    torch.cuda.empty_cache()
    _peaks_properties(ctx, orc, d_seq, seq_off, n)
    torch.cuda.empty_cache()
    _influence_properties(ctx, orc, d_seq, seq_off, 0.5, n_spot=64)
    torch.cuda.empty_cache()


def test_c3_arabidopsis_scale_wave_and_sw(ctx, orc):
    chr_lengths = [30_427_671, 19_698_289, 23_459_830, 18_585_056, 26_975_502, 154_478, 366_924]  # Atha: 119 667 750 bp
    n = int(sum(chr_lengths))
    d_seq, seq_off, ctgs = _wave_properties(ctx, orc, n, chr_lengths, seed=7, n_spot=6)
    h_seq = d_seq[:n].cpu().numpy()
    # `gams gen` on the whole genome (7 chromosomes, N gaps of 1-49 bases every 250 kbp, --fill 50): the device scan + host
    # fill / excise / piece cut against the oracle's numpy restatement, chromosome by chromosome
    chr_off = np.concatenate(([0], np.cumsum(chr_lengths))).astype(np.uint64)
    for piece, fill in ((PIECE, 50), (PIECE, 10)):
        gc_, gs_, ge_ = api.gen_ctgs(ctx, h_seq, chr_off, piece=piece, fill=fill, min_=5000)
        ref = [(k, a, b) for k in range(len(chr_lengths))
               for a, b in orc.gen_ctgs(h_seq[int(chr_off[k]):int(chr_off[k + 1])], piece=piece, fill=fill, min_=5000)]
        assert list(zip(gc_.tolist(), gs_.tolist(), ge_.tolist())) == ref
    assert len(ref) > len(ctgs)   # --fill 10 leaves the longer N gaps open: more, shorter regions
    # 10^6 point features uniform over the genome; chr coordinates per ctg as `gen` assigns them
    rng = np.random.default_rng(8)
    lens = np.diff(seq_off.astype(np.int64))
    chr_start = np.array([c[1] for c in ctgs], dtype=np.int32)
    n_feat = 1_000_000
    f_ctg = np.sort(rng.choice(len(ctgs), n_feat, p=lens / lens.sum())).astype(np.uint32)
    f_s = (chr_start[f_ctg] + (rng.random(n_feat) * lens[f_ctg]).astype(np.int64)).astype(np.int32)
    f_e = np.minimum(f_s + rng.integers(0, 3, n_feat), chr_start[f_ctg] + lens[f_ctg] - 1).astype(np.int32)
    r = api.sw_batch(ctx, h_seq, seq_off, chr_start, f_ctg, f_s, f_e)
    rows = np.diff(r["row_off"].astype(np.int64))
    assert rows.min() >= 1 and rows.max() <= 41 and int(rows.sum()) == r["gc"].size > 35_000_000
    assert np.all((r["end"] - r["start"] + 1 <= 100) & (r["end"] >= r["start"]))
    first = r["row_off"][:-1].astype(np.int64)
    assert np.all(r["type"][first] == 0) and np.all(r["dist"][first] == 0)
    assert np.all((r["gc"] >= 0) & (r["gc"] <= 1))
    ok = ~np.isnan(r["mean"])
    assert np.all((r["mean"][ok] >= 0) & (r["mean"][ok] <= 1))
    for f in rng.choice(n_feat, 1500, replace=False):  # oracle-exact rows on a sample of features
        c = int(f_ctg[f])
        seq = h_seq[int(seq_off[c]):int(seq_off[c + 1])]
        ref = orc.sw_feature(seq, int(chr_start[c]), int(chr_start[c] + lens[c] - 1), int(f_s[f]), int(f_e[f]))
        a, b = int(r["row_off"][f]), int(r["row_off"][f + 1])
        assert b - a == len(ref["gc"])
        for key in ("gc", "mean", "sd", "cv"):
            assert np.array_equal(r[key][a:b].view(np.uint32), ref[key].view(np.uint32)), (f, key)
        assert np.array_equal(r["start"][a:b], ref["start"]) and np.array_equal(r["type"][a:b].astype(np.int32), ref["type"])


def test_c5_overlap_counts_1e8_by_1e8(ctx, orc):
    dev = torch.device("cuda", 0)
    n = 3_100_000_000
    seq_off, _ = synth.ctg_layout(synth.human_like_lengths(n), PIECE)
    lens = np.diff(seq_off.astype(np.int64))
    n_ctg = lens.size
    per = 100_000_000 // n_ctg
    d_st, d_sp, iv_off = synth.rg_index_torch(lens, per, seed=3, device=dev)
    plan = api.CountPlan(ctx, iv_off)
    st = torch.cuda.current_stream().cuda_stream
    n_q = 100_000_000
    L = torch.as_tensor(lens, device=dev)

    def count(seg, a, b):
        out = torch.empty(seg.numel(), dtype=torch.int32, device=dev)
        plan.count(st, d_st, d_sp, seg, a.to(torch.int32), b.to(torch.int32), out)
        torch.cuda.synchronize()
        return out.to(torch.int64)

    g = torch.Generator(device=dev)
    g.manual_seed(5)
    seg = torch.randint(0, n_ctg, (n_q,), generator=g, device=dev, dtype=torch.int32)
    Lq = L[seg.long()]
    p = torch.sort((torch.rand(3, n_q, generator=g, device=dev, dtype=torch.float64) * Lq.to(torch.float64)[None, :]).to(torch.int64) + 1, dim=0).values
    a, b, c = p[0], p[1], p[2]
    # additivity for point intervals under gams' half-open call: count(a,c) = count(a,b) + count(b,c)
    ac, ab, bc = count(seg, a, c), count(seg, a, b), count(seg, b, c)
    assert torch.equal(ac, ab + bc)
    assert int(ac.sum()) > 10 * n_q  # not vacuous: thousands of SNPs per query on average
    # the whole ctg sees every interval of the ctg
    segs = torch.arange(n_ctg, device=dev, dtype=torch.int32)
    assert torch.equal(count(segs, torch.ones(n_ctg, device=dev, dtype=torch.int64), L + 1),
                       torch.full((n_ctg,), per, device=dev, dtype=torch.int64))
    # oracle-exact on a sample, arbitrary order incl. unknown segments
    k = 300_000
    sel = torch.randint(0, n_q, (k,), generator=g, device=dev)
    s2 = seg[sel].clone()
    s2[::97] = -1
    s2[5::97] = n_ctg
    got = count(s2, a[sel], c[sel]).cpu().numpy().astype(np.uint32)
    ref = orc.count_batch(d_st.cpu().numpy().view(np.uint32), d_sp.cpu().numpy().view(np.uint32), iv_off, s2.cpu().numpy(),
                          a[sel].cpu().numpy().astype(np.uint32), c[sel].cpu().numpy().astype(np.uint32), threads=8)
    assert np.array_equal(got, ref)
    plan.close()
    torch.cuda.empty_cache()
