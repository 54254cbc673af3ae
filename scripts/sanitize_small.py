"""Small calls through every kernel of libgams_cuda.so, for compute-sanitizer (memcheck): `PYTHONPATH=. compute-sanitizer
--tool memcheck python scripts/sanitize_small.py`.  Results are checked against the oracle as well."""
import numpy as np

from gams_b200 import api, synth
from oracle import oracle as orc

g = synth.genome(260_000, seed=5)
g[1000:1060] = ord("N")
seqs = [g[:120_001], g[120_001:120_001 + 1500], g[130_000:]]
cat, off = orc.pack_ctgs(seqs)
with api.Context([0]) as ctx:
    for size, step, lag, infl in ((100, 10, 100, 1.0), (100, 10, 100, 0.5), (97, 13, 31, 1.0), (16, 16, 8, 1.0), (64, 8, 100, 1.0)):
        woff, cnt, gc, sig = api.wave_batch(ctx, cat, off, size, step, lag, 3.0, infl)
        rw, rc, rg, rs = orc.wave_batch(cat, off, size, step, lag, 3.0, infl, threads=4)
        assert np.array_equal(cnt, rc) and np.array_equal(gc.view(np.uint32), rg.view(np.uint32)) and np.array_equal(sig, rs)
    pk = api.wave_peaks_batch(ctx, cat, off, 100, 10, 100, 3.0, 1.0, 0.2, cap=1)
    assert pk["ctg"].size > 0
    c, s, e = api.gen_ctgs(ctx, cat, off, piece=50_000, fill=50, min_=1000)
    assert list(zip(c.tolist(), s.tolist(), e.tolist())) == [(k, a, b) for k, q in enumerate(seqs) for a, b in orc.gen_ctgs(q, piece=50_000, fill=50, min_=1000)]
    lens = [len(q) for q in seqs]
    st, sp, ioff = synth.rg_index(lens, [4000, 30, 30000], seed=2, max_len=50)
    for grouped in (True, False):
        seg, qs, qe = synth.queries(lens, 40_000, seed=3, grouped=grouped)
        got = api.count_overlaps(ctx, st, sp, ioff, seg, qs, qe)
        assert np.array_equal(got, orc.count_batch(st, sp, ioff, seg, qs, qe))
    import torch
    dev = torch.device("cuda", 0)
    t = lambda a: torch.as_tensor(a.view(np.int32) if a.dtype != np.int32 else a, device=dev)
    seg, qs, qe = synth.queries(lens, 40_000, seed=4, grouped=True)
    plan = api.CountPlan(ctx, ioff)
    out = torch.empty(seg.size, dtype=torch.int32, device=dev)
    plan.count(torch.cuda.current_stream().cuda_stream, t(st), t(sp), t(seg), t(qs), t(qe), out, grouped=True)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), orc.count_batch(st, sp, ioff, seg, qs, qe))
    plan.close()
    cs = np.array([1, 1, 1], dtype=np.int32)
    res = api.sw_batch(ctx, cat, off, cs, [0, 2, 2], [5000, 100, 60_000], [5100, 100, 60_400])
    assert res["gc"].size > 0
    cn = api.gc_ranges(ctx, cat, off, [0, 1, 2], [0, 10, 77], [100, 1500, 130_000])
    assert cn[0] == int(np.isin(seqs[0][:100], np.frombuffer(b"cgCG", dtype=np.uint8)).sum())
    lk = api.peak_link_batch(ctx, cat, off, cs, [0, 0, 2], [10, 500, 7], [20, 600, 7], [1, -1, 1])
    assert lk["gc"].size == 3
    pr = api.anno_coverage(ctx, [5, 100], [50, 200], [0, 2], [0, 0, 1], [1, 1, 1], [1000, 1000, 10], [1, 40, 1], [300, 60, 5])
    assert pr.size == 3
    # rg counts along sliding windows and the index build (non-point intervals, an empty segment)
    st2, sp2, ioff2 = synth.rg_index(lens, [4000, 0, 30000], seed=9, max_len=50)
    for size, step in ((100, 10), (50, 7), (3000, 1500)):
        woff = np.concatenate(([0], np.cumsum([orc.nwin(n, size, step) for n in lens]))).astype(np.uint64)
        seg = np.concatenate([np.full(int(woff[c + 1] - woff[c]), c, np.int32) for c in range(3)])
        k = np.concatenate([np.arange(int(woff[c + 1] - woff[c])) for c in range(3)])
        ws = (1 + k * step).astype(np.uint32)
        got = api.count_windows(ctx, cs, woff, size, step, st2, sp2, ioff2)
        assert np.array_equal(got, orc.count_batch(st2, sp2, ioff2, seg, ws, ws + np.uint32(size - 1)))
    rng = np.random.default_rng(1)
    a = rng.integers(1, 100_000, 50_000).astype(np.uint32)
    b = a + rng.integers(1, 30, 50_000).astype(np.uint32)
    ix = api.build_idx_rg(ctx, a, b, np.array([0, 20_000, 20_000, 50_000], np.uint64))
    lp = orc.lapper_new([(int(x), int(y), "") for x, y in zip(a[:20_000], b[:20_000])])
    assert np.array_equal(ix["iv_start"][:20_000], lp["start"]) and np.array_equal(ix["stops"][:20_000], lp["stops"])
    assert ix["blob"][:int(ix["blob_off"][1])].tobytes() == orc.lapper_to_bincode(lp)
print("sanitize_small ok")
