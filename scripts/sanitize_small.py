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
print("sanitize_small ok")
