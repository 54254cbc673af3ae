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
