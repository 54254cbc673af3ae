"""Seeded synthetic inputs of the shapes BASELINE.json names: genomes with a slowly
varying GC bias, soft-masked (lowercase) runs and N gaps; ctg cuts per `gams gen`; rg intervals
and query ranges.  numpy versions for tests; torch-on-device versions for the full-size bench."""
from __future__ import annotations

import numpy as np

_UP = np.frombuffer(b"ATCG", dtype=np.uint8)  # index bit 1 = is GC


def genome(n_bases: int, seed: int = 20240613, lower_frac=0.02, n_gap_every=250_000) -> np.ndarray:
    """ASCII bases with GC target doing a random walk in [0.30, 0.65], re-drawn every 2-20 kbp."""
    rng = np.random.default_rng(seed)
    out = np.empty(n_bases, dtype=np.uint8)
    pos, gc = 0, 0.45
    while pos < n_bases:
        seg = int(rng.integers(2_000, 20_001))
        seg = min(seg, n_bases - pos)
        gc = float(np.clip(gc + rng.normal(0, 0.05), 0.30, 0.65))
        is_gc = rng.random(seg) < gc
        which = rng.integers(0, 2, seg)
        out[pos:pos + seg] = _UP[(is_gc.astype(np.int64) << 1) | which]
        pos += seg
    # soft-masked runs
    n_runs = max(1, int(n_bases * lower_frac / 300))
    for s in rng.integers(0, max(1, n_bases - 600), n_runs):
        ln = int(rng.integers(50, 600))
        out[s:s + ln] |= 0x20
    # N gaps: short ones survive `--fill 50`, a few long ones split ctgs (gen.rs:100-104)
    for s in range(n_gap_every, n_bases - 1000, n_gap_every):
        ln = int(rng.integers(1, 50)) if rng.random() < 0.8 else int(rng.integers(50, 400))
        out[s:s + ln] = ord("N") if rng.random() < 0.7 else ord("n")
    return out


def cut_pieces(length: int, piece: int = 500_000):
    """gen.rs:107-126 on one fully valid region [1, length]: fixed pieces, remainder merged into the last."""
    out, pos = [], 1
    while length - pos + 1 > piece:
        out.append([pos, pos + piece - 1])
        pos += piece
    if not out:
        out.append([pos, length])
    else:
        out[-1][1] = length
    return [tuple(x) for x in out]


def ctg_layout(chr_lengths, piece: int = 500_000):
    """→ seq_off (uint64, n_ctg+1) for chromosomes laid back to back, plus (chr, chr_start, chr_end) per ctg."""
    off, ctgs, base = [0], [], 0
    for ci, ln in enumerate(chr_lengths):
        for s, e in cut_pieces(int(ln), piece):
            ctgs.append((ci, s, e))
            off.append(base + e)
        base += int(ln)
    return np.asarray(off, dtype=np.uint64), ctgs


HUMAN_CHR_MBP = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47,
                 51, 156, 57]  # 24 chromosomes, ~3.1 Gbp


def human_like_lengths(total_bases: int = 3_100_000_000):
    w = np.asarray(HUMAN_CHR_MBP, dtype=np.float64)
    ln = np.floor(w / w.sum() * total_bases).astype(np.int64)
    ln[-1] += total_bases - ln.sum()
    return ln


def rg_index(ctg_lengths, n_per_ctg, seed=7, max_len=1):
    """Sorted endpoint arrays of random rg intervals per ctg (coords 1-based within the ctg's chr range
    starting at 1).  max_len == 1 gives SNP points chr:p → [p, p+1) (redis.rs:291-295)."""
    rng = np.random.default_rng(seed)
    starts, stops, off = [], [], [0]
    for ln, k in zip(ctg_lengths, n_per_ctg):
        s = rng.integers(1, int(ln) + 1, int(k)).astype(np.uint32)
        e = s + (rng.integers(0, max_len, int(k)).astype(np.uint32) if max_len > 1 else 0)  # inclusive end
        starts.append(np.sort(s))
        stops.append(np.sort(e + 1))
        off.append(off[-1] + int(k))
    return (np.concatenate(starts) if starts else np.zeros(0, np.uint32),
            np.concatenate(stops) if stops else np.zeros(0, np.uint32), np.asarray(off, dtype=np.uint64))


def queries(ctg_lengths, n_q, seed=11, max_len=10_000, grouped=True):
    """n_q query ranges: ctg ∝ length, uniform start, length U[1, max_len] → (q_seg, q_start, q_end_inclusive)."""
    rng = np.random.default_rng(seed)
    ln = np.asarray(ctg_lengths, dtype=np.float64)
    seg = rng.choice(len(ln), size=n_q, p=ln / ln.sum()).astype(np.int32)
    if grouped:
        seg.sort()
    L = np.asarray(ctg_lengths, dtype=np.int64)[seg]
    start = (rng.random(n_q) * L).astype(np.int64) + 1
    end = np.minimum(start + rng.integers(0, max_len, n_q), L)
    return seg, start.astype(np.uint32), end.astype(np.uint32)


# ------------------------------------------------------------------------------------------
# torch-on-device generators for the full-size configs (3.1 Gbp would take minutes in numpy)
# ------------------------------------------------------------------------------------------
def genome_torch(n_bases: int, seed: int = 20240613, device="cuda", capacity: int | None = None, block=8192):
    """uint8 ASCII genome on `device`: per-8 kbp-block GC target smoothed over 4 blocks (range about
    0.30-0.65), 2 % soft-masked 512-base runs, an N gap of 1-49 bases every 250 kbp.  `capacity`
    pads the allocation (zero bytes) for the kernels' 16-byte loads."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cap = max(capacity or 0, n_bases)
    out = torch.zeros(cap, dtype=torch.uint8, device=device)
    n_blk = (n_bases + block - 1) // block
    t = torch.rand(n_blk + 3, generator=g, device=device) * 0.35 + 0.30
    target = (t[:-3] + t[1:-2] + t[2:-1] + t[3:]) / 4
    lut = torch.tensor(list(b"ATCG"), dtype=torch.uint8, device=device)
    chunk = 1 << 27
    for s in range(0, n_bases, chunk):
        e = min(n_bases, s + chunk)
        idx = torch.arange(s, e, device=device)
        r = torch.rand(e - s, generator=g, device=device)
        is_gc = (r < target[idx // block]).to(torch.int64)
        which = (torch.rand(e - s, generator=g, device=device) < 0.5).to(torch.int64)
        out[s:e] = lut[is_gc * 2 + which]
        del idx, r, is_gc, which
    n_runs = max(1, int(n_bases * 0.02 / 512))
    starts = torch.randint(0, max(1, n_bases - 512), (n_runs,), generator=g, device=device)
    pos = (starts[:, None] + torch.arange(512, device=device)[None, :]).reshape(-1)
    out[pos] |= 0x20
    gaps = torch.arange(250_000, max(250_000, n_bases - 1000), 250_000, device=device)
    if gaps.numel():
        glen = torch.randint(1, 50, (gaps.numel(),), generator=g, device=device)
        off = torch.arange(49, device=device)[None, :]
        gp = (gaps[:, None] + off)[off < glen[:, None]]
        out[gp] = ord("N")
    return out


def rg_index_torch(ctg_lengths, n_per_ctg: int, seed=7, device="cuda"):
    """Uniform SNP points per ctg → sorted starts, stops = starts + 1, iv_off (numpy)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    L = torch.as_tensor(np.asarray(ctg_lengths, dtype=np.int64), device=device)
    n_ctg = L.numel()
    r = torch.rand(n_ctg, n_per_ctg, generator=g, device=device, dtype=torch.float64)
    p = (r * L[:, None].to(torch.float64)).to(torch.int64) + 1
    p, _ = torch.sort(p, dim=1)
    starts = p.reshape(-1).to(torch.int32)  # values < 2^31; reinterpret as uint32 on the C side
    stops = (p + 1).reshape(-1).to(torch.int32)
    iv_off = (np.arange(n_ctg + 1, dtype=np.uint64) * np.uint64(n_per_ctg))
    return starts, stops, iv_off


def queries_torch(ctg_lengths, n_q: int, seed=11, max_len=10_000, grouped=True, device="cuda"):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    L = torch.as_tensor(np.asarray(ctg_lengths, dtype=np.int64), device=device)
    seg = torch.multinomial(L.to(torch.float32), n_q, replacement=True, generator=g) if n_q <= (1 << 24) else \
        torch.searchsorted(torch.cumsum(L, 0).to(torch.float64) / float(L.sum()),
                           torch.rand(n_q, generator=g, device=device, dtype=torch.float64)).clamp_(max=L.numel() - 1)
    if grouped:
        seg, _ = torch.sort(seg)
    Lq = L[seg]
    start = (torch.rand(n_q, generator=g, device=device, dtype=torch.float64) * Lq.to(torch.float64)).to(torch.int64) + 1
    end = torch.minimum(start + torch.randint(0, max_len, (n_q,), generator=g, device=device), Lq)
    return seg.to(torch.int32), start.to(torch.int32), end.to(torch.int32)
