"""Small calls through the round-2 kernels (part-parallel gunzip, chained gzip paths, segmented influence walk), checked against zlib and
the oracle: `python scripts/sanitize_gz.py` (written for `compute-sanitizer --tool memcheck`, which this pool no longer runs)."""
import os
import sys
import zlib

sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))

import numpy as np

os.environ.setdefault("GAMS_CUDA_ZSEG", "300")
from gams_b200 import api, synth
from oracle import oracle as orc


def gzb(data, level=1, strategy=0):
    o = zlib.compressobj(level, zlib.DEFLATED, 31, 8, strategy)
    return o.compress(data) + o.flush()


rng = np.random.default_rng(5)
dna = synth.genome(400_000, seed=6).tobytes()
plain = [dna[:150_000], dna[150_000:150_900], b"", b"A", b"N" * 20_000, dna[200_000:330_000], bytes(rng.integers(0, 256, 9000, dtype=np.uint8)),
         (b"ACGT" * 5 + b"G") * 900]
blobs = [gzb(p, level=(1, 6, 9)[i % 3]) for i, p in enumerate(plain)]
blobs[6] = gzb(plain[6], level=0)
blobs[7] = gzb(plain[7], strategy=zlib.Z_FIXED)
with api.Context([0]) as ctx:
    for dec in (0, 1):
        ctx.set_gunzip_options(dec, True)
        out, off, st = api.gunzip_batch(ctx, *api.pack_blobs(blobs))
        assert not st.any(), st
        assert out.tobytes() == b"".join(plain)
    ctx.set_gunzip_options(0, True)
    bad = bytearray(blobs[0]); bad[len(bad) // 2] ^= 0x21
    _, _, st = api.gunzip_batch(ctx, *api.pack_blobs([bytes(bad), blobs[1]]), raise_on_error=False)
    assert st[0] != 0 and st[1] == 0
    seqs = [np.frombuffer(plain[0], dtype=np.uint8), np.frombuffer(plain[1], dtype=np.uint8), np.frombuffer(plain[5], dtype=np.uint8)]
    cat, soff = orc.pack_ctgs(seqs)
    gcat, goff = api.pack_blobs([blobs[0], blobs[1], blobs[5]])
    for infl in (1.0, 0.5):
        woff, cnt, gc, sig, st = api.wave_gz_batch(ctx, gcat, goff, soff, 100, 10, 50, 3.0, infl)
        rw, rc, rg, rs = orc.wave_batch(cat, soff, 100, 10, 50, 3.0, infl, threads=4)
        assert not st.any() and np.array_equal(cnt, rc) and np.array_equal(gc.view(np.uint32), rg.view(np.uint32)) and np.array_equal(sig, rs)
    ref = api.wave_peaks_batch(ctx, cat, soff, 100, 10, 50, 3.0, 1.0, 0.2)
    got, st = api.wave_peaks_gz_batch(ctx, gcat, goff, soff, 100, 10, 50, 3.0, 1.0, 0.2)
    assert not st.any() and all(np.array_equal(got[k], ref[k]) for k in ("ctg", "first", "last", "signal"))
print("sanitize_gz ok")
