"""Speed of the device gunzip on a human-scale ctg set (level-1 blobs of the bench genome): scripts/gz_probe.py [bases]"""
import sys, time, zlib
from concurrent.futures import ThreadPoolExecutor
import ctypes as C
import numpy as np, torch
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
from gams_b200 import api, synth, _lib
import bench

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else int(3.1e9)
L = _lib.load()
dev = torch.device("cuda", 0)
ctx = api.Context([0])
seq_off, _ = bench.layout(n)  # noqa
n_ctg = seq_off.size - 1
d_seq = synth.genome_torch(n, seed=20240613, device=dev, capacity=n + 64)
h = d_seq[:n].cpu().numpy()
t = time.perf_counter()
def comp(c):
    o = zlib.compressobj(1, zlib.DEFLATED, 31)
    return o.compress(h[int(seq_off[c]):int(seq_off[c + 1])].tobytes()) + o.flush()
with ThreadPoolExecutor(16) as ex:
    blobs = list(ex.map(comp, range(n_ctg)))
print(f"compressed {n_ctg} ctgs in {time.perf_counter() - t:.1f} s, ratio {n / sum(map(len, blobs)):.2f}", flush=True)
cat, goff = api.pack_blobs(blobs)
d_gz = torch.from_numpy(cat.copy()).to(dev)
d_goff = torch.from_numpy(goff.astype(np.int64)).to(dev)
d_ooff = torch.from_numpy(seq_off.astype(np.int64)).to(dev)
d_out = torch.zeros(n + 64, dtype=torch.uint8, device=dev)
d_st = torch.zeros(n_ctg, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
def run():
    api.check(L.gams_cuda_gunzip_dev(ctx.handle, 0, stream, n_ctg, d_gz.data_ptr(), d_goff.data_ptr(), d_out.data_ptr(), d_ooff.data_ptr(), d_st.data_ptr()))
import os
NOCHECK = bool(os.environ.get("GAMS_GZ_PROBE_NOCHECK"))     # timing experiments with kernels that are wrong on purpose
run(); torch.cuda.synchronize()
assert NOCHECK or (int(d_st.abs().sum()) == 0 and torch.equal(d_out[:n], d_seq[:n]))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for dec, crc in ((0, True), (0, False), (1, False)):
    ctx.set_gunzip_options(dec, crc)
    run(); torch.cuda.synchronize()
    assert NOCHECK or (int(d_st.abs().sum()) == 0 and torch.equal(d_out[:n], d_seq[:n]))
    e0.record()
    for _ in range(3): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"gunzip_dev decoder {dec} crc {crc}: {ms:.2f} ms per {n / 1e9:.2f} Gbp = {n / ms / 1e6:.1f} Gbp/s out, {cat.size / ms / 1e6:.1f} GB/s in; {"NOT CHECKED" if NOCHECK else "bit-exact"}")
