"""GAMS_CUDA_TRACE=1 python scripts/trace_gz_peaks.py -> stage times of gams_cuda_wave_peaks_gz_batch on the 3.1 Gbp bench genome"""
import sys, time, zlib, os
from concurrent.futures import ThreadPoolExecutor
import ctypes as C
import numpy as np, torch
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
from gams_b200 import api, synth, _lib
import bench
n = int(3.1e9)
L = _lib.load()
ctx = api.Context([0])
seq_off, _ = bench.layout(n)
n_ctg = seq_off.size - 1
d_seq = synth.genome_torch(n, seed=20240613, device=torch.device("cuda", 0), capacity=n + 64)
h = d_seq[:n].cpu().numpy()
def comp(c):
    o = zlib.compressobj(1, zlib.DEFLATED, 31)
    return o.compress(h[int(seq_off[c]):int(seq_off[c + 1])].tobytes()) + o.flush()
with ThreadPoolExecutor(16) as ex:
    blobs = list(ex.map(comp, range(n_ctg)))
goff = np.zeros(n_ctg + 1, dtype=np.uint64); goff[1:] = np.cumsum([len(b) for b in blobs])
n_gz = int(goff[-1])
p_gz = bench.pinned(n_gz, L)
bench.as_np(p_gz, n_gz, np.uint8)[:] = np.frombuffer(b"".join(blobs), dtype=np.uint8)
woff = api.win_offsets(seq_off, 100, 10)
cap = int(woff[-1]) // 16
pk = [np.empty(cap, dtype=dt) for dt in (np.uint32, np.uint32, np.uint32, np.float32, np.int8)]
npk = C.c_uint64(); st = np.zeros(n_ctg, dtype=np.int32)
vp = lambda a: a.ctypes.data_as(C.c_void_p)
for it in range(3):
    sys.stderr.write(f"---- call {it}\n")
    t = time.perf_counter()
    api.check(L.gams_cuda_wave_peaks_gz_batch(ctx.handle, n_ctg, p_gz, vp(goff), vp(seq_off), 100, 10, 100, 3.0, 1.0, 0.2, vp(woff), cap, *[vp(a) for a in pk], C.byref(npk), vp(st)))
    sys.stderr.write(f"call {it}: {(time.perf_counter() - t) * 1e3:.2f} ms, {npk.value} peaks\n")
