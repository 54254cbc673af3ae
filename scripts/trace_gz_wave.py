"""python scripts/trace_gz_wave.py [pieces ...] -> ms per call of gams_cuda_wave_gz_batch (gzip blobs in pinned host memory -> all
windows back in pinned host memory) on the 3.1 Gbp bench genome, for each number of decoder pieces (GAMS_CUDA_GZ_PIECES);
GAMS_CUDA_TRACE=1 adds the stage times of every call."""
import sys, time, zlib, os
from concurrent.futures import ThreadPoolExecutor
import ctypes as C
import numpy as np, torch
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
from gams_b200 import api, synth, _lib
import bench
pieces = [a for a in sys.argv[1:]] or ["1", "2", "3", "4"]
n = int(float(os.environ.get("BASES", "3.1e9")))
L = _lib.load()
ctx = api.Context([0])
seq_off, _ = bench.layout(n)
n_ctg = seq_off.size - 1
d_seq = synth.genome_torch(n, seed=20240613, device=torch.device("cuda", 0), capacity=n + 64)
h = d_seq[:n].cpu().numpy()
del d_seq
torch.cuda.empty_cache()
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
nw = int(woff[-1])
p_gc, p_sig = bench.pinned(nw * 4, L), bench.pinned(nw, L)
st = np.zeros(n_ctg, dtype=np.int32)
vp = lambda a: a.ctypes.data_as(C.c_void_p)
ref = None
for P in pieces:
    os.environ["GAMS_CUDA_GZ_PIECES"] = P
    ts = []
    for it in range(4):
        t = time.perf_counter()
        api.check(L.gams_cuda_wave_gz_batch(ctx.handle, n_ctg, p_gz, vp(goff), vp(seq_off), 100, 10, 100, C.c_float(3.0), C.c_float(1.0), vp(woff), None, p_gc, p_sig, vp(st)))
        ts.append((time.perf_counter() - t) * 1e3)
    sig = bench.as_np(p_sig, nw, np.int8).copy()
    gc = bench.as_np(p_gc, nw, np.uint32).copy()
    if ref is None: ref = (gc, sig)
    same = bool(np.array_equal(ref[0], gc) and np.array_equal(ref[1], sig))
    print(f"pieces {P}: " + " ".join(f"{t:.2f}" for t in ts) + f" ms; status ok {not st.any()}; equals first {same}", flush=True)
