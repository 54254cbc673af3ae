"""GAMS_CUDA_TRACE=1 python scripts/trace_wave.py -> stage times of gams_cuda_wave_batch (bases in pinned host memory, all windows back)"""
import sys, time
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
p_seq = bench.pinned(n + 64, L)
bench.as_np(p_seq, n, np.uint8)[:] = d_seq[:n].cpu().numpy()
woff = api.win_offsets(seq_off, 100, 10)
nw = int(woff[-1])
p_gc, p_sig = bench.pinned(nw * 4, L), bench.pinned(nw, L)
vp = lambda a: a.ctypes.data_as(C.c_void_p)
for it in range(3):
    sys.stderr.write(f"---- call {it}\n")
    t = time.perf_counter()
    api.check(L.gams_cuda_wave_batch(ctx.handle, n_ctg, p_seq, vp(seq_off), 100, 10, 100, C.c_float(3.0), C.c_float(1.0), vp(woff), None, p_gc, p_sig))
    sys.stderr.write(f"call {it}: {(time.perf_counter() - t) * 1e3:.2f} ms\n")
