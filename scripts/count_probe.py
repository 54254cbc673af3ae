"""Arbitrary-order overlap counts at C5 scale (1e8 queries x 1e8 SNP intervals on the 3.1 Gbp ctg layout), device-resident:
ms per call and a bit-exact sample check against the oracle.  GAMS_CUDA_LIB selects a variant library."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gams_b200 import api, synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

n_q = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
seq_off, _ = synth.ctg_layout(synth.human_like_lengths(3_100_000_000), 500_000)
lens = np.diff(seq_off.astype(np.int64))
d_st, d_sp, iv_off = synth.rg_index_torch(lens, 100_000_000 // lens.size, seed=7, device=dev)
ctx = api.Context([0])
plan = api.CountPlan(ctx, iv_off)
q_seg, q_s, q_e = synth.queries_torch(lens, n_q, seed=11, grouped=False, device=dev)
out = torch.empty(n_q, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    plan.count(st, d_st, d_sp, q_seg, q_s, q_e, out, grouped=False)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    plan.count(st, d_st, d_sp, q_seg, q_s, q_e, out, grouped=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
sel = torch.randint(0, n_q, (200_000,), device=dev)
ref = orc.count_batch(d_st.cpu().numpy().view(np.uint32), d_sp.cpu().numpy().view(np.uint32), iv_off, q_seg[sel].cpu().numpy(),
                      q_s[sel].cpu().numpy().view(np.uint32), q_e[sel].cpu().numpy().view(np.uint32), threads=os.cpu_count() or 1)
ok = bool(np.array_equal(out[sel].cpu().numpy().view(np.uint32), ref))
alg = n_q * 16 + int(iv_off[-1]) * 8
print(f"{os.environ.get('GAMS_CUDA_LIB', 'default').split('/')[-1]:24s} any-order {ms:7.4f} ms  {alg / ms / 1e6 / 6557.1:.3f} of HBM peak  bit-exact sample {ok}")
