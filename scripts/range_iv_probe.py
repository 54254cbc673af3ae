"""Grouped overlap counts when the rg intervals are ranges (not SNP points): the staged kernel's two-array form."""
import numpy as np, torch
from gams_b200 import api, synth
import bench
dev = torch.device("cuda", 0)
ctx = api.Context([0])
g_off, _ = bench.layout(int(3.1e9))
lens = np.diff(g_off.astype(np.int64)); n_ctg = lens.size
per = 100_000_000 // n_ctg
stream = torch.cuda.current_stream().cuda_stream
for max_len in (1, 50, 2000):
    d_st, d_sp, iv_off = synth.rg_index_torch(lens, per, seed=7, device=dev)
    if max_len > 1:
        gen = torch.Generator(device=dev); gen.manual_seed(5)
        ln = torch.randint(1, max_len + 1, (d_st.numel(),), generator=gen, device=dev, dtype=torch.int32)
        e = (d_st + ln).view(n_ctg, per)
        d_sp = torch.sort(e, dim=1)[0].reshape(-1).contiguous()
    plan = api.CountPlan(ctx, iv_off)
    q_seg, q_s, q_e = synth.queries_torch(lens, 100_000_000, seed=11, grouped=True, device=dev)
    out = torch.empty(q_seg.numel(), dtype=torch.int32, device=dev)
    for _ in range(3): plan.count(stream, d_st, d_sp, q_seg, q_s, q_e, out, grouped=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): plan.count(stream, d_st, d_sp, q_seg, q_s, q_e, out, grouped=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    from oracle import oracle as orc
    sel = torch.randint(0, q_seg.numel(), (100_000,), device=dev)
    ref = orc.count_batch(d_st.cpu().numpy().view(np.uint32), d_sp.cpu().numpy().view(np.uint32), iv_off, q_seg[sel].cpu().numpy(),
                          q_s[sel].cpu().numpy().view(np.uint32), q_e[sel].cpu().numpy().view(np.uint32), threads=8)
    print(f"max_len {max_len}: {ms:.3f} ms per 1e8 grouped queries, bit-exact sample {np.array_equal(out[sel].cpu().numpy().view(np.uint32), ref)}", flush=True)
    plan.close(); del q_seg, q_s, q_e, out, d_st, d_sp
