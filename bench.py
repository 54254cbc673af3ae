#!/usr/bin/env python3
"""bench.py — the gams hot path on B200: GC-wave Gbp/s (headline) and rg overlap counts/s.

One "step" = `passes_per_step` passes of the hot path over the synthetic batch (so many that K steps keep the device
busy for >= 0.6 s: one pass takes under a millisecond, and a timed region of a few ms sees neither sustained clocks nor
more than one NVML sample):
  wave  : the fused GC-wave kernel (sequence bytes -> per-window GC + z-score signals) over a human-scale
          synthetic genome (3.1 Gbp, 24 chromosomes, `gen --piece 500000` ctgs, wave 100/10, lag 100,
          threshold 3, influence 1) resident in HBM  -> `value` (Gbp/s);
          the same through the host C-ABI call gams_cuda_wave_batch with pinned HOST buffers, H2D of the
          sequence and D2H of gc+signal inside the timed region                      -> `e2e`;
          and through gams_cuda_wave_gz_batch, fed with the gzip blobs Redis holds (seq:{ctg}), which are
          inflated on the device                                                      -> `e2e_gz`.
  count : batched `locate --count`: 1e8 query ranges against 1e8 SNP intervals on the same ctg layout,
          device-resident and end-to-end, reported under "overlap" in the same JSON line.

Multi-GPU: one process per GPU (torchrun), ctg-sharded, no data-path collective; torch.distributed is
used only for the barrier and the max-over-ranks of the device time.  scaling "weak": every rank
processes its own 3.1 Gbp genome (a cohort of N genomes); --scaling strong splits one genome by ctg.

`--impl reference` times the reference's CPU algorithm (the oracle port: the reference is Rust and
cannot be built in this image) on the box's host cores on a bounded sample of the same workload, the
per-ctg gunzip of Conn::get_seq (redis.rs:142-161) included, as `gams wave` pays it.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SIZE, STEP, LAG, THRESHOLD, INFLUENCE = 100, 10, 100, 3.0, 1.0
PIECE = 500_000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bases", type=float, default=3.1e9, help="bases per genome (per rank under weak scaling)")
    ap.add_argument("--queries", type=float, default=1e8, help="overlap queries per genome")
    ap.add_argument("--intervals", type=float, default=1e8, help="rg intervals per genome")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--influence", type=float, default=INFLUENCE)
    ap.add_argument("--skip-count", action="store_true")
    ap.add_argument("--skip-sw", action="store_true")
    ap.add_argument("--sw-features", type=float, default=1e6, help="features of the sw section (configs[2])")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-gz", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-frontend", action="store_true", help="skip the parse / locate-text / index-build records")
    ap.add_argument("--skip-multi", action="store_true", help="skip the strong-scaling and single-process records (N > 1)")
    ap.add_argument("--passes", type=int, default=0, help="passes over the genome per step (0: as many as make K steps >= 0.6 s)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ helpers
def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", d
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", {}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML every 5 ms; the recipe's
    nvidia-smi query is the fallback).  Samples carry a timestamp so only those inside [mark_start, mark_end]
    are summarised."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.stop, self.th, self.how = gpu_index, [], False, None, None
        self.t0 = self.t1 = None

    def _nvml_sample(self):
        nv, h = self.nv, self.h
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        self.rows.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), float(self.mx),
                          [k for k, b in bits.items() if r & b]))

    def _nvml_loop(self):
        while not self.stop:
            try:
                self._nvml_sample()
            except Exception:
                pass
            time.sleep(0.004)

    def _smi_loop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                 "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
        for line in proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((time.perf_counter(), float(r[1]), float(r[2]),
                                  [n for n, v in zip(names, r[5:9]) if v.lower().startswith("active")]))
            except Exception:
                pass
            if self.stop:
                break
        proc.terminate()

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self._nvml_sample()
            self.how, target = "nvml", self._nvml_loop
        except Exception:
            self.how, target = "nvidia-smi", self._smi_loop
        self.th = threading.Thread(target=target, daemon=True)
        self.th.start()
        return self

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        if self.how == "nvml":  # the GPU is still busy here (the caller has not synchronised yet)
            try:
                self._nvml_sample()
            except Exception:
                pass
        self.t1 = time.perf_counter()

    def __exit__(self, *a):
        self.stop = True
        if self.th:
            self.th.join(timeout=2)

    def summary(self):
        inside = [r for r in self.rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1]
        rows = inside if len(inside) >= 3 else self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "how": self.how}
        reasons = sorted({x for r in rows for x in r[3]})
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": float(max(r[2] for r in rows)),
                "reasons": reasons, "samples": len(rows), "samples_in_timed_region": len(inside), "how": self.how}


def ncu_traffic(kernel: str):
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/traffic.json), or None.  An entry names the
    source files of its kernel with their sha1 at capture time: if one of them has changed since, the number is stale and
    None is returned instead (the capture is repeated with scripts/round_profile.sh)."""
    import hashlib
    try:
        d = json.loads((ROOT / "profiles" / "traffic.json").read_text())
        ent = d[kernel]
        for f, h in (ent.get("source_sha1") or {}).items():
            if hashlib.sha1((ROOT / "gams_b200" / "csrc" / f).read_bytes()).hexdigest() != h:
                return None
        return float(ent["dram_bytes_per_launch"])
    except Exception:
        return None


def layout(n_bases: int):
    from gams_b200 import synth
    seq_off, ctgs = synth.ctg_layout(synth.human_like_lengths(n_bases), PIECE)
    return seq_off, ctgs


def pinned(nbytes: int, L):
    p = L.gams_cuda_host_alloc(max(int(nbytes), 1))
    if not p:
        raise MemoryError("gams_cuda_host_alloc failed")
    return p


def as_np(ptr, n, dtype):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(int(n),))


# ------------------------------------------------------------------------------- reference (CPU) arm
def cpu_wave_rate(seq_host: np.ndarray, seq_off: np.ndarray, threads: int, target_s: float, influence: float):
    """Oracle (faithful port of wave.rs:138-155) on a bounded sample of ctgs → (Gbp/s, sample description)."""
    from oracle import oracle as orc
    n_ctg = seq_off.size - 1
    cal = min(n_ctg, max(threads, 1))
    sub_off = (seq_off[:cal + 1] - seq_off[0]).astype(np.uint64)
    t0 = time.perf_counter()
    orc.wave_batch(seq_host, sub_off, SIZE, STEP, LAG, THRESHOLD, influence, threads=threads)
    dt = time.perf_counter() - t0
    k = int(min(n_ctg, max(cal, cal * target_s / max(dt, 1e-3))))
    k = max(threads, (k // max(threads, 1)) * max(threads, 1)) if k >= threads else k
    k = min(k, n_ctg)
    sub_off = (seq_off[:k + 1] - seq_off[0]).astype(np.uint64)
    t0 = time.perf_counter()
    orc.wave_batch(seq_host, sub_off, SIZE, STEP, LAG, THRESHOLD, influence, threads=threads)
    dt = time.perf_counter() - t0
    bases = int(sub_off[-1])
    return bases / dt / 1e9, f"first {k} ctgs ({bases / 1e6:.1f} Mbp) of the workload, wave {SIZE}/{STEP} lag {LAG}, {dt:.1f} s wall"


def gz_blobs(h_seq: np.ndarray, seq_off: np.ndarray, threads: int):
    """The seq:{ctg} values as encode_gz writes them (redis.rs:149-154: gzip, Compression::fast() = level 1), one per ctg."""
    import zlib
    from concurrent.futures import ThreadPoolExecutor

    def comp(c):
        o = zlib.compressobj(1, zlib.DEFLATED, 31)
        return o.compress(h_seq[int(seq_off[c]):int(seq_off[c + 1])].tobytes()) + o.flush()

    with ThreadPoolExecutor(max(1, threads)) as ex:
        return list(ex.map(comp, range(seq_off.size - 1)))


def cpu_gunzip_into(blobs, out: np.ndarray, seq_off: np.ndarray, threads: int):
    """decode_gz per ctg (redis.rs:156-161) with zlib on `threads` host threads (zlib releases the GIL), into `out`."""
    import zlib
    from concurrent.futures import ThreadPoolExecutor

    def dec(c):
        out[int(seq_off[c]):int(seq_off[c + 1])] = np.frombuffer(zlib.decompress(blobs[c], 31), dtype=np.uint8)

    with ThreadPoolExecutor(max(1, threads)) as ex:
        list(ex.map(dec, range(len(blobs))))


def run_reference(args, rank):
    """bench.py --impl reference: the reference's CPU path (oracle port) with all host threads: per step every ctg of the
    sample is gunzipped (Conn::get_seq -> decode_gz, redis.rs:142-161: zlib here, flate2 there) and goes through wave."""
    if rank != 0:
        return
    from gams_b200 import synth
    from oracle import oracle as orc
    orc.lib()
    cores = os.cpu_count() or 1
    seq_off, _ = layout(int(args.bases))
    # bounded sample: enough ctgs for a few seconds per step, generated on the host
    n_cal = min(seq_off.size - 1, cores)
    g = synth.genome(int(seq_off[n_cal]), seed=20240613)
    rate, _ = cpu_wave_rate(g, seq_off[:n_cal + 1], cores, 1.0, args.influence)
    per_step_s = min(6.0, 150.0 / max(args.steps + args.warmup, 1))
    k = int(min(seq_off.size - 1, max(n_cal, rate * 1e9 * per_step_s / PIECE), max(n_cal, 600e6 / PIECE)))  # <= 600 Mbp
    g = synth.genome(int(seq_off[k]), seed=20240613)
    sub_off = seq_off[:k + 1].astype(np.uint64)
    blobs = gz_blobs(g, sub_off, cores)
    buf = np.empty_like(g)
    t_gz = 0.0

    def step():
        nonlocal t_gz
        t = time.perf_counter()
        cpu_gunzip_into(blobs, buf, sub_off, cores)
        t_gz += time.perf_counter() - t
        orc.wave_batch(buf, sub_off, SIZE, STEP, LAG, THRESHOLD, args.influence, threads=cores)

    for _ in range(args.warmup):
        step()
    t_gz = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    assert np.array_equal(buf, g)
    bases = int(sub_off[-1])
    val = bases * args.steps / dt / 1e9
    sample = f"{k} ctgs ({bases / 1e6:.1f} Mbp) per step of the {args.bases / 1e9:.2f} Gbp workload, gunzip of the seq blobs included"
    line = {
        "impl": "reference", "metric": "gc_wave_throughput", "value": val, "unit": "Gbp/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "u8->u32 counts, f32 statistics", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": val, "unit": "Gbp/s", "cores": cores, "kind": "port", "sample": sample,
                         "gunzip_share_of_time": t_gz / dt},
        "e2e": {"value": val, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is Rust (no cargo in this image): timed the faithful C port oracle/gams_oracle.c, "
                "one worker per ctg over all host threads (the reference's --parallel model); the seq blobs are in memory "
                "(no Redis GET) and are inflated per step with zlib, as Conn::get_seq does with flate2",
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"human-scale synthetic genome ({args.bases / 1e9:.2f} Gbp, 24 chromosomes, gen --piece {PIECE}) "
                        f"GC-wave --size {SIZE} --step {STEP} --lag {LAG} --threshold {THRESHOLD} --influence {args.influence}",
            "baseline_config": "configs[3]", "size": SIZE, "step": STEP, "lag": LAG, "threshold": THRESHOLD,
            "influence": args.influence, "piece": PIECE, "bases_per_genome": int(args.bases),
            "l2": "inputs larger than L2 (3.1 GB sequence + 1.55 GB outputs per step vs 126 MB L2)",
            "parallelism": f"ctg-sharded x{args.gpus}, {args.scaling} scaling, no collective"}


# ------------------------------------------------------------------------------------------- our arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from gams_b200 import _lib, api, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; gams_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        gloo = dist.new_group(backend="gloo")   # a barrier that keeps the GPUs idle (an NCCL barrier is a kernel on every device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    L = _lib.load()
    ctx = api.Context([local])
    hbm_peak, peak_src, peaks_json = peaks()

    # ---- workload: this rank's ctgs ------------------------------------------------------------
    g_off, g_ctgs = layout(int(args.bases))
    if args.scaling == "strong" and world > 1:
        from gams_b200 import shard
        _, _, seq_off = shard.shard_for_rank(g_off, rank, world)
    else:
        seq_off = g_off
    n_ctg = seq_off.size - 1
    plan = api.WavePlan(ctx, seq_off, SIZE, STEP, LAG)
    n_bases, n_win = plan.total_bases, plan.total_windows
    d_seq = synth.genome_torch(n_bases, seed=20240613 + rank, device=dev, capacity=plan.seq_capacity)
    d_gc = torch.empty(n_win, dtype=torch.float32, device=dev)
    d_sig = torch.empty(n_win, dtype=torch.int8, device=dev)
    d_filt = torch.empty(n_win, dtype=torch.float32, device=dev) if args.influence != 1.0 else None
    stream = torch.cuda.current_stream().cuda_stream

    fused = plan.fused(THRESHOLD, args.influence)

    def one_pass():
        # one pass of the hot path over the rank's genome: the fused kernel when the statistics are the defaults
        # (influence 1), else Kernel A followed by Kernel B
        plan.wave(stream, d_seq, THRESHOLD, args.influence, d_gc, d_sig, None, d_filt)

    # passes per step: so many that the K timed steps keep the device busy for >= 0.6 s (sustained clocks, a clock record of
    # more than one sample); every rank uses the same number
    one_pass()
    torch.cuda.synchronize()
    c0e, c1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0e.record()
    for _ in range(3):
        one_pass()
    c1e.record()
    torch.cuda.synchronize()
    ms_cal = max_over_ranks(c0e.elapsed_time(c1e) / 3)
    passes = args.passes if args.passes > 0 else int(min(256, max(1, np.ceil(600.0 / max(args.steps, 1) / max(ms_cal, 1e-3)))))

    def step():
        for _ in range(passes):
            one_pass()

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    with ClockSampler(local) as clk:
        for _ in range(args.warmup):
            step()
        l0 = plan.launches
        barrier()
        clk.mark_start()
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        t_start.record()
        for k in range(args.steps):
            ev[k][0].record()
            step()
            ev[k][1].record()
        t_end.record()
        clk.mark_end()
        barrier()
    launches = plan.launches - l0
    z_uncertain, _ = ctx.zscore_stats(reset=True)
    z_uncertain /= max(1, (args.steps + args.warmup) * passes + 4)
    ms_total = max_over_ranks(t_start.elapsed_time(t_end))
    ms_step = float(np.mean([e[0].elapsed_time(e[1]) for e in ev])) / passes     # one pass = one launch of the fused kernel
    total_bases = sum_over_ranks(float(n_bases))
    value = total_bases * args.steps * passes / (ms_total * 1e-3) / 1e9
    clocks = clk.summary()

    # the two-kernel path (Kernel A, Kernel B) timed separately, outside the headline region, for the per-kernel table
    n2 = max(1, min(args.steps, 10))
    ev2 = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n2)]
    plan.gc(stream, d_seq, d_gc)
    plan.signal(stream, d_gc, THRESHOLD, args.influence, d_sig, d_filt)
    for k in range(n2):
        ev2[k][0].record()
        plan.gc(stream, d_seq, d_gc)
        ev2[k][1].record()
        plan.signal(stream, d_gc, THRESHOLD, args.influence, d_sig, d_filt)
        ev2[k][2].record()
    torch.cuda.synchronize()
    ms_gc = float(np.mean([e[0].elapsed_time(e[1]) for e in ev2]))
    ms_z = float(np.mean([e[1].elapsed_time(e[2]) for e in ev2]))
    z2_uncertain, _ = ctx.zscore_stats(reset=True)
    z2_uncertain /= (n2 + 1)
    # `wave --influence 0.5` (stat.rs:40-48: the filtered series depends on earlier decisions): Kernel AB for GC and the influence-1
    # decisions, then the walk (one thread per segment, zscore_tseg_walk_kernel / zscore_tseg_fix_kernel); timed outside the
    # headline region, signals of three ctgs checked against the oracle
    wave_infl = None
    if args.influence == 1.0 and not args.skip_cpu:
        d_sig_i = torch.empty(n_win, dtype=torch.int8, device=dev)
        d_filt_i = torch.empty(n_win, dtype=torch.float32, device=dev)
        plan.wave(stream, d_seq, THRESHOLD, 0.5, d_gc, d_sig_i, None, d_filt_i)
        li0 = plan.launches
        ei0, ei1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ei0.record()
        for _ in range(5):
            plan.wave(stream, d_seq, THRESHOLD, 0.5, d_gc, d_sig_i, None, d_filt_i)
        ei1.record()
        torch.cuda.synchronize()
        ms_i = max_over_ranks(ei0.elapsed_time(ei1) / 5)
        ok_i = None
        if rank == 0:
            from oracle import oracle as orc
            woff_i = api.win_offsets(seq_off, SIZE, STEP)
            ok_i = True
            for c in sorted({0, n_ctg // 2, n_ctg - 1}):
                sq = d_seq[int(seq_off[c]):int(seq_off[c + 1])].cpu().numpy()
                _, _, _, rs = orc.wave_batch(sq, np.array([0, sq.size], dtype=np.uint64), SIZE, STEP, LAG, THRESHOLD, 0.5, threads=1)
                ok_i &= bool(np.array_equal(d_sig_i[int(woff_i[c]):int(woff_i[c + 1])].cpu().numpy(), rs))
        wave_infl = {"influence": 0.5, "ms_per_pass": ms_i, "value": total_bases / (ms_i * 1e-3) / 1e9, "unit": "Gbp/s",
                     "gpu_launches_per_pass": int((plan.launches - li0) // 5), "bit_exact_sample": ok_i,
                     "path": "gc_wave_fused_kernel + zscore_tseg_walk_kernel + zscore_tseg_fix_kernel (+ zscore_seq_ring_kernel for flagged series)"}
        del d_sig_i, d_filt_i
    step()  # leave the headline path's outputs in d_gc / d_sig for the parity spot check
    torch.cuda.synchronize()
    ctx.zscore_stats(reset=True)

    full = int(args.bases) == 3_100_000_000

    def roof(kernel, nbytes, ms):
        return {"bound": "hbm", "kernel": kernel, "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_peak, "traffic": ncu_traffic(kernel) if full else None,
                "algorithmic_bytes_per_launch": nbytes, "ms_per_launch": ms, "peak_source": peak_src}

    gc_bytes = n_bases * 1 + n_win * 4            # 1 B/base read + 4 B/window (f32 gc) written
    z_bytes = n_win * 5                           # 4 B gc read + 1 B signal written
    z_flops = n_win * (4 * LAG + 6)               # non-fusable f32 ops the reference's order forces
    roof_gc = roof("gc_window_kernel", gc_bytes, ms_gc)
    kern = {
        "gc_window_kernel": roof_gc,
        "zscore_kernel": {"bound": "issue (f64 certificate per window + exact f32 replay of inconclusive windows; not HBM, no tensor cores)",
                          "ms_per_launch": ms_z, "windows_replayed_exactly_per_launch": z2_uncertain,
                          "replay_fraction": z2_uncertain / max(1, n_win), "achieved_gbs": z_bytes / (ms_z * 1e-3) / 1e9,
                          "hbm_frac": z_bytes / (ms_z * 1e-3) / 1e9 / hbm_peak,
                          "traffic": ncu_traffic("zscore_kernel") if full else None,
                          "reference_equivalent_gflops": z_flops / (ms_z * 1e-3) / 1e9,
                          "algorithmic_bytes_per_launch": z_bytes},
    }
    if fused:
        # 1 B/base read + 4 B/window gc + 1 B/window signal written: the GC values never come back from HBM
        roof_main = roof("gc_wave_fused_kernel", n_bases + n_win * 5, ms_step)
        roof_main["windows_replayed_exactly_per_launch"] = z_uncertain
        roof_main["replay_fraction"] = z_uncertain / max(1, n_win)
        kern["gc_wave_fused_kernel"] = roof_main
    else:
        roof_main = roof_gc

    # ---- sanity: spot-check a few ctgs against the oracle (outside every timed region) ------------
    parity = None
    if rank == 0:
        from oracle import oracle as orc
        woff = api.win_offsets(seq_off, SIZE, STEP)
        pick = sorted({0, n_ctg // 2, n_ctg - 1})
        ok = True
        for c in pick:
            s = d_seq[int(seq_off[c]):int(seq_off[c + 1])].cpu().numpy()
            _, rc, rg, rs = orc.wave_batch(s, np.array([0, s.size], dtype=np.uint64), SIZE, STEP, LAG, THRESHOLD, args.influence, threads=1)
            a, b = int(woff[c]), int(woff[c + 1])
            ok &= bool(np.array_equal(d_gc[a:b].cpu().numpy().view(np.uint32), rg.view(np.uint32)))
            ok &= bool(np.array_equal(d_sig[a:b].cpu().numpy(), rs))
        parity = {"ctgs_checked": pick, "bit_exact": ok}

    # ---- e2e: host C-ABI call, pinned host buffers, H2D + D2H inside the timed region --------------
    e2e = e2e_gz = gz_host = None
    h_seq_ptr = None
    if not args.skip_e2e:
        woff = api.win_offsets(seq_off, SIZE, STEP)
        h_seq_ptr = pinned(n_bases, L)
        h_gc_ptr, h_sig_ptr = pinned(n_win * 4, L), pinned(n_win, L)
        h_seq = as_np(h_seq_ptr, n_bases, np.uint8)
        torch.from_numpy(h_seq).copy_(d_seq[:n_bases])
        torch.cuda.synchronize()

        def e2e_step():
            api.check(L.gams_cuda_wave_batch(ctx.handle, n_ctg, h_seq_ptr, seq_off.ctypes.data_as(C.c_void_p), SIZE, STEP,
                                             LAG, THRESHOLD, args.influence, woff.ctypes.data_as(C.c_void_p), None,
                                             h_gc_ptr, h_sig_ptr))

        # context for the e2e number: what this box's PCIe link gives for one plain pinned copy each way
        tpc = []
        api.check(L.gams_cuda_memcpy_probe(ctx.handle, h_seq_ptr, min(n_bases, 1 << 30), 1))   # untimed: reserves the device scratch
        for kind in (1, 2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            api.check(L.gams_cuda_memcpy_probe(ctx.handle, h_seq_ptr, min(n_bases, 1 << 30), kind))
            tpc.append(min(n_bases, 1 << 30) / (time.perf_counter() - t0) / 1e9)
        n_e2e = max(1, min(args.steps, 5))
        for _ in range(min(args.warmup, 2)):
            e2e_step()
        barrier()
        per_step = []
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            ts = time.perf_counter()
            e2e_step()
            per_step.append((time.perf_counter() - ts) * 1e3)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        if rank == 0:
            hs = as_np(h_sig_ptr, n_win, np.int8)
            parity["e2e_matches_device_path"] = bool(np.array_equal(hs, d_sig.cpu().numpy()) and np.array_equal(
                as_np(h_gc_ptr, n_win, np.float32).view(np.uint32), d_gc.cpu().numpy().view(np.uint32)))
        e2e = {"value": total_bases * n_e2e / dt / 1e9, "unit": "Gbp/s", "h2d_bytes_per_step": int(n_bases),
               "d2h_bytes_per_step": int(n_win * 5), "steps": n_e2e, "ms_per_step": dt / n_e2e * 1e3,
               "pcie_probe_gbs": {"h2d": tpc[0], "d2h": tpc[1]}, "ms_each_step_rank0": [round(x, 2) for x in per_step],
               "api": "gams_cuda_wave_batch (host pointers, pinned)", "timer": "host wall clock around the blocking call, max over ranks"}
        # ---- e2e_gz: the same step fed with what Redis holds: the gzip blobs of Conn::get_seq (redis.rs:142-161), inflated on the
        # device.  A third of the bytes cross the link; the decode costs GPU time instead (see `feed`).
        if not args.skip_gz:
            cores = os.cpu_count() or 1
            t = time.perf_counter()
            blobs = gz_blobs(h_seq, seq_off, max(1, cores // max(1, world)))
            t_comp = time.perf_counter() - t
            goff = np.zeros(n_ctg + 1, dtype=np.uint64)
            goff[1:] = np.cumsum([len(b) for b in blobs])
            n_gz = int(goff[-1])
            p_gz = pinned(n_gz, L)
            as_np(p_gz, n_gz, np.uint8)[:] = np.frombuffer(b"".join(blobs), dtype=np.uint8)
            del blobs
            st_gz = np.zeros(n_ctg, dtype=np.int32)
            as_np(h_sig_ptr, n_win, np.int8)[:] = 77
            vp = lambda a: a.ctypes.data_as(C.c_void_p)

            def gz_step():
                api.check(L.gams_cuda_wave_gz_batch(ctx.handle, n_ctg, p_gz, vp(goff), vp(seq_off), SIZE, STEP, LAG, THRESHOLD,
                                                    args.influence, vp(woff), None, h_gc_ptr, h_sig_ptr, vp(st_gz)))

            for _ in range(2):
                gz_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                gz_step()
            torch.cuda.synchronize()
            dtg = max_over_ranks(time.perf_counter() - t0)
            same = bool(np.array_equal(as_np(h_sig_ptr, n_win, np.int8), d_sig.cpu().numpy()) and np.array_equal(
                as_np(h_gc_ptr, n_win, np.float32).view(np.uint32), d_gc.cpu().numpy().view(np.uint32)))
            e2e_gz = {"value": total_bases * n_e2e / dtg / 1e9, "unit": "Gbp/s", "ms_per_step": dtg / n_e2e * 1e3,
                      "h2d_bytes_per_step": n_gz, "d2h_bytes_per_step": int(n_win * 5 + n_ctg * 4), "steps": n_e2e,
                      "compressed_bytes": n_gz, "ratio": n_bases / n_gz, "matches_device_path": same,
                      "api": "gams_cuda_wave_gz_batch (seq:{ctg} gzip blobs in pinned host memory -> device gunzip incl. CRC-32 -> fused GC-wave; decoded in two pieces, the windows of the first go down while the second is decoded)",
                      "host_compress_s": t_comp}
            gz_host = (p_gz, goff, n_gz)
        L.gams_cuda_host_free(h_gc_ptr)
        L.gams_cuda_host_free(h_sig_ptr)

        # the same through gams_cuda_wave_peaks_batch: what `gams wave` prints (merged peaks), so only the peak records come back
        cap = max(4096, n_win // 16)
        pk_dt = (np.uint32, np.uint32, np.uint32, np.float32, np.int8)
        pk_ptr = [pinned(cap * np.dtype(dt).itemsize, L) for dt in pk_dt]          # pinned, like the sequence and the window arrays
        pk = [as_np(p_, cap, dt) for p_, dt in zip(pk_ptr, pk_dt)]
        n_pk = C.c_uint64()

        def peaks_step():
            api.check(L.gams_cuda_wave_peaks_batch(ctx.handle, n_ctg, h_seq_ptr, seq_off.ctypes.data_as(C.c_void_p), SIZE, STEP, LAG,
                                                   THRESHOLD, args.influence, 0.2, woff.ctypes.data_as(C.c_void_p), cap,
                                                   *[a.ctypes.data_as(C.c_void_p) for a in pk], C.byref(n_pk)))

        peaks_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            peaks_step()
        dtp = max_over_ranks(time.perf_counter() - t0)
        e2e["peaks_only"] = {"value": total_bases * n_e2e / dtp / 1e9, "unit": "Gbp/s", "ms_per_step": dtp / n_e2e * 1e3,
                             "h2d_bytes_per_step": int(n_bases), "d2h_bytes_per_step": int(n_pk.value * 17 + 8),
                             "peak_records": int(n_pk.value), "api": "gams_cuda_wave_peaks_batch (host pointers, pinned; merge_ints on the device)"}
        if gz_host is not None:
            # ... and fed with the gzip blobs: `gams wave` between its GET and its print (get_seq + proc_ctg + merge_ints)
            pk2_ptr = [pinned(cap * np.dtype(dt).itemsize, L) for dt in pk_dt]
            pk2 = [as_np(p_, cap, dt) for p_, dt in zip(pk2_ptr, pk_dt)]
            n_pk2 = C.c_uint64()
            st_gz2 = np.zeros(n_ctg, dtype=np.int32)
            vp = lambda a: a.ctypes.data_as(C.c_void_p)

            def peaks_gz_step():
                api.check(L.gams_cuda_wave_peaks_gz_batch(ctx.handle, n_ctg, gz_host[0], vp(gz_host[1]), vp(seq_off), SIZE, STEP, LAG,
                                                          THRESHOLD, args.influence, 0.2, vp(woff), cap, *[vp(a) for a in pk2],
                                                          C.byref(n_pk2), vp(st_gz2)))

            peaks_gz_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                peaks_gz_step()
            dtq = max_over_ranks(time.perf_counter() - t0)
            k = int(n_pk.value)
            same_pk = bool(n_pk2.value == n_pk.value and all(np.array_equal(a[:k], b[:k]) for a, b in zip(pk, pk2)))
            e2e_gz["peaks_only"] = {"value": total_bases * n_e2e / dtq / 1e9, "unit": "Gbp/s", "ms_per_step": dtq / n_e2e * 1e3,
                                    "h2d_bytes_per_step": int(gz_host[2]), "d2h_bytes_per_step": int(n_pk2.value * 17 + 8 + n_ctg * 4),
                                    "peak_records": int(n_pk2.value), "equals_peaks_from_bases": same_pk,
                                    "api": "gams_cuda_wave_peaks_gz_batch (gzip blobs in pinned host memory -> device gunzip incl. CRC-32 -> "
                                           "fused GC-wave -> merge_ints on the device; peak records out)"}
            for p_ in pk2_ptr:
                L.gams_cuda_host_free(p_)
        for p_ in pk_ptr:
            L.gams_cuda_host_free(p_)

    # ---- feed step (a18): the seq:{ctg} blobs as Redis holds them (gzip, level 1) inflated on the device ------------
    feed = None
    if rank == 0 and world == 1 and gz_host is not None:
        feed = bench_gunzip(args, ctx, L, dev, gz_host, seq_off, d_seq, n_bases)
    if gz_host is not None:
        L.gams_cuda_host_free(gz_host[0])

    # ---- CPU baseline (rank 0, N = 1 only) --------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cores = os.cpu_count() or 1
        k = n_ctg
        h = d_seq[:int(seq_off[k])].cpu().numpy() if h_seq_ptr is None else as_np(h_seq_ptr, int(seq_off[k]), np.uint8)
        v, sample = cpu_wave_rate(h, seq_off[:k + 1], cores, args.cpu_seconds, args.influence)
        v1, sample1 = cpu_wave_rate(h, seq_off[:k + 1], 1, 3.0, args.influence)
        cpu = {"value": v, "unit": "Gbp/s", "cores": cores, "kind": "port", "sample": sample,
               "single_thread": {"value": v1, "cores": 1, "sample": sample1}}
    if h_seq_ptr is not None:
        L.gams_cuda_host_free(h_seq_ptr)

    del d_gc, d_sig, d_filt
    overlap = None if args.skip_count else bench_count(args, ctx, L, dev, rank, world, g_off, barrier, max_over_ranks,
                                                     sum_over_ranks, hbm_peak, peak_src)

    sw = None
    if rank == 0 and not args.skip_sw and not args.skip_e2e:
        sw = bench_sw(args, ctx, L, dev)

    frontend = None
    if rank == 0 and world == 1 and not args.skip_frontend and not args.skip_e2e:
        frontend = bench_frontend(args, ctx, L, dev, g_off, g_ctgs)

    multi = None
    if world > 1 and not args.skip_multi:
        multi = bench_multi(args, ctx, L, dev, rank, world, g_off, d_seq, barrier, max_over_ranks, gloo)

    if rank == 0:
        line = {
            "metric": "gc_wave_throughput", "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "passes_per_step": passes, "ms_per_pass": ms_total / args.steps / passes,
            "timed_region_ms": ms_total, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "u8->u32 counts, f32 statistics", "data": "synthetic",
            "config": workload_config(args), "clocks": clocks, "e2e": e2e, "e2e_gz": e2e_gz, "gpu_launches": int(launches),
            "roofline": roof_main, "kernels": kern, "wave_influence": wave_infl, "path": "fused (gc_wave_fused_kernel)" if fused else ("gc_wave_fused_kernel + zscore_tseg_walk_kernel + zscore_tseg_fix_kernel (influence != 1; GAMS_CUDA_ZWALK_WARPS=1: the warp-per-segment kernels)" if args.influence != 1.0 and plan.fused(THRESHOLD, 1.0) else "gc_window_kernel + zscore kernel"), "cpu_baseline": cpu, "overlap": overlap, "sw": sw, "feed": feed, "frontend": frontend, "multi_gpu": multi, "parity_spot_check": parity,
            "windows_per_step": int(n_win), "bases_per_rank": int(n_bases),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_gunzip(args, ctx, L, dev, gz_host, seq_off, d_seq, n_bases):
    """SURVEY a18 (Conn::get_seq -> decode_gz): every ctg compressed as encode_gz does (gzip level 1), inflated on the device
    with blobs and output resident in HBM; checked byte for byte against the genome it was made from.  Both decoders."""
    import zlib

    import torch

    from gams_b200 import api

    p_gz, goff, n_gz = gz_host
    n_ctg = seq_off.size - 1
    cat = as_np(p_gz, n_gz, np.uint8)
    k = min(n_ctg, 400)
    t = time.perf_counter()
    for c in range(k):
        zlib.decompress(cat[int(goff[c]):int(goff[c + 1])].tobytes(), 31)
    cpu_rate = float(seq_off[k]) / (time.perf_counter() - t) / 1e9
    d_gz = torch.from_numpy(cat).to(dev)
    d_goff = torch.from_numpy(goff.astype(np.int64)).to(dev)
    d_ooff = torch.from_numpy(seq_off.astype(np.int64)).to(dev)
    d_out = torch.zeros(n_bases + 64, dtype=torch.uint8, device=dev)
    d_st = torch.zeros(n_ctg, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def run():
        api.check(L.gams_cuda_gunzip_dev(ctx.handle, 0, stream, n_ctg, d_gz.data_ptr(), d_goff.data_ptr(), d_out.data_ptr(),
                                         d_ooff.data_ptr(), d_st.data_ptr()))

    res = {}
    n = max(1, min(args.steps, 3))
    for name, dec, crc in (("two_phase_crc", 0, True), ("two_phase", 0, False), ("one_pass", 1, False)):
        ctx.set_gunzip_options(dec, crc)
        d_out.zero_()
        run()
        torch.cuda.synchronize()
        ok = bool(int(d_st.abs().sum()) == 0 and torch.equal(d_out[:n_bases], d_seq[:n_bases]))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        res[name] = {"ms_per_step": ms, "value": n_bases / (ms * 1e-3) / 1e9, "bit_exact": ok}
    ctx.set_gunzip_options(0, True)
    main = res["two_phase_crc"]
    # the same through the host entry point (Conn::get_seq for a batch of ctgs: blobs in host memory, bases back in host memory;
    # both pinned): the batch goes in two pieces, upload / decode / download of different pieces overlap
    del d_out
    torch.cuda.empty_cache()
    p_out = pinned(n_bases, L)
    st_h = np.zeros(n_ctg, dtype=np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)

    def run_host():
        api.check(L.gams_cuda_gunzip_batch(ctx.handle, n_ctg, p_gz, vp(goff), vp(seq_off), p_out, vp(st_h)))

    run_host()
    ok_h = bool(not st_h.any() and torch.equal(torch.from_numpy(as_np(p_out, n_bases, np.uint8)).to(dev), d_seq[:n_bases]))
    t = time.perf_counter()
    for _ in range(n):
        run_host()
    ms_h = (time.perf_counter() - t) / n * 1e3
    L.gams_cuda_host_free(p_out)
    e2e_feed = {"value": n_bases / (ms_h * 1e-3) / 1e9, "unit": "Gbp/s", "ms_per_step": ms_h, "h2d_bytes_per_step": int(n_gz),
                "d2h_bytes_per_step": int(n_bases + 4 * n_ctg), "bit_exact": ok_h,
                "api": "gams_cuda_gunzip_batch (host pointers, pinned; two pieces)"}
    return {"e2e": e2e_feed, "what": "gunzip of the seq:{ctg} blobs on the device (gams_cuda_gunzip_dev), blobs and output in HBM; default decoder: block headers "
                    "found by a speculative search (one warp per group of 8 KB ranges) -> tokens (one thread per deflate block, verified "
                    "hand-over from block to block) -> bytes (one warp per blob) -> CRC-32",
            "value": main["value"], "unit": "Gbp/s", "ms_per_step": main["ms_per_step"], "blobs": int(n_ctg),
            "compressed_bytes": int(n_gz), "ratio": n_bases / n_gz, "bit_exact": all(r["bit_exact"] for r in res.values()),
            "decoders": res,
            "bound": "tokens: latency of the serial chain of one deflate symbol (index -> table load -> shift) over the longest block; bytes: "
                     "instruction issue (~96 warp instructions per 32 output bytes) and the serial chain of the longest blob",
            "cpu_baseline": {"value": cpu_rate, "unit": "Gbp/s", "cores": 1, "kind": "zlib",
                             "sample": f"zlib.decompress of the first {k} blobs, one thread"}}


def bench_sw(args, ctx, L, dev):
    """configs[2]: `gams sw` around 10^6 point features on an Arabidopsis-scale genome (120 Mbp), end to end through
    gams_cuda_sw_batch (host arrays in, the rows of sw.rs:141-191 out), with the oracle timed on a sample beside it."""
    import torch

    from gams_b200 import api, synth
    from oracle import oracle as orc

    chr_lengths = [30_427_671, 19_698_289, 23_459_830, 18_585_056, 26_975_502, 154_478, 366_924]   # Atha, 119.7 Mbp
    n = int(sum(chr_lengths))
    seq_off, ctgs = synth.ctg_layout(chr_lengths, PIECE)
    h_seq = synth.genome_torch(n, seed=7, device=dev, capacity=n + 64)[:n].cpu().numpy()
    rng = np.random.default_rng(8)
    lens = np.diff(seq_off.astype(np.int64))
    chr_start = np.array([c[1] for c in ctgs], dtype=np.int32)
    n_feat = int(args.sw_features)
    f_ctg = np.sort(rng.choice(len(ctgs), n_feat, p=lens / lens.sum())).astype(np.uint32)
    f_s = (chr_start[f_ctg] + (rng.random(n_feat) * lens[f_ctg]).astype(np.int64)).astype(np.int32)
    f_e = np.minimum(f_s + rng.integers(0, 3, n_feat), chr_start[f_ctg] + lens[f_ctg] - 1).astype(np.int32)
    # pinned host buffers, as for the wave path: the rows are 1.2 GB
    cap = n_feat * 41
    p_seq = pinned(n, L)
    as_np(p_seq, n, np.uint8)[:] = h_seq
    specs = [("start", np.int32), ("end", np.int32), ("type", np.int8), ("dist", np.int32), ("gc", np.float32), ("mean", np.float32),
             ("sd", np.float32), ("cv", np.float32)]
    ptrs = {k: pinned(cap * np.dtype(dt).itemsize, L) for k, dt in specs}
    row_off = np.zeros(n_feat + 1, dtype=np.uint64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)

    def sw_step():
        api.check(L.gams_cuda_sw_batch(ctx.handle, len(ctgs), p_seq, vp(seq_off), vp(chr_start), n_feat, vp(f_ctg), vp(f_s), vp(f_e),
                                       100, 20, 500, vp(row_off), ptrs["start"], ptrs["end"], ptrs["type"], ptrs["dist"], ptrs["gc"],
                                       ptrs["mean"], ptrs["sd"], ptrs["cv"]))

    sw_step()   # warm-up (allocations)
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        sw_step()
    dt = (time.perf_counter() - t0) / reps
    r = {k: as_np(ptrs[k], int(row_off[-1]), dt_) for k, dt_ in specs}
    r["row_off"] = row_off
    n_rows = int(r["gc"].size)
    # oracle on a sample of the features (single thread: sw::proc_ctg's per-feature loop, cache lookups replaced by recomputation)
    k = min(n_feat, 20_000)
    pick = rng.choice(n_feat, k, replace=False)
    t0 = time.perf_counter()
    rows_cpu, ok = 0, True
    for f in pick:
        c = int(f_ctg[f])
        ref = orc.sw_feature(h_seq[int(seq_off[c]):int(seq_off[c + 1])], int(chr_start[c]), int(chr_start[c] + lens[c] - 1),
                             int(f_s[f]), int(f_e[f]))
        rows_cpu += len(ref["gc"])
        if f in pick[:300]:
            a, b = int(r["row_off"][f]), int(r["row_off"][f + 1])
            ok &= bool(b - a == len(ref["gc"]) and all(np.array_equal(r[key][a:b].view(np.uint32), ref[key].view(np.uint32))
                                                      for key in ("gc", "mean", "sd", "cv")))
    dt_cpu = time.perf_counter() - t0
    # sw --action count: the same rows plus rg_count against 10^6 SNPs (chr coordinates), one call
    per = max(1, 1_000_000 // len(ctgs))
    st, sp, ioff = synth.rg_index(lens, [per] * len(ctgs), seed=9)
    shift = np.repeat((chr_start - 1).astype(np.uint32), per)
    st, sp = st + shift, sp + shift
    p_cnt = pinned(cap * 4, L)

    def sw_count_step():
        api.check(L.gams_cuda_sw_count_batch(ctx.handle, len(ctgs), p_seq, vp(seq_off), vp(chr_start), n_feat, vp(f_ctg), vp(f_s), vp(f_e),
                                             100, 20, 500, vp(st), vp(sp), vp(ioff), vp(row_off), ptrs["start"], ptrs["end"], ptrs["type"],
                                             ptrs["dist"], ptrs["gc"], ptrs["mean"], ptrs["sd"], ptrs["cv"], p_cnt))

    sw_count_step()
    t0 = time.perf_counter()
    for _ in range(reps):
        sw_count_step()
    dt_cnt = (time.perf_counter() - t0) / reps
    got_cnt = as_np(p_cnt, n_rows, np.uint32)
    sel = np.sort(rng.choice(n_rows, min(n_rows, 200_000), replace=False))
    seg_of_row = np.repeat(f_ctg.astype(np.int32), np.diff(row_off).astype(np.int64))
    ref_cnt = orc.count_batch(st, sp, ioff, seg_of_row[sel], r["start"][sel].astype(np.uint32), r["end"][sel].astype(np.uint32),
                              threads=os.cpu_count() or 1)
    sw_count = {"value": n_rows / dt_cnt, "unit": "sw rows/s", "ms_per_step": dt_cnt * 1e3, "rg_intervals": int(st.size),
                "d2h_bytes_per_step": int(n_rows * 33 + (n_feat + 1) * 8), "bit_exact_sample": bool(np.array_equal(got_cnt[sel], ref_cnt)),
                "api": "gams_cuda_sw_count_batch (sw rows + rg_count = count_rg of every row, staged Kernel C on the rows)"}
    L.gams_cuda_host_free(p_cnt)
    for pp in list(ptrs.values()) + [p_seq]:
        L.gams_cuda_host_free(pp)
    return {"sw_count": sw_count,
            "workload": "Arabidopsis-scale synthetic genome (119.7 Mbp) + sw --size 100 --max 20 --resize 500 around "
                        f"{n_feat} point features (configs[2])",
            "value": n_rows / dt, "unit": "sw rows/s", "features_per_s": n_feat / dt, "rows": n_rows, "ms_per_step": dt * 1e3,
            "h2d_bytes_per_step": int(n + n_feat * 12), "d2h_bytes_per_step": int(n_rows * 29 + (n_feat + 1) * 8),
            "api": "gams_cuda_sw_batch (host pointers, pinned; copies and row layout inside)",
            "bit_exact_sample": ok,
            "cpu_baseline": {"value": rows_cpu / dt_cpu, "unit": "sw rows/s", "cores": 1, "kind": "port",
                             "sample": f"{k} of the features through the oracle (ctypes call per feature)"}}


def bench_frontend(args, ctx, L, dev, g_off, ctgs):
    """The rows either side of the count kernel at scale, through the host C ABI (text / host arrays in, text / arrays out):
    a17 range-file parsing and `locate --count -f` text in / text out (locate.rs:80-141), a14 / f3 the rg index build
    (build_idx_rg, redis.rs:276-305), each with the oracle timed on a sample beside it."""
    import torch

    from gams_b200 import api, synth
    from oracle import oracle as orc

    out = {}
    rng = np.random.default_rng(3)
    lens = np.diff(g_off.astype(np.int64))
    n_ctg = lens.size
    chr_of = np.array([c[0] for c in ctgs])
    names = [f"chr{k + 1}" for k in range(int(chr_of.max()) + 1)]
    cs = np.array([c[1] for c in ctgs], dtype=np.uint32)
    ce = np.array([c[2] + 1 for c in ctgs], dtype=np.uint32)   # stop exclusive
    chr_off = np.concatenate(([0], np.cumsum(np.bincount(chr_of, minlength=len(names))))).astype(np.uint64)
    chr_len = np.array([int(ce[chr_of == k].max()) - 1 for k in range(len(names))])
    # ---- query text: 2e7 lines `chr:start-end`, arbitrary order
    n_lines, m = 20_000_000, 1_000_000
    c = rng.integers(0, len(names), m)
    qs = (rng.random(m) * chr_len[c]).astype(np.int64) + 1
    qe = np.minimum(qs + rng.integers(0, 10_000, m), chr_len[c])
    block = "\n".join(f"{names[a]}:{b}-{d}" for a, b, d in zip(c.tolist(), qs.tolist(), qe.tolist())) + "\n"
    text = (block * (n_lines // m)).encode()
    api.parse_ranges(ctx, text[:100_000], names)
    t = time.perf_counter()
    got = api.parse_ranges(ctx, text, names)
    dt = time.perf_counter() - t
    k = 100_000
    lines = block.splitlines()[:k]
    t = time.perf_counter()
    ref = [orc.range_from_str(ln) for ln in lines]
    dtc = time.perf_counter() - t
    ok = all(r is not None and (int(got["start"][i]), int(got["end"][i])) == (r[2], r[3]) and names[int(got["chr"][i])] == r[0]
             for i, r in enumerate(ref))
    out["parse_ranges"] = {"value": n_lines / dt, "unit": "lines/s", "ms_per_step": dt * 1e3, "lines": n_lines, "text_bytes": len(text),
                           "handed_back": int((got["chr"] == -3).sum()), "matches_oracle_sample": bool(ok),
                           "api": "gams_cuda_parse_ranges (host text in, arrays out, copies inside)",
                           "cpu_baseline": {"value": k / dtc, "unit": "lines/s", "cores": 1, "kind": "port",
                                            "sample": f"{k} lines through the oracle's Range::from_str restatement (python re)"}}
    del got
    # ---- the rg index: 1e8 SNPs in chr coordinates, unsorted as the loader sees them, built on the device
    per = max(1, int(args.intervals) // n_ctg)
    d_st, _, iv_off = synth.rg_index_torch(lens, per, seed=7, device=dev)
    d_st = d_st.view(n_ctg, per) + torch.as_tensor((cs.astype(np.int64) - 1), device=dev)[:, None].to(torch.int32)
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    perm = torch.argsort(torch.rand(n_ctg, per, generator=g, device=dev), dim=1)
    h_start = torch.gather(d_st, 1, perm).reshape(-1).cpu().numpy().view(np.uint32)     # file order inside each ctg
    h_sorted = d_st.reshape(-1).cpu().numpy().view(np.uint32)
    del d_st, perm
    h_stop = h_start + np.uint32(1)
    api.build_idx_rg(ctx, h_start[:int(iv_off[8])], h_stop[:int(iv_off[8])], iv_off[:9], want_blob=False)
    t = time.perf_counter()
    ix = api.build_idx_rg(ctx, h_start, h_stop, iv_off, want_blob=False)
    dt = time.perf_counter() - t
    ok = bool(np.array_equal(ix["starts"], h_sorted) and np.array_equal(ix["stops"], h_sorted + np.uint32(1)))
    kseg = 64
    t = time.perf_counter()
    for sgm in range(kseg):
        a, b = int(iv_off[sgm]), int(iv_off[sgm + 1])
        orc.lapper_new([(int(x), int(x) + 1, "") for x in h_start[a:b]])
    dtc = time.perf_counter() - t
    n_iv = int(iv_off[-1])
    out["build_idx_rg"] = {"value": n_iv / dt, "unit": "intervals/s", "ms_per_step": dt * 1e3, "intervals": n_iv, "segments": int(n_ctg),
                           "h2d_bytes_per_step": n_iv * 8, "d2h_bytes_per_step": n_iv * 16, "sorted_arrays_match": ok,
                           "api": "gams_cuda_build_idx_rg (host arrays in; sorted intervals, starts[], stops[], max_len out; no blob)",
                           "cpu_baseline": {"value": kseg * per / dtc, "unit": "intervals/s", "cores": 1, "kind": "port",
                                            "sample": f"Lapper::new restated (oracle.lapper_new) on the first {kseg} ctgs"}}
    # ---- locate --count -f: text in, text out, against that index
    st, sp = ix["starts"], ix["stops"]
    api.locate_count_text(ctx, text[:100_000], names, cs, ce, chr_off, st, sp, iv_off)
    t = time.perf_counter()
    txt, nl, npr, nb = api.locate_count_text(ctx, text, names, cs, ce, chr_off, st, sp, iv_off)
    dt = time.perf_counter() - t
    k = 2000
    ok = True
    for ln, gline in zip(lines[:k], txt[:200_000].decode().splitlines()[:k]):
        rg = orc.range_from_str(ln)
        ci = names.index(rg[0])
        j = int(np.searchsorted(ce[int(chr_off[ci]):int(chr_off[ci + 1])], rg[2], side="right")) + int(chr_off[ci])
        refc = orc.lapper_count(st[int(iv_off[j]):int(iv_off[j + 1])], sp[int(iv_off[j]):int(iv_off[j + 1])], rg[2], rg[3])
        ok &= gline == f"{ln}\t{refc}"
    out["locate_count_text"] = {"value": nl / dt, "unit": "lines/s", "ms_per_step": dt * 1e3, "lines": int(nl), "printed": int(npr),
                                "handed_back": int(nb), "h2d_bytes_per_step": len(text) + n_iv * 8, "d2h_bytes_per_step": len(txt),
                                "first_lines_match_oracle": bool(ok),
                                "api": "gams_cuda_locate_count_text (range-file text in, `rg\\tcount` text out; parse, find_one_idx, count_rg, "
                                       "formatting on the device)"}
    return out


def bench_multi(args, ctx, L, dev, rank, world, g_off, d_seq, barrier, max_over_ranks, gloo):
    """N > 1, next to the weak-scaling headline: (1) STRONG scaling as BASELINE configs[3] words it — ONE 3.1 Gbp genome, its ctgs
    sharded over the N ranks (shard.shard_for_rank = gams_cuda_partition), device-resident and end to end; (2) the library's own
    multi-device context, as a Rust host would use it: ONE process (rank 0) owns all N devices through gams_cuda_create and calls
    gams_cuda_wave_batch / gams_cuda_count_overlaps once; the other ranks wait on a CPU barrier with their GPUs idle."""
    import torch
    import torch.distributed as dist

    from gams_b200 import api, shard, synth

    out = {}
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    stream = torch.cuda.current_stream().cuda_stream
    # ---- (1) strong scaling: this rank's contiguous, byte-balanced run of ctgs of the one genome
    c0, c1, s_off = shard.shard_for_rank(g_off, rank, world)
    plan = api.WavePlan(ctx, s_off, SIZE, STEP, LAG)
    nb, nw = plan.total_bases, plan.total_windows
    d_gc = torch.empty(nw, dtype=torch.float32, device=dev)
    d_sig = torch.empty(nw, dtype=torch.int8, device=dev)
    d_filt = torch.empty(nw, dtype=torch.float32, device=dev) if args.influence != 1.0 else None
    seq = d_seq[:plan.seq_capacity]
    passes = 64
    for _ in range(3):
        plan.wave(stream, seq, THRESHOLD, args.influence, d_gc, d_sig, None, d_filt)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(passes):
        plan.wave(stream, seq, THRESHOLD, args.influence, d_gc, d_sig, None, d_filt)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / passes
    total = float(g_off[-1])
    woff = api.win_offsets(s_off, SIZE, STEP)
    h_seq_ptr, h_gc_ptr, h_sig_ptr = pinned(nb, L), pinned(nw * 4, L), pinned(nw, L)
    torch.from_numpy(as_np(h_seq_ptr, nb, np.uint8)).copy_(d_seq[:nb])
    torch.cuda.synchronize()

    def e2e_step():
        api.check(L.gams_cuda_wave_batch(ctx.handle, s_off.size - 1, h_seq_ptr, vp(s_off), SIZE, STEP, LAG, THRESHOLD, args.influence,
                                         vp(woff), None, h_gc_ptr, h_sig_ptr))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0) / 3
    out["strong"] = {"workload": f"ONE {total / 1e9:.2f} Gbp genome, ctgs {c0}..{c1} of {g_off.size - 1} on rank {rank} (contiguous, byte-balanced)",
                     "value": total / (ms * 1e-3) / 1e9, "unit": "Gbp/s", "ms_per_pass": ms, "scaling": "strong",
                     "e2e": {"value": total / dt / 1e9, "unit": "Gbp/s", "ms_per_step": dt * 1e3, "h2d_bytes_per_step_per_rank": int(nb),
                             "d2h_bytes_per_step_per_rank": int(nw * 5)},
                     "timer": "max over ranks (CUDA events for the resident passes, host clock around the blocking host calls)"}
    for pp in (h_seq_ptr, h_gc_ptr, h_sig_ptr):
        L.gams_cuda_host_free(pp)
    plan.close()
    del d_gc, d_sig, d_filt
    # ---- (2) one process, N devices (rank 0; the others wait with idle GPUs)
    torch.cuda.synchronize()
    dist.barrier(group=gloo)
    if rank == 0:
        n_ctg = g_off.size - 1
        n_bases = int(g_off[-1])
        gw = api.win_offsets(g_off, SIZE, STEP)
        n_win = int(gw[-1])
        p_seq, p_gc, p_sig, p_sig1 = pinned(n_bases, L), pinned(n_win * 4, L), pinned(n_win, L), pinned(n_win, L)
        torch.from_numpy(as_np(p_seq, n_bases, np.uint8)).copy_(d_seq[:n_bases])
        torch.cuda.synchronize()
        mctx = api.Context(list(range(world)))

        def call(c, sig_ptr):
            api.check(L.gams_cuda_wave_batch(c.handle, n_ctg, p_seq, vp(g_off), SIZE, STEP, LAG, THRESHOLD, args.influence, vp(gw), None,
                                             p_gc, sig_ptr))

        call(mctx, p_sig)
        t0 = time.perf_counter()
        for _ in range(3):
            call(mctx, p_sig)
        dtm = (time.perf_counter() - t0) / 3
        call(ctx, p_sig1)                   # the single-device result of the same call
        same = bool(np.array_equal(as_np(p_sig, n_win, np.int8), as_np(p_sig1, n_win, np.int8)))
        rec = {"devices": world, "wave_batch": {"value": n_bases / dtm / 1e9, "unit": "Gbp/s", "ms_per_step": dtm * 1e3,
                                                "h2d_bytes_per_step": n_bases, "d2h_bytes_per_step": n_win * 5,
                                                "equals_single_device_result": same}}
        for pp in (p_seq, p_gc, p_sig, p_sig1):
            L.gams_cuda_host_free(pp)
        # overlap counts through the same context: grouped queries shard by ctg, arbitrary ones replicate the index
        lens = np.diff(g_off.astype(np.int64))
        per = max(1, int(args.intervals) // lens.size)
        d_st, d_sp, iv_off = synth.rg_index_torch(lens, per, seed=7, device=dev)
        h_st, h_sp = d_st.cpu().numpy().view(np.uint32), d_sp.cpu().numpy().view(np.uint32)
        del d_st, d_sp
        n_q = int(args.queries)
        for name, grouped in (("grouped_by_ctg", True), ("arbitrary_order", False)):
            q_seg, q_s, q_e = synth.queries_torch(lens, n_q, seed=11, grouped=grouped, device=dev)
            hq = [t.cpu().numpy() for t in (q_seg, q_s, q_e)]
            del q_seg, q_s, q_e
            res = {}
            for label, c in (("multi", mctx), ("single", ctx)):
                o = np.empty(n_q, dtype=np.uint32)
                api.check(L.gams_cuda_count_overlaps(c.handle, iv_off.size - 1, vp(h_st), vp(h_sp), vp(iv_off), n_q, vp(hq[0]),
                                                     vp(hq[1].view(np.uint32)), vp(hq[2].view(np.uint32)), vp(o)))
                t0 = time.perf_counter()
                api.check(L.gams_cuda_count_overlaps(c.handle, iv_off.size - 1, vp(h_st), vp(h_sp), vp(iv_off), n_q, vp(hq[0]),
                                                     vp(hq[1].view(np.uint32)), vp(hq[2].view(np.uint32)), vp(o)))
                res[label] = (time.perf_counter() - t0, o)
            rec[f"count_overlaps_{name}"] = {"value": n_q / res["multi"][0], "unit": "counts/s", "ms_per_step": res["multi"][0] * 1e3,
                                             "ms_per_step_one_device": res["single"][0] * 1e3,
                                             "equals_single_device_result": bool(np.array_equal(res["multi"][1], res["single"][1])),
                                             "host_buffers": "pageable numpy arrays (not pinned)"}
        mctx.close()
        out["single_process_ctx"] = rec
    dist.barrier(group=gloo)
    return out


def bench_count(args, ctx, L, dev, rank, world, g_off, barrier, max_over_ranks, sum_over_ranks, hbm_peak, peak_src):
    """configs[4]: batched `locate --count` — queries in ARBITRARY order (the order locate.rs:111-141 walks the range file in:
    this is the configs[4] distribution, uniform random starts) and, for comparison, grouped by ctg (the order the `gams rg`
    loader's read_range yields); SNP intervals; index resident in HBM as it is in the reference's BTreeMap."""
    import torch

    from gams_b200 import api, synth

    lens = np.diff(g_off.astype(np.int64))
    n_ctg = lens.size
    per_ctg = max(1, int(args.intervals) // n_ctg)
    n_q = int(args.queries)
    d_st, d_sp, iv_off = synth.rg_index_torch(lens, per_ctg, seed=7 + rank, device=dev)
    plan = api.CountPlan(ctx, iv_off)
    stream = torch.cuda.current_stream().cuda_stream
    out = {}
    n_iv = int(iv_off[-1])
    for name, grouped in (("grouped_by_ctg", True), ("arbitrary_order", False)):
        q_seg, q_s, q_e = synth.queries_torch(lens, n_q, seed=11 + rank, grouped=grouped, device=dev)
        d_out = torch.empty(n_q, dtype=torch.int32, device=dev)
        for _ in range(args.warmup):
            plan.count(stream, d_st, d_sp, q_seg, q_s, q_e, d_out, grouped=grouped)
        l0 = plan.launches
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            plan.count(stream, d_st, d_sp, q_seg, q_s, q_e, d_out, grouped=grouped)
        t1.record()
        barrier()
        ms = max_over_ranks(t0.elapsed_time(t1)) / args.steps
        tot_q = sum_over_ranks(float(n_q))
        alg = n_q * 16 + n_iv * 8   # 12 B query + 4 B result per query, 8 B per interval
        res = {"value": tot_q / (ms * 1e-3), "unit": "counts/s", "ms_per_step": ms, "queries": n_q, "intervals": n_iv,
               "gpu_launches": int(plan.launches - l0),
               "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": alg / (ms * 1e-3) / 1e9 / hbm_peak,
                            "kernel": "seg_query_bounds_kernel + seg_items_kernel + count_overlaps_staged_kernel" if grouped else "part_sort_kernel + part_lookup_kernel + part_unpermute_kernel",
                            "traffic": ncu_traffic("count_overlaps:grouped_by_ctg" if grouped else "count_overlaps:arbitrary_order") if n_q == 100_000_000 else None,
                            "algorithmic_bytes_per_launch": alg, "peak_source": peak_src}}
        tr = res["roofline"]["traffic"]
        if tr:   # the same time against the bytes DRAM really moved (ncu): what the kernel achieves on its own traffic
            res["roofline"]["frac_of_peak_on_dram_traffic"] = tr / (ms * 1e-3) / 1e9 / hbm_peak
        if rank == 0:
            from oracle import oracle as orc
            k = 200_000
            sel = torch.randint(0, n_q, (k,), device=dev)
            ref = orc.count_batch(d_st.cpu().numpy().view(np.uint32), d_sp.cpu().numpy().view(np.uint32), iv_off,
                                  q_seg[sel].cpu().numpy(), q_s[sel].cpu().numpy().view(np.uint32),
                                  q_e[sel].cpu().numpy().view(np.uint32), threads=os.cpu_count() or 1)
            res["bit_exact_sample"] = bool(np.array_equal(d_out[sel].cpu().numpy().view(np.uint32), ref))
        out[name] = res
        if not grouped and not args.skip_e2e:
            # end to end through gams_cuda_count_overlaps: index + queries in pinned HOST memory, H2D and D2H inside
            ptrs, hs = [], []
            for t in (d_st, d_sp, q_seg, q_s, q_e):
                pp = pinned(t.numel() * 4, L)
                h = as_np(pp, t.numel(), np.int32)
                torch.from_numpy(h).copy_(t)
                ptrs.append(pp), hs.append(h)
            p_out = pinned(n_q * 4, L)
            torch.cuda.synchronize()
            ivp = iv_off.ctypes.data_as(C.c_void_p)

            def count_e2e():
                api.check(L.gams_cuda_count_overlaps(ctx.handle, iv_off.size - 1, ptrs[0], ptrs[1], ivp, n_q, ptrs[2], ptrs[3],
                                                     ptrs[4], p_out))

            n_e = max(1, min(args.steps, 3))
            count_e2e()
            barrier()
            t = time.perf_counter()
            for _ in range(n_e):
                count_e2e()
            dt = max_over_ranks(time.perf_counter() - t) / n_e
            h_out = as_np(p_out, n_q, np.uint32)
            out["e2e"] = {"value": tot_q / dt, "unit": "counts/s", "ms_per_step": dt * 1e3,
                          "h2d_bytes_per_step": int(n_q * 12 + n_iv * 8), "d2h_bytes_per_step": int(n_q * 4),
                          "api": "gams_cuda_count_overlaps (host pointers, pinned; index upload and rank-table build inside)",
                          "matches_device_path": bool(np.array_equal(h_out, d_out.cpu().numpy().view(np.uint32)))}
            hs = [h.copy() if (rank == 0 and world == 1 and not args.skip_cpu) else None for h in hs]
            for pp in ptrs + [p_out]:
                L.gams_cuda_host_free(pp)
            if rank == 0 and world == 1 and not args.skip_cpu:
                from oracle import oracle as orc
                cores = os.cpu_count() or 1
                k = min(n_q, 20_000_000)
                t = time.perf_counter()
                orc.count_batch(hs[0].view(np.uint32), hs[1].view(np.uint32), iv_off, hs[2][:k], hs[3][:k].view(np.uint32),
                                hs[4][:k].view(np.uint32), threads=cores)
                dtc = time.perf_counter() - t
                k1 = min(n_q, 2_000_000)
                t = time.perf_counter()
                orc.count_batch(hs[0].view(np.uint32), hs[1].view(np.uint32), iv_off, hs[2][:k1], hs[3][:k1].view(np.uint32),
                                hs[4][:k1].view(np.uint32), threads=1)
                dt1 = time.perf_counter() - t
                out["cpu_baseline"] = {"value": k / dtc, "unit": "counts/s", "cores": cores, "kind": "port",
                                       "sample": f"first {k} arbitrary-order queries against the full index",
                                       "single_thread": {"value": k1 / dt1, "sample": f"first {k1} queries"}}
        del q_seg, q_s, q_e, d_out
    plan.close()
    # rg counts along every sliding window of the genome (gams::sliding x count_rg): no query arrays, 4 B per window out
    chr_start = np.ones(n_ctg, dtype=np.int32)
    woff = np.zeros(n_ctg + 1, dtype=np.uint64)
    woff[1:] = np.cumsum([max(0, (int(x) - SIZE) // STEP + 1) if x >= SIZE else 0 for x in lens])
    wplan = api.CountWindowsPlan(ctx, chr_start, woff, SIZE, STEP, iv_off)
    n_w = int(woff[-1])
    d_cnt = torch.empty(n_w, dtype=torch.int32, device=dev)
    for _ in range(args.warmup):
        wplan.count(stream, d_st, d_sp, d_cnt)
    lw0 = wplan.launches
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        wplan.count(stream, d_st, d_sp, d_cnt)
    t1.record()
    barrier()
    ms = max_over_ranks(t0.elapsed_time(t1)) / args.steps
    alg = n_w * 4 + n_iv * 8
    res = {"value": sum_over_ranks(float(n_w)) / (ms * 1e-3), "unit": "window counts/s", "ms_per_step": ms, "windows": n_w,
           "intervals": n_iv, "gpu_launches": int(wplan.launches - lw0),
           "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg / (ms * 1e-3) / 1e9 / hbm_peak, "kernel": "cw_tile_kernel (per-tile descriptors built once per index by cw_desc_kernel)",
                        "traffic": ncu_traffic("count_windows"), "algorithmic_bytes_per_launch": alg, "peak_source": peak_src}}
    if rank == 0:
        from oracle import oracle as orc
        k = 200_000
        sel = np.sort(np.random.default_rng(5).integers(0, n_w, k))
        c = np.searchsorted(woff, sel, side="right") - 1
        ws = (1 + (sel - woff[c].astype(np.int64)) * STEP).astype(np.uint32)
        ref = orc.count_batch(d_st.cpu().numpy().view(np.uint32), d_sp.cpu().numpy().view(np.uint32), iv_off, c.astype(np.int32), ws,
                              (ws + np.uint32(SIZE - 1)), threads=os.cpu_count() or 1)
        res["bit_exact_sample"] = bool(np.array_equal(d_cnt[torch.as_tensor(sel, device=dev)].cpu().numpy().view(np.uint32), ref))
    out["sliding_windows"] = res
    wplan.close()
    return out


if __name__ == "__main__":
    main()
