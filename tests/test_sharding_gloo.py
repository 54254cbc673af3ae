"""N>1 host logic on CPU: two gloo ranks shard a ctg set with gams_cuda_partition, each computes only its
shard (the oracle stands in for the GPU here — this test is about the sharding, not the kernels), rank 0
gathers, and the result must equal the unsharded computation.  No collective on the data path."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gams_b200 import shard, synth
    from oracle import oracle as orc

    g = synth.genome(900_000, seed=77)
    cuts = [0, 130_001, 130_001 + 77_777, 400_000, 400_000 + 5_003, 650_000, 900_000]
    seq_off = np.asarray(cuts, dtype=np.uint64)
    c0, c1, loc = shard.shard_for_rank(seq_off, rank, world)
    lo, hi = int(seq_off[c0]), int(seq_off[c1])
    woff, cnt, gc, sig = orc.wave_batch(g[lo:hi], loc, 100, 10, 100, 3.0, 1.0)
    all_cnt = shard.gather_to_rank0(cnt, rank, world)
    all_sig = shard.gather_to_rank0(sig, rank, world)
    owned = shard.gather_to_rank0(np.array([c0, c1]), rank, world)
    dist.barrier()
    if rank == 0:
        _, rcnt, _, rsig = orc.wave_batch(g, seq_off, 100, 10, 100, 3.0, 1.0)
        ok = np.array_equal(all_cnt, rcnt) and np.array_equal(all_sig, rsig)
        cover = owned.reshape(world, 2)
        ok &= cover[0, 0] == 0 and cover[-1, 1] == len(cuts) - 1 and all(cover[i, 1] == cover[i + 1, 0] for i in range(world - 1))
        np.save(out_path, np.array([int(ok), all_cnt.size]))
    dist.destroy_process_group()


def test_two_rank_ctg_sharding_matches_unsharded(tmp_path):
    out = str(tmp_path / "ok.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    ok, n = np.load(out)
    assert ok == 1 and n > 80_000
