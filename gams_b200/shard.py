"""ctg sharding across ranks (one process per GPU).  The path shards by ctg with no exchange step
(SURVEY §8e): every rank takes a contiguous, byte-balanced run of ctgs (gams_cuda_partition), computes its
windows independently, and results are only gathered for output.  No collective is on the data path."""
from __future__ import annotations

import numpy as np

from . import api


def shard_for_rank(seq_off, rank: int, world: int):
    """→ (c0, c1, local_seq_off): this rank owns ctgs [c0, c1); local offsets start at 0."""
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
    part = api.partition(seq_off, world)
    c0, c1 = int(part[rank]), int(part[rank + 1])
    return c0, c1, (seq_off[c0:c1 + 1] - seq_off[c0]).astype(np.uint64)


def gather_to_rank0(local: np.ndarray, rank: int, world: int):
    """Concatenate per-rank result arrays in rank order on rank 0 (output assembly only; uses
    torch.distributed's object gather so it works on gloo and nccl alike)."""
    if world == 1:
        return local
    import torch.distributed as dist

    bufs = [None] * world if rank == 0 else None
    dist.gather_object(local, bufs, dst=0)
    return np.concatenate(bufs) if rank == 0 else None
