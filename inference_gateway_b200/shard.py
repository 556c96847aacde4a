"""Per-connection sharding across the GPUs of one box.

Connections are independent units: gpu = hash(conn_id) % n_gpu, stable for the connection's lifetime so that the
carry state and the per-connection FIFO stay on one device. No collective is needed on the data path; the only
cross-GPU quantities are a handful of counters summed for reporting.
"""
from __future__ import annotations

import numpy as np


def shard_of(conn_id, n_shards: int):
    """Fibonacci hash of the connection id (works on ints and numpy arrays)."""
    h = (np.asarray(conn_id, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(33)
    return (h % np.uint64(n_shards)).astype(np.int64)


def shard_connections(n_per_shard: int, n_shards: int, shard: int) -> np.ndarray:
    """Global ids of the first n_per_shard connections that hash to `shard` (weak scaling: fixed work per GPU)."""
    out = []
    got = 0
    lo = 0
    step = max(1024, n_per_shard * n_shards)
    while got < n_per_shard:
        ids = np.arange(lo, lo + step, dtype=np.uint64)
        mine = ids[shard_of(ids, n_shards) == shard]
        out.append(mine)
        got += len(mine)
        lo += step
    return np.concatenate(out)[:n_per_shard]


def reduce_counters(local: dict, world_size: int) -> dict:
    """Sums per-rank integer counters over the process group (gloo on CPU, nccl on GPU); identity at world 1."""
    if world_size == 1:
        return dict(local)
    import torch
    import torch.distributed as dist
    keys = sorted(local)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([int(local[k]) for k in keys], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: int(v) for k, v in zip(keys, t.tolist())}


def max_over_ranks(value: float, world_size: int) -> float:
    if world_size == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
