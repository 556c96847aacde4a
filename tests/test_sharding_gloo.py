"""CPU, world_size 2 over gloo: the per-connection sharding and the reporting reductions of the multi-GPU path."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from inference_gateway_b200 import shard as sh


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = sh.shard_connections(1000, world, rank)
    assert len(ids) == 1000 and np.all(sh.shard_of(ids, world) == rank)
    tot = sh.reduce_counters({"frames": 10 + rank, "bytes": 100 * (rank + 1)}, world)
    mx = sh.max_over_ranks(1.5 + rank, world)
    q.put((rank, ids[:50].tolist(), tot, mx))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    got.sort()
    assert set(got[0][1]).isdisjoint(got[1][1])
    for _, _, tot, mx in got:
        assert tot == {"frames": 21, "bytes": 300} and mx == 2.5


def test_hash_is_stable_and_balanced():
    ids = np.arange(1 << 16, dtype=np.uint64)
    for n in (2, 4, 8):
        s = sh.shard_of(ids, n)
        counts = np.bincount(s, minlength=n)
        assert counts.min() > 0.9 * len(ids) / n and counts.max() < 1.1 * len(ids) / n
        assert np.array_equal(s, sh.shard_of(ids, n))
    assert sh.reduce_counters({"a": 3}, 1) == {"a": 3} and sh.max_over_ranks(2.0, 1) == 2.0
