"""CPU suite: the N>1 plumbing (sharding + reductions) with world_size-2 gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rabbitsalign_b200 import sharding, workload as W


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1000, 1048576):
        for world in (1, 2, 3, 4, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = W.extension_pairs_fast(2000, seed=5)
    lo, hi = sharding.shard_range(b.n, rank, world)
    shard = b.slice(lo, hi)
    dist.barrier()
    t, c = sharding.reduce_step(10.0 + 5.0 * rank, float(shard.cells), dist)
    out[rank] = (t, c, lo, hi, float(b.cells), sharding.rank_seed(43, rank))
    dist.destroy_process_group()


def test_two_rank_reduction_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    (t0, c0, lo0, hi0, total, s0), (t1, c1, lo1, hi1, _, s1) = out[0], out[1]
    assert t0 == t1 == 15.0                # max over ranks
    assert c0 == c1 == total               # the shards' cells add up to the whole batch
    assert (lo0, hi0, lo1, hi1) == (0, 1000, 1000, 2000)
    assert s0 != s1
