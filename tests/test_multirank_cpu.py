"""N>1 host logic on CPU: world_size-2 gloo group, stream sharding and the max-over-ranks timing reduction."""
import os
import sys

import pytest

from conftest import ROOT, load_pkg


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import importlib

    import torch.distributed as dist

    sh = importlib.import_module("gd-slam_b200.sharding")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sh.shard_streams(8, world, rank)
    dist.barrier()
    elapsed = 1.0 + rank  # rank 1 is the slow one
    mx = sh.max_over_ranks(elapsed, dist)
    rate = sh.aggregate_rate(units_per_rank=len(mine) * 10, world=world, max_seconds=mx)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, mine, mx, rate))


def test_two_rank_gloo_sharding_and_reduction():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, m0, a0), (r1, s1, m1, a1) = res
    assert s0 == [0, 2, 4, 6] and s1 == [1, 3, 5, 7]  # stream s -> rank s mod G, disjoint and complete
    assert m0 == m1 == 2.0  # max over ranks, identical on both
    assert a0 == a1 == 4 * 10 * 2 / 2.0


def test_sharding_edges():
    sh = load_pkg("sharding")
    assert sh.shard_streams(3, 8, 5) == []  # more GPUs than streams: idle rank
    assert sh.shard_streams(64, 8, 7) == list(range(7, 64, 8))
    assert sh.stream_seed(3, 2, 32) == 98
    with pytest.raises(ValueError):
        sh.shard_streams(4, 2, 2)
    assert sh.max_over_ranks(1.5) == 1.5
