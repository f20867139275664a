"""CPU test of the multi-GPU host logic: world_size-2 gloo run of the tile scheduler."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hcore_b200.scheduler import TileScheduler


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ntiles, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sch = TileScheduler(ntiles)
        assert (sch.world, sch.rank) == (world, rank)
        # every rank "computes" rank = 10 + global tile index, status = tile % 3 on its shard only
        tiles = torch.arange(sch.first, sch.first + sch.count, dtype=torch.int32)
        ranks, stats = sch.gather_results(10 + tiles, tiles % 3)
        q.put((rank, sch.first, sch.count, ranks.numpy().copy(), stats.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_scheduler_world2_gloo():
    world, ntiles = 2, 13          # uneven split: 7 + 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ntiles, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 7), (7, 6)]
    for r in res:
        np.testing.assert_array_equal(r[3], 10 + np.arange(ntiles))
        np.testing.assert_array_equal(r[4], np.arange(ntiles) % 3)


def test_scheduler_single_process():
    sch = TileScheduler(10, world=1, rank=0)
    assert (sch.first, sch.count) == (0, 10)
    a, b = sch.gather_results(torch.arange(10, dtype=torch.int32), torch.zeros(10, dtype=torch.int32))
    assert a.tolist() == list(range(10))
    assert [sch.owner(t) for t in (0, 9)] == [0, 0]
    s4 = TileScheduler(10, world=4, rank=2)
    assert (s4.first, s4.count) == (6, 2) and s4.owner(5) == 1 and s4.owner(9) == 3
