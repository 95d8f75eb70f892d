"""CPU, world_size 2, gloo: the host logic of the N>1 path -- stream ownership, the single
init-time table broadcast, max-over-ranks rate aggregation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gr_lora_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gr_lora_b200 as G
    mine = sharding.shard_streams(384, world, rank)                  # config 4: 64 channels x 6 SF
    # every rank builds its own blob, then rank 0's replaces it: after the broadcast all are identical
    blob = G.tables_build_host(sf=9)
    ref = blob.copy()
    if rank != 0:
        blob[:] = 0                                                  # prove the bytes really travel
    sharding.broadcast_tables(blob, dist)
    same = bool(np.array_equal(blob, ref))
    rate = sharding.aggregate_rate(1000, 10.0 * (rank + 1), dist)    # slowest rank: 20 ms
    gathered = [None] * world
    dist.all_gather_object(gathered, mine.tolist())
    q.put((rank, same, rate, gathered))
    dist.destroy_process_group()


def test_two_rank_sharding_and_table_broadcast():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    out = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(30) for p in ps]
    for rank, same, rate, gathered in out:
        assert same, "table blob differs after the broadcast"
        assert rate == pytest.approx(2 * 1000 / 20e-3)
        all_ids = sorted(i for part in gathered for i in part)
        assert all_ids == list(range(384))                           # every stream owned exactly once
        assert len(gathered[0]) == len(gathered[1]) == 192
    # every stream costs the same bytes/s whatever its SF (8 MB/s at 1 MS/s), so balance = equal counts
    for g in (1, 2, 4, 8):
        counts = [sharding.shard_streams(384, g, r).size for r in range(g)]
        assert max(counts) - min(counts) <= 1 and sum(counts) == 384


def test_shard_edges():
    assert sharding.shard_streams(5, 8, 7).size == 0
    assert list(sharding.shard_streams(5, 2, 1)) == [1, 3]
    assert sharding.owner_of(13, 8) == 5
    with pytest.raises(ValueError):
        sharding.shard_streams(4, 2, 2)
