"""CPU, world_size 2 over gloo: the N>1 host logic (stream ownership, max-over-ranks timing, counter gather)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from satdump_b200 import shard


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard.streams_of_rank(8, rank, world)
        tmax = shard.max_over_ranks([1.0 + rank, 5.0 - rank])
        thr = shard.aggregate_throughput(1000, 4, 2.0 + rank)
        cnt = shard.gather_counters([rank, 10 * rank + 1])
        ok = shard.all_true(rank == 0)
        q.put((rank, mine, tmax, thr, cnt, ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5, 7]
    for r in res:
        assert r[2] == [2.0, 5.0]                    # max over ranks
        assert abs(r[3] - 8000 / 3.0) < 1e-9         # all samples / slowest rank
        assert r[4] == [[0, 1], [1, 11]]
        assert r[5] is False


def test_single_process_defaults():
    assert shard.world() == 1
    assert shard.max_over_ranks([3.0]) == [3.0]
    assert shard.stream_seed(3, 2) == 0xB2000000 + 50
    assert shard.gather_counters([4, 5]) == [[4, 5]]
