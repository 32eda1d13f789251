"""Multi-rank host logic on CPU (gloo, world_size 2): the pair list is sharded
round-robin with no data-path collective; the only exchange is the timing
reduction bench.py performs (max over ranks)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_pairs, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from posfeat_b200.pairs import shard
    mine = shard(range(n_pairs), rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # bench.py's reduction: elapsed time = max over ranks, throughput = all pairs / that
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put((gathered, float(t)))
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_sharding_two_ranks():
    world, n_pairs = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = sorted(x for part in gathered for x in part)
    assert flat == list(range(n_pairs))                       # complete, disjoint
    assert abs(len(gathered[0]) - len(gathered[1])) <= 1      # balanced
    assert gathered[0] == list(range(0, n_pairs, 2))
    assert tmax == 11.0


def test_shard_edge_cases():
    from posfeat_b200.pairs import shard
    assert shard([], 0, 4) == []
    assert shard(range(3), 3, 4) == []
    assert sum(len(shard(range(100), r, 8)) for r in range(8)) == 100
