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


# ---- training configuration: one bucketed gradient all-reduce per step -------------------
def _grad_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from posfeat_b200.dist import GradAllReducer
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(12, 20), torch.nn.Tanh(), torch.nn.Linear(20, 3))
    red = GradAllReducer(model.parameters(), bucket_mb=0.0005)     # ~131 floats per bucket -> several buckets
    x = torch.randn(8, 12, generator=torch.Generator().manual_seed(5))
    y = torch.randn(8, 3, generator=torch.Generator().manual_seed(6))
    xs, ys = x[rank::world], y[rank::world]                          # this rank's shard of the batch
    for step in range(2):                                            # twice: buckets are reused, zeroed in between
        red.zero_()
        loss = ((model(xs) - ys) ** 2).sum() / x.shape[0] * world   # so that the mean over ranks is the full-batch loss
        loss.backward()
        red.start()
        red.finish()
    grads = [p.grad.clone() for p in model.parameters()]
    if rank == 0:
        out.put((len(red.buckets), red.nbytes, [g.numpy() for g in grads]))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_grad_allreduce_equals_full_batch_gradient():
    import numpy as np
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    nb, nbytes, grads = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(12, 20), torch.nn.Tanh(), torch.nn.Linear(20, 3))
    x = torch.randn(8, 12, generator=torch.Generator().manual_seed(5))
    y = torch.randn(8, 3, generator=torch.Generator().manual_seed(6))
    (((model(x) - y) ** 2).sum() / x.shape[0]).backward()
    assert nb >= 2 and nbytes == 4 * sum(p.numel() for p in model.parameters())
    for g, p in zip(grads, model.parameters()):
        np.testing.assert_allclose(g, p.grad.numpy(), rtol=1e-5, atol=1e-7)


def test_grad_allreducer_single_process_is_identity():
    from posfeat_b200.dist import GradAllReducer
    w = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7))]
    w[0].grad = torch.ones(5, 3)
    red = GradAllReducer(w, bucket_mb=1)
    assert len(red.buckets) == 1 and torch.equal(w[0].grad, torch.ones(5, 3))     # existing gradients are kept
    (w[0].sum() * 2 + (w[1] ** 2).sum()).backward()
    red.start()
    red.finish()
    assert torch.equal(w[0].grad, torch.full((5, 3), 3.0))
    assert torch.allclose(w[1].grad, 2 * w[1].detach())
    assert w[0].grad.data_ptr() >= red.buckets[0].data_ptr()                       # the gradient IS a view of the bucket


def test_shard_by_group_keeps_groups_whole_and_balanced():
    from posfeat_b200.dist import shard_by_group
    # HPatches-shaped: 116 sequences x 5 ref->target pairs; Aachen-shaped: queries with 20 retrieval pairs each
    pairs = [(f"seq{s}", k) for s in range(116) for k in range(2, 7)]
    parts = [shard_by_group(pairs, lambda p: p[0], r, 8) for r in range(8)]
    assert sorted(x for part in parts for x in part) == sorted(pairs)
    for part in parts:
        seqs = {p[0] for p in part}
        assert len(part) == 5 * len(seqs)                          # whole sequences only
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 5
    ragged = [("q%d" % q, k) for q in range(10) for k in range(1 + (q * 7) % 20)]
    parts = [shard_by_group(ragged, lambda p: p[0], r, 3) for r in range(3)]
    assert sorted(x for part in parts for x in part) == sorted(ragged)
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 20
    assert shard_by_group([], lambda p: p, 0, 4) == []
