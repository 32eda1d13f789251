"""Multi-GPU plumbing for the two ways the path scales:

* extraction / matching shard with NO collective: ``shard`` (static round-robin), ``shard_by_group`` (HPatches:
  by sequence, so a rank extracts a reference image once; Aachen: by query, so its 20 retrieval pairs share one
  query extraction) and ``gather_objects`` (results to rank 0 over the host -- file lists / match counts, not
  tensors);
* the training configuration needs exactly one collective per step: the sum of the parameter gradients.
  ``GradAllReducer`` keeps the gradients in a few large flat buckets (the views ARE the ``.grad`` tensors, so no
  copy in or out) and launches one NCCL all-reduce per bucket on a side stream, so that the reduction of what the
  backward pass has finished runs under what it is still computing.  No SyncBatchNorm, no unused-parameter graph
  walk: the two things the reference's DDP set-up (networks/PoSFeat_model.py:48-55) pays for on every step.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def shard(items: Sequence, rank: int, world: int) -> list:
    """Static round-robin sharding of an image / pair list (no communication)."""
    return list(items)[rank::world]


def shard_by_group(items: Sequence, key, rank: int, world: int) -> list:
    """Shard whole groups: all items with the same ``key(item)`` (HPatches sequence, Aachen query) go to one rank;
    groups are dealt largest first to the least loaded rank (deterministic: ties by group key)."""
    groups = {}
    for it in items:
        groups.setdefault(key(it), []).append(it)
    load = [0] * world
    mine = []
    for k in sorted(groups, key=lambda g: (-len(groups[g]), str(g))):
        r = min(range(world), key=lambda i: (load[i], i))
        load[r] += len(groups[k])
        if r == rank:
            mine.extend(groups[k])
    return mine


def gather_objects(obj, dst: int = 0, group=None):
    """Python objects (per-rank result summaries) to rank ``dst``; returns the list there, None elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [obj]
    out = [None] * dist.get_world_size(group) if dist.get_rank(group) == dst else None
    dist.gather_object(obj, out, dst=dst, group=group)
    return out


class GradAllReducer:
    """Bucketed gradient all-reduce (sum, then divide by the world size) for data-parallel training.

    ``params``: the parameters (or plain tensors standing for their gradients).  Their gradients live in flat
    buckets of at most ``bucket_mb`` megabytes, filled in REVERSE parameter order -- the order a backward pass
    produces them in.  ``start()`` queues the all-reduces of all buckets on the side stream behind everything
    already queued on the current stream; ``start_bucket(i)`` does so for one bucket (call it from a gradient
    hook as soon as the bucket is complete); ``finish()`` makes the current stream wait for them and applies
    the 1/world scale.  On the CPU (gloo) the collectives are issued asynchronously and waited for in finish().
    """

    def __init__(self, params: Iterable[torch.Tensor], bucket_mb: float = 25.0, group=None):
        self.group = group
        self.params: List[torch.Tensor] = [p for p in params]
        if not self.params:
            raise ValueError("GradAllReducer needs at least one parameter")
        dev = self.params[0].device
        dt = self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise ValueError("all parameters must share one device and dtype")
        self.device = dev
        cap = max(int(bucket_mb * 2 ** 20) // self.params[0].element_size(), 1)
        self.buckets: List[torch.Tensor] = []
        self.bucket_of = {}
        cur, cur_n = [], 0
        plan = []
        for p in reversed(self.params):
            if cur and cur_n + p.numel() > cap:
                plan.append(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        plan.append(cur)
        for bi, ps in enumerate(plan):
            flat = torch.zeros(sum(p.numel() for p in ps), dtype=dt, device=dev)
            off = 0
            for p in ps:
                view = flat[off:off + p.numel()].view_as(p)
                if p.grad is not None:
                    view.copy_(p.grad)
                p.grad = view                     # autograd accumulates into the bucket in place
                self.bucket_of[id(p)] = bi
                off += p.numel()
            self.buckets.append(flat)
        self.side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._pending = []

    @property
    def nbytes(self) -> int:
        return sum(b.numel() * b.element_size() for b in self.buckets)

    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def start_bucket(self, i: int):
        if self.world() == 1:
            return
        b = self.buckets[i]
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream(self.device))     # the gradients queued so far are final
            with torch.cuda.stream(self.side):
                self._pending.append(dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self._pending.append(dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def start(self):
        for i in range(len(self.buckets)):
            self.start_bucket(i)

    def finish(self):
        w = self.world()
        for work in self._pending:
            work.wait()                       # NCCL: makes the calling stream wait; gloo: blocks the host
        self._pending = []
        if self.side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.side)
        if w > 1:
            for b in self.buckets:
                b.div_(w)

    def zero_(self):
        for b in self.buckets:
            b.zero_()
