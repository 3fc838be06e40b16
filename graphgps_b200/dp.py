"""Data-parallel plumbing for the GPSLayer hot path.

Graph mini-batches are independent units: every rank runs its own 256-graph batch through the full
layer with per-replica BatchNorm statistics (the reference has no SyncBN), and the only exchange is
one all-reduce (mean) of the parameter gradients per step over NCCL/NVLink.  The gradients of a
layer (13d^2+22d floats, 4.8 MB at d=304) are flattened into one bucket so the collective is a
single launch-latency-bound call.  The reference itself has no distributed code at all.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_graph_range(num_graphs: int, rank: int, world: int):
    """Graphs [lo, hi) of a global batch that rank `rank` owns (contiguous, balanced)."""
    base, rem = divmod(num_graphs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_gradients(params, bucket=None, group=None):
    """All-reduce (mean) the .grad of `params` through one flat bucket. Returns the bucket for reuse."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return bucket
    n = sum(g.numel() for g in grads)
    if bucket is None or bucket.numel() != n or bucket.device != grads[0].device:
        bucket = torch.empty(n, dtype=grads[0].dtype, device=grads[0].device)
    views, off = [], 0
    for g in grads:
        views.append(bucket[off:off + g.numel()].view_as(g))
        off += g.numel()
    torch._foreach_copy_(views, grads)
    world = dist.get_world_size(group)
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    bucket.mul_(1.0 / world)
    torch._foreach_copy_(grads, views)
    return bucket
