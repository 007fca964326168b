"""Batch sharding helpers (one process per GPU).

Graphs in a batch are independent (reference gcm.py:274-314 indexes everything by `b`), so the hot path
shards over the batch with NO data-path collective: every rank owns a contiguous slice of the graphs and
their hidden state never leaves its GPU.  Two places exchange data:
  * training: one all-reduce (sum) of the GNN weight gradients per optimiser step (`allreduce_grads`);
  * EuclideanEdge only: the reference's distance averages over the current observation of EVERY graph in
    the batch (edge_selectors/distance.py:48-49), so the ranks all-gather x before the step
    (`gather_current_obs`).
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the transport."""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a global batch owned by `rank`; sizes differ by at most one."""
    assert 0 <= rank < world
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t: torch.Tensor, rank: int, world: int, dim: int = 0) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[dim], rank, world)
    return t.narrow(dim, lo, hi - lo)


def allreduce_grads(params: Iterable[torch.nn.Parameter], group=None, average: bool = False) -> int:
    """One flattened all-reduce over every parameter gradient (a few hundred KB at most: latency-bound,
    so a single bucket).  Returns the number of elements reduced."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sum(g.numel() for g in grads)
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        g.copy_(flat[off: off + g.numel()].view_as(g))
        off += g.numel()
    return flat.numel()


def gather_current_obs(x: torch.Tensor, group=None, sizes=None) -> torch.Tensor:
    """All ranks' current observations [sum_r B_r, F] in batch order (EuclideanEdge under batch sharding).
    sizes: the per-rank batch sizes when the caller already knows them (DenseGCM caches them: a rollout's shards do not
    change size), which spares the size exchange and its host sync; equal shards take a single all_gather into one
    buffer."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    world = dist.get_world_size(group)
    if sizes is None:
        sizes = shard_sizes(x.shape[0], x.device, group)
    if len(set(sizes)) == 1:
        out = x.new_empty(world * x.shape[0], x.shape[1])
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        return out
    cap = max(sizes)                                   # all_gather needs equal shapes: pad, then trim
    mine = x.new_zeros(cap, x.shape[1])
    mine[: x.shape[0]] = x
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def shard_sizes(local_batch: int, device, group=None):
    """Every rank's local batch size (one small all_gather + host read; callers cache the result)."""
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.long, device=device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local_batch], dtype=torch.long, device=device), group=group)
    return [int(s) for s in sizes]
