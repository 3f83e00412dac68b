"""Batch-sharded multi-GPU inference.

The forward pass is independent per image and the packed weights are tiny (1.4 MB for ResNet-18),
so the path shards by batch: one process per GPU, rank r runs images [lo_r, hi_r) through the
single-GPU engine, and the ONLY collective is one all-gather of the logits (NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests), issued on the compute stream right
behind the last kernel -- no host synchronisation in between.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shards; the first ``total % world`` ranks get one extra image."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_logits(local: torch.Tensor, total: Optional[int] = None, group=None) -> torch.Tensor:
    """All-gather per-rank logits [n_r, classes] into [total, classes] on every rank."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    if total is None or total % world == 0:
        out = local.new_empty((local.shape[0] * world,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # ragged shards: pad to the largest shard, gather, trim
    biggest = -(-total // world)
    padded = local.new_zeros((biggest,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    out = local.new_empty((biggest * world,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded, group=group)
    pieces = []
    for r in range(world):
        lo, hi = shard_bounds(total, r, world)
        pieces.append(out[r * biggest: r * biggest + (hi - lo)])
    return torch.cat(pieces, 0)


class ShardedInference:
    """``engine(global_batch)`` -> logits for the whole batch on every rank."""

    def __init__(self, model: torch.nn.Module, group=None) -> None:
        self.model, self.group = model, group

    @torch.no_grad()
    def run_local(self, local_batch: torch.Tensor, total: Optional[int] = None) -> torch.Tensor:
        return gather_logits(self.model(local_batch), total, self.group)

    @torch.no_grad()
    def __call__(self, global_batch: torch.Tensor) -> torch.Tensor:
        local = shard_batch(global_batch, dist.get_rank(self.group), dist.get_world_size(self.group))
        return self.run_local(local, global_batch.shape[0])
