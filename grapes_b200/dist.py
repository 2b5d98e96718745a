"""Data-parallel plumbing (new work defined by BASELINE.json; the reference is single-process).

One process per GPU, CSR + features replicated, target-node batches sharded by index, and ONE collective per
step: the mean all-reduce of the flat gradient buffer (gcn_c | gcn_gf | gcn_z; 91 k floats on products-shape).
``loss_c.detach()`` inside the GFlowNet loss is rank-local (main.py:274), so nothing else is exchanged.
The trajectory-balance loss is the square of a per-batch scalar, so W-rank data parallelism equals averaging
W per-batch gradients -- not one W x B batch."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_batches(num_batches: int, rank: int, world: int) -> List[int]:
    """Batch i of the un-shuffled loader (main.py:125-126) goes to rank i mod W; every rank runs the same
    number of steps (the tail that does not fill a full round is dropped so no rank waits at the all-reduce)."""
    rounds = num_batches // world
    return [r * world + rank for r in range(rounds)]


def allreduce_mean_(flat: torch.Tensor) -> torch.Tensor:
    """In-place mean over ranks of a flat buffer (NCCL: one ReduceOp.AVG call; gloo: SUM then divide)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return flat
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
    return flat


def flatten_grads(named_grads) -> torch.Tensor:
    return torch.cat([g.reshape(-1) for _, g in sorted(named_grads.items())])
