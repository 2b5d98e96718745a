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


def ranks_agree(ok: bool, device, group=None) -> bool:
    """True iff EVERY rank of ``group`` passed ok=True (all-reduce MIN of a flag; also a synchronisation point of the
    ranks).  Used where a rank-local decision would deadlock the group: whether the peer-memory exchange could be set up
    (``GrapesEngine.enable_data_parallel``)."""
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(int(flag.item()))


def flatten_grads(named_grads) -> torch.Tensor:
    return torch.cat([g.reshape(-1) for _, g in sorted(named_grads.items())])


class PeerGradExchange:
    """Symmetric (peer-mapped) gradient buffers for ``grapes_step_tail`` / ``grapes_allreduce_adam_peer``: every rank
    allocates the same buffer with ``torch.distributed._symmetric_memory`` and receives the peers' device pointers, so the
    mean all-reduce and both Adam updates run as ONE kernel over NVLink inside the step's CUDA graph (no host-side
    collective call)."""

    def __init__(self, n_floats: int, device: torch.device, group=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        from ._lib import lib
        group = group or dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise RuntimeError("peer exchange is built for one NVSwitch box (<= 8 ranks)")
        total = int(lib().cdll.grapes_peer_buffer_floats(int(n_floats), self.world))
        self.buf = symm_mem.empty(total, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group.group_name if hasattr(group, "group_name") else group)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        assert ptrs[self.rank] == self.buf.data_ptr()
        self.peer_ptrs = (ctypes.c_void_p * self.world)(*ptrs)              # HOST array handed to the C ABI
        self.state = torch.zeros(int(lib().cdll.grapes_peer_state_words()), dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        # NOTE: no collective here.  The caller must synchronise the ranks once before the first exchange (every buffer is
        # zeroed before anyone publishes): GrapesEngine.enable_data_parallel does it with the all-reduce that also makes
        # the ranks agree on whether the peer mapping worked everywhere.
