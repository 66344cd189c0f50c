"""Data-parallel training step: sequences are independent, so the batch is sharded by sequence across ranks
(one process per GPU) and the only exchange is one sum all-reduce of the flat gradient buffer (+ the loss
scalars riding at its end) and a 16-byte fp64 all-reduce for the physics constants.  SURVEY section 8(e).

The reference has no distributed code at all (single device, runners/torch_run_physics.py:78)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a global batch for `rank`: the first batch % world ranks get one extra
    sequence (100 over 8 -> 4 x 13 + 4 x 12)."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_step(flat_grad: torch.Tensor, phys_grad: torch.Tensor, group=None) -> None:
    """Sum the step's gradients (and the 4 loss scalars at the tail of flat_grad) over the job."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    if phys_grad is not None and phys_grad.numel():
        dist.all_reduce(phys_grad, op=dist.ReduceOp.SUM, group=group)


class DataParallelStep:
    """Wraps a PhysicsNet: `step(x_local)` runs the fused LIVE step on this rank's shard with the job-wide loss
    normalisation, then all-reduces.  Every rank ends with the gradient of the global-batch loss."""

    def __init__(self, net, global_batch: int, group=None):
        self.net = net
        self.group = group
        self.global_batch = int(global_batch)
        net.batch_global = self.global_batch

    def step(self, x_local: torch.Tensor) -> torch.Tensor:
        losses = self.net.train_step(x_local)
        allreduce_step(self.net.flat_gradients(), self.net._phys_grad, self.group)
        return losses
