"""Data-parallel training step: sequences are independent, so the batch is sharded by sequence across ranks
(one process per GPU) and the only exchange is the sum all-reduce of the gradients.  SURVEY section 8(e).

The flat gradient buffer is ordered [everything but the UNet conv layers | UNet conv layers | 4 loss scalars]
(``PhysicsNet.flat_order``).  The first part -- VariableFromNetworks, encoder MLP, velocity MLP: 98 % of the bytes -- is
final before the UNet backward starts; the library records an event at that point (``paig_set_early_grad_event``) and
with ``overlap=True`` the all-reduce of that prefix, and of the 16-byte fp64 physics gradients, is launched on a side
stream behind the event, underneath the UNet backward and weight-gradient kernels, leaving only the 34 k conv gradients
+ losses (138 KB) for after the step.

MEASURED (2 x B200, profiles/r2i_bench_n2*.json): the overlapped form is SLOWER -- 3.31 ms per step against 2.75 ms
with one all-reduce after the step (2.71 ms on one GPU).  The UNet backward is a persistent kernel that wants every SM
(one 512-thread CTA with 214 KB of shared memory each); the NCCL kernel's CTAs take some of them first, the frames those
SMs own start late, and both ranks then wait on each other's late kernels.  So the default is the plain all-reduce
after the step (98.3 % weak-scaling efficiency at N = 2); the overlap path stays selectable (PAIG_DP_OVERLAP=1) for
configurations whose backward does not fill the GPU.

The reference has no distributed code at all (single device, runners/torch_run_physics.py:78)."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a global batch for `rank`: the first batch % world ranks get one extra
    sequence (100 over 8 -> 4 x 13 + 4 x 12)."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _active(group) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def allreduce_step(flat_grad: torch.Tensor, phys_grad: torch.Tensor, group=None) -> None:
    """Sum the step's gradients (and the 4 loss scalars at the tail of flat_grad) over the job, in line."""
    if not _active(group):
        return
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    if phys_grad is not None and phys_grad.numel():
        dist.all_reduce(phys_grad, op=dist.ReduceOp.SUM, group=group)


class DataParallelStep:
    """Wraps a PhysicsNet: `step(x_local)` runs the fused LIVE step on this rank's shard with the job-wide loss
    normalisation and all-reduces.  Every rank ends with the gradient of the global-batch loss.  With `overlap` the bulk
    of the all-reduce runs underneath the UNet backward (off by default: measured slower, module docstring)."""

    def __init__(self, net, global_batch: int, group=None, overlap: bool | None = None):
        self.net = net
        self.group = group
        self.global_batch = int(global_batch)
        net.batch_global = self.global_batch
        if overlap is None:
            overlap = os.environ.get("PAIG_DP_OVERLAP", "0") == "1"
        self.overlap = bool(overlap) and net.device.type == "cuda"
        self._side = self._event = None

    def _setup(self):
        self._side = torch.cuda.Stream(self.net.device)
        self._event = torch.cuda.Event()
        self._event.record(torch.cuda.current_stream(self.net.device))      # materialise the cudaEvent_t behind it

    def arm(self) -> bool:
        """Before enqueuing a fused step: arm the early-gradient event.  False: the all-reduce will run in line."""
        if not (self.overlap and _active(self.group)):
            return False
        from . import _lib
        if self._side is None:
            self._setup()
        self.net.flat_gradients()
        _lib.load().paig_set_early_grad_event(self._event.cuda_event)
        return True

    def reduce(self, armed: bool) -> None:
        """After the step was enqueued: launch the all-reduces (the early part behind the event on the side stream)."""
        net = self.net
        flat = net.flat_gradients()
        if not armed:
            allreduce_step(flat, net._phys_grad, self.group)
            return
        main = torch.cuda.current_stream(net.device)
        self._side.wait_event(self._event)
        with torch.cuda.stream(self._side):
            dist.all_reduce(flat[:net.flat_early], op=dist.ReduceOp.SUM, group=self.group)
            if net._phys_grad.numel():
                dist.all_reduce(net._phys_grad, op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(flat[net.flat_early:], op=dist.ReduceOp.SUM, group=self.group)    # UNet conv gradients + losses
        main.wait_stream(self._side)

    def step(self, x_local: torch.Tensor) -> torch.Tensor:
        armed = self.arm()
        losses = self.net.train_step(x_local)             # enqueues the whole step; the library records the event mid-way
        self.reduce(armed)
        return losses
