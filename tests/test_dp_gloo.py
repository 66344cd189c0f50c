"""Data-parallel host logic on CPU: 2 processes over gloo.

Each rank takes its shard of a global batch, forms the gradient of the *globally normalised* loss on the shard
(here with the oracle standing in for the CUDA step, which needs a GPU), and `parallel.allreduce_step` sums the
flat buffer.  The result must equal the full-batch gradient, and the loss scalars riding at the buffer's tail must
sum to the full-batch losses -- the contract `paig_task.batch_global` + one all-reduce implements on the GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import physicsnet_oracle as po
from paig_reproduction_b200.parallel import allreduce_step, shard_bounds

TASK, B, ALPHA = "spring_color", 5, 3.0      # 5 over 2 ranks -> shards of 3 and 2 (uneven on purpose)


def test_shard_bounds_partition():
    for batch, world in [(100, 8), (100, 1), (5, 2), (7, 8), (8192, 8)]:
        cuts = [shard_bounds(batch, world, r) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == batch
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in cuts]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert [hi - lo for lo, hi in (shard_bounds(100, 8, r) for r in range(8))] == [13] * 4 + [12] * 4


def _flat(grads, keys):
    return torch.cat([grads[k].reshape(-1).float() for k in keys])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        spec = po.TASKS[TASK]
        sd = po.init_state_dict(spec, 0)
        x = po.synthetic_frames(spec, B, spec.seq_len, 0)
        lo, hi = shard_bounds(B, world, rank)
        _, ls, grads = po.live_step(sd, x[lo:hi], spec, ALPHA)
        scale = (hi - lo) / B                       # means over the shard -> means over the global batch
        f32 = sorted(k for k, g in grads.items() if g.dtype == torch.float32)
        f64 = sorted(k for k, g in grads.items() if g.dtype == torch.float64)
        flat = torch.cat([_flat(grads, f32) * scale,
                          torch.stack([ls[k].detach().float() * scale for k in ("train", "pred", "extrap", "recons")])])
        phys = torch.stack([grads[k].reshape(()) for k in f64]) * scale
        allreduce_step(flat, phys)
        if rank == 0:
            torch.save({"flat": flat, "phys": phys, "f32": f32, "f64": f64}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_step_allreduce_equals_full_batch(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    spec = po.TASKS[TASK]
    sd = po.init_state_dict(spec, 0)
    x = po.synthetic_frames(spec, B, spec.seq_len, 0)
    _, ls, grads = po.live_step(sd, x, spec, ALPHA)
    want = torch.cat([_flat(grads, got["f32"]),
                      torch.stack([ls[k].detach().float() for k in ("train", "pred", "extrap", "recons")])])
    err = (got["flat"] - want).abs().max() / want.abs().max()
    assert err < 1e-5, err
    want64 = torch.stack([grads[k].reshape(()) for k in got["f64"]])
    # dk is a small difference of large per-sequence terms accumulated in fp32: bound it by the vector's scale
    assert (got["phys"] - want64).abs().max() < 1e-3 * want64.abs().max(), (got["phys"], want64)


def test_allreduce_step_is_a_noop_without_a_process_group():
    flat = torch.arange(4.0)
    allreduce_step(flat, torch.zeros(2, dtype=torch.float64))
    assert flat.tolist() == [0.0, 1.0, 2.0, 3.0]
