"""BASELINE config 5: spring_color_half long-horizon evaluation (test-mode rollout, T=30, no_grad) with the batch
swept 100..8192, on one GPU or sharded by sequence over the ranks of a torchrun job (evaluation needs no collective
on the data path; the three loss sums are all-reduced at the end).  One JSON line per batch size on rank 0:
sequences/s of `net.output = net(x); net.compute_loss()` -- the call eval_performance makes (base.py:195-196) -- with
every frame materialised, max over ranks.
    python tools/eval_sweep.py [--batches 100,512,2048,8192]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P tools/eval_sweep.py"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from oracle import physicsnet_oracle as po
from paig_reproduction_b200.physics_models import PhysicsNet
from paig_reproduction_b200.parallel import shard_bounds

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="100,512,2048,8192")
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
os.environ.setdefault("NCCL_DEBUG", "WARN")
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
spec = po.TASKS["spring_color_half"]
T = 30
BYTES_PER_SEQ = T * 3 * 32 * 32 * 4 + 36 * 3 * 32 * 32 * 4       # input read once + 36 decoded frames written
net = PhysicsNet("spring_color_half", 100, 1, "spring_ode_cell", T, spec.input_steps, spec.pred_steps, 3.0, False, True,
                 32 * 32, "conv_encoder", "conv_st_decoder", device=dev)
net.load_state_dict(po.init_state_dict(spec, 0), strict=True)
net.eval()
try:
    HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except (OSError, KeyError, ValueError):
    HBM = 6650.0
for B in [int(b) for b in a.batches.split(",")]:
    lo, hi = shard_bounds(B, world, rank)
    net.batch_global = B
    x = torch.rand(hi - lo, T, 3, 32, 32, device=dev)
    torch.cuda.reset_peak_memory_stats(dev)
    with torch.no_grad():
        for _ in range(2):
            net.output = net(x); net.compute_loss()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            net.output = net(x)
            _, evals = net.compute_loss()
            if world > 1:
                sums = torch.stack([v.detach().float() for v in evals])
                dist.all_reduce(sums)
        e1.record(); torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if rank == 0:
        sps = B / ms * 1e3
        print(json.dumps({"workload": "spring_color_half test-mode rollout T=30 (36 decoded frames/seq, frames materialised, inference plan)",
                          "n_gpus": world, "batch": B, "batch_this_rank": hi - lo, "ms_per_batch": ms, "sequences_per_s": sps,
                          "hbm_fraction_per_gpu": sps / world * BYTES_PER_SEQ / (HBM * 1e9),
                          "peak_mem_gb_this_rank": torch.cuda.max_memory_allocated(dev) / 2**30}), flush=True)
    del x
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
