"""Training-step throughput (LIVE fused step, fwd+bwd) for every BASELINE.json task, on 1 GPU or data-parallel:

    python tools/bench_tasks.py [tasks...] [--batch 100] [--scaling weak|strong] [--steps 10] [--profile]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_tasks.py 3bp_color --scaling strong          # BASELINE config 3: global batch 100 over N GPUs

weak: `--batch` sequences per GPU; strong: `--batch` is the job's batch, sharded 13/12/... per rank
(parallel.shard_bounds).  One JSON line per task on rank 0 (CUDA events, max over ranks); spring_color weak is what
bench.py reports.  --profile adds the per-kernel breakdown (rank 0, an instrumented pass after the timed one)."""
import argparse, ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from oracle import physicsnet_oracle as po
from paig_reproduction_b200.physics_models import PhysicsNet
from paig_reproduction_b200.parallel import DataParallelStep, shard_bounds
from paig_reproduction_b200 import _lib

ALPHA = {"spring_color": 3.0, "bouncing_balls": 2.0, "3bp_color": 5.0, "mnist_spring_color": 3.0}
ap = argparse.ArgumentParser()
ap.add_argument("tasks", nargs="*", default=["spring_color", "bouncing_balls", "3bp_color", "mnist_spring_color"])
ap.add_argument("--batch", type=int, default=100)
ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--profile", action="store_true")
a = ap.parse_args()
os.environ.setdefault("NCCL_DEBUG", "WARN")
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
for task in a.tasks:
    spec = po.TASKS[task]
    net = PhysicsNet(task, 100, 1, po.CELL_TYPE_NAMES[spec.cell], spec.seq_len, spec.input_steps, spec.pred_steps, ALPHA[task],
                     False, True, spec.H * spec.H, "conv_encoder", "conv_st_decoder", device=dev)
    net.load_state_dict(po.init_state_dict(spec, 0), strict=True)
    if a.scaling == "weak":
        b_local, b_global = a.batch, a.batch * world
    else:
        lo, hi = shard_bounds(a.batch, world, rank)
        b_local, b_global = hi - lo, a.batch
    dp = DataParallelStep(net, b_global)
    pool = [torch.rand(b_local, spec.seq_len, 3, spec.H, spec.H, device=dev) for _ in range(4)]

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
    for i in range(3):
        dp.step(pool[i % 4])
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        dp.step(pool[i % 4])
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    line = {"task": task, "H": spec.H, "n_objs": spec.n_objs, "T": spec.seq_len, "n_gpus": world, "scaling": a.scaling,
            "global_batch": b_global, "batch_this_rank": b_local, "ms_per_step": ms, "sequences_per_s": b_global / ms * 1e3,
            "allreduce": ("overlapped" if dp.overlap else "in line") if world > 1 else None}
    if a.profile and rank == 0:
        lib.paig_profile_begin()                  # per-launch CUDA events: an instrumented pass after the timed one
        for i in range(4):
            net.train_step(pool[i % 4])
        buf = ctypes.create_string_buffer(1 << 16)
        _lib.check(lib.paig_profile_end(buf, len(buf)))
        kern = {}
        for ln in buf.value.decode().strip().splitlines():
            name, cnt, tot = ln.rsplit(" ", 2)
            kern[name] = round(float(tot) / 4, 4)
        line["kernels_ms_per_step"] = dict(sorted(kern.items(), key=lambda kv: -kv[1]))
    if rank == 0:
        print(json.dumps(line), flush=True)
    sync()
    del net, pool, dp
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
