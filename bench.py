"""bench.py -- train sequences/sec (fwd+bwd, LIVE step) of the PhysicsNet hot path, spring_color, batch 100 per GPU.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on rank 0 (contract: task prompt + BASELINE.json):
  value     whole-job sequences/s with the inputs resident in HBM (CUDA events, max over ranks, barrier + sync)
  e2e       the same metric through paig_step_fused_host: pinned HOST input copied in, 4 loss scalars copied out
  roofline  the dominant kernel's achieved TFLOP/s from live per-launch CUDA-event timing vs the measured peak
  cpu_baseline  the oracle (CPU restatement of the reference step) timed on this box's host cores (rank 0, N=1)
`--impl reference` times that CPU path alone (the reference is pure Python and does not travel to the GPU box;
the oracle is its pinned restatement -- DESIGN.md).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TASK = "spring_color"
B_PER_GPU = 100
ALPHA = 3.0                      # README.md:63-67 value for spring_color
BYTES_PER_SEQ = 147456           # T*3*H*H*4, input read once
FLOPS_PER_SEQ = 696.2e6          # GEMM+conv FLOPs fwd+bwd per sequence
FP32_FMA_NOMINAL_TFLOPS = 74.4   # 148 SM x 128 lanes x 2 x 1.965 GHz


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.rows, self.stop_flag, self.index = [], False, index

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def cpu_leg(steps, warmup, budget_s=25.0):
    """The reference's own LIVE step (net.output = net(x); compute_loss(); backward()) on all host threads, through its
    public API, from the ignored copy oracle/_ref/reference (oracle/fetch_reference.py); when that copy is absent, the
    oracle port.  Returns (seq/s, cores, sample description, seconds per step, kind)."""
    import torch

    from oracle import fetch_reference
    from oracle import physicsnet_oracle as po
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = po.TASKS[TASK]
    sd = po.init_state_dict(spec, 0)
    x = po.synthetic_frames(spec, B_PER_GPU, spec.seq_len, 0)
    kind = "port"
    if fetch_reference.reference_path() is not None:
        try:
            pm = fetch_reference.import_reference()
            net = pm.PhysicsNet(TASK, 100, 1, "spring_ode_cell", spec.seq_len, spec.input_steps, spec.pred_steps, ALPHA,
                                False, True, spec.H * spec.H, "conv_encoder", "conv_st_decoder")
            net.load_state_dict(sd, strict=True)
            kind = "reference"
        except Exception as e:                      # e.g. a dependency of the reference missing on this box
            sys.stderr.write("bench: reference import failed (%s: %s); timing the oracle port\n" % (type(e).__name__, e))

    def one_step():
        if kind == "reference":
            net.zero_grad(set_to_none=True)
            inp = x.clone().requires_grad_(True)    # base.py:141
            net.output = net(inp)                   # LIVE: what base.py:195 does in eval
            loss, _ = net.compute_loss()
            loss.backward()
        else:
            po.live_step(sd, x, spec, ALPHA)

    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one_step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > budget_s and len(times) >= 2:
            break
    med = statistics.median(times)
    what = "the unmodified reference (oracle/_ref)" if kind == "reference" else "the oracle port"
    return (B_PER_GPU / med, cores, "%d timed LIVE steps of %s B=%d through %s (median), %d warm-up"
            % (len(times), TASK, B_PER_GPU, what, warmup), med, kind)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(2, min(args.steps, 8))
    val, cores, sample, med, kind = cpu_leg(steps, max(1, min(args.warmup, 2)), budget_s=90.0)
    line = {"impl": "reference", "metric": "train sequences/sec (fwd+bwd, spring_color, bs=100)", "value": val,
            "unit": "sequences/s", "n_gpus": args.gpus, "steps": steps, "warmup": max(1, min(args.warmup, 2)),
            "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "spring_color PhysicsNet LIVE training step (fwd+bwd, all parameter gradients), "
                                   "2 objects, 32x32 RGB, T=12, batch 100, on the host CPU", "task": TASK,
                       "batch": B_PER_GPU, "note": "the reference's own PhysicsNet through its public API "
                                                   "(net(x); compute_loss(); backward()) on all host threads"
                       if kind == "reference" else "oracle port (reference copy absent on this box)"},
            "cpu_baseline": {"value": val, "unit": "sequences/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sustained / drop-in / strong-scaling extra keys")
    args = ap.parse_args()
    import warnings
    warnings.filterwarnings("ignore")
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)

    os.environ.setdefault("NCCL_DEBUG", "WARN")         # keep stdout to the one JSON line
    import torch
    import torch.distributed as dist

    from oracle import physicsnet_oracle as po          # only for the synthetic inputs / weights and the cpu_baseline leg
    from paig_reproduction_b200 import _abi, _lib
    from paig_reproduction_b200.parallel import DataParallelStep, allreduce_step
    from paig_reproduction_b200.physics_models import PhysicsNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    spec = po.TASKS[TASK]
    T, H = spec.seq_len, spec.H
    net = PhysicsNet(TASK, 100, 1, "spring_ode_cell", T, spec.input_steps, spec.pred_steps, ALPHA, False, True, H * H,
                     "conv_encoder", "conv_st_decoder", device=dev)
    net.load_state_dict(po.init_state_dict(spec, 0), strict=True)
    net.batch_global = B_PER_GPU * world            # loss means are over the job's batch (weak scaling: 100 per GPU)
    # inputs rotate over a pool larger than L2 (126 MB): 12 x 14.7 MB device batches, same pool pinned on the host
    POOL = 12
    g = torch.Generator().manual_seed(100 + rank)
    host_pool = [torch.rand(B_PER_GPU, T, 3, H, H, generator=g).pin_memory() for _ in range(POOL)]
    dev_pool = [h.to(dev) for h in host_pool]
    flat = net.flat_gradients()
    stream = torch.cuda.current_stream(dev)

    dp = DataParallelStep(net, B_PER_GPU * world)   # overlapped gradient all-reduce (parallel.py); a plain step at N = 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def eager_step(i):
        dp.step(dev_pool[i % POOL])

    def graph_step(i):                            # the same launches replayed from a CUDA graph; the all-reduce stays outside it
        net.train_step_graph(dev_pool[i % POOL])
        dp.reduce(False)

    # `value` is timed on graph replays (PhysicsNet.train_step_graph: one graph per input buffer of the pool, recorded here,
    # outside the timed region) unless recording fails, the overlapped all-reduce is selected or PAIG_BENCH_GRAPH=0
    use_graph = os.environ.get("PAIG_BENCH_GRAPH", "1") != "0" and not dp.overlap
    if use_graph:
        try:
            for i in range(POOL):                 # no collective in here: a rank that fails to record cannot strand the others
                net.train_step_graph(dev_pool[i])
            torch.cuda.synchronize(dev)
            use_graph = all(v is not False for v in net._step_graphs.values())
        except Exception:                         # noqa: BLE001 -- eager launches are always available
            use_graph = False
    if world > 1:                                 # every rank times the same path
        flag = torch.tensor([1 if use_graph else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        use_graph = bool(flag.item())
    step = graph_step if use_graph else eager_step

    def launch_count():
        return lib.paig_launch_count() + getattr(net, "graph_replay_launches", 0)

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    launches = launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    losses = net._loss_view.detach().cpu().tolist()

    # ---- e2e: host buffers in, losses out, every step (paig_step_fused_host) ----
    tk = net._task(T)
    ws = net._workspace(T, B_PER_GPU, fresh=False)
    params = net._params_now()
    P, G = net._param_table(params), net._param_table(net._grad_views)
    losses_host = torch.empty(2, 4).pin_memory()  # one row per in-flight step
    done = [torch.cuda.Event(), torch.cuda.Event()]

    copy_stream = torch.cuda.Stream(dev)

    def stage(i):                                 # H2D of batch i into slot i % 2 on the copy stream
        _lib.check(lib.paig_stage_input_host(ctypes.byref(tk), host_pool[i % POOL].data_ptr(), B_PER_GPU, i % 2, ws.data_ptr(),
                                             copy_stream.cuda_stream))

    seen = []

    def step_host(i, first):
        # every step: its own input comes from pinned host memory and its four losses go back to the host and are READ by
        # the caller -- one step behind, the way a training loop logs: step i is enqueued, then the host waits for step
        # i-1, reads its losses, and only then (its input slot is free again) stages batch i+1 under step i's kernels
        armed = dp.arm()
        _lib.check(lib.paig_step_fused_staged(ctypes.byref(tk), ctypes.byref(P), ctypes.byref(G), B_PER_GPU, i % 2,
                                              losses_host[i % 2].data_ptr(), ws.data_ptr(), stream.cuda_stream))
        dp.reduce(armed)
        done[i % 2].record(stream)
        if not first:
            done[(i - 1) % 2].synchronize()
            seen.append(float(losses_host[(i - 1) % 2][0]))
        stage(i + 1)

    def drain_host(i):                            # the last step's losses
        done[i % 2].synchronize()
        seen.append(float(losses_host[i % 2][0]))

    stage(0)
    for i in range(3):
        step_host(i, i == 0)
    drain_host(2)
    barrier()
    seen.clear()
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(3, 3 + args.steps):            # the batch counter runs on: slot (i % 2) holds batch i
        step_host(i, i == 3)
    drain_host(3 + args.steps - 1)
    e1.record(stream)
    barrier()
    assert len(seen) == args.steps and all(v == v for v in seen), "e2e: every step's losses must have been read"
    ms_e2e = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms_e2e.item())
    sampler.stop_flag = True
    sampler.join(timeout=2)

    extras = {}
    if not args.no_extras:
        # (1) sustained: the same step for >= 2 s back to back (the headline region is a ~0.1 s burst)
        n_sus = max(args.steps, int(2.2e3 / max(ms_total / args.steps, 1e-3)))
        barrier()
        e0.record(stream)
        for i in range(n_sus):
            step(i)
        e1.record(stream)
        barrier()
        ms_sus = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms_sus, op=dist.ReduceOp.MAX)
        extras["sustained"] = {"steps": n_sus, "seconds": float(ms_sus.item()) * 1e-3,
                               "value": B_PER_GPU * world * n_sus / (float(ms_sus.item()) * 1e-3), "unit": "sequences/s"}
        # (1b) the other launch mode of the same step: eager launches when `value` was timed on graph replays, and vice versa
        try:
            other = eager_step if use_graph else graph_step
            for i in range(POOL + 3):
                other(i)
            barrier()
            e0.record(stream)
            for i in range(args.steps):
                other(i)
            e1.record(stream)
            barrier()
            ms_g = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms_g, op=dist.ReduceOp.MAX)
            extras["eager_launches" if use_graph else "cuda_graph"] = {
                "value": B_PER_GPU * world * args.steps / (float(ms_g.item()) * 1e-3), "unit": "sequences/s",
                "ms_per_step": float(ms_g.item()) / args.steps,
                "api": "net.train_step (paig_step_fused, 44 launches enqueued per step)" if use_graph else
                       "PhysicsNet.train_step_graph (paig_step_fused recorded once per input buffer, replayed)"}
        except Exception as exc:                      # noqa: BLE001 -- an extra key must never cost the bench line
            extras["eager_launches" if use_graph else "cuda_graph"] = {"error": str(exc)[:200]}
        # (2) the reference's own call sequence on the drop-in module (what base.py:142-151 / :195 does):
        #     net.output = net(inp); loss, _ = net.compute_loss(); loss.backward()  -- frames materialised, torch autograd
        def dropin_step(i):
            inp = dev_pool[i % POOL].requires_grad_(True)
            net.output = net(inp)
            loss, _ = net.compute_loss()
            for p_ in net.parameters():
                p_.grad = None
            loss.backward()
        for i in range(3 if world == 1 else 0):
            dropin_step(i)
        barrier()
        n_drop = max(5, min(args.steps, 20))
        e0.record(stream)
        for i in range(n_drop if world == 1 else 0):
            dropin_step(i)
        e1.record(stream)
        barrier()
        ms_drop = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world == 1:
            extras["dropin_module_path"] = {"value": B_PER_GPU * world * n_drop / (float(ms_drop.item()) * 1e-3), "unit": "sequences/s",
                                        "ms_per_step": float(ms_drop.item()) / n_drop,
                                            "api": "net.output = net(inp); net.compute_loss(); loss.backward() (paig_step_forward "
                                                   "+ paig_frame_sse_* + paig_step_backward under torch autograd), 1 GPU"}
        for p_ in net.parameters():
            p_.grad = None
        net._flat_grad = None
        flat = net.flat_gradients()
        # (3) strong scaling: BASELINE config 3's reading of "batch 100 on N GPUs" -- the GLOBAL batch stays 100
        if world > 1:
            from paig_reproduction_b200.parallel import shard_bounds
            lo, hi = shard_bounds(B_PER_GPU, world, rank)
            net.batch_global = B_PER_GPU
            shard = [d_[lo:hi].contiguous() for d_ in dev_pool]

            def strong_step(i):
                dp.step(shard[i % POOL])
            for i in range(args.warmup):
                strong_step(i)
            barrier()
            e0.record(stream)
            for i in range(args.steps):
                strong_step(i)
            e1.record(stream)
            barrier()
            ms_st = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(ms_st, op=dist.ReduceOp.MAX)
            extras["strong_scaling"] = {"global_batch": B_PER_GPU, "sequences_per_rank": "%d-%d" % (B_PER_GPU // world, -(-B_PER_GPU // world)),
                                        "value": B_PER_GPU * args.steps / (float(ms_st.item()) * 1e-3), "unit": "sequences/s",
                                        "ms_per_step": float(ms_st.item()) / args.steps}
            net.batch_global = B_PER_GPU * world

    # ---- live per-kernel timing (separate instrumented pass right after the timed region; events between launches
    #      would otherwise perturb the headline number) ----
    prof_steps = min(args.steps, 10)
    lib.paig_profile_begin()
    for i in range(prof_steps):
        net.train_step(dev_pool[i % POOL])
    buf = ctypes.create_string_buffer(1 << 16)
    _lib.check(lib.paig_profile_end(buf, len(buf)))
    kern = {}
    for ln in buf.value.decode().strip().splitlines():
        name, cnt, tot = ln.rsplit(" ", 2)
        kern[name] = {"launches_per_step": int(cnt) / prof_steps, "ms_per_step": float(tot) / prof_steps}
    kernel_ms = sum(v["ms_per_step"] for v in kern.values())
    # the 12 weight-gradient launches of a step are one group for the roofline whichever kernel a layer takes (the 32 -> 32
    # layer c6 runs on the tcgen05 kernel of csrc/wgrad_tc.cu, the rest on the TMA / mma.sync kernels of csrc/wgrad_tma.cu)
    wg_tc = kern.pop("conv3x3_wgrad_tc", None)
    if wg_tc and "conv3x3_wgrad" in kern:
        kern["conv3x3_wgrad"] = {"launches_per_step": kern["conv3x3_wgrad"]["launches_per_step"] + wg_tc["launches_per_step"],
                                 "ms_per_step": kern["conv3x3_wgrad"]["ms_per_step"] + wg_tc["ms_per_step"]}
    elif wg_tc:
        kern["conv3x3_wgrad_tc"] = wg_tc

    if rank == 0:
        pk, pk_src = peaks()
        seqs = B_PER_GPU * world
        value = seqs * args.steps / (ms_total * 1e-3)
        e2e_val = seqs * args.steps / (ms_e2e * 1e-3)
        # dominant kernel group and its algorithmic FLOPs per step (shallow UNet, N = 1000 frames of 32x32)
        N = B_PER_GPU * spec.enc_steps
        convs = [(3, 8, 32), (8, 8, 32), (8, 16, 16), (16, 16, 16), (16, 32, 8), (32, 32, 8), (32, 16, 16), (32, 16, 16),
                 (16, 16, 16), (16, 16, 32), (24, 8, 32), (8, 8, 32)]
        f_fwd = sum(2.0 * 9 * ci * co * s * s for ci, co, s in convs) * N
        f_dgrad = sum(2.0 * 9 * ci * co * s * s for ci, co, s in convs[1:]) * N       # no input gradient for c1
        # fused kernels (unet_fused.cu) or, under PAIG_UNET_LAYERWISE=1, the per-layer conv3x3 kernel
        flops = {"unet_fused_fwd": f_fwd, "unet_tc_fwd": f_fwd, "unet_fused_bwd": f_dgrad, "unet_tc_bwd": f_dgrad, "conv3x3": f_fwd + f_dgrad, "conv3x3_wgrad": f_fwd,
                 "sgemm": 2.0 * 3 * (2 * N) * (3072 * 200 + 200 * 200 + 200 * 2)}
        # the MLP GEMMs are profiled under several names (sgemm, sgemm_l1_fwd, ...): one group for the roofline
        sg = [k for k in kern if k.startswith("sgemm")]
        if sg:
            kern["sgemm"] = {"launches_per_step": sum(kern[k]["launches_per_step"] for k in sg),
                             "ms_per_step": sum(kern[k]["ms_per_step"] for k in sg)}
        parts = {k: kern.pop(k) for k in sg if k != "sgemm"}
        top = max(kern, key=lambda k: kern[k]["ms_per_step"]) if kern else None
        roof = None
        if top in flops:
            ach = flops[top] / (kern[top]["ms_per_step"] * 1e-3) / 1e12
            traffic = None
            try:
                # DRAM bytes (read + write) of this kernel group per step, from the committed `ncu --set full` capture
                traffic = json.load(open(os.path.join(ROOT, "profiles", "r2z_ncu_traffic.json"))).get(top)
            except (OSError, ValueError):
                pass
            roof = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["bf16_tflops_sustained"], "traffic": traffic, "peak_source": pk_src + " bf16 sustained",
                    "note": ("%s; fp32 results (1e-4 gradient parity excludes single-pass TF32: every product is 3 tensor-core "
                             "products, so the tensor pipe does 3x the algorithmic FLOPs counted here); of the nominal FP32-FMA "
                             "peak %.1f TFLOP/s the algorithmic rate is %.3f" % (
                                 {"unet_tc_fwd": "tcgen05 3xTF32 persistent kernel", "unet_tc_bwd": "tcgen05 3xTF32 persistent kernel",
                                  "conv3x3_wgrad": "12 launches: mma.sync 3xTF32 for layers with 16-channel m-tiles, tcgen05 3xTF32 for the 32 -> 32 layer, FP32 FMA for the rest"
                                  }.get(top, "fp32 CUDA-core kernel"), FP32_FMA_NOMINAL_TFLOPS, ach / FP32_FMA_NOMINAL_TFLOPS)),
                    "share_of_step": kern[top]["ms_per_step"] / kernel_ms if kernel_ms else None,
                    "algorithmic_flops_per_step": flops[top], "kernel_ms_per_step": kern[top]["ms_per_step"],
                    "unet_kernels_tflops": {k: round(flops[k] / (kern[k]["ms_per_step"] * 1e-3) / 1e12, 1)
                                            for k in ("unet_tc_fwd", "unet_tc_bwd", "unet_fused_fwd", "unet_fused_bwd", "conv3x3_wgrad")
                                            if k in kern}}
        line = {"metric": "train sequences/sec (fwd+bwd, spring_color, bs=100)", "value": value, "unit": "sequences/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "spring_color PhysicsNet LIVE training step (fwd+bwd, all parameter gradients), "
                                       "2 objects, 32x32 RGB, T=12, batch 100 per GPU", "task": TASK, "batch_per_gpu": B_PER_GPU,
                           "global_batch": seqs, "parallelism": "dp%d" % world, "alpha": ALPHA,
                           "l2": "inputs rotate over a %d x 14.7 MB pool (> 126 MB L2); ~0.8 GB of saved activations and "
                                 "gradients stream through L2 every step" % POOL,
                           "grad_allreduce": ("NCCL sum of the flat fp32 gradient buffer in two parts: everything but the UNet conv "
                                              "layers (98 % of the bytes) + 16 B fp64 on a side stream behind the library's "
                                              "early-gradient event, underneath the UNet backward; conv layers + losses after"
                                              if dp.overlap else "NCCL sum of one flat fp32 buffer + 16 B fp64, after the step")
                           if world > 1 else "none (1 GPU)",
                           "launch": ("CUDA graph replay of paig_step_fused's 44 launches (PhysicsNet.train_step_graph, one graph per "
                                      "input buffer, recorded before the timed region); e2e and the per-kernel pass launch eagerly"
                                      if use_graph else "eager launches"),
                           "optimizer": "not part of the metric (fwd+bwd, BASELINE.json)"},
                "clocks": sampler.summary(),
                "e2e": {"value": e2e_val, "unit": "sequences/s", "h2d_bytes_per_step": B_PER_GPU * T * 3 * H * H * 4,
                        "d2h_bytes_per_step": 16, "ms_per_step": ms_e2e / args.steps,
                        "api": "paig_stage_input_host + paig_step_fused_staged (pinned host input of every step copied on a side stream one step ahead; the losses of every step copied back and read by the host one step behind, so the next step is already enqueued while the host waits)"},
                "gpu_launches": int(launches),
                "gpu_launches_per_step": launches / args.steps,
                "roofline": roof,
                "hbm_fraction": value / world * BYTES_PER_SEQ / (pk["hbm_gbs"] * 1e9),
                "flop_fraction_fp32_fma_nominal": value / world * FLOPS_PER_SEQ / (FP32_FMA_NOMINAL_TFLOPS * 1e12),
                "kernels_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms_per_step"])},
                "sgemm_parts_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in parts.items()},
                "losses": losses}
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            val, cores, sample, _, kind = cpu_leg(6, 1)
            line["cpu_baseline"] = {"value": val, "unit": "sequences/s", "cores": cores, "kind": kind, "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
