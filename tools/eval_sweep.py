"""BASELINE config 5: spring_color_half long-horizon evaluation (test-mode rollout, T=30, no_grad) with the batch
swept 100..8192 on one GPU.  Prints one JSON line per batch size: sequences/s of net(x) + compute_loss().
    python tools/eval_sweep.py [--batches 100,512,2048,8192]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import physicsnet_oracle as po
from paig_reproduction_b200.physics_models import PhysicsNet

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="100,512,2048,8192")
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
spec = po.TASKS["spring_color_half"]
T = 30
net = PhysicsNet("spring_color_half", 100, 1, "spring_ode_cell", T, spec.input_steps, spec.pred_steps, 3.0, False, True,
                 32 * 32, "conv_encoder", "conv_st_decoder", device="cuda:0")
net.load_state_dict(po.init_state_dict(spec, 0), strict=True)
net.eval()
for B in [int(b) for b in a.batches.split(",")]:
    x = torch.rand(B, T, 3, 32, 32, device="cuda:0")
    with torch.no_grad():
        for _ in range(2):
            net.output = net(x); net.compute_loss()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            net.output = net(x)
            _, evals = net.compute_loss()
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"workload": "spring_color_half test-mode rollout T=30 (36 decoded frames/seq, frames materialised)",
                      "batch": B, "ms_per_batch": ms, "sequences_per_s": B / ms * 1e3,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}), flush=True)
