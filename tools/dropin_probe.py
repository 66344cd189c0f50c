"""Where does the drop-in module path (net(inp); compute_loss(); backward()) spend its time?  CPU enqueue time per phase
(no synchronisation) against GPU time per phase (CUDA events), spring_color B = 100."""
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import physicsnet_oracle as po
from paig_reproduction_b200.physics_models import PhysicsNet

spec = po.TASKS["spring_color"]
T, H, B = spec.seq_len, spec.H, 100
dev = torch.device("cuda", 0)
net = PhysicsNet("spring_color", 100, 1, "spring_ode_cell", T, spec.input_steps, spec.pred_steps, 3.0, False, True, H * H,
                 "conv_encoder", "conv_st_decoder", device=dev)
net.load_state_dict(po.init_state_dict(spec, 0), strict=True)
g = torch.Generator().manual_seed(1)
pool = [torch.rand(B, T, 3, H, H, generator=g).to(dev) for _ in range(12)]


def step(i, marks=None):
    inp = pool[i % 12].requires_grad_(True)
    t = [time.perf_counter()]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    net.output = net(inp)
    ev[1].record(); t.append(time.perf_counter())
    loss, _ = net.compute_loss()
    for p_ in net.parameters():
        p_.grad = None
    ev[2].record(); t.append(time.perf_counter())
    loss.backward()
    ev[3].record(); t.append(time.perf_counter())
    if marks is not None:
        marks.append((t, ev))


for i in range(5):
    step(i)
torch.cuda.synchronize()
marks = []
t0 = time.perf_counter()
for i in range(20):
    step(i, marks)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
cpu = [sum(m[0][k + 1] - m[0][k] for m in marks) / len(marks) * 1e3 for k in range(3)]
gpu = [sum(m[1][k].elapsed_time(m[1][k + 1]) for m in marks) / len(marks) for k in range(3)]
print({"cpu_enqueue_ms": dict(zip(("forward", "compute_loss", "backward"), cpu)), "gpu_ms": dict(zip(("forward", "compute_loss", "backward"), gpu)),
       "cpu_total_per_step_ms": (t1 - t0) / 20 * 1e3, "wall_per_step_ms": (t2 - t0) / 20 * 1e3})
from paig_reproduction_b200 import _lib
lib = _lib.load()
import ctypes
lib.paig_profile_begin()
for i in range(5):
    step(i)
buf = ctypes.create_string_buffer(1 << 16)
lib.paig_profile_end(buf, len(buf))
rows = []
for ln in buf.value.decode().strip().splitlines():
    name, cnt, tot = ln.rsplit(" ", 2)
    rows.append((float(tot) / 5, name, int(cnt) / 5))
print(sorted(rows, reverse=True)[:40])
