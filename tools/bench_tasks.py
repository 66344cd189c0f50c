"""Training-step throughput (LIVE fused step, fwd+bwd, batch 100, one GPU) for every BASELINE.json task.
One JSON line per task; spring_color is what bench.py reports."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import physicsnet_oracle as po
from paig_reproduction_b200.physics_models import PhysicsNet
from paig_reproduction_b200 import _lib

ALPHA = {"spring_color": 3.0, "bouncing_balls": 2.0, "3bp_color": 5.0, "mnist_spring_color": 3.0}
TASKS = sys.argv[1:] or ["spring_color", "bouncing_balls", "3bp_color", "mnist_spring_color"]
for task in TASKS:
    spec = po.TASKS[task]
    net = PhysicsNet(task, 100, 1, po.CELL_TYPE_NAMES[spec.cell], spec.seq_len, spec.input_steps, spec.pred_steps, ALPHA[task],
                     False, True, spec.H * spec.H, "conv_encoder", "conv_st_decoder", device="cuda:0")
    net.load_state_dict(po.init_state_dict(spec, 0), strict=True)
    pool = [torch.rand(100, spec.seq_len, 3, spec.H, spec.H, device="cuda:0") for _ in range(4)]
    for i in range(3):
        net.train_step(pool[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for i in range(steps):
        net.train_step(pool[i % 4])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    lib = _lib.load()
    lib.paig_profile_begin()                      # per-launch CUDA events: an instrumented pass after the timed one
    for i in range(4):
        net.train_step(pool[i % 4])
    buf = ctypes.create_string_buffer(1 << 16)
    _lib.check(lib.paig_profile_end(buf, len(buf)))
    kern = {}
    for ln in buf.value.decode().strip().splitlines():
        name, cnt, tot = ln.rsplit(" ", 2)
        kern[name] = round(float(tot) / 4, 4)
    kern = dict(sorted(kern.items(), key=lambda kv: -kv[1]))
    print(json.dumps({"task": task, "H": spec.H, "n_objs": spec.n_objs, "T": spec.seq_len, "batch": 100, "ms_per_step": ms,
                      "sequences_per_s": 100 / ms * 1e3, "kernels_ms_per_step": kern}), flush=True)
    del net, pool
    torch.cuda.empty_cache()
