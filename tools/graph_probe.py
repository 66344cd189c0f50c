"""Does recording the fused training step into a CUDA graph pay?  Captures paig_step_fused (44 launches, side streams
included) once per input buffer and compares replay time with eager launches.  spring_color, B = 100."""
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import physicsnet_oracle as po
from paig_reproduction_b200.physics_models import PhysicsNet

TASK, B = "spring_color", 100
spec = po.TASKS[TASK]
T, H = spec.seq_len, spec.H
dev = torch.device("cuda", 0)
net = PhysicsNet(TASK, 100, 1, "spring_ode_cell", T, spec.input_steps, spec.pred_steps, 3.0, False, True, H * H,
                 "conv_encoder", "conv_st_decoder", device=dev)
net.load_state_dict(po.init_state_dict(spec, 0), strict=True)
POOL = 12
g = torch.Generator().manual_seed(100)
pool = [torch.rand(B, T, 3, H, H, generator=g).to(dev) for _ in range(POOL)]
for i in range(5):
    net.train_step(pool[i % POOL])
torch.cuda.synchronize()
ref = net.flat_gradients().clone()
net.train_step(pool[4])
torch.cuda.synchronize()
ref = net.flat_gradients().clone()


def timed(fn, n=50):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


eager = timed(lambda i: net.train_step(pool[i % POOL]))
graphs = []
s = torch.cuda.Stream(dev)
s.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(s):
    for i in range(POOL):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s, capture_error_mode="relaxed"):
            net.train_step(pool[i])
        graphs.append(gr)
torch.cuda.current_stream(dev).wait_stream(s)
torch.cuda.synchronize()
graphs[4].replay()
torch.cuda.synchronize()
same = torch.equal(net.flat_gradients(), ref)
replay = timed(lambda i: graphs[i % POOL].replay())
print({"eager_ms": eager, "graph_ms": replay, "bit_identical": same})
