"""fused optimizer step, device-resident batch gather, eval loop.
CPU: the iterator's index / epoch bookkeeping equals the reference's DataIterator under the same numpy seed.
GPU: optimizer kernels vs torch.optim (all four entries of base.py:12-17), gather vs numpy bit-exact, a short fused
training run that lowers the loss, eval_performance vs the drop-in's own compute_loss."""
import os

import numpy as np
import pytest
import torch

from oracle import physicsnet_oracle as po


class _RefIterator:
    """iterators.py:4-40 restated (index logic only) -- checked against the reference itself when it is importable."""

    def __init__(self, n):
        self.num_examples, self.epochs_completed, self.indices = n, 0, np.arange(n)
        self.reset_iteration()

    def reset_iteration(self):
        np.random.shuffle(self.indices)
        self.start_idx = 0

    def next_idx(self, bs):
        idx = self.indices[self.start_idx:self.start_idx + bs].copy()
        self.start_idx += bs
        if self.start_idx + bs > self.num_examples:
            self.reset_iteration()
            self.epochs_completed += 1
        return idx


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference checkout only exists in the build container")
def test_ref_iterator_restatement_matches_reference():
    import sys
    sys.path.insert(0, "/root/reference")
    from nn.datasets.iterators import DataIterator
    X = np.arange(23 * 2, dtype=np.float32).reshape(23, 2)
    np.random.seed(5)                                  # both draw from numpy's global RNG: run them one after the other
    ref = DataIterator(X)
    want = []
    for _ in range(12):
        bx, _ = ref.next_batch(5)
        want.append((bx.copy(), ref.epochs_completed))
    np.random.seed(5)
    mine = _RefIterator(23)
    for bx, ep in want:
        idx = mine.next_idx(5)
        assert np.array_equal(bx, X[idx]) and ep == mine.epochs_completed


def test_device_iterator_bookkeeping_on_cpu_indices():
    """DeviceIterator's index / epoch logic without touching the GPU (gather is patched out)."""
    from paig_reproduction_b200.train_loop import DeviceIterator
    X = np.zeros((23, 2, 4, 4, 3), np.uint8)
    np.random.seed(9)
    it = DeviceIterator.__new__(DeviceIterator)
    it.num_examples, it.epochs_completed, it.indices = 23, 0, np.arange(23)
    seen = []
    it.gather = lambda idx: seen.append(np.array(idx)) or None
    it.reset_iteration()
    epochs = []
    for _ in range(12):
        it.next_batch(5)
        epochs.append(it.epochs_completed)
    np.random.seed(9)
    ref = _RefIterator(23)
    for k in range(12):
        assert np.array_equal(seen[k], ref.next_idx(5)) and epochs[k] == ref.epochs_completed
    del X


def test_optimizer_names_follow_reference_table():
    from paig_reproduction_b200.train_loop import OPT_KINDS
    assert sorted(OPT_KINDS) == ["adam", "momentum", "rmsprop", "sgd"]            # base.py:12-17


gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("name", ["sgd", "momentum", "rmsprop", "adam"])
def test_fused_optimizer_matches_torch_optim(name):
    from paig_reproduction_b200 import _lib
    from paig_reproduction_b200.train_loop import OPT_KINDS
    lib = _lib.load()
    torch.manual_seed(0)
    n, lr = 100003, 3e-3
    make = {"adam": lambda p: torch.optim.Adam(p, lr=lr), "rmsprop": lambda p: torch.optim.RMSprop(p, lr=lr),
            "momentum": lambda p: torch.optim.SGD(p, momentum=0.9, lr=lr), "sgd": lambda p: torch.optim.SGD(p, lr=lr)}
    for dtype, fn in ((torch.float32, lib.paig_optimizer_step), (torch.float64, lib.paig_optimizer_step_f64)):
        p0 = torch.randn(n, dtype=dtype, device="cuda")
        ref = torch.nn.Parameter(p0.clone())
        opt = make[name]([ref])
        mine, s0, s1 = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
        for step in range(1, 5):
            g = torch.randn(n, dtype=dtype, device="cuda") * (10.0 ** (step - 2))
            ref.grad = g.clone()
            opt.step()
            _lib.check(fn(OPT_KINDS[name], mine.data_ptr(), g.data_ptr(), s0.data_ptr(), s1.data_ptr(), n, lr, step,
                          torch.cuda.current_stream().cuda_stream))
            err = (mine - ref.detach()).abs().max().item() / ref.detach().abs().max().item()
            assert err < (2e-6 if dtype == torch.float32 else 1e-13), (name, dtype, step, err)


@gpu
def test_gather_batch_is_bit_exact_and_is_a_reshape():
    from paig_reproduction_b200.train_loop import DeviceIterator
    rng = np.random.RandomState(1)
    X = rng.randint(0, 256, size=(37, 12, 32, 32, 3)).astype(np.uint8)
    np.random.seed(3)
    ref = _RefIterator(37)
    order = [ref.next_idx(10) for _ in range(5)]
    np.random.seed(3)
    it = DeviceIterator(X, "cuda:0")
    want_all = X.astype(np.float32).reshape(X.shape[:2] + (3, 32, 32)) / 255          # iterators.py:57-64
    for idx in order:
        got, _ = it.next_batch(10)
        assert got.shape == (10, 12, 3, 32, 32)
        assert np.array_equal(got.cpu().numpy(), want_all[idx])


@gpu
def test_fused_training_loop_and_eval():
    from paig_reproduction_b200.physics_models import PhysicsNet
    from paig_reproduction_b200.train_loop import DeviceIterator, FusedOptimizer, eval_performance, train_epochs
    spec = po.TASKS["spring_color"]
    frames = po.synthetic_frames(spec, 40, spec.seq_len, 2)                            # [N,T,C,H,W] in [0,1)
    X = (frames.numpy() * 255).astype(np.uint8).reshape(40, spec.seq_len, 32, 32, 3)   # any bytes: layout is a reshape
    np.random.seed(0)
    it = DeviceIterator(X, "cuda:0")
    net = PhysicsNet("spring_color", 100, 1, "spring_ode_cell", spec.seq_len, spec.input_steps, spec.pred_steps, 3.0, False,
                     True, 32 * 32, "conv_encoder", "conv_st_decoder", device="cuda:0")
    net.load_state_dict(po.init_state_dict(spec, 0), strict=True)
    sd_keys = list(net.state_dict().keys())
    opt = FusedOptimizer(net, "rmsprop", 3e-4)
    assert list(net.state_dict().keys()) == sd_keys                                   # re-homing keeps the checkpoint format
    first = net.train_step(it.gather(np.arange(20))).cpu().clone()
    last = train_epochs(net, it, opt, epochs=4, batch_size=20, anneal_lr=True).cpu()
    assert torch.isfinite(last).all() and last[0] < first[0]                          # the loss goes down
    assert abs(opt.lr - 3e-4 * 0.2) < 1e-12                                            # the anneal reached the optimizer
    # the same parameters step under torch.optim from the same start: one step, same update
    net2 = PhysicsNet("spring_color", 100, 1, "spring_ode_cell", spec.seq_len, spec.input_steps, spec.pred_steps, 3.0, False,
                      True, 32 * 32, "conv_encoder", "conv_st_decoder", device="cuda:0")
    net2.load_state_dict(po.init_state_dict(spec, 0), strict=True)
    net3 = PhysicsNet("spring_color", 100, 1, "spring_ode_cell", spec.seq_len, spec.input_steps, spec.pred_steps, 3.0, False,
                      True, 32 * 32, "conv_encoder", "conv_st_decoder", device="cuda:0")
    net3.load_state_dict(po.init_state_dict(spec, 0), strict=True)
    x = it.gather(np.arange(20))
    net2.build_optimizer(3e-4, "rmsprop")
    net2.train_step(x)
    net2.optimizer.step()
    o3 = FusedOptimizer(net3, "rmsprop", 3e-4)
    net3.train_step(x)
    o3.step()
    a, b = net2.state_dict(), net3.state_dict()
    for k in a:
        if a[k].is_floating_point():
            d = (a[k].double() - b[k].double()).abs().max().item()
            assert d <= 2e-6 * max(1.0, a[k].abs().max().item()), (k, d)
    # N3: eval loop
    np.random.seed(1)
    means = eval_performance(net, it, 20)
    assert set(means) == {"eval_pred_loss", "eval_extrap_loss", "eval_recons_loss"} and all(np.isfinite(v) for v in means.values())
