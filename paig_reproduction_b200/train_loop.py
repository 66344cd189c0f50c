"""The steps either side of the hot path in the reference's training / evaluation loop:

  N1  FusedOptimizer   optimizer.step() of base.py:152 as ONE kernel over a flat parameter / gradient buffer
                       (torch.optim defaults of the base.py:12-17 table), plus an LR anneal that actually changes
                       the step size (the reference's mutates self.lr only, SURVEY Q7 -- opt-in here).
  N2  DeviceIterator   iterators.py:4-40 with the uint8 dataset resident on the GPU: same shuffling / epoch
                       bookkeeping, but next_batch() gathers and normalises on the device (no per-step H2D copy).
  N3  eval_performance base.py:171-216: batched no-grad evaluation, mean of the three losses, outputs.npz.

Host logic is Python like the reference's; the arithmetic runs in libpaig_b200.so (no fallback)."""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib

OPT_KINDS = {"sgd": 0, "momentum": 1, "rmsprop": 2, "adam": 3}     # base.py:12-17


class FusedOptimizer:
    """Drop-in for ``net.optimizer``: ``zero_grad()`` / ``step()``.  Re-homes every live parameter into one flat fp32
    buffer laid out like ``net.flat_gradients()`` (``p.data`` becomes a view, so ``state_dict`` / checkpoints are
    unchanged); the fp64 physics scalars get a second tiny launch."""

    def __init__(self, net, optimizer: str = "rmsprop", lr: float = 3e-4):
        if optimizer not in OPT_KINDS:
            raise KeyError(optimizer)                                   # the reference indexes OPTIMIZERS[...] the same way
        self.net, self.kind, self.lr, self.steps = net, OPT_KINDS[optimizer], float(lr), 0
        self.flat_grad = net.flat_gradients()
        params = dict(net.named_parameters())
        names = [k for k in net.live_parameter_names()]
        self.f32 = net.flat_order()                                     # same order as the flat gradient buffer
        self.f64 = [k for k in names if params[k].dtype == torch.float64]
        n = sum(params[k].numel() for k in self.f32)
        dev = self.flat_grad.device
        self.flat_param = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for k in self.f32:
                p = params[k]
                view = self.flat_param[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view                                           # same storage order as the gradient views
                off += p.numel()
        self.n = n
        self.phys_param = torch.zeros(len(self.f64), dtype=torch.float64, device=dev)
        with torch.no_grad():
            for i, k in enumerate(self.f64):
                self.phys_param[i] = params[k]
                params[k].data = self.phys_param[i].view(())
        self.state = [torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(2 if self.kind == 3 else 1)]
        self.state64 = [torch.zeros(len(self.f64), dtype=torch.float64, device=dev) for _ in range(2 if self.kind == 3 else 1)]

    def zero_grad(self, set_to_none: bool = True):                      # gradients are written, not accumulated
        pass

    def step(self):
        lib = _lib.load()
        self.steps += 1
        st = torch.cuda.current_stream(self.flat_param.device).cuda_stream
        s1 = self.state[1].data_ptr() if len(self.state) > 1 else None
        _lib.check(lib.paig_optimizer_step(self.kind, self.flat_param.data_ptr(), self.flat_grad.data_ptr(),
                                           self.state[0].data_ptr(), s1, self.n, self.lr, self.steps, st), "optimizer_step")
        if len(self.f64):
            s1 = self.state64[1].data_ptr() if len(self.state64) > 1 else None
            _lib.check(lib.paig_optimizer_step_f64(self.kind, self.phys_param.data_ptr(), self.net._phys_grad.data_ptr(),
                                                   self.state64[0].data_ptr(), s1, len(self.f64), self.lr, self.steps, st),
                       "optimizer_step_f64")

    def anneal(self, factor: float = 0.2):
        """base.py:135-137 divides self.lr by 5 at 75 % of the epochs but never tells the optimizer; this does."""
        self.lr *= factor


class DeviceIterator:
    """iterators.py:4-40 (DataIterator) over a uint8 array [N, T, H, W, C] kept on the device.  Index shuffling uses
    numpy's global RNG exactly like the reference, so a seeded run visits the same sequences in the same order.  The epoch
    bookkeeping (reset_iteration / get_epoch / reset_epoch / the wrap-around in next_batch) is a near-verbatim restatement of
    the reference class, assert message included: identical RNG draws and epoch boundaries require identical control flow;
    what differs is where the data lives and how a batch is gathered (paig_gather_batch_u8)."""

    def __init__(self, X_u8: np.ndarray, device, conv: bool = True):
        assert X_u8.dtype == np.uint8 and X_u8.ndim == 5
        self.shape = X_u8.shape
        n, t, h, w, c = X_u8.shape
        # iterators.py:57-64: reshape (not permute) to [N, T, C, H, W] for conv nets, flat otherwise
        self.out_shape = (t, c, h, w) if conv else (t, h * w * c)
        self.seq_elems = t * h * w * c
        assert self.seq_elems % 4 == 0
        self.data = torch.from_numpy(np.ascontiguousarray(X_u8)).to(device)
        self.num_examples = n
        self.epochs_completed = 0
        self.indices = np.arange(n)
        self.reset_iteration()

    def reset_iteration(self):
        np.random.shuffle(self.indices)
        self.start_idx = 0

    def get_epoch(self):
        return self.epochs_completed

    def reset_epoch(self):
        self.reset_iteration()
        self.epochs_completed = 0

    @property
    def X(self):                                                        # base.py:190 reads iterator.X.shape[0]
        return self.data

    def next_batch(self, batch_size, data_type="train", shuffle=True):
        assert data_type in ["train", "val", "test"], "data_type must be 'train', 'val', or 'test'."
        idx = self.indices[self.start_idx:self.start_idx + batch_size]
        batch_x = self.gather(idx)
        self.start_idx += batch_size
        if self.start_idx + batch_size > self.num_examples:
            self.reset_iteration()
            self.epochs_completed += 1
        return batch_x, None

    def gather(self, idx) -> torch.Tensor:
        lib = _lib.load()
        dev = self.data.device
        idx_d = torch.as_tensor(np.asarray(idx, dtype=np.int64), device=dev)
        out = torch.empty((len(idx),) + self.out_shape, dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.paig_gather_batch_u8(self.data.data_ptr(), self.seq_elems, idx_d.data_ptr(), len(idx), out.data_ptr(),
                                            st), "gather_batch_u8")
        return out


def get_iterators(file: str, device, conv: bool = True):
    """iterators.py:49-69 with device-resident splits."""
    data = np.load(file)
    return tuple(DeviceIterator(data[k], device, conv) for k in ("train_x", "valid_x", "test_x"))


def eval_performance(net, iterator, batch_size: int, save_dir: Optional[str] = None) -> Dict[str, float]:
    """base.py:171-216: one no-grad pass over `iterator`; mean over batches of (train-aliased pred, extrap, recons)
    (eval_pred_loss is the in-place alias pred + alpha * recons); optionally writes outputs.npz."""
    net.eval()
    names = ["eval_pred_loss", "eval_extrap_loss", "eval_recons_loss"]
    results: Dict[str, List[np.ndarray]] = {k: [] for k in names}
    inputs, outputs = [], []
    with torch.no_grad():
        iterator.reset_epoch()
        while iterator.get_epoch() < 1:
            if iterator.X.shape[0] < 100:                               # base.py:190-191
                batch_size = iterator.X.shape[0]
            inp, _ = iterator.next_batch(batch_size, "test")
            net.output = net.conv_feedforward(inp)                      # base.py:195
            _, evals = net.compute_loss()
            vals = [v.detach().cpu().numpy() for v in evals]
            for k, v in zip(names, vals):
                results[k].append(v)
            if save_dir is not None:
                inputs.append(inp.cpu().numpy())
                outputs.append(vals)
    means = {k: float(np.mean(v, axis=0)) for k, v in results.items()}
    if save_dir is not None:
        os.makedirs(save_dir, exist_ok=True)
        np.savez_compressed(os.path.join(save_dir, "outputs.npz"), input=np.concatenate(inputs, axis=0),
                            output=np.array(outputs))
    return means


def train_epochs(net, train_it, optimizer: FusedOptimizer, epochs: int, batch_size: int, anneal_lr: bool = False):
    """base.py:131-152 with the fused LIVE step: get_batch -> paig_step_fused -> fused optimizer step.  Returns the last
    [train, pred, extrap, recons] losses (device tensor).  `anneal_lr` applies the 75 % / divide-by-5 schedule for real."""
    losses = None
    for ep in range(1, epochs + 1):
        if anneal_lr and ep == int(0.75 * epochs):
            optimizer.anneal(0.2)
        while train_it.epochs_completed < ep:
            inp, _ = train_it.next_batch(batch_size)
            losses = net.train_step(inp)
            optimizer.step()
    return losses
