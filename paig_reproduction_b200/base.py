"""Host-side training / evaluation driver of the drop-in: the surface of the reference's ``BaseNetTorch``
(nn/network/base.py:20-218) that ``runners/torch_run_physics.py:90-98,108-117`` calls on the model --
``get_data``, ``initialize_graph``, ``train_model`` (which calls ``eval_performance``, ``run_extra_fns``,
``add_train_logger``) -- with the same artefacts in ``save_dir``: ``log.txt`` (same logger name and line format,
nn/utils/misc.py:6-9), ``code.zip``, ``model.ckpt`` (plain ``state_dict``), ``outputs.npz``.

Control flow is the reference's, including the two behaviours SURVEY section 0 calls out:
  Q1  ``train_model`` does NOT assign ``self.output``; ``compute_loss`` therefore scores the output of the last
      evaluation pass ("STALE" mode) and only the reconstruction path trains.  ``live_training = True`` (an
      attribute, default False so the unmodified runner behaves as it does today) assigns it.
  Q7  the LR anneal changes ``self.lr`` only.  ``anneal_reaches_optimizer = True`` also updates the optimizer.
Only host logic lives here; every tensor op of the step is in libpaig_b200.so (physics_models.py).
"""
from __future__ import annotations

import logging
import os
import shutil
import sys
import zipfile

import numpy as np
import torch

logger = logging.getLogger("torch")          # base.py:9 -- the runner attaches its stream handler to this name
PKG_ROOT = os.path.dirname(os.path.abspath(__file__))

# base.py:12-17
OPTIMIZERS = {
    "adam": lambda params, lr: torch.optim.Adam(params, lr=lr),
    "rmsprop": lambda params, lr: torch.optim.RMSprop(params, lr=lr),
    "momentum": lambda params, lr: torch.optim.SGD(params, momentum=0.9, lr=lr),
    "sgd": lambda params, lr: torch.optim.SGD(params, lr=lr),
}


def log_metrics(log, prefix, metrics):
    """nn/utils/misc.py:6-9: ``<prefix> k1=v1 k2=v2`` with keys sorted; values print as torch / numpy would."""
    log.info(prefix + " " + " ".join("%s=%s" % (k, metrics[k]) for k in sorted(metrics.keys())))


def zip_sources(src_root: str, save_dir: str) -> str:
    """nn/utils/misc.py:22-33 (zipdir): every .py under `src_root` into save_dir/code.zip."""
    path = os.path.join(save_dir, "code.zip")
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
        for root, _dirs, files in os.walk(src_root):
            for f in files:
                if f.endswith(".py"):
                    full = os.path.join(root, f)
                    z.write(full, os.path.relpath(full, os.path.join(src_root, "..")))
    return path


class BaseNetTorch(torch.nn.Module):
    """nn/network/base.py:20-218."""

    live_training = False               # Q1: False = the reference's loop as shipped
    anneal_reaches_optimizer = False    # Q7: False = the reference's no-op anneal

    def __init__(self):
        super().__init__()
        self.train_metrics = {}
        self.eval_metrics = {}
        # (fn, args, kwargs) triples run after a train step / a validation pass / a test pass  (base.py:26-35)
        self.extra_train_fns = []
        self.extra_valid_fns = []
        self.extra_test_fns = []

    # ------------------------------------------------------------------ small helpers (base.py:37-63,96-110)
    def run_extra_fns(self, type):
        fns = {"train": self.extra_train_fns, "valid": self.extra_valid_fns}.get(type, self.extra_test_fns)
        for fn, args, kwargs in fns:
            fn(*args, **kwargs)

    def conv_feedforward(self, inp):
        raise NotImplementedError

    def compute_loss(self):
        raise NotImplementedError

    def get_data(self, data_iterators):
        self.train_iterator, self.valid_iterator, self.test_iterator = data_iterators

    def get_batch(self, batch_size, iterator):
        batch_x, batch_y = iterator.next_batch(batch_size)
        feed = {"input": batch_x} if batch_y is None else {"input": batch_x, "target": batch_y}
        return feed, (batch_x, batch_y)

    def get_iterator(self, type):
        return {"train": self.train_iterator, "valid": self.valid_iterator, "test": self.test_iterator}[type]

    def add_train_logger(self):
        handler = logging.FileHandler(os.path.join(self.save_dir, "log.txt"))
        handler.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(message)s"))
        logger.addHandler(handler)
        self._log_handler = handler

    # ------------------------------------------------------------------ checkpoint directory (base.py:65-94)
    def initialize_graph(self, save_dir, use_ckpt, ckpt_dir=""):
        self.save_dir = save_dir
        restore_dir = None
        if os.path.exists(save_dir):
            if use_ckpt:
                restore_dir = ckpt_dir if ckpt_dir else save_dir
            else:
                logger.info("Folder exists, deleting...")
                shutil.rmtree(save_dir)
                os.makedirs(save_dir)
        else:
            os.makedirs(save_dir)
            if use_ckpt:
                restore_dir = ckpt_dir
        if restore_dir is not None:
            print(f"Loading model from: {restore_dir + '/model.ckpt'}")
            self.load_state_dict(torch.load(os.path.join(restore_dir, "model.ckpt"), map_location=self.device))

    def _as_input(self, batch, requires_grad):
        """base.py:141,194: torch.tensor(feed_dict["input"], device=self.device).  numpy batches (the reference's
        DataIterator) are copied host->device; device batches (train_loop.DeviceIterator) are used as they are."""
        if isinstance(batch, torch.Tensor):
            t = batch.detach().to(self.device, torch.float32)
        else:
            t = torch.as_tensor(np.asarray(batch), dtype=torch.float32).to(self.device, non_blocking=True)
        return t.requires_grad_(requires_grad)

    # ------------------------------------------------------------------ training loop (base.py:112-172)
    def train_model(self, epochs, batch_size, save_every_n_epochs, eval_every_n_epochs, print_interval, debug=False):
        self.train()
        self.batch_size = batch_size
        self.add_train_logger()
        zip_sources(PKG_ROOT, self.save_dir)
        logger.info("\n".join(sys.argv))
        step = 0
        if not debug and epochs > 0:                      # one validation pass before training
            log_metrics(logger, "valid - epoch=%s" % 0, self.eval_performance(batch_size, type="valid"))
        for ep in range(1, epochs + 1):
            if self.anneal_lr and ep == int(0.75 * epochs):
                self.lr = self.lr / 5
                if self.anneal_reaches_optimizer:
                    self._set_optimizer_lr(self.lr)
            while self.train_iterator.epochs_completed < ep:
                feed_dict, _ = self.get_batch(batch_size, self.train_iterator)
                inp = self._as_input(feed_dict["input"], True)
                result_sequence = self.forward(inp)
                if self.live_training:
                    self.output = result_sequence
                self.train_loss, self.eval_losses = self.compute_loss()
                self.train_metrics["train_loss"] = self.train_loss
                self.eval_metrics["eval_pred_loss"] = self.eval_losses[0]
                self.eval_metrics["eval_extrap_loss"] = self.eval_losses[1]
                self.eval_metrics["eval_recons_loss"] = self.eval_losses[2]
                self.loss = self.train_loss
                self.optimizer.zero_grad(set_to_none=True)
                self.loss.backward()
                self.optimizer.step()
                self.run_extra_fns("train")
                if step % print_interval == 0:
                    log_metrics(logger, "train - iter=%s" % step, self.train_metrics)
                step += 1
            if ep % eval_every_n_epochs == 0:
                print("eval running")
                log_metrics(logger, "valid - epoch=%s" % ep, self.eval_performance(batch_size, type="valid"))
            if ep % save_every_n_epochs == 0:
                print("saving")
                torch.save(self.state_dict(), os.path.join(self.save_dir, "model.ckpt"))
        log_metrics(logger, "test - epoch=%s" % epochs, self.eval_performance(batch_size, type="test"))

    def _set_optimizer_lr(self, lr):
        opt = getattr(self, "optimizer", None)
        if opt is None:
            return
        if hasattr(opt, "param_groups"):
            for group in opt.param_groups:
                group["lr"] = lr
        elif hasattr(opt, "lr"):
            opt.lr = lr

    # ------------------------------------------------------------------ evaluation pass (base.py:174-218)
    def eval_performance(self, batch_size, type="valid"):
        self.eval()
        with torch.no_grad():
            for k in ("eval_pred_loss", "eval_extrap_loss", "eval_recons_loss"):
                self.eval_metrics[k] = torch.tensor([0], device=self.device)
            results = {k: [] for k in self.eval_metrics.keys()}
            inputs, outputs = [], []
            iterator = self.get_iterator(type)
            iterator.reset_epoch()
            while iterator.get_epoch() < 1:
                if iterator.X.shape[0] < 100:
                    batch_size = iterator.X.shape[0]
                feed_dict, _ = self.get_batch(batch_size, iterator)
                inp = self._as_input(feed_dict["input"], False)
                self.output = self.conv_feedforward(inp)
                self.train_loss, self.eval_losses = self.compute_loss()
                self.train_metrics["train_loss"] = self.train_loss
                self.eval_metrics["eval_pred_loss"] = self.eval_losses[0]
                self.eval_metrics["eval_extrap_loss"] = self.eval_losses[1]
                self.eval_metrics["eval_recons_loss"] = self.eval_losses[2]
                self.loss = self.train_loss
                for k in self.eval_metrics.keys():
                    results[k].append(self.eval_metrics[k])
                batch = feed_dict["input"]
                inputs.append(batch.detach().cpu().numpy() if isinstance(batch, torch.Tensor) else np.asarray(batch))
                outputs.append(self.eval_losses)
            means = {k: np.mean([v.detach().cpu().numpy() for v in vals], axis=0) for k, vals in results.items()}
            np.savez_compressed(os.path.join(self.save_dir, "outputs.npz"), input=np.concatenate(inputs, axis=0),
                                output=np.array([[v.detach().cpu().numpy() for v in row] for row in outputs]))
            self.run_extra_fns(type)
            return means
