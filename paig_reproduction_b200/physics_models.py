"""Drop-in PhysicsNet whose per-sequence training step runs in libpaig_b200.so (hand-written sm_100a kernels).

Mirrors the Python surface of the reference's ``nn/network/physics_models.py:PhysicsNet`` (constructor
argument order, ``forward`` / ``conv_feedforward`` / ``compute_loss`` / ``build_optimizer``, the cached
attributes and every ``state_dict`` key), so ``runners/torch_run_physics.py`` can construct it unchanged.
The parameters live in ordinary ``nn.Module`` holders (same construction order as the reference, hence the
same default initialisation under the same seed), but no torch operator runs on the hot path: ``forward``
is one call into the C ABI (include/paig_b200.h), wrapped in one ``torch.autograd.Function``.

There is no CPU path and no PyTorch fallback: tensors must be CUDA tensors and the library must load.
"""
from __future__ import annotations

import ctypes
import logging
import os
import weakref
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _abi, _lib
from .base import OPTIMIZERS, BaseNetTorch

logger = logging.getLogger("tf")          # physics_models.py:19 (the reference logs pos_vel_seq rows under this name)

# physics_models.py:31-37
COORD_UNITS = {"bouncing_balls": 8, "spring_color": 8, "spring_color_half": 8, "3bp_color": 12,
               "mnist_spring_color": 8}
# cells.py class names accepted by the runner's --cell_type table (runners/torch_run_physics.py:49-75)
CELLS = {"spring_ode_cell": "spring", "bouncing_ode_cell": "bouncing", "gravity_ode_cell": "gravity"}

# ---- parameter holders (never executed with torch ops; they exist for state_dict / optimizer parity) ----
class VariableFromNetwork(nn.Module):
    """blocks.py:311-322."""

    def __init__(self, shape):
        super().__init__()
        self.shape = list(shape)
        self.l1 = nn.Linear(10, 200)
        self.l2 = nn.Linear(200, int(np.prod(shape)))


def _conv(cin, cout, k=3):
    return nn.Conv2d(cin, cout, kernel_size=k, padding="same")


class ShallowUNet(nn.Module):
    """blocks.py:240-276 (layer shapes only)."""

    def __init__(self, in_channels, h, out_features):
        super().__init__()
        chans = [(in_channels, h), (h, h), (h, 2 * h), (2 * h, 2 * h), (2 * h, 4 * h), (4 * h, 4 * h), (4 * h, 2 * h),
                 (4 * h, 2 * h), (2 * h, 2 * h), (2 * h, 2 * h), (3 * h, h), (h, h)]
        for i, (ci, co) in enumerate(chans, start=1):
            setattr(self, "c%d" % i, _conv(ci, co))
        self.c13 = _conv(h, out_features, 1)


class UNet(nn.Module):
    """blocks.py:106-170 (layer shapes only)."""

    def __init__(self, in_channels, h, out_features):
        super().__init__()
        chans = [(in_channels, h), (h, h), (h, 2 * h), (2 * h, 2 * h), (2 * h, 4 * h), (4 * h, 4 * h), (4 * h, 8 * h),
                 (8 * h, 8 * h), (8 * h, 2 * h), (6 * h, 4 * h), (4 * h, 4 * h), (4 * h, 2 * h), (4 * h, 2 * h),
                 (2 * h, 2 * h), (2 * h, 2 * h), (3 * h, h), (h, h)]
        for i, (ci, co) in enumerate(chans, start=1):
            setattr(self, "c%d" % i, _conv(ci, co))
        self.c18 = _conv(h, out_features, 1)


class ConvolutionalEncoder(nn.Module):
    """blocks.py:52-75: both UNets are always constructed; only one is used."""

    def __init__(self, in_features, hidden_dim, out_features, n_objects):
        super().__init__()
        c, h, _ = in_features
        self.shallow_unet = ShallowUNet(c, 8, n_objects)
        self.unet = UNet(c, 16, n_objects)
        l1_in = h * h * c if h < 40 else (h // 2) * (h // 2) * c
        self.l1 = nn.Linear(l1_in, hidden_dim)
        self.l2 = nn.Linear(hidden_dim, hidden_dim)
        self.l3 = nn.Linear(hidden_dim, out_features)


class VelocityEncoder(nn.Module):
    """blocks.py:8-29."""

    def __init__(self, alt_vel, input_steps, n_objs, coord_units):
        super().__init__()
        if alt_vel:
            self.init_vel_linear = nn.Linear((input_steps - 1) * 2, 2)
        else:
            self.init_vel_mlp = nn.Sequential(nn.Linear(input_steps * coord_units // n_objs // 2, 100), nn.Tanh(),
                                              nn.Linear(100, 100), nn.Tanh(),
                                              nn.Linear(100, coord_units // n_objs // 2))


class ODECell(nn.RNNCell):
    """cells.py:6-8,24-29,55-58,87-94: an RNNCell subclass, so it carries the (dead) weight_ih/hh, bias_ih/hh."""

    def __init__(self, kind, size):
        super().__init__(size, size)
        self.kind = kind
        self.dt = nn.Parameter(torch.tensor(0.5 if kind == "gravity" else 0.3), requires_grad=False)
        if kind == "spring":
            self.k = nn.Parameter(torch.tensor(np.log(1.0)), requires_grad=True)           # float64, as the reference
            self.equil = nn.Parameter(torch.tensor(np.log(1.0)), requires_grad=True)
        elif kind == "gravity":
            self.g = nn.Parameter(torch.tensor(np.log(1.0)), requires_grad=True)
            self.m = nn.Parameter(torch.tensor(np.log(1.0)), requires_grad=False)
            # the reference caches A = exp(g) exp(2m) here once; the kernels recompute it every step


class _Step(torch.autograd.Function):
    """conv_feedforward as one autograd node: forward = paig_step_forward, backward = paig_step_backward."""

    @staticmethod
    def forward(ctx, net, inp, need_backward, *params):
        out = net._run_forward(inp, need_backward=need_backward)
        ctx.net = net
        ctx.inp = inp
        ctx.lease = out.pop("_lease")             # returns the workspace to the net's pool when this node dies
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(out["enc_masks"], out["masked_objs"])
        net._last = out
        return out["output_seq"], out["recons_out"], out["enc_pos"], out["pos_vel_seq"], out["enc_masks"], out["masked_objs"]

    @staticmethod
    def backward(ctx, d_out, d_rec, d_enc_pos, d_seq, _dm, _dmo):
        grads = ctx.net._run_backward(ctx.inp, ctx.lease.ws, d_out, d_rec, d_enc_pos, d_seq)
        return (None, None, None) + tuple(grads)


class _Lease:
    """A workspace on loan to one autograd node: saved activations live in it between forward and backward.  When the
    node is released (after backward, or when the graph is dropped) the buffer goes back to the net's pool, so a
    training loop re-uses ONE workspace instead of allocating ~1 GB per forward."""

    def __init__(self, pool, ws):
        self.pool, self.ws = pool, ws

    def __del__(self):
        pool = self.pool() if self.pool is not None else None
        if pool is not None and len(pool) < 2:
            pool.append(self.ws)


class _Pool(list):
    pass                                            # a list that can be weakly referenced


class PhysicsNet(BaseNetTorch):
    """physics_models.py:40-111.  Same positional constructor as the reference, same base class surface
    (paig_reproduction_b200/base.py mirrors nn/network/base.py), so runners/torch_run_physics.py drives it unchanged."""

    def __init__(self, task="", recurrent_units=128, lstm_layers=1, cell_type="", seq_len=20, input_steps=3,
                 pred_steps=5, autoencoder_loss=0.0, alt_vel=False, color=False, input_size=36 * 36,
                 encoder_type="conv_encoder", decoder_type="conv_st_decoder", device="cuda"):
        super().__init__()
        assert task in COORD_UNITS                                   # physics_models.py:84
        if cell_type not in CELLS:
            raise KeyError(cell_type)                                # the reference looks the class up by name (:75-77)
        if decoder_type != "conv_st_decoder" or encoder_type != "conv_encoder":
            raise KeyError("%s / %s" % (encoder_type, decoder_type))
        if not color:
            # torch.tile(template, [1,3,1,1]) hard-codes 3 channels: grayscale is broken in the reference
            raise NotImplementedError("only --color configurations are supported (as in the reference, SURVEY Q12)")
        self.device = torch.device(device)
        self.task = task
        self.recurrent_units, self.lstm_layers = recurrent_units, lstm_layers
        self.cell_type = cell_type
        self.cell_kind = CELLS[cell_type]
        self.seq_len, self.input_steps, self.pred_steps = seq_len, input_steps, pred_steps
        assert seq_len > input_steps + pred_steps and input_steps >= 1 and pred_steps >= 1       # :59,85-86
        self.extrap_steps = seq_len - input_steps - pred_steps
        self.autoencoder_loss = autoencoder_loss
        self.alt_vel = bool(alt_vel)
        self.color = color
        self.input_size = input_size
        side = int(np.sqrt(input_size))
        self.conv_ch = 3
        self.input_shape = [3, side, side]
        self.conv_input_shape = [3, side, side]
        self.coord_units = COORD_UNITS[task]
        self.n_objs = self.coord_units // 4
        self.output_shape = self.conv_input_shape
        self.log_sig = 1.0
        self.decoder = self.conv_st_decoder                          # physics_models.py:78-80 (looked up by name there)
        self.extra_valid_fns.append((self.visualize_sequence, [], {}))   # physics_models.py:98-99
        self.extra_test_fns.append((self.visualize_sequence, [], {}))
        tmpl = side // 2
        # construction order = the reference's (physics_models.py:106-111): same default init under the same seed
        self.var_net_content = VariableFromNetwork([self.n_objs, 3, tmpl, tmpl])
        self.var_net_background = VariableFromNetwork([1, 3, side, side])
        self.var_net_template = VariableFromNetwork([self.n_objs, 1, tmpl, tmpl])
        self.encoder = ConvolutionalEncoder(self.conv_input_shape, 200, 2, self.n_objs)
        self.velocity_encoder = VelocityEncoder(self.alt_vel, input_steps, self.n_objs, self.coord_units)
        self.rollout_cell = ODECell(self.cell_kind, self.coord_units // 2)
        self.to(self.device)

        self.H = side
        self.deep = side >= 40                                       # blocks.py:79-82
        self._unet = "unet" if self.deep else "shallow_unet"
        self._n_convs = 18 if self.deep else 13
        self.batch_global = 0           # set by the data-parallel wrapper: loss normalisers use the job's batch
        # the reference's gravity cell evaluates A = exp(g) exp(2m) once, in its constructor, so a checkpoint
        # with g != 0 still rolls out with A = 1 there.  Default here: A follows g (dL/dg flows).  Set
        # freeze_gravity_A = True to roll out with the constructor-time A like the reference does.
        self.freeze_gravity_A = False
        self._gravity_A0 = 1.0          # exp(log 1) * exp(2 log 1): the constructor-time value (cells.py:91-94)
        self._ws_nograd: Dict[int, torch.Tensor] = {}
        self._ws_pools: Dict[tuple, _Pool] = {}
        self._last_ws = None
        self._last: Optional[dict] = None
        self._flat_grad: Optional[torch.Tensor] = None
        self.optimizer = None

    # ------------------------------------------------------------------ C-ABI plumbing
    def _task(self, T: int, inference: bool = False) -> _abi.Task:
        frozen = self._gravity_A0 if (self.freeze_gravity_A and self.cell_kind == "gravity") else 0.0
        return _abi.Task(_abi.CELL_IDS[self.cell_kind], self.n_objs, self.H, T, self.input_steps, self.pred_steps,
                         int(self.alt_vel), int(self.deep), float(self.autoencoder_loss), int(self.batch_global),
                         float(frozen), _abi.FLAG_INFERENCE if inference else 0)

    def live_parameter_names(self, with_rollout: bool = True) -> List[str]:
        """state_dict keys that receive a gradient in a LIVE step, in state_dict order."""
        cache = self.__dict__.setdefault("_live_names", {})
        if with_rollout in cache:
            return list(cache[with_rollout])
        names = cache[with_rollout] = []
        for k, _ in self._named():
            if k.startswith("encoder.") and not (k.startswith("encoder." + self._unet + ".") or k.startswith("encoder.l")):
                continue                                              # the unused UNet
            if k.startswith("rollout_cell."):
                if not with_rollout or k.rsplit(".", 1)[1] not in ("k", "equil", "g"):
                    continue                                          # RNNCell weights, dt, m
            if k.startswith("velocity_encoder.") and not with_rollout:
                continue
            names.append(k)
        return list(names)

    # The table is ~95 pointers; filled field by field through ctypes attribute access (and with named_parameters() walked
    # three times per step) the Python side of the drop-in path cost more than the GPU side (2.4 ms of host time per 2.2 ms
    # of kernels, tools/dropin_probe.py).  paig_params is a flat array of 8-byte pointers, so every state_dict key is
    # resolved ONCE to its slot index (by filling a probe struct with sentinels through _abi.fill_params, the one place that
    # knows the layout) and a table is then ~95 integer stores.
    def _slot_of(self) -> Dict[str, int]:
        bind = self.__dict__.get("_slot_index")
        if bind is None:
            names = [k for k, _ in self._named()]
            probe = _abi.Params()
            sent = {k: 0x1000 + 8 * i for i, k in enumerate(names)}
            _abi.fill_params(probe, lambda k: sent[k], names, self._unet, self._n_convs, self.alt_vel, self.cell_kind)
            arr = (ctypes.c_uint64 * (ctypes.sizeof(_abi.Params) // 8)).from_buffer(probe)
            where = {int(v): i for i, v in enumerate(arr) if v}
            bind = self.__dict__["_slot_index"] = {k: where[sent[k]] for k in names if sent[k] in where}
        return bind

    def _named(self):
        """(name, parameter) pairs in state_dict order, walked once: the modules of a PhysicsNet never change."""
        named = self.__dict__.get("_named_params")
        if named is None:
            named = self.__dict__["_named_params"] = list(self.named_parameters())
        return named

    def _param_table(self, tensors: Dict[str, torch.Tensor]) -> _abi.Params:
        p = _abi.Params()
        arr = (ctypes.c_uint64 * (ctypes.sizeof(_abi.Params) // 8)).from_buffer(p)
        slot = self._slot_of()
        for k, t in tensors.items():
            i = slot.get(k)
            if i is None:
                continue
            if not (t.is_cuda and t.is_contiguous()):
                raise _lib.PaigError("parameters must be contiguous CUDA tensors (no CPU path exists)")
            arr[i] = t.data_ptr()
        return p

    def _params_now(self) -> Dict[str, torch.Tensor]:
        return {k: v.data for k, v in self._named()}

    def _workspace(self, T: int, B: int, fresh: bool, inference: bool = False) -> torch.Tensor:
        lib = _lib.load()
        tk = self._task(T, inference)
        n = lib.paig_workspace_bytes(ctypes.byref(tk), B)
        if n == 0:
            raise _lib.PaigError(lib.paig_last_error().decode())
        key = (T, B, inference)
        if fresh:
            pool = self._ws_pools.setdefault(key, _Pool())
            ws = pool.pop() if pool else torch.empty(n // 4 + 64, dtype=torch.float32, device=self.device)
            return _Lease(weakref.ref(pool), ws)
        ws = self._ws_nograd.get(key)
        if ws is None:
            ws = self._ws_nograd[key] = torch.empty(n // 4 + 64, dtype=torch.float32, device=self.device)
        return ws

    def _check_input(self, inp: torch.Tensor):
        if not inp.is_cuda:
            raise _lib.PaigError("PhysicsNet (B200) needs CUDA input tensors: there is no CPU fallback")
        if inp.dim() != 5 or inp.shape[2] != 3 or inp.shape[3] != self.H or inp.shape[4] != self.H:
            raise ValueError("expected input [B, T, 3, %d, %d], got %s" % (self.H, self.H, tuple(inp.shape)))
        if inp.shape[1] <= self.input_steps + self.pred_steps:
            raise ValueError("sequence too short")

    def _run_forward(self, inp: torch.Tensor, need_backward: bool) -> dict:
        lib = _lib.load()
        x = inp.detach().contiguous().float()
        B, T = x.shape[0], x.shape[1]
        n, H, e, steps = self.n_objs, self.H, self.input_steps + self.pred_steps, T - self.input_steps
        dev = self.device
        # under no_grad (eval_performance, base.py:179) nothing is kept for a backward pass: inference plan, small workspace
        lease = self._workspace(T, B, fresh=True) if need_backward else _Lease(None, self._workspace(T, B, False, inference=True))
        ws = lease.ws
        out = dict(output_seq=torch.empty(B, steps, 3, H, H, device=dev), recons_out=torch.empty(B, e, 3, H, H, device=dev),
                   enc_pos=torch.empty(B, e, 2 * n, device=dev), pos_vel_seq=torch.empty(B, steps + 1, 4 * n, device=dev),
                   enc_masks=torch.empty(B * e, n + 1, H, H, device=dev),
                   masked_objs=torch.empty(n, B * e, 3, H, H, device=dev),
                   templates=torch.empty(n * (H // 2) ** 2 * 4 + 3 * H * H, device=dev), losses=torch.empty(4, device=dev))
        O = _abi.Outputs(*[out[k].data_ptr() for k in ("output_seq", "recons_out", "enc_pos", "pos_vel_seq", "enc_masks",
                                                       "masked_objs", "templates", "losses")])
        tk = self._task(T, inference=not need_backward)
        params = self._params_now()
        P = self._param_table(params)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.paig_step_forward(ctypes.byref(tk), ctypes.byref(P), x.data_ptr(), B, ctypes.byref(O),
                                         ws.data_ptr(), stream), "paig_step_forward")
        out["_lease"] = lease
        out["_x"] = x
        self._last_ws = weakref.ref(ws)           # parity tests read the ReLU decisions of this forward back (tests/stage_checks.py)
        return out

    def _alloc_grads(self, params, live):
        """Fresh gradient tensors for the live parameters: one allocation per dtype, carved into views (90 separate
        torch.empty_like calls were 0.3 ms of host time per step)."""
        by_dtype: Dict[torch.dtype, list] = {}
        for k in live:
            by_dtype.setdefault(params[k].dtype, []).append(k)
        grads = {}
        for dt, names in by_dtype.items():
            sizes = [(params[k].numel() + 3) & ~3 for k in names]          # every view starts 16-byte aligned (fp32) or better
            flat = torch.empty(sum(sizes), dtype=dt, device=self.device)
            for k, chunk in zip(names, flat.split_with_sizes(sizes)):
                grads[k] = chunk[:params[k].numel()].view(params[k].shape)
        return grads

    def _run_backward(self, inp, ws, d_out, d_rec, d_enc_pos, d_seq):
        lib = _lib.load()
        x = inp.detach().contiguous().float()
        B, T = x.shape[0], x.shape[1]
        with_rollout = d_out is not None or d_seq is not None
        live = self.live_parameter_names(with_rollout=True)
        params = self._params_now()
        if self.input_steps == 1:
            # physics_models.py:222-223: vel = zeros, the velocity encoder is not on the graph -> grad stays None
            live = [k for k in live if not k.startswith("velocity_encoder.")]
        grads = self._alloc_grads(params, live)
        P, G = self._param_table(params), self._param_table(grads)
        tk = self._task(T)

        def ptr(t):
            return None if t is None else t.contiguous().float().data_ptr()
        keep = [t.contiguous().float() if t is not None else None for t in (d_out, d_rec, d_enc_pos, d_seq)]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(lib.paig_step_backward(ctypes.byref(tk), ctypes.byref(P), ctypes.byref(G), x.data_ptr(), B,
                                          *[None if t is None else t.data_ptr() for t in keep], ws.data_ptr(), stream),
                   "paig_step_backward")
        result = []
        for k, p in self._named():
            g = grads.get(k)
            if g is not None and not with_rollout and (k.startswith("velocity_encoder.") or k.startswith("rollout_cell.")):
                g = None                 # STALE mode: nothing flows through the rollout, grads stay None
            result.append(g)
        return result

    # ------------------------------------------------------------------ reference surface
    def forward(self, input):                                         # physics_models.py:201-202
        return self.conv_feedforward(input)

    def conv_feedforward(self, inp):                                  # physics_models.py:204-245
        self._check_input(inp)
        self.input = inp
        params = [p for _, p in self._named()]
        output_seq, recons_out, enc_pos, pos_vel_seq, enc_masks, masked = _Step.apply(self, inp, torch.is_grad_enabled(), *params)
        last = self._last
        self.recons_out = recons_out
        self.enc_pos = enc_pos
        self.pos_vel_seq = pos_vel_seq
        self.enc_masks = enc_masks
        self.masked_objs = [masked[o] for o in range(self.n_objs)]
        n, t, H = self.n_objs, self.H // 2, self.H
        raw = last["templates"]
        self.template = raw[:n * t * t].view(n, 1, t, t)
        self.contents = raw[n * t * t:4 * n * t * t].view(n, 3, t, t)
        self.background_content = torch.sigmoid(raw[4 * n * t * t:].view(1, 3, H, H))
        self.step_losses = last["losses"]          # [train, pred, extrap, recons] reduced in-kernel (no autograd)
        self._layers = None                        # transf_contents / transf_masks of the last decoder call, on demand
        self._last_decode_loc = pos_vel_seq.detach()[:, -1, :2 * n]
        return output_seq

    # ---- the decoder as a method (physics_models.py:78-80,151-199) and the per-layer tensors it caches ----
    def _decoder_constants(self):
        """paig_templates_forward: raw VariableFromNetwork outputs and [template+5 | sigmoid(contents) | sigmoid(bg)]."""
        lib = _lib.load()
        n, t, H = self.n_objs, self.H // 2, self.H
        CN = n * t * t * 4 + 3 * H * H
        raw = torch.empty(CN, device=self.device)
        consts = torch.empty(CN, device=self.device)
        hidden = torch.empty(3 * 200, device=self.device)
        tk, P = self._task(self.seq_len), self._param_table(self._params_now())
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(lib.paig_templates_forward(ctypes.byref(tk), ctypes.byref(P), raw.data_ptr(), consts.data_ptr(),
                                              hidden.data_ptr(), stream), "paig_templates_forward")
        return raw, consts

    def conv_st_decoder(self, inp):
        """physics_models.py:151-199 for positions ``inp`` [N, 2n] -> frames [N, 3, H, H] (no autograd through this
        stand-alone call; inside ``forward`` the decoder runs -- and is differentiated -- within the fused step)."""
        lib = _lib.load()
        if not inp.is_cuda:
            raise _lib.PaigError("conv_st_decoder needs CUDA tensors: there is no CPU fallback")
        loc = inp.detach().contiguous().float()
        N, n, t, H = loc.shape[0], self.n_objs, self.H // 2, self.H
        raw, consts = self._decoder_constants()
        self.template = raw[:n * t * t].view(n, 1, t, t)
        self.contents = raw[n * t * t:4 * n * t * t].view(n, 3, t, t)
        self.background_content = consts[4 * n * t * t:].view(1, 3, H, H)
        frames = torch.empty(N, 3, H, H, device=self.device)
        tk = self._task(self.seq_len)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(lib.paig_decode_forward(ctypes.byref(tk), consts.data_ptr(), loc.data_ptr(), N, frames.data_ptr(), None, 0,
                                           1, None, stream), "paig_decode_forward")
        self._layers = None
        self._last_decode_loc = loc
        return frames

    def _decode_layers(self):
        if getattr(self, "_layers", None) is None:
            lib = _lib.load()
            loc = self._last_decode_loc.contiguous().float()
            N, n, H = loc.shape[0], self.n_objs, self.H
            _, consts = self._decoder_constants()
            tc = torch.empty(n + 1, N, 3, H, H, device=self.device)
            tm = torch.empty(n + 1, N, 3, H, H, device=self.device)
            tk = self._task(self.seq_len)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(lib.paig_decode_layers(ctypes.byref(tk), consts.data_ptr(), loc.data_ptr(), N, tc.data_ptr(),
                                              tm.data_ptr(), stream), "paig_decode_layers")
            self._layers = (tc, tm)
        return self._layers

    @property
    def transf_contents(self):
        """physics_models.py:190: list of n+1 tensors [N,3,H,H] of the LAST decoder call (the final rollout step):
        each object's sampled sigmoid(content), then the tiled background.  Materialised on first access."""
        return list(self._decode_layers()[0].unbind(0))

    @property
    def transf_masks(self):
        """physics_models.py:192-196: tuple of n+1 softmax masks [N,3,H,H] of the last decoder call."""
        return tuple(self._decode_layers()[1].unbind(0))

    def visualize_sequence(self):
        """physics_models.py:247-330 (registered as extra valid / test fn, :98-99): example%d.jpg, animation%d.gif,
        templates.jpg and extra_outputs.npz in save_dir."""
        from .viz import visualize_sequence
        return visualize_sequence(self, logger)

    def compute_loss(self):                                           # physics_models.py:119-142
        from .losses import frame_sse
        e, i, p = self.input_steps + self.pred_steps, self.input_steps, self.pred_steps
        x = self.input.detach()
        Bg = self.batch_global or x.shape[0]
        scale = float(x.shape[0]) / Bg                                # shards of a data-parallel job sum to the job's mean
        recons = frame_sse(x, 0, e, self.recons_out)                  # [B, e]
        self.recons_loss = torch.mean(recons) * scale
        loss = frame_sse(x, i, x.shape[1] - i, self.output)           # [B, T-in]
        self.pred_loss = torch.mean(loss[:, :p]) * scale
        self.extrap_loss = torch.mean(loss[:, p:]) * scale
        train_loss = self.pred_loss
        if self.autoencoder_loss > 0.0:
            train_loss += self.autoencoder_loss * self.recons_loss    # in place, as the reference
        eval_losses = [self.pred_loss, self.extrap_loss, self.recons_loss]
        return train_loss, eval_losses

    def build_optimizer(self, base_lr, optimizer="rmsprop", anneal_lr=True):     # physics_models.py:144-149
        self.base_lr = base_lr
        self.anneal_lr = anneal_lr
        self.lr = base_lr
        self.optimizer = OPTIMIZERS[optimizer](self.parameters(), self.lr)

    def get_batch(self, batch_size, iterator):                        # physics_models.py:113-117
        batch_x, _ = iterator.next_batch(batch_size)
        return {"input": batch_x}, (batch_x, None)

    # ------------------------------------------------------------------ fused LIVE step (the fast path)
    def flat_order(self) -> List[str]:
        """fp32 live parameters in flat-buffer order: everything EXCEPT the UNet conv layers first (their gradients are
        final before the UNet backward starts, so a data-parallel job can all-reduce that prefix underneath it), the UNet
        conv layers last."""
        params = dict(self.named_parameters())
        names = [k for k in self.live_parameter_names() if params[k].dtype == torch.float32]
        conv = "encoder." + self._unet + "."
        return [k for k in names if not k.startswith(conv)] + [k for k in names if k.startswith(conv)]

    def flat_gradients(self) -> torch.Tensor:
        """One contiguous fp32 buffer holding the gradient of every live fp32 parameter (+ the 4 loss scalars at the
        end), so a data-parallel job needs one all-reduce (or two: ``flat_early`` floats that are final early, then the
        rest).  ``p.grad`` of each parameter is a view into it."""
        if self._flat_grad is None:
            params = dict(self.named_parameters())
            names = self.flat_order()
            conv = "encoder." + self._unet + "."
            total = sum(params[k].numel() for k in names)
            total_al = (total + 3) // 4 * 4
            flat = torch.zeros(total_al + 4, dtype=torch.float32, device=self.device)
            off = 0
            self._grad_views = {}
            self.flat_early = 0
            for k in names:
                n = params[k].numel()
                self._grad_views[k] = flat[off:off + n].view_as(params[k])
                off += n
                if not k.startswith(conv):
                    self.flat_early = off
            self._loss_view = flat[total_al:total_al + 4]
            self._phys_grad = torch.zeros(2, dtype=torch.float64, device=self.device)
            i = 0
            for k in self.live_parameter_names():
                if params[k].dtype == torch.float64:
                    self._grad_views[k] = self._phys_grad[i].view(())
                    i += 1
            self._flat_grad = flat
        return self._flat_grad

    def train_step(self, inp: torch.Tensor) -> torch.Tensor:
        """LIVE step in one library call (paig_step_fused): forward, losses and every parameter gradient.
        Sets ``p.grad`` (views into ``flat_gradients()``) and returns the device tensor
        [train, pred, extrap, recons]; the caller runs ``optimizer.step()``.  No frames are materialised."""
        self._check_input(inp)
        lib = _lib.load()
        x = inp.detach().contiguous().float()
        B, T = x.shape[0], x.shape[1]
        self.flat_gradients()
        params = self._params_now()
        P, G = self._param_table(params), self._param_table(self._grad_views)
        tk = self._task(T)
        ws = self._workspace(T, B, fresh=False)
        O = _abi.Outputs(None, None, None, None, None, None, None, self._loss_view.data_ptr())
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(lib.paig_step_fused(ctypes.byref(tk), ctypes.byref(P), ctypes.byref(G), x.data_ptr(), B,
                                       ctypes.byref(O), ws.data_ptr(), stream), "paig_step_fused")
        for k, p in self._named():
            v = self._grad_views.get(k)
            if v is not None:
                p.grad = v
        return self._loss_view

    def train_step_graph(self, inp: torch.Tensor) -> torch.Tensor:
        """``train_step`` replayed from a CUDA graph: the step's 44 launches, side streams included, are recorded once per
        (input buffer, shape, parameter / gradient / workspace addresses, loss normalisation) and replayed afterwards, which
        removes the launch gaps between its dependent kernels (spring_color B = 100: 2.05 -> 2.02 ms, bit-identical
        gradients; tests/test_gpu_module.py).  The graph bakes in addresses, not values: in-place parameter updates
        (optimizer.step, load_state_dict) are seen by the next replay; anything that moves a tensor records a new graph.
        Falls back to the eager ``train_step`` when recording fails."""
        self._check_input(inp)
        if not inp.is_contiguous() or inp.dtype != torch.float32:
            return self.train_step(inp)
        self.flat_gradients()
        params = self._params_now()
        ws = self._workspace(inp.shape[1], inp.shape[0], fresh=False)
        key = (inp.data_ptr(), tuple(inp.shape), ws.data_ptr(), self._loss_view.data_ptr(),
               int(self.batch_global), float(self.autoencoder_loss), bool(self.freeze_gravity_A),     # by-value fields of paig_task
               tuple(v.data_ptr() for v in params.values()), tuple(v.data_ptr() for v in self._grad_views.values()))
        cache = self.__dict__.setdefault("_step_graphs", {})
        entry = cache.get(key)
        if entry is None:
            entry = cache[key] = self._record_step_graph(inp)
            if len(cache) > 64:                      # a caller cycling through fresh buffers: do not hoard graphs
                cache.pop(next(iter(cache)))
        if entry is False:
            return self.train_step(inp)
        graph, launches = entry
        graph.replay()
        self.graph_replay_launches = getattr(self, "graph_replay_launches", 0) + launches
        for k, p in self._named():
            v = self._grad_views.get(k)
            if v is not None:
                p.grad = v
        return self._loss_view

    def _record_step_graph(self, inp: torch.Tensor):
        lib = _lib.load()
        self.train_step(inp)                         # eager once: lazily created streams / events / attributes exist
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        try:
            graph = torch.cuda.CUDAGraph()
            n0 = lib.paig_launch_count()
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side, capture_error_mode="relaxed"):
                    self.train_step(inp)
            launches = int(lib.paig_launch_count() - n0)
        except Exception:                            # noqa: BLE001 -- any capture failure: the eager path is always valid
            torch.cuda.synchronize(self.device)
            return False
        finally:
            cur.wait_stream(side)
        return graph, launches
