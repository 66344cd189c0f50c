"""Per-frame squared error of compute_loss (physics_models.py:122-131) as an autograd node over the
library's frame_sse kernels (no torch elementwise ops on full frames)."""
from __future__ import annotations

import torch

from . import _lib


class _FrameSSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, first, count, pred):
        lib = _lib.load()
        if not (x.is_cuda and pred.is_cuda):
            raise _lib.PaigError("frame_sse needs CUDA tensors: there is no CPU fallback")
        x = x.contiguous().float()
        predc = pred.detach().contiguous().float()
        B, T = x.shape[0], x.shape[1]
        chw = x.shape[2] * x.shape[3] * x.shape[4]
        if predc.shape[0] != B or predc.shape[1] != count or predc[0, 0].numel() != chw or first + count > T:
            raise ValueError("frame_sse: shape mismatch")
        sse = torch.empty(B, count, device=x.device, dtype=torch.float32)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(lib.paig_frame_sse_forward(x.data_ptr(), T * chw, first, predc.data_ptr(), B, count, chw,
                                              sse.data_ptr(), stream), "paig_frame_sse_forward")
        ctx.save_for_backward(x, predc)
        ctx.first, ctx.count = first, count
        return sse

    @staticmethod
    def backward(ctx, d_sse):
        lib = _lib.load()
        x, pred = ctx.saved_tensors
        B, T = x.shape[0], x.shape[1]
        chw = x.shape[2] * x.shape[3] * x.shape[4]
        d_pred = torch.empty_like(pred)
        d = d_sse.contiguous().float()
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(lib.paig_frame_sse_backward(x.data_ptr(), T * chw, ctx.first, pred.data_ptr(), B, ctx.count, chw,
                                               d.data_ptr(), d_pred.data_ptr(), stream), "paig_frame_sse_backward")
        return None, None, None, d_pred


def frame_sse(x, first, count, pred):
    """[B, count]: sum over (c,h,w) of (x[:, first:first+count] - pred)^2; differentiable w.r.t. pred."""
    return _FrameSSE.apply(x, first, count, pred)
