"""Drop-in for ``generative.losses.JukeboxLoss`` (monai-generative) as the reference uses it
(``src/train_autoencoderkl.py:158,208``: ``JukeboxLoss(spatial_dims=1, reduction="sum")``): batched cuFFT R2C + one custom
magnitude / squared-error kernel; the gradient w.r.t. ``input`` comes from one C2R (``eegldm_jukebox_loss``)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._module import check_cuda_f32


class _JukeboxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, target, reduction):
        # inp / target arrive fp32 and contiguous (JukeboxLoss.forward normalises them OUTSIDE this Function, where autograd
        # tracks the cast / copy); whether a gradient is wanted is a property of the argument, not of a detached copy
        B, Cc, N = inp.shape
        loss = torch.empty((), device=inp.device, dtype=torch.float32)
        grad = torch.empty_like(inp) if ctx.needs_input_grad[0] else None
        with torch.cuda.device(inp.device):
            _lib.check(_lib.lib().eegldm_jukebox_loss(
                C.c_void_p(inp.data_ptr()), C.c_void_p(target.data_ptr()), int(B), int(Cc), int(N), int(reduction),
                C.c_void_p(loss.data_ptr()), C.c_void_p(grad.data_ptr()) if grad is not None else None,
                C.c_void_p(_lib.current_stream_ptr(inp.device))))
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g if grad is not None else None), None, None


class JukeboxLoss(torch.nn.Module):
    def __init__(self, spatial_dims: int = 1, fft_signal_size=None, fft_norm: str = "ortho", reduction: str = "mean"):
        super().__init__()
        if spatial_dims != 1 or fft_signal_size is not None or fft_norm != "ortho":
            raise NotImplementedError("the reference uses JukeboxLoss(spatial_dims=1) with the default ortho norm")
        if reduction not in ("sum", "mean"):
            raise NotImplementedError("reduction must be 'sum' or 'mean'")
        self.reduction = reduction

    def forward(self, input, target):
        if input.shape != target.shape or input.dim() != 3:
            raise ValueError("input/target must both be [B, C, N]")
        # dtype / contiguity normalisation happens here, under autograd (a sliced reconstruction such as recon[:, :, 36:-36]
        # or a half tensor under autocast keeps its gradient path)
        input, target = check_cuda_f32(input, "input"), check_cuda_f32(target.detach(), "target")
        return _JukeboxFn.apply(input, target, 0 if self.reduction == "sum" else 1)
