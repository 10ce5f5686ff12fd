"""Fused `square_pad` + `TF.resize` (reference utils_general.py:32-43,
MF-training-UNI.py:54-73, Dropblock_Uncertainty.py:52-61): one kernel, the padded square is never materialised.
Differentiable: the multi-fidelity TRAINING steps resize the segmentation back up before the loss
(MF-training-UNI.py:66-69), so the gradient flows through the resize -- `b2u_square_pad_resize_bwd` is the adjoint of
the same anti-aliased filter (same weights, gather form)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


def _sizes(tensor, size):
    oh, ow = (int(size), int(size)) if isinstance(size, int) else (int(size[0]), int(size[1]))
    h, w = tensor.shape[-2], tensor.shape[-1]
    return oh, ow, h, w, tensor.numel() // (h * w)


class _SquarePadResize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tensor, oh, ow, square_pad):
        x = tensor.detach().to(torch.float32).contiguous()
        h, w = x.shape[-2], x.shape[-1]
        planes = x.numel() // (h * w)
        out = torch.empty(*x.shape[:-2], oh, ow, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            call("b2u_square_pad_resize", ptr(x), ptr(out), planes, h, w, int(square_pad), oh, ow, stream_ptr())
        ctx.geom = (tuple(x.shape), planes, h, w, oh, ow, int(square_pad), tensor.dtype)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        shape, planes, h, w, oh, ow, square_pad, dtype = ctx.geom
        g = grad_out.detach().to(torch.float32).contiguous()
        gin = torch.empty(shape, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            call("b2u_square_pad_resize_bwd", ptr(g), ptr(gin), planes, h, w, square_pad, oh, ow, stream_ptr())
        return gin.to(dtype), None, None, None


def square_pad_resize(tensor: torch.Tensor, size, square_pad: bool = True) -> torch.Tensor:
    """`TF.resize(square_pad(tensor), size=(s, s))` (square_pad=False: plain `TF.resize`) for a CUDA tensor [..., H, W];
    `size` int or (oh, ow).  fp32 result; differentiable w.r.t. `tensor`."""
    if not tensor.is_cuda:
        raise _lib.B2uError("square_pad_resize runs on CUDA tensors only (no CPU path)")
    oh, ow, _, _, _ = _sizes(tensor, size)
    return _SquarePadResize.apply(tensor, oh, ow, bool(square_pad))
