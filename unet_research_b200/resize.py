"""Fused `square_pad` + `TF.resize` (reference utils_general.py:32-43,
MF-training-UNI.py:54-73, Dropblock_Uncertainty.py:52-61) for inference-side use: one kernel, the padded square is
never materialised.  Forward only -- the multi-fidelity TRAINING steps that differentiate through the resize keep
calling torchvision, which composes with `UNet`'s autograd bridge unchanged."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


def square_pad_resize(tensor: torch.Tensor, size, square_pad: bool = True) -> torch.Tensor:
    """`TF.resize(square_pad(tensor), size=(s, s))` for an fp32 CUDA tensor [..., H, W]; `size` int or (oh, ow)."""
    if not tensor.is_cuda:
        raise _lib.B2uError("square_pad_resize runs on CUDA tensors only (no CPU path)")
    oh, ow = (int(size), int(size)) if isinstance(size, int) else (int(size[0]), int(size[1]))
    x = tensor.detach().to(torch.float32).contiguous()
    h, w = x.shape[-2], x.shape[-1]
    planes = x.numel() // (h * w)
    out = torch.empty(*x.shape[:-2], oh, ow, dtype=torch.float32, device=x.device)
    call("b2u_square_pad_resize", ptr(x), ptr(out), planes, h, w, int(square_pad), oh, ow, stream_ptr())
    return out
