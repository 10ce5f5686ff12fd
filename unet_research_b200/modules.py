"""Drop-in mirrors of the reference's custom layers (unet_code/utils/utils_modules.py).

`DropBlock2D(drop_prob, block_size)` keeps the reference's constructor order (:30), its mutable
`.drop_prob` / `.block_size` attributes and its "active only while `self.training`" rule (:42-43), so
`model.apply(set_dropblock_on)` (Dropblock_Uncertainty.py:22-25) works unchanged.  Inside a `UNet` the
layer is never called as a module: the U-Net forward builds all 22 masks of a pass with the fused CUDA
mask kernels and folds mask-apply and rescale into the GroupNorm/ReLU kernels.  Called stand-alone on a
CUDA NCHW tensor, `forward` runs the same two mask kernels and applies the mask with the exact
reference arithmetic (`x * mask * numel / sum`, :61-64).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import nn

from . import _lib
from ._lib import DropblockCall, call, ptr, stream_ptr


class DropBlock2D(nn.Module):
    def __init__(self, drop_prob, block_size):
        super().__init__()
        self.drop_prob = drop_prob
        self.block_size = block_size

    def extra_repr(self):
        return f"drop_prob={self.drop_prob}, block_size={self.block_size}"

    def block_mask(self, x: torch.Tensor):
        """Bit-exact reproduction of the reference mask for the current CUDA generator state: returns
        (block_mask [N,C,H,W] float 0/1, keep_count int64[1]) and advances the generator like torch.rand."""
        from .engine import dropblock_gamma, philox_thresholds, rand_grid, rand_offset_increment, device_rand_geometry
        if not x.is_cuda:
            raise _lib.B2uError("DropBlock2D needs a CUDA tensor: there is no CPU path")
        n, c, h, w = x.shape
        bs = int(self.block_size)
        if bs % 2 == 0 or bs > 31:
            raise NotImplementedError("CUDA DropBlock supports odd block_size <= 31")
        if c % 32 != 0:
            raise NotImplementedError("CUDA DropBlock needs channels to be a multiple of 32")
        sms, mt = device_rand_geometry()
        numel = n * c * (h - bs + 1) * (w - bs + 1)
        gen = torch.cuda.default_generators[x.device.index if x.device.index is not None else torch.cuda.current_device()]
        seed, offset = gen.initial_seed(), gen.get_offset()
        d = DropblockCall()
        d.philox_offset, d.center_word_off, d.mask_word_off = offset, 0, 0
        d.numel, d.grid = numel, rand_grid(numel, sms, mt)
        d.thresh_lo, d.thresh_hi = philox_thresholds(dropblock_gamma(float(self.drop_prob), bs, h, w))
        d.n_img, d.c, d.h, d.w, d.block_size, d.count_index = n, c, h, w, bs, 0
        table = torch.from_numpy(np.frombuffer(bytes(d), dtype=np.uint8).copy()).to(x.device)
        centers = torch.zeros((numel + 31) // 32 + 4, dtype=torch.int32, device=x.device)
        bits = torch.empty(n * h * w * (c // 32), dtype=torch.int32, device=x.device)
        keep = torch.zeros(1, dtype=torch.int64, device=x.device)
        call("b2u_dropblock_centers", ptr(table), 1, C.c_uint64(seed & (2 ** 64 - 1)), None, ptr(centers), stream_ptr())
        call("b2u_dropblock_dilate", ptr(table), 1, C.byref(d), ptr(centers), ptr(bits), ptr(keep), stream_ptr())
        gen.set_offset(offset + rand_offset_increment(numel, sms, mt))
        shifts = torch.arange(32, device=x.device, dtype=torch.int32)
        m = ((bits.view(n, h, w, c // 32, 1) >> shifts) & 1).reshape(n, h, w, c).permute(0, 3, 1, 2)
        return m.to(x.dtype), keep

    def forward(self, x):
        if not self.training or self.drop_prob == 0.:
            return x
        block_mask, _ = self.block_mask(x)
        out = x * block_mask
        out = out * block_mask.numel() / block_mask.sum()
        return out


class LinearScheduler(nn.Module):
    """`dropblock==0.3.0` LinearScheduler (third-party; imported by the reference at
    utils_modules.py:1 and used at utils_unet.py:128-132, 410-411).  Same attribute names
    (`dropblock`, `i`, `drop_values`) and the same `step()` rule."""

    def __init__(self, dropblock, start_value, stop_value, nr_steps):
        super().__init__()
        self.dropblock = dropblock
        self.i = 0
        self.drop_values = np.linspace(start=start_value, stop=stop_value, num=int(nr_steps))

    def forward(self, x):
        return self.dropblock(x)

    def step(self):
        if self.i < len(self.drop_values):
            self.dropblock.drop_prob = self.drop_values[self.i]
        self.i += 1
