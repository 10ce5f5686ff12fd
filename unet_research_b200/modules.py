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


def _dilate(table, d, centers, bits, keep):
    """Centre bitmap -> NHWC keep mask + keep count for a one-entry call table: the v2 kernels (sparse NHWC scatter +
    word-parallel 7x7 OR) for the reference's block size 7, the generic bit-transpose kernel otherwise (B2U_DILATE=v1
    forces it)."""
    import os
    if d.block_size == 7 and os.environ.get("B2U_DILATE", "v2") != "v1":
        scatter = torch.empty_like(bits)
        call("b2u_dropblock_dilate_v2", ptr(table), 1, C.byref(d), ptr(centers), ptr(scatter), bits.numel(), ptr(bits), ptr(keep),
             stream_ptr())
    else:
        call("b2u_dropblock_dilate", ptr(table), 1, C.byref(d), ptr(centers), ptr(bits), ptr(keep), stream_ptr())


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
        _dilate(table, d, centers, bits, keep)
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


class Dropblock2d_ichan(nn.Module):
    """`Dropblock2d_ichan(drop_prob=0.5, block_size=3)` (reference utils_modules.py:86-139), the variant the six
    multi-fidelity scripts train with (MF-training-UNI.py:244): centres are `torch.bernoulli(gamma)` draws over the
    full feature map with a border of block_size//2 zeroed, gamma is clamped to 1, and the rescale is skipped when
    everything was dropped.  Same setters/getters as the reference; inside a `UNet` the masks come from the fused CUDA
    kernels (bit-exact with torch.bernoulli's CUDA stream), stand-alone `forward` applies them with the reference's
    arithmetic."""

    def __init__(self, drop_prob=0.5, block_size=3):
        super().__init__()
        self.drop_prob = drop_prob
        self.block_size = block_size

    def set_drop_prob(self, p):
        self.drop_prob = p

    def get_drop_prob(self):
        return self.drop_prob

    def get_gamma(self, feat_x, feat_y):
        keep_prob = 1 - self.drop_prob
        gamma = (1 - keep_prob) / (self.block_size ** 2) * (feat_x * feat_y) / ((feat_x - self.block_size + 1) * (feat_y - self.block_size + 1))
        return min(gamma, 1)

    def extra_repr(self):
        return f"drop_prob={self.drop_prob}, block_size={self.block_size}"

    def block_mask(self, x: torch.Tensor):
        """(keep mask [N,C,H,W] float 0/1, keep_count int64[1]) for the current CUDA generator state; advances the
        generator exactly like the reference's single torch.bernoulli call."""
        from .engine import MaskPlan, BERNOULLI_OFFSET_INCREMENT
        if not x.is_cuda:
            raise _lib.B2uError("Dropblock2d_ichan needs a CUDA tensor: there is no CPU path")
        n, c, h, w = x.shape
        if c % 32 != 0:
            raise NotImplementedError("CUDA DropBlock needs channels to be a multiple of 32")
        plan = _SingleSitePlan(n, c, h, w, float(self.drop_prob), int(self.block_size), x.device, "ichan")
        gen = torch.cuda.default_generators[x.device.index if x.device.index is not None else torch.cuda.current_device()]
        plan.run(gen.initial_seed(), gen.get_offset())
        gen.set_offset(gen.get_offset() + BERNOULLI_OFFSET_INCREMENT)
        return plan.mask_nchw(x.dtype), plan.keep

    def forward(self, tensor):
        if not self.training or self.drop_prob == 0.:
            return tensor
        mask, _ = self.block_mask(tensor)
        tensor *= mask                                             # in place, as the reference (:128)
        total = mask.numel()
        scale_denominator = 1. - torch.true_divide(total - torch.sum(mask), total)
        if scale_denominator != 0:
            tensor *= 1. / scale_denominator
        return tensor


class _SingleSitePlan:
    """One DropBlock call outside a U-Net (stand-alone layer use): table of one entry + its bitmaps."""

    def __init__(self, n, c, h, w, drop_prob, block_size, device, mode):
        from .engine import _thresholds_cached
        if block_size % 2 == 0 or block_size > 31:
            raise NotImplementedError("CUDA DropBlock supports odd block_size <= 31")
        self.n, self.c, self.h, self.w, self.mode = n, c, h, w, mode
        hc, wc = h - block_size + 1, w - block_size + 1
        if hc <= 0 or wc <= 0:
            raise ValueError(f"feature map {h}x{w} smaller than block_size {block_size}")
        d = DropblockCall()
        d.philox_offset, d.center_word_off, d.mask_word_off = 0, 0, 0
        d.numel = n * c * h * w if mode == "ichan" else n * c * hc * wc
        d.thresh_lo, d.thresh_hi = _thresholds_cached(drop_prob, block_size, h, w, mode)
        d.n_img, d.c, d.h, d.w, d.block_size, d.count_index = n, c, h, w, block_size, 0
        self.d = d
        self.words = (n * c * hc * wc + 31) // 32 + 4
        self.centers = torch.zeros(self.words, dtype=torch.int32, device=device)
        self.bits = torch.empty(n * h * w * (c // 32), dtype=torch.int32, device=device)
        self.keep = torch.zeros(1, dtype=torch.int64, device=device)
        self.device = device

    def run(self, seed: int, offset: int):
        self.d.philox_offset = offset
        table = torch.from_numpy(np.frombuffer(bytes(self.d), dtype=np.uint8).copy()).to(self.device)
        self.keep.zero_()
        call("b2u_dropblock_centers_ichan", ptr(table), 1, C.byref(self.d), C.c_uint64(seed & (2 ** 64 - 1)), None,
             ptr(self.centers), self.words, stream_ptr())
        _dilate(table, self.d, self.centers, self.bits, self.keep)

    def mask_nchw(self, dtype):
        shifts = torch.arange(32, device=self.device, dtype=torch.int32)
        m = ((self.bits.view(self.n, self.h, self.w, self.c // 32, 1) >> shifts) & 1).reshape(self.n, self.h, self.w, self.c)
        return m.permute(0, 3, 1, 2).to(dtype)


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
