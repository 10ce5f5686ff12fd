"""Deterministic synthetic workload: weights, DRIVE-shaped image, ground truth and FOV mask.

There is no network for the DRIVE dataset or trained checkpoints, so every test and the
bench use tensors produced here.  Shapes and statistics follow SURVEY.md section 8(d):
584x565 one-channel image in [0,1), vessel fraction 8.6 %, centred FOV disc of radius 270 px.

The state-dict key layout is the reference's (`utils_unet.py:98-115`, probed key list in
SURVEY.md section 8b); values are drawn from per-key seeded CPU generators so that the
same weights can be rebuilt anywhere (here, on the GPU box, in the golden-vector script)
without shipping a 124 MB checkpoint.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch

H_DRIVE, W_DRIVE = 584, 565


def unet_param_shapes(init_channels: int = 1, filters: int = 64, output_channels: int = 1,
                      model_depth: int = 4) -> "OrderedDict[str, tuple]":
    """Key -> shape, in the reference's registration order (`utils_unet.py:98-115`)."""
    shapes: "OrderedDict[str, tuple]" = OrderedDict()

    def conv_unit(prefix: str, idx: int, cin: int, cout: int):
        shapes[f"{prefix}.{idx}.weight"] = (cout, cin, 3, 3)
        shapes[f"{prefix}.{idx + 1}.weight"] = (cout,)
        shapes[f"{prefix}.{idx + 1}.bias"] = (cout,)

    f = filters
    cin = init_channels
    for lvl in range(model_depth):
        cout = f if lvl == 0 else f * 2
        conv_unit(f"down_blocks.{lvl}.0", 0, cin, cout)
        conv_unit(f"down_blocks.{lvl}.0", 4, cout, cout)
        shapes[f"down_blocks.{lvl}.1.1.weight"] = (cout,)
        shapes[f"down_blocks.{lvl}.1.1.bias"] = (cout,)
        f = cout
        cin = cout
    conv_unit("conn_block", 0, f, f * 2)
    conv_unit("conn_block", 4, f * 2, f * 2)
    f *= 2
    for u in range(model_depth):
        shapes[f"up_blocks.{u}.0.0.weight"] = (f, f // 2, 2, 2)
        f //= 2
        shapes[f"up_blocks.{u}.0.1.weight"] = (f,)
        shapes[f"up_blocks.{u}.0.1.bias"] = (f,)
        conv_unit(f"up_blocks.{u}.1", 0, f * 2, f)
        conv_unit(f"up_blocks.{u}.1", 4, f, f)
    shapes["output_conv.0.weight"] = (output_channels, f, 1, 1)
    return shapes


def make_state_dict(init_channels: int = 1, filters: int = 64, output_channels: int = 1,
                    model_depth: int = 4, seed: int = 1234) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic fp32 state dict.  Conv weights ~ U(-b, b), b = sqrt(3 / fan_in)
    (unit-variance-preserving), GroupNorm weight ~ 1 + 0.1 N(0,1), bias ~ 0.1 N(0,1) so the
    affine path is exercised (the default init, weight 1 / bias 0, would hide bugs there)."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape in unet_param_shapes(init_channels, filters, output_channels, model_depth).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63))
        if len(shape) == 4:
            if key.startswith("up_blocks") and key.endswith(".0.0.weight"):
                fan_in = shape[0]            # ConvTranspose k2 s2: one tap per output pixel
            else:
                fan_in = shape[1] * shape[2] * shape[3]
            b = math.sqrt(3.0 / fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif key.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            t = 0.1 * torch.randn(shape, generator=g)
        sd[key] = t.float()
    return sd


def make_image(h: int = H_DRIVE, w: int = W_DRIVE, channels: int = 1, seed: int = 1234,
               batch: int = 1) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, channels, h, w, generator=g)


def make_gt(h: int = H_DRIVE, w: int = W_DRIVE, seed: int = 1234, batch: int = 1,
            vessel_fraction: float = 0.086) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed + 1)
    return (torch.rand(batch, 1, h, w, generator=g) < vessel_fraction).float()


def make_fov_mask(h: int = H_DRIVE, w: int = W_DRIVE, batch: int = 1) -> torch.Tensor:
    """Centred disc, radius 270/584 of the height (about 69 % coverage at DRIVE size)."""
    r = 270.0 * h / H_DRIVE
    yy = torch.arange(h, dtype=torch.float32).view(h, 1) - (h - 1) / 2
    xx = torch.arange(w, dtype=torch.float32).view(1, w) - (w - 1) / 2
    m = ((yy * yy + xx * xx) <= r * r).float()
    return m.view(1, 1, h, w).repeat(batch, 1, 1, 1)
