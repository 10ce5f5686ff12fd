"""ORACLE (test infrastructure, not product code).

A plain-PyTorch fp32/fp64 restatement of the reference's U-Net hot path, written as pure
functions over a reference-layout state dict.  Only `tests/`, `__graft_entry__.smoke()` and
the `cpu_baseline` / `--impl reference` legs of `bench.py` may import this module; the
product package `unet_research_b200` never does.

Every function cites the reference lines it restates (R = /root/reference/Unet_research/
unet_code).  The restatement is pinned against the real reference modules, imported in the
build container by `oracle/ref_shims.py`, through the golden vectors under `tests/golden/`
(`tests/golden/make_golden.py` generates them, `tests/test_oracle_golden.py` checks them).
The one boundary that stays "parity unpinned" is `LinearScheduler` (third-party
`dropblock==0.3.0`, absent from /root/reference and from this image): restated from the
published package source, see `linear_scheduler_values`.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------- DropBlock
def dropblock_gamma(drop_prob: float, block_size: int, h: int, w: int) -> float:
    """R/utils/utils_modules.py:81-82 (`DropBlock2D._compute_gamma`)."""
    return drop_prob * h * w / ((block_size ** 2) * (h - block_size + 1) * (w - block_size + 1))


def dropblock_block_mask(mask_center: Tensor, block_size: int) -> Tensor:
    """R/utils/utils_modules.py:51-58,68-79: zero-pad the centre mask by bs//2, 7x7 stride-1
    max-pool with padding bs//2, then `1 - .` (1 = keep, 0 = dropped)."""
    mask = F.pad(mask_center, (block_size // 2,) * 4)
    if block_size % 2 == 0:
        mask = mask[:, :, :-1, :-1]
    bm = F.max_pool2d(mask, kernel_size=(block_size, block_size), stride=(1, 1), padding=block_size // 2)
    if block_size % 2 == 0:
        bm = bm[:, :, :-1, :-1]
    return 1 - bm


def dropblock2d(x: Tensor, drop_prob: float, block_size: int, training: bool = True,
                rand_fn: Callable = torch.rand, record: Optional[list] = None) -> Tensor:
    """R/utils/utils_modules.py:36-66 (`DropBlock2D.forward`).  `rand_fn` is the uniform source
    (the reference calls `torch.rand(N, C, H-bs+1, W-bs+1, device=x.device)`, :49)."""
    if not training or drop_prob == 0.:
        return x
    gamma = dropblock_gamma(drop_prob, block_size, x.shape[2], x.shape[3])
    mask_center = (rand_fn(x.shape[0], x.shape[1], x.shape[2] - block_size + 1,
                           x.shape[3] - block_size + 1, device=x.device) < gamma).float()
    block_mask = dropblock_block_mask(mask_center, block_size)
    if record is not None:
        record.append(block_mask)
    out = x * block_mask
    out = out * block_mask.numel() / block_mask.sum()          # :64, evaluated left to right
    return out


def dropblock2d_ichan(x: Tensor, drop_prob: float, block_size: int, training: bool = True,
                      bernoulli_fn: Callable = torch.bernoulli, record: Optional[list] = None) -> Tensor:
    """R/utils/utils_modules.py:106-139 (`Dropblock2d_ichan.forward`), out of place.  `bernoulli_fn` is the draw
    (`torch.bernoulli(torch.ones_like(tensor) * gamma)`, :113-114)."""
    if not training or drop_prob == 0.:
        return x
    fx, fy = x.shape[2], x.shape[3]
    keep_prob = 1 - drop_prob
    gamma = (1 - keep_prob) / (block_size ** 2) * (fx * fy) / ((fx - block_size + 1) * (fy - block_size + 1))   # :96-99
    gamma = min(gamma, 1)
    mask = bernoulli_fn(torch.ones_like(x) * gamma).clone()
    ex = block_size // 2                                                                    # :116-121
    mask[:, :, :ex] = 0
    mask[:, :, :, :ex] = 0
    mask[:, :, fx - ex:] = 0
    mask[:, :, :, fy - ex:] = 0
    shp = mask.shape
    mp = F.max_pool2d(mask.view(-1, 1, shp[2], shp[3]), kernel_size=(block_size, block_size), stride=(1, 1),
                      padding=block_size // 2).view(shp)                                      # :123-127
    mask = 1 - mp
    if record is not None:
        record.append(mask)
    out = x * mask
    total = mask.numel()
    den = 1. - torch.true_divide(total - torch.sum(mask), total)                            # :133-137
    if den != 0:
        out = out * (1. / den)
    return out


def square_pad(tensor: Tensor) -> Tensor:
    """R/utils/utils_general.py:32-43, restated with F.pad: rows (top = d//2, bottom = d - d//2); columns get
    LEFT = d - d//2 and right = d//2 (TF.pad's padding tuple is (left, top, right, bottom) and the reference passes its
    variable named `left` = total - total//2 first)."""
    size = max(tensor.shape[-2], tensor.shape[-1])
    tp = size - tensor.shape[-2]
    top, bot = tp // 2, tp - tp // 2
    tw = size - tensor.shape[-1]
    right = tw // 2
    left = tw - right
    return F.pad(tensor, (left, right, top, bot))


def square_pad_resize(tensor: Tensor, size: int) -> Tensor:
    """`TF.resize(square_pad(x), size=(s, s))` as called at Dropblock_Uncertainty.py:52-61 / MF-training-UNI.py:54-66:
    torchvision on tensors = F.interpolate(bilinear, align_corners=False, antialias=True) (torchvision 0.26)."""
    return F.interpolate(square_pad(tensor), size=(size, size), mode="bilinear", align_corners=False, antialias=True)


def linear_scheduler_values(start_value: float, stop_value: float, nr_steps: int) -> np.ndarray:
    """`dropblock==0.3.0` `LinearScheduler.__init__`: `np.linspace(start, stop, int(nr_steps))`;
    `step()` assigns `drop_values[i]` to `dropblock.drop_prob` while `i < len` and increments i.
    PARITY UNPINNED: the package is not vendored in the reference nor installed here."""
    return np.linspace(start=start_value, stop=stop_value, num=int(nr_steps))


# ----------------------------------------------------------------------------- U-Net forward
class DropBlockCfg:
    """The single shared DropBlock instance of the reference (`utils_unet.py:117-134`): one
    (drop_prob, block_size) pair used at all 22 sites, each site drawing its own mask."""

    def __init__(self, drop_prob: float = 0.0, block_size: int = 7, training: bool = True,
                 rand_fn: Callable = torch.rand, mode: str = "dropblock2d", bernoulli_fn: Callable = torch.bernoulli):
        self.drop_prob, self.block_size, self.training, self.rand_fn = drop_prob, block_size, training, rand_fn
        self.mode, self.bernoulli_fn = mode, bernoulli_fn          # "ichan": Dropblock2d_ichan (utils_modules.py:86-139)
        self.masks: Optional[list] = None      # set to [] to record block masks in call order

    def __call__(self, x: Tensor) -> Tensor:
        if self.mode == "ichan":
            return dropblock2d_ichan(x, self.drop_prob, self.block_size, self.training, self.bernoulli_fn, self.masks)
        return dropblock2d(x, self.drop_prob, self.block_size, self.training, self.rand_fn, self.masks)


def _conv_unit(x, sd, prefix, idx, groups, db, taps, name):
    """Conv3x3(same, no bias) -> GroupNorm -> DropBlock -> ReLU (R/utils/utils_unet.py:166-182)."""
    y = F.conv2d(x, sd[f"{prefix}.{idx}.weight"], None, stride=1, padding=1)
    if taps is not None:
        taps[f"{name}.conv"] = y
    y = F.group_norm(y, groups, sd[f"{prefix}.{idx + 1}.weight"], sd[f"{prefix}.{idx + 1}.bias"], eps=1e-5)
    if db is not None:
        y = db(y)
    return F.relu(y)


def autopad(x: Tensor, model_depth: int):
    """R/utils/utils_unet.py:451-458: zero-pad bottom/right to a multiple of 2**depth."""
    mult = 2 ** model_depth
    h, w = x.shape[-2:]
    pb = math.ceil(h / mult) * mult - h
    pr = math.ceil(w / mult) * mult - w
    return F.pad(x, (0, pr, 0, pb)), (h, w)


def unet_forward(sd: Dict[str, Tensor], x: Tensor, model_depth: int = 4, num_groups: int = 32,
                 dropblock: Optional[DropBlockCfg] = None, taps: Optional[dict] = None,
                 return_logits: bool = False) -> Tensor:
    """R/utils/utils_unet.py:408-449 (`UNet.forward`) for the canonical options
    (pool 'max', up 'upconv', connection 'cat', same padding, 2 convs per block, GroupNorm).
    `taps`, when a dict, receives named intermediate tensors (raw conv outputs, pooled,
    pre-sigmoid logits) for per-layer parity checks."""
    x, (h0, w0) = autopad(x, model_depth)
    skips: List[Tensor] = []
    for lvl in range(model_depth):
        p = f"down_blocks.{lvl}.0"
        x = _conv_unit(x, sd, p, 0, num_groups, dropblock, taps, f"d{lvl}.c1")
        x = _conv_unit(x, sd, p, 4, num_groups, dropblock, taps, f"d{lvl}.c2")
        skips.append(x.clone())                                                   # :420
        x = F.max_pool2d(x, 2, 2)                                                 # :265-266
        if taps is not None:
            taps[f"d{lvl}.pool"] = x
        x = F.group_norm(x, num_groups, sd[f"down_blocks.{lvl}.1.1.weight"],
                         sd[f"down_blocks.{lvl}.1.1.bias"], eps=1e-5)              # :282 (no ReLU)
    x = _conv_unit(x, sd, "conn_block", 0, num_groups, dropblock, taps, "b.c1")
    x = _conv_unit(x, sd, "conn_block", 4, num_groups, dropblock, taps, "b.c2")
    for u in range(model_depth):
        x = F.conv_transpose2d(x, sd[f"up_blocks.{u}.0.0.weight"], None, stride=2)   # :311-315
        if taps is not None:
            taps[f"u{u}.up"] = x
        x = F.group_norm(x, num_groups, sd[f"up_blocks.{u}.0.1.weight"], sd[f"up_blocks.{u}.0.1.bias"], eps=1e-5)
        x = F.relu(x)                                                             # :320-322
        x = torch.cat([x, skips[model_depth - 1 - u]], dim=1)                     # :382 [up, skip]
        if dropblock is not None:
            x = dropblock(x)                                                      # :383 (no ReLU after)
        p = f"up_blocks.{u}.1"
        x = _conv_unit(x, sd, p, 0, num_groups, dropblock, taps, f"u{u}.c1")
        x = _conv_unit(x, sd, p, 4, num_groups, dropblock, taps, f"u{u}.c2")
    logits = F.conv2d(x, sd["output_conv.0.weight"], None)                        # :397-402
    if taps is not None:
        taps["logits"] = logits
    if return_logits:
        return logits[:, :, :h0, :w0]
    y = torch.sigmoid(logits)                                                     # :404
    y = y[:, :, :h0, :w0]                                                         # :440, :460-463
    y = y.clamp(0, 1)                                                             # :443
    y = torch.where(y != y, torch.zeros_like(y), y)                               # :444
    return y


def dropblock_site_shapes(h: int, w: int, filters: int = 64, model_depth: int = 4):
    """(C, H, W) of the 22 DropBlock sites in the reference's RNG call order
    (R/utils/utils_unet.py:417-433).  h, w are the PADDED sizes."""
    sites = []
    c = filters
    for lvl in range(model_depth):
        hh, ww = h >> lvl, w >> lvl
        sites += [(c, hh, ww), (c, hh, ww)]
        c *= 2
    hh, ww = h >> model_depth, w >> model_depth
    sites += [(c, hh, ww), (c, hh, ww)]
    for u in range(model_depth):
        lvl = model_depth - 1 - u
        c //= 2
        hh, ww = h >> lvl, w >> lvl
        sites += [(2 * c, hh, ww), (c, hh, ww), (c, hh, ww)]
    return sites


# ----------------------------------------------------------------------------- MC DropBlock
def mc_dropblock(sd, im: Tensor, mask: Tensor, num_iterations: int, return_num: int,
                 drop_prob: float, block_size: int = 7, num_groups: int = 32, model_depth: int = 4,
                 rand_fn: Callable = torch.rand):
    """R/uncertainty_tests/Dropblock_Uncertainty.py:48-72 (`DropBlockEval.predict_step`, mode
    'save', resize -1): T stochastic forwards with only the DropBlock layers in training mode,
    each multiplied by the FOV mask, stacked; per-pixel mean and UNBIASED std (:66-67)."""
    db = DropBlockCfg(drop_prob, block_size, True, rand_fn)
    with torch.no_grad():
        tensors = torch.vstack([(unet_forward(sd, im, model_depth, num_groups, db) * mask).unsqueeze(0)
                                for _ in range(num_iterations)])
        mean = tensors.mean(0)
        std = tensors.std(0)
    return mean, std, tensors[0:min(return_num, num_iterations)].clone()


# ----------------------------------------------------------------------------- rotation
def rotate_bilinear(img: Tensor, angle: float) -> Tensor:
    """torchvision 0.26 `TF.rotate(img, angle, BILINEAR, fill=0)` on a float tensor, restated
    (functional.py:1066-1131 -> `_get_inverse_affine_matrix([0,0], -angle, [0,0], 1, [0,0])`,
    :1006-1063; `_functional_tensor.py` `rotate` :654-669, `_gen_affine_grid` :579-602,
    `_apply_grid_transform` :545-576).  For zero shear/translate/scale 1 the inverse matrix is
    [cos t, -sin t, 0, sin t, cos t, 0] with t = radians(angle) evaluated in Python doubles;
    a ones channel is sampled alongside and the result is img*m + (1-m)*fill."""
    n, c, h, w = img.shape
    rot = math.radians(-angle)
    a = math.cos(rot)
    b = -math.sin(rot)                   # shear 0: a=cos, b=-sin, c=sin, d=cos (functional.py:1035-1041)
    cc = math.sin(rot)
    d = math.cos(rot)
    matrix = [d, -b, 0.0, -cc, a, 0.0]   # inverse of [[a,b],[cc,d]] / scale(=1)  (:1043-1051)
    theta = torch.tensor(matrix, dtype=img.dtype, device=img.device).reshape(1, 2, 3)
    # _gen_affine_grid
    d5 = 0.5
    base = torch.empty(1, h, w, 3, dtype=img.dtype, device=img.device)
    xg = torch.linspace(-w * 0.5 + d5, w * 0.5 + d5 - 1, steps=w, device=img.device, dtype=img.dtype)
    base[..., 0].copy_(xg)
    yg = torch.linspace(-h * 0.5 + d5, h * 0.5 + d5 - 1, steps=h, device=img.device, dtype=img.dtype).unsqueeze(-1)
    base[..., 1].copy_(yg)
    base[..., 2].fill_(1)
    rescaled = theta.transpose(1, 2) / torch.tensor([0.5 * w, 0.5 * h], dtype=img.dtype, device=img.device)
    grid = base.view(1, h * w, 3).bmm(rescaled).view(1, h, w, 2)
    # _apply_grid_transform with fill
    ones = torch.ones(n, 1, h, w, dtype=img.dtype, device=img.device)
    src = torch.cat((img, ones), dim=1)
    g = grid.expand(n, h, w, 2)
    out = F.grid_sample(src, g, mode="bilinear", padding_mode="zeros", align_corners=False)
    m = out[:, -1:, :, :].expand(n, c, h, w)
    out = out[:, :-1, :, :]
    return out * m                        # fill = 0: img*m + (1-m)*0


def rotation_ensemble(sd, im: Tensor, mask: Tensor, num_iterations: int, return_num: int,
                      num_groups: int = 32, model_depth: int = 4):
    """R/uncertainty_tests/Rotational_Uncertainty.py:36-68 (`RotationEval.predict_step`):
    angles 1..num_iterations degrees, rotate in, eval forward (DropBlock is Identity because
    `set_dropblock` is never called, :122), rotate back by -angle, times FOV mask; mean/std."""
    runs = []
    with torch.no_grad():
        for it in range(1, num_iterations + 1):
            rot = rotate_bilinear(im, float(it))
            seg = unet_forward(sd, rot, model_depth, num_groups, None)
            seg = rotate_bilinear(seg, float(-it))
            runs.append((seg * mask).unsqueeze(0))
        tensors = torch.vstack(runs)
        mean = tensors.mean(0)
        std = tensors.std(0)
    return mean, std, tensors[0:min(return_num, num_iterations)].clone()


# ----------------------------------------------------------------------------- train step
def train_step_loss(sd, im: Tensor, gt: Tensor, mask: Tensor, dropblock: Optional[DropBlockCfg] = None,
                    num_groups: int = 32, model_depth: int = 4) -> Tensor:
    """R/utils/utils_training.py:21-39 (`BaseUNetTraining.training_step`) with `nn.BCELoss()`
    (R/base_model_tests/training.py:195): seg*mask, gt*mask, mean BCE over all pixels, then
    times numel / count_nonzero(mask)."""
    seg = unet_forward(sd, im, model_depth, num_groups, dropblock)
    seg = seg * mask
    gt = gt * mask
    loss = F.binary_cross_entropy(seg, gt)
    loss = loss * (seg.numel() / mask.count_nonzero())
    return loss


def dice_score(pred: Tensor, gt: Tensor, mask: Tensor, thresh: float = 0.5) -> float:
    """Dice == F1 for binary masks (R/utils/utils_metrics.py:157-173 uses sklearn f1_score on
    the thresholded prediction inside the FOV)."""
    sel = mask > 0
    p = (pred[sel] > thresh).double()
    g = (gt[sel] > 0.5).double()
    tp = (p * g).sum()
    den = p.sum() + g.sum()
    return float(2 * tp / den) if den > 0 else 1.0
