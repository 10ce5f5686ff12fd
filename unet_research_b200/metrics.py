"""On-device evaluation metrics: `get_accuracy_metrics(segmentation, gt, mask)` of the
reference (unet_code/utils/utils_metrics.py:157-173) without shipping the tensors to numpy / scikit-learn.

The reference masks with `np.ma.array(..., mask=mask.long())` and keeps the elements where that mask is TRUE, i.e.
the pixels inside the field of view, rounds the probabilities (numpy: half to even) and calls
`sklearn.metrics.f1_score / roc_auc_score / accuracy_score`.  Here:

  * F1 (= Dice of the vessel class) and accuracy come from ONE pass of `b2u_confusion_counts` (TP, FP, FN, TN);
  * AUROC is the tie-aware Mann-Whitney statistic on the device (sort + average ranks, fp64 sums), which is what
    `roc_auc_score` computes for a binary target.

Three scalars cross PCIe instead of three 330k-pixel tensors per image.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


def confusion_counts(segmentation: torch.Tensor, gt: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """int64[4] (TP, FP, FN, TN) on the device."""
    if not segmentation.is_cuda:
        raise _lib.B2uError("metrics run on CUDA tensors only (no CPU path)")
    seg = segmentation.detach().to(torch.float32).contiguous()
    g = gt.detach().to(device=seg.device, dtype=torch.float32).expand_as(seg).contiguous()
    m = mask.detach().to(device=seg.device, dtype=torch.float32).expand_as(seg).contiguous()
    counts = torch.empty(4, dtype=torch.int64, device=seg.device)
    call("b2u_confusion_counts", ptr(seg), ptr(g), ptr(m), seg.numel(), ptr(counts), stream_ptr())
    return counts


def auroc(segmentation: torch.Tensor, gt: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Area under the ROC curve of the in-FOV pixels (0-dim fp64 tensor on the device); NaN when one class is absent
    (scikit-learn raises there)."""
    seg = segmentation.detach().to(torch.float32).reshape(-1)
    keep = mask.detach().to(seg.device).expand_as(segmentation).reshape(-1).long() != 0
    s = seg[keep]
    y = (gt.detach().to(seg.device).expand_as(segmentation).reshape(-1).long() == 1)[keep]
    s, order = torch.sort(s)
    y = y[order]
    new = torch.ones_like(s, dtype=torch.bool)
    new[1:] = s[1:] != s[:-1]
    gid = torch.cumsum(new.long(), 0) - 1
    cnt = torch.bincount(gid)
    end = torch.cumsum(cnt, 0)
    avg_rank = (end - cnt).double() + (cnt.double() + 1.0) / 2.0          # mean of the 1-based ranks of a tie group
    rank = avg_rank[gid]
    n_pos = y.sum().double()
    n_neg = y.numel() - n_pos
    u = rank[y].sum() - n_pos * (n_pos + 1.0) / 2.0
    return u / (n_pos * n_neg)


def get_accuracy_metrics(segmentation: torch.Tensor, gt: torch.Tensor, mask: torch.Tensor) -> Tuple[float, float, float]:
    """(f1 score vessel, auroc, accuracy) -- same order and meaning as the reference function."""
    c = confusion_counts(segmentation, gt, mask).double()
    tp, fp, fn, tn = c[0], c[1], c[2], c[3]
    f1 = 2 * tp / torch.clamp(2 * tp + fp + fn, min=1.0)
    acc = (tp + tn) / torch.clamp(c.sum(), min=1.0)
    out = torch.stack([f1, auroc(segmentation, gt, mask), acc]).cpu()      # one 24-byte device->host copy
    return float(out[0]), float(out[1]), float(out[2])
