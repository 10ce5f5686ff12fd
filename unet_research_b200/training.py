"""Training-step mirror of `BaseUNetTraining` (unet_code/utils/utils_training.py:8-78).

The forward runs the B200 kernel schedule; the backward pass (dgrad / wgrad tcgen05 GEMMs, GroupNorm,
DropBlock, max-pool and up-conv gradients) is the next row of the scope table and is not built yet, so
a forward under autograd raises instead of silently falling back to PyTorch ops.
"""
from __future__ import annotations

import torch
from torch import nn


def unet_autograd_forward(model, x):
    raise NotImplementedError(
        "unet_research_b200: the backward pass of the U-Net is not implemented yet; run inference under "
        "torch.no_grad() (there is deliberately no PyTorch fallback)")


class BaseUNetTraining(nn.Module):
    """LightningModule-shaped shell: `self._model` (so checkpoints keep the `_model.` key prefix),
    `forward`, `training_step`, `validation_step`, `test_step`, `predict_step` with the reference's
    masked-BCE rescale (utils_training.py:21-57)."""

    def __init__(self, model, loss_fcn, optimizer):
        super().__init__()
        self._model = model
        self._loss_fcn = loss_fcn
        self._optimizer = optimizer

    def forward(self, x):
        return self._model(x)

    def log(self, *a, **k):
        pass

    def _masked_loss(self, batch):
        im_batch, gt, mask = batch
        segmentation = self._model(im_batch)
        segmentation = segmentation * mask
        gt = gt * mask
        loss = self._loss_fcn(segmentation, gt)
        loss = loss * (segmentation.numel() / mask.count_nonzero())
        return loss

    def training_step(self, batch, batch_idx):
        batch[0].requires_grad = True
        return self._masked_loss(batch)

    def validation_step(self, batch, batch_idx):
        return self._masked_loss(batch)

    def configure_optimizers(self):
        return self._optimizer

    def test_step(self, batch, batch_idx):
        im_batch, _, mask = batch
        return self._model(im_batch) * mask

    def predict_step(self, batch, batch_idx):
        im_batch, gt, mask = batch
        segmentation = self._model(im_batch) * mask
        return batch_idx, segmentation, im_batch, gt, mask
