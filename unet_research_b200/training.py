"""Training-step mirror of `BaseUNetTraining` (unet_code/utils/utils_training.py:8-78) and the autograd bridge
of `UNet.forward`: forward and backward both run the B200 kernel schedules (engine.py / backward.py); PyTorch
only carries the loss (`nn.BCELoss` on the returned probabilities, training.py:195) and the optimiser.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib


class _UNetFunction(torch.autograd.Function):
    """out = UNet(x); backward returns one gradient per parameter, in `model.parameters()` order."""

    @staticmethod
    def forward(ctx, model, x, *params):
        from .backward import TrainBuffers
        eng = model._get_engine(x.device)
        eng.enable_training()
        n, _, h0, w0 = x.shape
        model._original_size = (h0, w0)
        ws = eng.workspace(n, h0, w0)
        tb = getattr(ws, "train_buffers", None)
        if tb is None:
            tb = TrainBuffers(eng, ws)
            ws.train_buffers = tb
        xin = x.detach().to(torch.float32).contiguous()
        active, p, bs = model._dropblock_state()
        masks = None
        if active:
            masks = model._mask_plan(eng, 1, n, ws, p, bs)
            idx = x.device.index if x.device.index is not None else torch.cuda.current_device()
            gen = torch.cuda.default_generators[idx]
            masks.set_stream_position(gen.get_offset())
            masks.generate(gen.initial_seed())
            gen.set_offset(gen.get_offset() + masks.offset_per_call)
        out = eng.forward(xin, ws, masks, argmax=tb.argmax).clone()
        ctx.model, ctx.eng, ctx.ws, ctx.tb, ctx.masks, ctx.xin = model, eng, ws, tb, masks, xin
        ctx.keys = [k for k, _ in model.named_parameters()]
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        from .backward import unet_backward
        (out,) = ctx.saved_tensors
        import torch.distributed as dist
        dp = bool(getattr(ctx.model, "data_parallel", False)) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        grads = unet_backward(ctx.eng, ctx.ws, ctx.tb, ctx.masks, ctx.xin, out, grad_out, data_parallel=dp)
        missing = [k for k in ctx.keys if k not in grads]
        if missing:
            raise _lib.B2uError(f"backward produced no gradient for {missing[:3]}...")
        return (None, None) + tuple(grads[k] for k in ctx.keys)


def unet_autograd_forward(model, x):
    """The workspace is shared between forward and backward: one forward/backward pair at a time per
    (batch, H, W), exactly like the reference's activation-checkpointed blocks hold one set of activations."""
    if model.compute_dtype != "bf16":
        raise NotImplementedError("training runs with compute_dtype='bf16'")
    params = [p for _, p in model.named_parameters()]
    return _UNetFunction.apply(model, x, *params)


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
