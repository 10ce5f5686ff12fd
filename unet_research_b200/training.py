"""Training-step mirror of `BaseUNetTraining` (unet_code/utils/utils_training.py:8-78) and the autograd bridge
of `UNet.forward`: forward and backward both run the B200 kernel schedules (engine.py / backward.py); PyTorch
only carries the loss (`nn.BCELoss` on the returned probabilities, training.py:195) and the optimiser.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib


def _dp_active(model) -> bool:
    import torch.distributed as dist
    return (bool(getattr(model, "data_parallel", False)) and dist.is_available() and dist.is_initialized()
            and dist.get_world_size() > 1)


class TrainStep:
    """One training forward/backward of a fixed (batch, H, W) as three CUDA graphs (BASELINE configs[1]: batch 1
    is launch-bound when ~290 kernels are issued one by one from Python):

      forward  graph: repack the conv weights from the live parameters -> build the DropBlock masks -> forward schedule
      backward graph 0: head + decoder + bottleneck     graph 1: encoder levels 3, 2     graph 2: encoder levels 1, 0

    Everything step-dependent lives in device memory the graphs read: the input copy, the upstream gradient copy,
    the Philox offset and the DropBlock thresholds (the scheduler changes drop_prob every step).  With
    data-parallel training the decoder-side gradients (85 % of the bytes, flat and contiguous) are all-reduced
    over NCCL while graph 1 runs.  The first `WARMUP` calls run eagerly (kernel attributes, lazy workspaces)."""

    WARMUP = 2

    def __init__(self, model, eng, ws, tb):
        self.eng, self.ws, self.tb = eng, ws, tb
        dev = eng.device
        self.x = torch.empty(ws.n, eng.init_channels, ws.h0, ws.w0, dtype=torch.float32, device=dev)
        self.go = torch.empty(ws.n, 1, ws.h0, ws.w0, dtype=torch.float32, device=dev)
        self.named = [(k, p) for k, p in model.named_parameters()]
        self.keys = [k for k, _ in self.named]
        self.sd = {k: p.detach() for k, p in self.named}
        tb.bind_parameters([(k, tuple(p.shape)) for k, p in self.named])
        offs = tb.goffsets
        self.tail = min(o for k, o in offs.items() if not k.startswith("down_blocks"))
        if any(o >= self.tail for k, o in offs.items() if k.startswith("down_blocks")):
            self.tail = 0                                  # unexpected parameter order: one all-reduce over everything
        # second bucket boundary inside the encoder: levels >= SPLIT_LEVEL (94 % of the encoder's bytes) are reduced while
        # the shallow, full-resolution levels -- most of the encoder backward's TIME -- still run
        from .backward import SPLIT_LEVEL
        deep = [o for k, o in offs.items() if k.startswith("down_blocks.") and int(k.split(".")[1]) >= SPLIT_LEVEL]
        shallow = [o for k, o in offs.items() if k.startswith("down_blocks.") and int(k.split(".")[1]) < SPLIT_LEVEL]
        self.mid = min(deep) if deep and shallow and self.tail > 0 and max(shallow) < min(deep) else 0
        self.sig = None
        self.calls = 0
        self.fwd_graph = None
        self.bwd_graphs = None
        self.masks = None
        self.plan = None                                   # the MaskPlan this step's CUDA graphs read: owned here, never evicted
        self.plan_key = None
        self.seed = 0
        self.generation = 0                                # forwards run on this workspace (a backward must match the latest)

    def _signature(self, active, bs, seed, mode):
        return (active, bs, mode, seed if active else 0, tuple(p.data_ptr() for _, p in self.named))

    def _fwd_body(self):
        self.eng.load_weights(self.sd, sync=False)
        if self.masks is not None:
            self.masks.generate(self.seed)
        self.eng.forward(self.x, self.ws, self.masks, argmax=self.tb.argmax)

    def forward(self, model, xin: torch.Tensor) -> torch.Tensor:
        eng, ws = self.eng, self.ws
        active, p, bs = model._dropblock_state()
        idx = xin.device.index if xin.device.index is not None else torch.cuda.current_device()
        gen = torch.cuda.default_generators[idx]
        seed = gen.initial_seed()
        mode = model._dropblock_mode()
        sig = self._signature(active, bs, seed, mode)
        if sig != self.sig:
            self.sig, self.calls, self.fwd_graph, self.bwd_graphs = sig, 0, None, None
            self.sd = {k: q.detach() for k, q in self.named}
        self.masks, self.seed = None, seed
        self.generation += 1
        if active:
            if self.plan is None or self.plan_key != (bs, mode):
                # a new plan means new device pointers: the signature above changed with (bs, mode), so the graphs that
                # captured the old plan are already dropped
                from .engine import MaskPlan
                self.plan = MaskPlan(1, ws.n, ws.h, ws.w, eng.filters, eng.depth, p, bs, eng.device, mode=mode)
                self.plan_key = (bs, mode)
            self.plan.set_drop_prob(p)
            self.masks = self.plan
            self.masks.set_stream_position(gen.get_offset())
            gen.set_offset(gen.get_offset() + self.masks.offset_per_call)
        self.x.copy_(xin)
        use_graph = bool(getattr(model, "use_cuda_graph", True))
        if use_graph and self.calls >= self.WARMUP and self.fwd_graph is None:
            torch.cuda.synchronize(eng.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._fwd_body()
            self.fwd_graph = g
        if use_graph and self.fwd_graph is not None:
            self.fwd_graph.replay()
        else:
            self._fwd_body()
        self.calls += 1
        model._mark_weights_current(eng)                             # the packed weights now match the live parameters
        return ws.out.clone()

    def backward(self, model, grad_out: torch.Tensor):
        from .backward import unet_backward
        eng, ws, tb = self.eng, self.ws, self.tb
        self.go.copy_(grad_out)
        dp = _dp_active(model)
        use_graph = bool(getattr(model, "use_cuda_graph", True)) and self.fwd_graph is not None
        if not use_graph:
            grads = unet_backward(eng, ws, tb, self.masks, self.x, ws.out, self.go, data_parallel=dp)
            return tuple(grads[k] for k in self.keys)
        if self.bwd_graphs is None:
            torch.cuda.synchronize(eng.device)
            g0, g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g0):
                unet_backward(eng, ws, tb, self.masks, self.x, ws.out, self.go, phase=0)
            with torch.cuda.graph(g1, pool=g0.pool()):
                unet_backward(eng, ws, tb, self.masks, self.x, ws.out, self.go, phase=11)
            with torch.cuda.graph(g2, pool=g0.pool()):
                unet_backward(eng, ws, tb, self.masks, self.x, ws.out, self.go, phase=12)
            self.bwd_graphs = (g0, g1, g2)
        g0, g1, g2 = self.bwd_graphs
        if dp:
            # three NCCL buckets of the flat gradient buffer, each all-reduced (AVG) while the next graph segment runs:
            # decoder + bottleneck (85 % of the bytes) | deep encoder | shallow encoder (1 MB: the only exposed exchange)
            import torch.distributed as dist
            avg = dist.ReduceOp.AVG
            handles = []
            g0.replay()
            if self.tail > 0:
                handles.append(dist.all_reduce(tb.flat[self.tail:], op=avg, async_op=True))
            g1.replay()
            if self.tail > 0 and self.mid > 0:
                handles.append(dist.all_reduce(tb.flat[self.mid:self.tail], op=avg, async_op=True))
            g2.replay()
            last = tb.flat[:self.mid] if (self.tail > 0 and self.mid > 0) else (tb.flat[:self.tail] if self.tail > 0 else tb.flat)
            handles.append(dist.all_reduce(last, op=avg, async_op=True))
            for h in handles:
                h.wait()
            tb.last_allreduce_bytes = tb.flat.numel() * 4
        else:
            g0.replay()
            g1.replay()
            g2.replay()
        return tuple(tb.gviews[k] for k in self.keys)


class _UNetFunction(torch.autograd.Function):
    """out = UNet(x); backward returns one gradient per parameter, in `model.parameters()` order."""

    @staticmethod
    def forward(ctx, model, x, *params):
        from .backward import TrainBuffers
        n, _, h0, w0 = x.shape
        eng = model._get_engine(x.device, repack=False, training=True)
        if not eng.training_weights:
            eng.enable_training()
        model._original_size = (h0, w0)
        ws = eng.workspace(n, h0, w0)
        tb = getattr(ws, "train_buffers", None)
        if tb is None:
            tb = TrainBuffers(eng, ws)
            ws.train_buffers = tb
        ts = getattr(ws, "train_step", None)
        if ts is None or [id(p) for _, p in ts.named] != [id(p) for p in params]:
            ts = TrainStep(model, eng, ws, tb)
            ws.train_step = ts
        xin = x.detach().to(torch.float32).contiguous()
        out = ts.forward(model, xin)
        ctx.model, ctx.ts, ctx.generation = model, ts, ts.generation
        return out

    @staticmethod
    def backward(ctx, grad_out):
        """The parameter gradients live in ONE flat fp32 buffer owned by the workspace (TrainBuffers.flat).  They are
        delivered by aliasing: `p.grad` becomes a view of that buffer (no 124 MB accumulate-copy per step, stable
        pointers for the fused optimiser and the NCCL buckets).  Autograd's accumulate semantics are kept: a gradient
        that already exists and is NOT that view (another loss term, another workspace) is added to; one that IS the
        view (no `zero_grad(set_to_none=True)` since the last backward: gradient accumulation, or an in-place
        `zero_grad(set_to_none=False)`) is staged before the buffer is overwritten and added back afterwards.
        Only `model.data_parallel = True` averages gradients over ranks; parameter hooks / a DistributedDataParallel
        wrapper never see these gradients (they are returned as None to autograd)."""
        ts = ctx.ts
        if ctx.generation != ts.generation:
            raise RuntimeError("UNet backward: the activations of this forward were overwritten by a later training forward "
                               "of the same (batch, H, W) -- forward and backward share one workspace per shape, run them "
                               "as pairs (the reference's checkpointed blocks hold one set of activations, too)")
        flat = ts.tb.flat
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * flat.element_size()
        staged = None
        if any(p.grad is not None and lo <= p.grad.data_ptr() < hi for _, p in ts.named):
            staged = flat.clone()                          # unconsumed (or in-place zeroed) gradients of earlier backwards
        grads = ts.backward(ctx.model, grad_out)
        if staged is not None:
            # only the views that are still bound as p.grad carry over; the others were consumed (set to None)
            for (_, p), g in zip(ts.named, grads):
                if p.grad is not None and p.grad.data_ptr() == g.data_ptr():
                    off = (g.data_ptr() - lo) // flat.element_size()
                    g.add_(staged[off:off + g.numel()].view_as(g))
        for (_, p), g in zip(ts.named, grads):
            if p.grad is None:
                p.grad = g
            elif p.grad.data_ptr() != g.data_ptr():
                p.grad.add_(g)
        return (None, None) + (None,) * len(grads)


def unet_autograd_forward(model, x):
    """The workspace is shared between forward and backward: one forward/backward pair at a time per
    (batch, H, W), exactly like the reference's activation-checkpointed blocks hold one set of activations."""
    if model._resolve_dtype(True) != _lib.BF16:
        raise NotImplementedError("training runs with compute_dtype='bf16' (or 'auto')")
    params = [p for _, p in model.named_parameters()]
    return _UNetFunction.apply(model, x, *params)


class _MaskedBCE(torch.autograd.Function):
    """loss = mean(BCE(out * mask, gt * mask)) * numel / count_nonzero(mask) (utils_training.py:28-33 with nn.BCELoss) as
    one fused kernel pair; the backward is one scaled copy of the gradient the forward already computed."""

    @staticmethod
    def forward(ctx, out, gt, mask):
        from ._lib import call, ptr, stream_ptr
        n = out.numel()
        o, g, m = out.detach().contiguous(), gt.detach().contiguous(), mask.detach().contiguous()
        grad_u = torch.empty_like(o)
        loss = torch.empty((), dtype=torch.float32, device=out.device)
        scale = torch.empty((), dtype=torch.float32, device=out.device)
        import ctypes as C
        lib = _lib.load()
        partials = torch.empty(2 * lib.b2u_masked_bce_blocks(), dtype=torch.float64, device=out.device)
        with torch.cuda.device(out.device):
            call("b2u_masked_bce_fwd", ptr(o), ptr(g), ptr(m), C.c_longlong(n), ptr(grad_u), ptr(partials), ptr(loss), ptr(scale),
                 stream_ptr())
        ctx.save_for_backward(grad_u, scale)
        return loss

    @staticmethod
    def backward(ctx, upstream):
        from ._lib import call, ptr, stream_ptr
        import ctypes as C
        grad_u, scale = ctx.saved_tensors
        up = upstream.detach().to(torch.float32).contiguous()
        grad = torch.empty_like(grad_u)
        with torch.cuda.device(grad_u.device):
            call("b2u_masked_bce_bwd", ptr(grad_u), ptr(up), ptr(scale), ptr(grad), C.c_longlong(grad_u.numel()), stream_ptr())
        return grad, None, None


def _fusable_bce(loss_fcn, seg, gt, mask) -> bool:
    return (type(loss_fcn) is nn.BCELoss and loss_fcn.reduction == "mean" and loss_fcn.weight is None
            and seg.is_cuda and seg.dtype == torch.float32 and gt.dtype == torch.float32 and mask.dtype == torch.float32
            and seg.shape == gt.shape == mask.shape and not gt.requires_grad and not mask.requires_grad)


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
        if _fusable_bce(self._loss_fcn, segmentation, gt, mask):
            return _MaskedBCE.apply(segmentation, gt, mask)      # the same arithmetic as the five lines below, fused
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
