"""Monte-Carlo inference loops: drop-in mirrors of `DropBlockEval.predict_step`
(unet_code/uncertainty_tests/Dropblock_Uncertainty.py:27-72) and `RotationEval.predict_step`
(unet_code/uncertainty_tests/Rotational_Uncertainty.py:21-68), re-designed for B200:

* `iter_batch` Monte-Carlo iterations run as ONE batched forward (same image, independent DropBlock
  masks), so the deep, small-M layers fill the 148 SMs and launch overhead is amortised; the whole
  step (mask build -> forward -> head accumulate -> Philox advance) is captured in a CUDA graph.
* the [T,1,1,H,W] stack of the reference (1.3 GB at T=1000) is never materialised: the head kernel
  accumulates per-pixel sum / sum-of-squares in fp64 and stores only the first `return_num` samples.
* with torch.distributed initialised, iterations (or angles) are sharded over ranks -- iteration t
  always consumes the Philox window of global index t, so results do not depend on the world size --
  and closed by ONE all-reduce of the fp64 [2,H,W] accumulator (5.28 MB at 584x565).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib
from ._lib import call, ptr, stream_ptr
from .engine import MaskPlan, UNetEngine
from .modules import DropBlock2D


def set_dropblock_on(layer):
    """Dropblock_Uncertainty.py:22-25."""
    if type(layer) == DropBlock2D:
        layer.training = True


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition [start, stop) of `total` work items; rank 0 owns the first items so the
    first `return_num` samples need no gather when return_num <= its share."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def _gen_for(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    return torch.cuda.default_generators[idx]


class _MCBase(nn.Module):
    """LightningModule-shaped shell (BaseUNetTraining, utils_training.py:8-19): `self._model`, `forward`."""

    def __init__(self, model):
        super().__init__()
        self._model = model

    def forward(self, x):
        return self._model(x)


class DropBlockEval(_MCBase):
    def __init__(self, model, num_iterations=1000, return_num=25, mode='save', resize=-1, iter_batch: int = 5,
                 use_cuda_graph: bool = True):
        super().__init__(model)
        self.num_iterations = num_iterations
        self.return_num = min(return_num, num_iterations)
        self.set_mode(mode)
        self.resize = resize
        self.iter_batch = iter_batch
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}

    def set_mode(self, mode):
        self.mode = mode
        assert self.mode in ['save', 'evaluate']

    # -------------------------------------------------------------------------------------------
    def mc_statistics(self, im: torch.Tensor, mask: torch.Tensor, num_iterations: Optional[int] = None):
        """Per-pixel mean / unbiased std of `model(im) * mask` over the Monte-Carlo iterations plus the
        first `return_num` samples.  im: [1,C,H,W], mask: [1,1,H,W] (CUDA)."""
        model = self._model
        if not im.is_cuda:
            raise _lib.B2uError("MC-DropBlock runs on CUDA only")
        if im.shape[0] != 1:
            raise NotImplementedError("the reference loop is batch 1 (Dropblock_Uncertainty.py:84-85)")
        T = int(num_iterations if num_iterations is not None else self.num_iterations)
        active, p, bs = model._dropblock_state()
        dist, rank, world = _dist()
        t0, t1 = shard_range(T, rank, world)
        dev = im.device
        eng: UNetEngine = model._get_engine(dev)
        _, _, h0, w0 = im.shape
        npix = h0 * w0
        acc = torch.zeros(2, h0, w0, dtype=torch.float64, device=dev)
        R = self.return_num
        samples = torch.zeros(max(R, 1), h0, w0, dtype=torch.float32, device=dev)
        fov = mask.reshape(h0, w0).to(torch.float32).contiguous()
        iter_base = torch.full((1,), t0, dtype=torch.int64, device=dev)
        gen = _gen_for(dev)
        seed, stream_start = gen.initial_seed(), gen.get_offset()
        x1 = im.detach().to(torch.float32).contiguous()

        done = t0
        per_iter_offset = 0
        while done < t1:
            nb = min(self.iter_batch, t1 - done)
            ws = eng.workspace(nb, h0, w0)
            masks = model._mask_plan(eng, nb, 1, ws, p, bs) if active else None
            if masks is not None:
                per_iter_offset = masks.offset_per_call
                masks.set_stream_position(stream_start + done * per_iter_offset)
            xb = x1.expand(nb, -1, -1, -1).contiguous()
            mc = {"acc": acc, "fov": fov, "samples": samples if R > 0 else None, "iter_base": iter_base, "return_num": R}
            steps = (t1 - done) // nb

            def step():
                if masks is not None:
                    masks.generate(seed)
                eng.forward(xb, ws, masks, head_out=False, mc=mc)
                if masks is not None:
                    masks.advance(nb)
                call("b2u_advance_counter", ptr(iter_base), nb, stream_ptr())

            if self.use_cuda_graph and steps >= 3:
                step()                                   # warm-up (also sets kernel attributes) outside capture
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    step()
                # the capture itself does not execute; one eager step is done, replay the rest
                for _ in range(steps - 1):
                    g.replay()
                self._last_graph = g
            else:
                for _ in range(steps):
                    step()
            done += steps * nb
        if active:
            # leave the torch generator where T reference forwards would have left it
            gen.set_offset(stream_start + T * per_iter_offset)
        if dist is not None and world > 1:
            dist.all_reduce(acc)                       # the ONE exchange step of the path (fp64 [2,H,W])
            if R > 0:
                dist.all_reduce(samples)               # disjoint writers (zeros elsewhere): a sum is a gather
        mean = torch.empty(1, 1, h0, w0, dtype=torch.float32, device=dev)
        std = torch.empty(1, 1, h0, w0, dtype=torch.float32, device=dev)
        if T > 1:
            call("b2u_mc_finalize", ptr(acc), ptr(mean), ptr(std), npix, T, stream_ptr())
        else:
            mean.copy_(acc[0].view(1, 1, h0, w0))
            std.fill_(float("nan"))                    # torch.std of one sample
        tensors = samples[:R].view(R, 1, 1, h0, w0).clone()
        return mean, std, tensors

    def predict_step(self, batch, batch_idx):
        im, gt, mask = batch
        self._model.apply(set_dropblock_on)            # Dropblock_Uncertainty.py:50
        if self.resize != -1:
            raise NotImplementedError("on-the-fly resize (Dropblock_Uncertainty.py:52-61) is a 'next' row (SURVEY 8f)")
        mean, std, tensors = self.mc_statistics(im, mask)
        if self.mode == 'save':
            return batch_idx, (mean, std, tensors)
        return batch_idx, mean, im, gt, mask


class RotationEval(_MCBase):
    """Angles 1..num_iterations degrees: rotate in -> eval forward -> rotate back -> * mask -> mean/std."""

    def __init__(self, model, num_iterations=1000, return_num=25, resize=-1, angle_batch: int = 5):
        super().__init__(model)
        self.num_iterations = num_iterations
        self.return_num = min(return_num, num_iterations)
        self.resize = resize
        self.angle_batch = angle_batch

    def rotation_statistics(self, im: torch.Tensor, mask: torch.Tensor, num_iterations: Optional[int] = None):
        import ctypes as C
        model = self._model
        if not im.is_cuda:
            raise _lib.B2uError("the rotation ensemble runs on CUDA only")
        if im.shape[0] != 1:
            raise NotImplementedError("the reference loop is batch 1")
        T = int(num_iterations if num_iterations is not None else self.num_iterations)
        dist, rank, world = _dist()
        t0, t1 = shard_range(T, rank, world)
        dev = im.device
        eng: UNetEngine = model._get_engine(dev)
        _, cin, h0, w0 = im.shape
        npix = h0 * w0
        acc = torch.zeros(2, h0, w0, dtype=torch.float64, device=dev)
        R = self.return_num
        samples = torch.zeros(max(R, 1), h0, w0, dtype=torch.float32, device=dev)
        fov = mask.reshape(h0, w0).to(torch.float32).contiguous()
        iter_base = torch.full((1,), t0, dtype=torch.int64, device=dev)
        x1 = im.detach().to(torch.float32).contiguous()
        done = t0
        while done < t1:
            nb = min(self.angle_batch, t1 - done)
            ws = eng.workspace(nb, h0, w0)
            rot_in = torch.empty(nb, cin, h0, w0, dtype=torch.float32, device=dev)
            rot_out = torch.empty(nb, 1, h0, w0, dtype=torch.float32, device=dev)
            steps = (t1 - done) // nb
            for s in range(steps):
                first = done + s * nb + 1                                   # angles are 1-based (:51)
                ang_in = (C.c_double * nb)(*[float(first + k) for k in range(nb)])
                ang_out = (C.c_double * nb)(*[-float(first + k) for k in range(nb)])
                call("b2u_rotate_bilinear", ptr(x1), ptr(rot_in), nb, cin, h0, w0, ang_in, 1, stream_ptr())
                seg = eng.forward(rot_in, ws, None)                          # eval forward, DropBlock is Identity (:122)
                call("b2u_rotate_bilinear", ptr(seg), ptr(rot_out), nb, 1, h0, w0, ang_out, 0, stream_ptr())
                call("b2u_mc_accumulate", ptr(rot_out), ptr(fov), ptr(acc), ptr(samples) if R > 0 else None, ptr(iter_base),
                     nb, npix, R, stream_ptr())
                call("b2u_advance_counter", ptr(iter_base), nb, stream_ptr())
            done += steps * nb
        if dist is not None and world > 1:
            dist.all_reduce(acc)
            if R > 0:
                dist.all_reduce(samples)
        mean = torch.empty(1, 1, h0, w0, dtype=torch.float32, device=dev)
        std = torch.empty(1, 1, h0, w0, dtype=torch.float32, device=dev)
        call("b2u_mc_finalize", ptr(acc), ptr(mean), ptr(std), npix, T, stream_ptr())
        tensors = samples[:R].view(R, 1, 1, h0, w0).clone()
        return mean, std, tensors

    def predict_step(self, batch, batch_idx):
        im, gt, mask = batch
        if self.resize != -1:
            raise NotImplementedError("on-the-fly resize is a 'next' row (SURVEY 8f)")
        mean, std, tensors = self.rotation_statistics(im, mask)
        return batch_idx, (mean, std, tensors)
