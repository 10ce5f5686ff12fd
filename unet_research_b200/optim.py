"""Fused optimiser step for the reference training recipe.

`torch.optim.SGD(model.parameters(), lr, momentum)` (unet_code/base_model_tests/training.py:32) followed by
Lightning's `gradient_clip_val=0.5` (global L2-norm clip, `torch.nn.utils.clip_grad_norm_`) costs ~10 multi-tensor
passes over the 31 M parameters; `FusedSGD.step()` does the clip and the update in two kernel launches
(csrc/optim.cu).  It is a `torch.optim.Optimizer`: `param_groups[...]['lr']` is read at every step, so
`ReduceLROnPlateau` (training.py:34-44) and checkpointing of `state_dict()` keep working.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


class FusedSGD(torch.optim.Optimizer):
    """SGD with momentum (dampening 0, no Nesterov, no weight decay -- the reference's settings) with an optional
    fused global-norm gradient clip.  CUDA fp32 parameters only; raises otherwise (there is no CPU path)."""

    def __init__(self, params, lr: float = 1e-3, momentum: float = 0.0, max_grad_norm: Optional[float] = None):
        if lr < 0 or momentum < 0:
            raise ValueError("lr and momentum must be non-negative")
        super().__init__(params, dict(lr=lr, momentum=momentum, max_grad_norm=max_grad_norm))
        self._tables = {}
        self.last_grad_norm: Optional[torch.Tensor] = None

    def _table(self, gi: int, plist: List[torch.Tensor], mlist: List[torch.Tensor]):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), m.data_ptr(), p.numel()) for p, m in zip(plist, mlist))
        ent = self._tables.get(gi)
        if ent is not None and ent[0] == key:
            return ent[1:]
        dev = plist[0].device
        chunk = int(_lib.load().b2u_sgd_chunk_elems())
        trows = np.zeros((len(plist), 4), dtype=np.int64)
        crows = []
        for i, (p, m) in enumerate(zip(plist, mlist)):
            trows[i] = (p.data_ptr(), p.grad.data_ptr(), m.data_ptr(), p.numel())
            for s in range(0, p.numel(), chunk):
                crows.append((i, min(chunk, p.numel() - s), s))
        cdt = np.dtype([("tensor", np.int32), ("count", np.int32), ("start", np.int64)])
        carr = np.array(crows, dtype=cdt)
        t_dev = torch.from_numpy(trows.view(np.uint8).reshape(-1).copy()).to(dev)
        c_dev = torch.from_numpy(carr.view(np.uint8).reshape(-1).copy()).to(dev)
        partial = torch.zeros(len(crows), dtype=torch.float64, device=dev)
        norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self._tables[gi] = (key, t_dev, c_dev, len(crows), partial, norm)
        return t_dev, c_dev, len(crows), partial, norm

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            if not plist[0].is_cuda:
                raise _lib.B2uError("FusedSGD needs contiguous fp32 CUDA parameters and gradients (no CPU path)")
            with torch.cuda.device(plist[0].device):      # libb2u launches on the current device's current stream
                self._step_group(gi, group, plist)
        return loss

    def _step_group(self, gi, group, plist):
        for p in plist:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()
                    and p.grad.dtype == torch.float32):
                raise _lib.B2uError("FusedSGD needs contiguous fp32 CUDA parameters and gradients (no CPU path)")
        # new momentum buffers are zero-initialised, so the regular update m = momentum * m + g gives torch's first-step
        # result (m = g) exactly, per tensor -- no group-wide "first step" flag that would reset existing buffers
        first = False
        mlist = []
        for p in plist:
            st = self.state[p]
            if "momentum_buffer" not in st:
                st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            mlist.append(st["momentum_buffer"])
        t_dev, c_dev, n_chunks, partial, norm = self._table(gi, plist, mlist)
        mgn = group.get("max_grad_norm")
        call("b2u_sgd_step", ptr(t_dev), ptr(c_dev), n_chunks, ptr(partial), float(group["lr"]), float(group["momentum"]),
             float(mgn) if mgn else 0.0, int(first), ptr(norm), stream_ptr())
        self.last_grad_norm = norm if mgn else None
        # the kernels wrote the parameters (and the clipped gradients) behind autograd's back: bump the version
        # counters so every consumer that caches by version (UNet's packed-weight cache) sees the update
        torch.autograd.graph.increment_version(plist)
