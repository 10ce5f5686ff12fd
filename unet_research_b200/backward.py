"""Backward schedule of the U-Net (`loss.backward()` through `UNet.forward`, reference
utils_training.py:21-39 / utils_unet.py:408-449), on the same NHWC workspaces the forward filled.

Per conv unit, in reverse order:
    unit_bwd_stats -> unit_bwd_finalize -> unit_bwd_apply      (GroupNorm/DropBlock/ReLU backward, dgamma/dbeta)
    b2u_wgrad(dY, unit input)                                   (tcgen05, K = pixels)
    b2u_conv3x3_fwd(dY, rotated/transposed weights)             (data gradient = the forward v2 kernel)
The max-pool backward, the concat-site DropBlock backward, the skip-connection fan-in and the head (1x1 conv +
sigmoid + crop) are gradient SOURCES of the unit kernels, never materialised tensors.  The transposed conv's
gradients are 1x1 GEMMs over the space-to-depth layout its unit kernel writes.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import ConvDesc, UnitBwdDesc, WgradDesc, call, ptr, stream_ptr
from .engine import MaskPlan, UNetEngine, Workspace


class TrainBuffers:
    """Gradient-side buffers of one workspace (allocated once, on the first training step)."""

    def __init__(self, eng: UNetEngine, ws: Workspace):
        n, d, f = ws.n, eng.depth, eng.filters
        dev = eng.device
        bf = torch.bfloat16
        self.argmax = {}
        self.gY, self.gA, self.gCAT, self.gP, self.gPA, self.gT = {}, {}, {}, {}, {}, {}
        c = f
        max_part = 0
        for lvl in range(d + 1):
            hh, ww = ws.h >> lvl, ws.w >> lvl
            self.gY[lvl] = torch.empty(n, hh, ww, c, dtype=bf, device=dev)
            self.gA[lvl] = torch.empty(n, hh, ww, c, dtype=bf, device=dev)
            rows = C.c_int(0)
            call("b2u_unit_bwd_rows", hh, ww, c, C.byref(rows))
            max_part = max(max_part, n * rows.value * c * 3)
            if lvl < d:
                self.argmax[lvl] = torch.empty(n, hh // 2, ww // 2, c, dtype=torch.uint8, device=dev)
                self.gCAT[lvl] = torch.empty(n, hh, ww, 2 * c, dtype=bf, device=dev)
                self.gP[lvl] = torch.empty(n, hh // 2, ww // 2, c, dtype=bf, device=dev)
                self.gPA[lvl] = torch.empty(n, hh // 2, ww // 2, c, dtype=bf, device=dev)
                self.gT[lvl] = torch.empty(n, hh // 2, ww // 2, 4, c, dtype=bf, device=dev)
                rows2 = C.c_int(0)
                call("b2u_unit_bwd_rows", hh // 2, ww // 2, c, C.byref(rows2))
                max_part = max(max_part, n * rows2.value * c * 3)
            c *= 2
        self.partials = torch.empty(max_part, dtype=torch.float32, device=dev)
        self.gcoef = torch.empty(n * eng.num_groups * 2, dtype=torch.float32, device=dev)
        self.wg_ws: Optional[torch.Tensor] = None
        self.reducer: Optional["GradReducer"] = None
        self.first_ws = torch.empty(n * _lib.load().b2u_wgrad_first_rows() * f * eng.init_channels * 9, dtype=torch.float32, device=dev)
        self.flat: Optional[torch.Tensor] = None        # all parameter gradients in ONE fp32 buffer (named_parameters order)
        self.gviews: Dict[str, torch.Tensor] = {}
        self.goffsets: Dict[str, int] = {}

    def bind_parameters(self, named_shapes):
        """Lay the gradients out in one flat buffer, `named_parameters()` order, every tensor 256-byte aligned:
        one NCCL all-reduce (or two, split at the encoder/decoder boundary) and one fused optimiser pass cover them."""
        if self.flat is not None and list(self.gviews.keys()) == [k for k, _ in named_shapes]:
            return
        off = 0
        offs = []
        for k, shape in named_shapes:
            numel = 1
            for v in shape:
                numel *= int(v)
            offs.append((k, tuple(shape), off, numel))
            off += (numel + 63) // 64 * 64
        self.flat = torch.zeros(off, dtype=torch.float32, device=self.gY[0].device)
        self.gviews = {k: self.flat[o:o + nel].view(shape) for k, shape, o, nel in offs}
        self.goffsets = {k: o for k, _, o, _ in offs}

    def grad(self, key: str, shape) -> torch.Tensor:
        v = self.gviews.get(key)
        if v is not None:
            return v.view(shape)
        return torch.empty(shape, dtype=torch.float32, device=self.gY[0].device)

    def wgrad_workspace(self, floats: int, dev) -> torch.Tensor:
        if self.wg_ws is None or self.wg_ws.numel() < floats:
            self.wg_ws = torch.empty(floats, dtype=torch.float32, device=dev)
        return self.wg_ws


def _unit_bwd(eng: UNetEngine, ws: Workspace, tb: TrainBuffers, grads: Dict[str, torch.Tensor], *, n, h, w, c, yname, sname,
              gkey, relu, masks: Optional[MaskPlan], site: Optional[int], dy: torch.Tensor, s2d=False, grad_a=None,
              mask2_site=None, grad_pool=None, head=None):
    """grad_a = (tensor, cstride, coffset); mask2_site = (site, cstride, coffset); grad_pool = (tensor, argmax);
    head = (grad_out, out, w_head_key, h0, w0)."""
    G = eng.num_groups
    st = ws.stat[sname]
    d = UnitBwdDesc()
    d.n, d.h, d.w, d.c, d.dtype, d.relu, d.num_groups, d.s2d = n, h, w, c, eng.dtype, int(relu), G, int(s2d)
    d.y, d.coef, d.mean_rstd, d.gamma = ptr(ws.buf[yname]), ptr(st.coef), ptr(st.mr), ptr(eng.w[gkey + ".weight"])
    d.images_per_call1, d.images_per_call2 = 1, 1
    if masks is not None and site is not None:
        d.mask1, d.keep_counts1 = masks.mask_ptr(site), masks.keep_ptr(site)
        d.images_per_call1, d.numel_per_call1 = masks.ipc, masks.numel_per_call[site]
    if grad_a is not None:
        t, cs, co = grad_a
        d.grad_a, d.a_cstride, d.a_coffset = ptr(t), cs, co
    if masks is not None and mask2_site is not None:
        s2, cs, co = mask2_site
        d.mask2, d.keep_counts2 = masks.mask_ptr(s2), masks.keep_ptr(s2)
        d.mask2_cstride, d.mask2_coffset = cs, co
        d.images_per_call2, d.numel_per_call2 = masks.ipc, masks.numel_per_call[s2]
    if grad_pool is not None:
        d.grad_pool, d.argmax = ptr(grad_pool[0]), ptr(grad_pool[1])
    dwh = None
    if head is not None:
        go, out, whkey, h0, w0 = head
        d.grad_out, d.out, d.w_head, d.h0, d.w0 = ptr(go), ptr(out), ptr(eng.w[whkey]), h0, w0
        dwh = tb.grad(whkey, (c,))
        grads[whkey] = dwh
    rows = C.c_int(0)
    call("b2u_unit_bwd_rows", h, w, c, C.byref(rows))
    dgamma = tb.grad(gkey + ".weight", (c,))
    dbeta = tb.grad(gkey + ".bias", (c,))
    grads[gkey + ".weight"], grads[gkey + ".bias"] = dgamma, dbeta
    sp = stream_ptr()
    call("b2u_unit_bwd_stats", C.byref(d), ptr(tb.partials), sp)
    call("b2u_unit_bwd_finalize", ptr(tb.partials), n, rows.value, c, G, ptr(eng.w[gkey + ".weight"]), float((c // G) * h * w),
         ptr(tb.gcoef), ptr(dgamma), ptr(dbeta), ptr(dwh), sp)
    call("b2u_unit_bwd_apply", C.byref(d), ptr(tb.gcoef), ptr(dy), sp)


class GradReducer:
    """Data-parallel gradient exchange (the implicit Lightning DDP of the reference): every large
    weight gradient is all-reduced over NCCL as soon as its wgrad kernel has been enqueued, so the transfer
    over NVLink overlaps the rest of the backward; the small GroupNorm / head gradients travel in one flat
    bucket at the end.  `finish()` waits for the handles and averages (DDP semantics)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.handles = []
        self.big = []
        self.bytes = 0

    def submit(self, t: torch.Tensor):
        self.big.append(t)
        self.bytes += t.numel() * 4
        self.handles.append(self.dist.all_reduce(t, group=self.group, async_op=True))

    def finish(self, grads: Dict[str, torch.Tensor]):
        big_ids = {id(t) for t in self.big}
        small = [(k, v) for k, v in grads.items() if id(v) not in big_ids]
        if small:
            flat = torch.cat([v.reshape(-1) for _, v in small])
            self.bytes += flat.numel() * 4
            self.dist.all_reduce(flat, group=self.group)
            off = 0
            for k, v in small:
                v.copy_(flat[off:off + v.numel()].view_as(v))
                off += v.numel()
        for h in self.handles:
            h.wait()
        inv = 1.0 / self.world
        torch._foreach_mul_(list(grads.values()), inv)


def _wgrad(eng, tb, grads, key, g: torch.Tensor, x: torch.Tensor, n, h, w, cg, cx, x_cstride, taps, layout, shape):
    d = WgradDesc()
    d.n, d.h, d.w, d.cg, d.cx, d.x_cstride, d.taps, d.layout, d.dtype = n, h, w, cg, cx, x_cstride, taps, layout, eng.dtype
    fl = C.c_longlong(0)
    call("b2u_wgrad_workspace_floats", C.byref(d), C.byref(fl))
    wsb = tb.wgrad_workspace(fl.value, eng.device)
    dw = tb.grad(key, shape)
    call("b2u_wgrad", ptr(g), ptr(x), ptr(wsb), ptr(dw), C.byref(d), stream_ptr())
    grads[key] = dw
    if tb.reducer is not None:
        tb.reducer.submit(dw)


def _dgrad3x3(eng, key, g: torch.Tensor, out: torch.Tensor, n, h, w, cout_fwd, cin_fwd):
    """dX = conv3x3(dY, W rotated 180 and transposed): the forward kernel with cin/cout swapped, no statistics."""
    d = ConvDesc()
    d.n, d.h, d.w, d.cin, d.cout, d.dtype, d.num_groups, d.x_cstride = n, h, w, cout_fwd, cin_fwd, eng.dtype, 0, cout_fwd
    call("b2u_conv3x3_fwd", ptr(g), ptr(eng.w[key + "#dgrad"]), ptr(out), None, C.byref(d), stream_ptr())


SPLIT_LEVEL = 2        # encoder levels >= SPLIT_LEVEL hold 94 % of the encoder's parameters and finish first


def unet_backward(eng: UNetEngine, ws: Workspace, tb: TrainBuffers, masks: Optional[MaskPlan], xin: torch.Tensor,
                  out: torch.Tensor, grad_out: torch.Tensor, data_parallel: bool = False,
                  phase: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Returns {state-dict key: fp32 gradient in the PyTorch parameter layout}.  Launches only (plus, with
    data_parallel, NCCL all-reduces overlapped with the remaining backward kernels).
    phase: None = everything; 0 = head + decoder + bottleneck; 1 = encoder; 11 = deep encoder levels (>= SPLIT_LEVEL);
    12 = shallow encoder levels (< SPLIT_LEVEL) (the CUDA-graph segments of
    training.TrainStep: the decoder-side gradients travel over NVLink while the encoder segment runs)."""
    tb.reducer = GradReducer() if data_parallel else None
    n, f, dpt = ws.n, eng.filters, eng.depth
    B = ws.buf
    grads: Dict[str, torch.Tensor] = {}
    go = grad_out.detach().to(torch.float32).contiguous()
    outc = out.detach().contiguous()
    m = masks

    # ---------------- decoder, last level first
    c = f
    for u in (range(dpt - 1, -1, -1) if phase in (None, 0) else ()):
        lvl = dpt - 1 - u
        c = f << lvl
        hh, ww = ws.h >> lvl, ws.w >> lvl
        scat, s1, s2 = 2 * dpt + 2 + 3 * u, 2 * dpt + 3 + 3 * u, 2 * dpt + 4 + 3 * u
        p = f"up_blocks.{u}.1"
        # conv 2 unit: upstream = head (last level) or the next up-conv's data gradient
        if u == dpt - 1:
            _unit_bwd(eng, ws, tb, grads, n=n, h=hh, w=ww, c=c, yname=f"u{u}.raw2", sname=f"u{u}.c2", gkey=p + ".5", relu=True,
                      masks=m, site=s2, dy=tb.gY[lvl], head=(go, outc, "output_conv.0.weight", ws.h0, ws.w0))
        else:
            _unit_bwd(eng, ws, tb, grads, n=n, h=hh, w=ww, c=c, yname=f"u{u}.raw2", sname=f"u{u}.c2", gkey=p + ".5", relu=True,
                      masks=m, site=s2, dy=tb.gY[lvl], grad_a=(tb.gA[lvl], c, 0))
        _wgrad(eng, tb, grads, p + ".4.weight", tb.gY[lvl], B[f"u{u}.act1"], n, hh, ww, c, c, c, 9, 0, (c, c, 3, 3))
        _dgrad3x3(eng, p + ".4.weight", tb.gY[lvl], tb.gA[lvl], n, hh, ww, c, c)
        # conv 1 unit (input = concat buffer, 2c channels)
        _unit_bwd(eng, ws, tb, grads, n=n, h=hh, w=ww, c=c, yname=f"u{u}.raw1", sname=f"u{u}.c1", gkey=p + ".1", relu=True,
                  masks=m, site=s1, dy=tb.gY[lvl], grad_a=(tb.gA[lvl], c, 0))
        _wgrad(eng, tb, grads, p + ".0.weight", tb.gY[lvl], B[f"cat{lvl}"], n, hh, ww, c, 2 * c, 2 * c, 9, 0, (c, 2 * c, 3, 3))
        _dgrad3x3(eng, p + ".0.weight", tb.gY[lvl], tb.gCAT[lvl], n, hh, ww, c, 2 * c)
        # up-conv unit: upstream = first half of dCAT through the concat-site DropBlock; ReLU; GroupNorm; no own mask
        _unit_bwd(eng, ws, tb, grads, n=n, h=hh, w=ww, c=c, yname=f"u{u}.rawT", sname=f"u{u}.up", gkey=f"up_blocks.{u}.0.1", relu=True,
                  masks=m, site=None, dy=tb.gT[lvl], s2d=True, grad_a=(tb.gCAT[lvl], 2 * c, 0), mask2_site=(scat, 2 * c, 0))
        prev = f"u{u - 1}.act2" if u > 0 else "b.act2"
        cin = 2 * c
        # transposed-conv weight gradient [Cin, C, 2, 2] and data gradient: 1x1 GEMMs over the space-to-depth dY
        _wgrad(eng, tb, grads, f"up_blocks.{u}.0.0.weight", tb.gT[lvl], B[prev], n, hh // 2, ww // 2, 4 * c, cin, cin, 1, 1, (cin, c, 2, 2))
        dsc = ConvDesc()
        dsc.n, dsc.h, dsc.w, dsc.cin, dsc.cout, dsc.dtype, dsc.num_groups, dsc.x_cstride = n, hh // 2, ww // 2, 4 * c, cin, eng.dtype, 0, 4 * c
        call("b2u_gemm1x1_fwd", ptr(tb.gT[lvl]), ptr(eng.w[f"up_blocks.{u}.0.0.weight#dgrad"]), ptr(tb.gA[lvl + 1]), C.byref(dsc), stream_ptr())

    # ---------------- bottleneck
    c = f << dpt
    hh, ww = ws.h >> dpt, ws.w >> dpt
    if phase in (None, 0):
        _unit_bwd(eng, ws, tb, grads, n=n, h=hh, w=ww, c=c, yname="b.raw2", sname="b.c2", gkey="conn_block.5", relu=True, masks=m,
                  site=2 * dpt + 1, dy=tb.gY[dpt], grad_a=(tb.gA[dpt], c, 0))
        _wgrad(eng, tb, grads, "conn_block.4.weight", tb.gY[dpt], B["b.act1"], n, hh, ww, c, c, c, 9, 0, (c, c, 3, 3))
        _dgrad3x3(eng, "conn_block.4.weight", tb.gY[dpt], tb.gA[dpt], n, hh, ww, c, c)
        _unit_bwd(eng, ws, tb, grads, n=n, h=hh, w=ww, c=c, yname="b.raw1", sname="b.c1", gkey="conn_block.1", relu=True, masks=m,
                  site=2 * dpt, dy=tb.gY[dpt], grad_a=(tb.gA[dpt], c, 0))
        _wgrad(eng, tb, grads, "conn_block.0.weight", tb.gY[dpt], B[f"d{dpt - 1}.pact"], n, hh, ww, c, c // 2, c // 2, 9, 0, (c, c // 2, 3, 3))
        _dgrad3x3(eng, "conn_block.0.weight", tb.gY[dpt], tb.gPA[dpt - 1], n, hh, ww, c, c // 2)

    # ---------------- encoder, deepest level first
    enc_levels = {None: range(dpt - 1, -1, -1), 1: range(dpt - 1, -1, -1), 11: range(dpt - 1, SPLIT_LEVEL - 1, -1),
                  12: range(min(SPLIT_LEVEL, dpt) - 1, -1, -1)}.get(phase, ())
    for lvl in enc_levels:
        c = f << lvl
        hh, ww = ws.h >> lvl, ws.w >> lvl
        s1, s2, scat = 2 * lvl, 2 * lvl + 1, 2 * dpt + 2 + 3 * (dpt - 1 - lvl)
        p = f"down_blocks.{lvl}.0"
        # GroupNorm after the pool (no ReLU, no DropBlock): d(pact) -> d(praw)
        _unit_bwd(eng, ws, tb, grads, n=n, h=hh // 2, w=ww // 2, c=c, yname=f"d{lvl}.praw", sname=f"d{lvl}.pool",
                  gkey=f"down_blocks.{lvl}.1.1", relu=False, masks=None, site=None, dy=tb.gP[lvl], grad_a=(tb.gPA[lvl], c, 0))
        # conv 2 unit: upstream = max-pool backward + skip half of dCAT through the concat-site DropBlock
        _unit_bwd(eng, ws, tb, grads, n=n, h=hh, w=ww, c=c, yname=f"d{lvl}.raw2", sname=f"d{lvl}.c2", gkey=p + ".5", relu=True, masks=m,
                  site=s2, dy=tb.gY[lvl], grad_a=(tb.gCAT[lvl], 2 * c, c), mask2_site=(scat, 2 * c, c),
                  grad_pool=(tb.gP[lvl], tb.argmax[lvl]))
        _wgrad(eng, tb, grads, p + ".4.weight", tb.gY[lvl], B[f"d{lvl}.act1"], n, hh, ww, c, c, c, 9, 0, (c, c, 3, 3))
        _dgrad3x3(eng, p + ".4.weight", tb.gY[lvl], tb.gA[lvl], n, hh, ww, c, c)
        # conv 1 unit
        _unit_bwd(eng, ws, tb, grads, n=n, h=hh, w=ww, c=c, yname=f"d{lvl}.raw1", sname=f"d{lvl}.c1", gkey=p + ".1", relu=True, masks=m,
                  site=s1, dy=tb.gY[lvl], grad_a=(tb.gA[lvl], c, 0))
        if lvl > 0:
            _wgrad(eng, tb, grads, p + ".0.weight", tb.gY[lvl], B[f"d{lvl - 1}.pact"], n, hh, ww, c, c // 2, c // 2, 9, 0, (c, c // 2, 3, 3))
            _dgrad3x3(eng, p + ".0.weight", tb.gY[lvl], tb.gPA[lvl - 1], n, hh, ww, c, c // 2)
        else:
            dw = tb.grad(p + ".0.weight", (c, eng.init_channels, 3, 3))
            call("b2u_wgrad_first", ptr(tb.gY[0]), ptr(xin), ptr(tb.first_ws), ptr(dw), n, eng.init_channels, ws.h0, ws.w0, hh, ww, c,
                 eng.dtype, stream_ptr())
            grads[p + ".0.weight"] = dw
    if "output_conv.0.weight" in grads:
        grads["output_conv.0.weight"] = grads["output_conv.0.weight"].view(1, f, 1, 1)
    if tb.reducer is not None:
        tb.reducer.finish(grads)
        tb.last_allreduce_bytes = tb.reducer.bytes
        tb.reducer = None
    return grads
