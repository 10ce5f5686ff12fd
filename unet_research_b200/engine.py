"""Host-side runtime of the U-Net forward: packed weights, HBM workspaces and the kernel schedule.

One `UNetEngine` serves one set of weights; `Workspace` objects (one per (batch, H0, W0)) own every
activation / statistics / mask buffer, so a forward launches kernels only -- no allocation, no host
synchronisation -- and can be captured in a CUDA graph.  The schedule mirrors `UNet.forward`
(reference utils_unet.py:408-449):

  conv (tcgen05 implicit GEMM, raw output + GroupNorm partials)
    -> gn_finalize (per-(image, channel) affine, DropBlock rescale folded in)
    -> gn_apply (normalise + DropBlock mask + ReLU [+ 2x2 max-pool + skip store into the concat buffer])

Data layout in HBM: activations NHWC (bf16, or fp32 in TF32 mode); the decoder's concat buffer
[N,H,W,2C] is written in halves by its two producers (up-conv apply -> channels [0,C), encoder
apply_pool -> channels [C,2C)), so `torch.cat` and `x.clone()` never run.  DropBlock keep-masks are
bit-packed NHWC (uint32 [N,H,W,C/32]).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import ApplyDesc, ConvDesc, DropblockCall, HeadDesc, UnitBwdDesc, WgradDesc, call, ptr, stream_ptr

GN_EPS = 1e-5


def _torch_dtype(dtype: int):
    return {_lib.F32: torch.float32, _lib.F16: torch.float16}.get(dtype, torch.bfloat16)


def device_rand_geometry() -> Tuple[int, int]:
    sms, mt = C.c_int(0), C.c_int(0)
    call("b2u_device_info", C.byref(sms), C.byref(mt))
    return sms.value, mt.value


# ----------------------------------------------------------------------------- DropBlock host logic
def dropblock_gamma(drop_prob: float, block_size: int, h: int, w: int) -> float:
    """gamma of reference utils_modules.py:81-82."""
    return drop_prob * h * w / ((block_size ** 2) * (h - block_size + 1) * (w - block_size + 1))


def _uniform_f32(x: np.ndarray) -> np.ndarray:
    xf = x.astype(np.float32)
    return xf * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)


def philox_thresholds(gamma: float) -> Tuple[int, int]:
    """Raw-word equivalents of `curand_uniform(word) < float32(gamma)` including torch's 1.0 -> 0.0
    wrap: centre = word < lo or word >= hi.  Found by bisection on the exact fp32 arithmetic."""
    g = np.float32(gamma)
    lo, hi = 0, 1 << 32
    while lo < hi:
        mid = (lo + hi) // 2
        if _uniform_f32(np.array([mid], dtype=np.uint32))[0] < g:
            lo = mid + 1
        else:
            hi = mid
    a, b = 0, (1 << 32) - 1
    while a < b:
        mid = (a + b) // 2
        if _uniform_f32(np.array([mid], dtype=np.uint32))[0] == np.float32(1.0):
            b = mid
        else:
            a = mid + 1
    wrap = a if g > 0 else (1 << 32) - 1
    return min(lo, (1 << 32) - 1), min(wrap, (1 << 32) - 1)


_THRESH_CACHE: Dict[Tuple[float, int, int, int], Tuple[int, int]] = {}


def _thresholds_cached(drop_prob: float, block_size: int, h: int, w: int, mode: str = "dropblock2d") -> Tuple[int, int]:
    key = (float(drop_prob), block_size, h, w, mode)
    v = _THRESH_CACHE.get(key)
    if v is None:
        if len(_THRESH_CACHE) > 4096:
            _THRESH_CACHE.clear()
        if mode == "ichan":                      # utils_modules.py:98-102: min(gamma, 1), compared with `<=`
            t = philox_threshold_le(min(dropblock_gamma(drop_prob, block_size, h, w), 1.0))
            v = (max(t, 0), 0) if t >= 0 else (0, 0)
        else:
            v = philox_thresholds(dropblock_gamma(drop_prob, block_size, h, w))
        _THRESH_CACHE[key] = v
    return v


def philox_threshold_le(p: float) -> int:
    """Largest raw Philox word whose `curand_uniform` value is <= float32(p) (torch.bernoulli's `rand <= p`,
    DistributionTemplates.h:622-641); -1 when no word qualifies.  Bisection on the exact fp32 arithmetic."""
    pf = np.float32(p)
    lo, hi = 0, 1 << 32                      # first word with uniform > p
    while lo < hi:
        mid = (lo + hi) // 2
        if _uniform_f32(np.array([mid], dtype=np.uint32))[0] <= pf:
            lo = mid + 1
        else:
            hi = mid
    return lo - 1


BERNOULLI_OFFSET_INCREMENT = 12              # gen->philox_cuda_state(10), rounded up to a multiple of 4


def rand_grid(numel: int, sms: int, max_threads: int) -> int:
    return min(sms * (max_threads // 256), (numel + 255) // 256)


def rand_offset_increment(numel: int, sms: int, max_threads: int) -> int:
    tn = 256 * rand_grid(numel, sms, max_threads)
    return ((numel - 1) // (tn * 4) + 1) * 4


def site_shapes(h: int, w: int, filters: int, depth: int) -> List[Tuple[int, int, int]]:
    """(C, H, W) of the DropBlock sites in the reference's call order (utils_unet.py:417-433)."""
    sites = []
    c = filters
    for lvl in range(depth):
        sites += [(c, h >> lvl, w >> lvl)] * 2
        c *= 2
    sites += [(c, h >> depth, w >> depth)] * 2
    for u in range(depth):
        lvl = depth - 1 - u
        c //= 2
        sites += [(2 * c, h >> lvl, w >> lvl), (c, h >> lvl, w >> lvl), (c, h >> lvl, w >> lvl)]
    return sites


class MaskPlan:
    """All DropBlock masks of one forward batch: `n_calls` independent reference forwards of
    `images_per_call` images each (Monte-Carlo: n_calls = batched iterations, 1 image each; training:
    one call covering the whole batch).  Holds the call table, bitmaps and keep counters."""

    def __init__(self, n_calls: int, images_per_call: int, h: int, w: int, filters: int, depth: int,
                 drop_prob: float, block_size: int, device, mode: str = "dropblock2d"):
        """mode "dropblock2d": DropBlock2D (torch.rand over the interior grid, utils_modules.py:49);
        mode "ichan": Dropblock2d_ichan (torch.bernoulli over the full tensor, border zeroed, :113-121)."""
        if block_size % 2 == 0 or block_size > 31:
            raise NotImplementedError("CUDA DropBlock supports odd block_size <= 31 (reference default 7)")
        if mode not in ("dropblock2d", "ichan"):
            raise ValueError(f"unknown DropBlock mode {mode!r}")
        self.mode = mode
        self.n_calls, self.ipc = n_calls, images_per_call
        self.drop_prob, self.block_size = float(drop_prob), int(block_size)
        self.sites = site_shapes(h, w, filters, depth)
        self.n_sites = len(self.sites)
        sms, mt = device_rand_geometry()
        n_img = n_calls * images_per_call
        calls = (DropblockCall * (n_calls * self.n_sites))()
        center_words = 0
        mask_words = 0
        self.mask_site_off = []
        self.numel_per_call = []
        site_prefix = []
        per_iter = 0
        for (c, hh, ww) in self.sites:
            if hh < block_size or ww < block_size:
                raise ValueError(f"feature map {hh}x{ww} smaller than block_size {block_size} "
                                 "(the reference fails here too: negative dimension)")
            site_prefix.append(per_iter)
            if mode == "ichan":
                per_iter += BERNOULLI_OFFSET_INCREMENT
            else:
                per_iter += rand_offset_increment(images_per_call * c * (hh - block_size + 1) * (ww - block_size + 1), sms, mt)
        self.offset_per_call = per_iter
        k = 0
        for s, (c, hh, ww) in enumerate(self.sites):
            self.mask_site_off.append(mask_words)
            hc, wc = hh - block_size + 1, ww - block_size + 1
            numel = images_per_call * c * hc * wc                  # bits of the compact centre bitmap
            draws = images_per_call * c * hh * ww if mode == "ichan" else numel
            if draws >= 2 ** 32:
                raise NotImplementedError("one DropBlock call is limited to 2^32 - 1 random draws")
            lo, hi = _thresholds_cached(drop_prob, block_size, hh, ww, mode)
            self.numel_per_call.append(float(images_per_call * c * hh * ww))
            for b in range(n_calls):
                d = calls[k]
                k += 1
                d.philox_offset = b * per_iter + site_prefix[s]
                d.center_word_off = center_words
                d.mask_word_off = mask_words + b * images_per_call * hh * ww * (c // 32)
                d.numel = draws
                d.grid = rand_grid(numel, sms, mt)
                d.thresh_lo, d.thresh_hi = lo, hi
                d.n_img, d.c, d.h, d.w = images_per_call, c, hh, ww
                d.block_size = block_size
                d.count_index = s * n_calls + b
                center_words += ((numel + 31) // 32 + 2 + 3) // 4 * 4     # 16-byte aligned: the v2 scatter reads uint4 chunks
            mask_words += n_img * hh * ww * (c // 32)
        self.host_table = calls
        call("b2u_dropblock_plan", calls, n_calls * self.n_sites, None)     # flat dilate grid: first block of every call
        raw = np.frombuffer(bytes(calls), dtype=np.uint8).copy()
        self.table = torch.from_numpy(raw).to(device)
        self.center_words = center_words + 4
        self.center_bits = torch.zeros(center_words + 4, dtype=torch.int32, device=device)
        self.mask_bits = torch.empty(mask_words, dtype=torch.int32, device=device)
        # v2 dilation (block size 7): sparse scatter of the centres into an NHWC word bitmap + word-parallel 7x7 OR
        self.mask_words = mask_words
        self.short_blocks = os.environ.get("B2U_CENTERS_SHORT", "0") != "0"     # measured: no gain
        self.dilate_v2 = block_size == 7 and os.environ.get("B2U_DILATE", "v2") != "v1"
        self.scatter_bits = torch.empty(mask_words, dtype=torch.int32, device=device) if self.dilate_v2 else None
        self.keep_counts = torch.zeros(self.n_sites * n_calls, dtype=torch.int64, device=device)
        self.offset_base = torch.zeros(1, dtype=torch.int64, device=device)

    def set_drop_prob(self, drop_prob: float):
        """Re-threshold the call table for a new drop_prob (LinearScheduler ramps it every training step,
        reference utils_unet.py:410-411): same buffers, same device pointers (captured CUDA graphs stay valid)."""
        drop_prob = float(drop_prob)
        if drop_prob == self.drop_prob:
            return
        self.drop_prob = drop_prob
        k = 0
        for s, (c, hh, ww) in enumerate(self.sites):
            lo, hi = _thresholds_cached(drop_prob, self.block_size, hh, ww, self.mode)
            for b in range(self.n_calls):
                self.host_table[k].thresh_lo, self.host_table[k].thresh_hi = lo, hi
                k += 1
        raw = np.frombuffer(bytes(self.host_table), dtype=np.uint8).copy()
        self.table.copy_(torch.from_numpy(raw), non_blocking=False)

    def set_stream_position(self, philox_offset: int):
        self.offset_base.fill_(int(philox_offset))

    def generate(self, seed: int):
        """Launch both mask phases for the whole table (reads the stream position from `offset_base`)."""
        self.centers(seed)
        self.dilate()

    def centers(self, seed: int):
        """Phase 1: the Philox draws of every call -> compact centre bitmap (ALU-bound; reads `offset_base`)."""
        n = self.n_calls * self.n_sites
        if self.mode == "ichan":
            call("b2u_dropblock_centers_ichan", ptr(self.table), n, self.host_table, C.c_uint64(seed & (2 ** 64 - 1)),
                 ptr(self.offset_base), ptr(self.center_bits), self.center_words, stream_ptr())
        else:
            if self.short_blocks:
                call("b2u_dropblock_centers_ex", ptr(self.table), n, self.host_table, C.c_uint64(seed & (2 ** 64 - 1)),
                     ptr(self.offset_base), ptr(self.center_bits), stream_ptr())
            else:
                call("b2u_dropblock_centers", ptr(self.table), n, C.c_uint64(seed & (2 ** 64 - 1)), ptr(self.offset_base),
                     ptr(self.center_bits), stream_ptr())

    def dilate(self):
        """Phase 2: centre bitmap -> NHWC keep masks + keep counts (memory-bound)."""
        self.keep_counts.zero_()
        n = self.n_calls * self.n_sites
        if self.dilate_v2:
            call("b2u_dropblock_dilate_v2", ptr(self.table), n, self.host_table, ptr(self.center_bits), ptr(self.scatter_bits),
                 self.mask_words, ptr(self.mask_bits), ptr(self.keep_counts), stream_ptr())
        else:
            call("b2u_dropblock_dilate", ptr(self.table), n, self.host_table, ptr(self.center_bits), ptr(self.mask_bits),
                 ptr(self.keep_counts), stream_ptr())

    def generate_partial(self, seed: int, skip: str):
        """Timing diagnostic (tests/exp notes in DESIGN.md): rebuild only part of the masks; the rest stays stale."""
        n = self.n_calls * self.n_sites
        if skip == "dilate":
            call("b2u_dropblock_centers", ptr(self.table), n, C.c_uint64(seed & (2 ** 64 - 1)), ptr(self.offset_base),
                 ptr(self.center_bits), stream_ptr())
        elif skip == "centers":
            self.keep_counts.zero_()
            call("b2u_dropblock_dilate_v2", ptr(self.table), n, self.host_table, ptr(self.center_bits), ptr(self.scatter_bits),
                 self.mask_words, ptr(self.mask_bits), ptr(self.keep_counts), stream_ptr())

    def advance(self, n_calls: Optional[int] = None):
        call("b2u_advance_counter", ptr(self.offset_base), (n_calls or self.n_calls) * self.offset_per_call, stream_ptr())

    def mask_ptr(self, site: int) -> int:
        return self.mask_bits.data_ptr() + 4 * self.mask_site_off[site]

    def keep_ptr(self, site: int) -> int:
        return self.keep_counts.data_ptr() + 8 * site * self.n_calls


# ----------------------------------------------------------------------------- workspace
@dataclass
class _Stat:
    partials: torch.Tensor
    rows: int
    sgs: int
    coef: torch.Tensor
    mr: Optional[torch.Tensor] = None          # float2[n][G] (mean, rstd), kept for the backward pass


class Workspace:
    """Every device buffer one forward of a fixed (batch, H0, W0) needs."""

    def __init__(self, eng: "UNetEngine", n: int, h0: int, w0: int):
        self.n, self.h0, self.w0 = n, h0, w0
        mult = 2 ** eng.depth
        self.h = -(-h0 // mult) * mult
        self.w = -(-w0 // mult) * mult
        dev, dt = eng.device, _torch_dtype(eng.dtype)
        self.buf: Dict[str, torch.Tensor] = {}
        self.stat: Dict[str, _Stat] = {}
        f, d = eng.filters, eng.depth

        def act(name, hh, ww, c):
            self.buf[name] = torch.empty(n, hh, ww, c, dtype=dt, device=dev)

        def stat(name, rows, sgs, c):
            self.stat[name] = _Stat(torch.empty(n, rows, c // sgs, 2, dtype=torch.float32, device=dev), rows, sgs,
                                    torch.empty(n, c, 2, dtype=torch.float32, device=dev),
                                    torch.empty(n, eng.num_groups, 2, dtype=torch.float32, device=dev))

        c = f
        for lvl in range(d):
            hh, ww = self.h >> lvl, self.w >> lvl
            cin = eng.init_channels if lvl == 0 else c // 2
            for j in (1, 2):
                act(f"d{lvl}.raw{j}", hh, ww, c)
            act(f"d{lvl}.act1", hh, ww, c)
            act(f"cat{lvl}", hh, ww, 2 * c)
            act(f"d{lvl}.praw", hh // 2, ww // 2, c)
            act(f"d{lvl}.pact", hh // 2, ww // 2, c)
            if lvl == 0:
                r, s = eng.first_stat_layout(hh, ww, c)
            else:
                r, s = eng.conv_stat_layout(n, hh, ww, cin, c, False)
            stat(f"d{lvl}.c1", r, s, c)
            r, s = eng.conv_stat_layout(n, hh, ww, c, c, False)
            stat(f"d{lvl}.c2", r, s, c)
            r, s = eng.pool_stat_layout(hh, ww, c)
            stat(f"d{lvl}.pool", r, s, c)
            c *= 2
        hh, ww = self.h >> d, self.w >> d
        for j in (1, 2):
            act(f"b.raw{j}", hh, ww, c)
            act(f"b.act{j}", hh, ww, c)
        r, s = eng.conv_stat_layout(n, hh, ww, c // 2, c, False)
        stat("b.c1", r, s, c)
        r, s = eng.conv_stat_layout(n, hh, ww, c, c, False)
        stat("b.c2", r, s, c)
        for u in range(d):
            lvl = d - 1 - u
            cin = c
            c //= 2
            hh, ww = self.h >> lvl, self.w >> lvl
            act(f"u{u}.rawT", hh, ww, c)
            for j in (1, 2):
                act(f"u{u}.raw{j}", hh, ww, c)
                act(f"u{u}.act{j}", hh, ww, c)
            r, s = eng.conv_stat_layout(n, hh // 2, ww // 2, cin, c, True)
            stat(f"u{u}.up", r, s, c)
            r, s = eng.conv_stat_layout(n, hh, ww, 2 * c, c, False)
            stat(f"u{u}.c1", r, s, c)
            r, s = eng.conv_stat_layout(n, hh, ww, c, c, False)
            stat(f"u{u}.c2", r, s, c)
        self.out = torch.empty(n, 1, h0, w0, dtype=torch.float32, device=dev)
        self.logits: Optional[torch.Tensor] = None
        self.masks: Optional[MaskPlan] = None

    def nbytes(self) -> int:
        t = sum(b.numel() * b.element_size() for b in self.buf.values())
        t += sum(s.partials.numel() * 4 + s.coef.numel() * 4 for s in self.stat.values())
        return t


# ----------------------------------------------------------------------------- engine
class UNetEngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], init_channels: int, filters: int, depth: int,
                 num_groups: int, dtype: int = _lib.BF16, device=None):
        _lib.load()
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise _lib.B2uError("UNetEngine needs a CUDA device: there is no CPU path")
        if filters % 64 != 0:
            raise NotImplementedError("tensor-core path needs filters to be a multiple of 64")
        if init_channels not in (1, 3):
            raise NotImplementedError("first-layer kernel supports init_channels 1 or 3")
        self.init_channels, self.filters, self.depth, self.num_groups, self.dtype = init_channels, filters, depth, num_groups, dtype
        self.w: Dict[str, torch.Tensor] = {}
        self._aliased = set()
        self._workspaces: Dict[Tuple[int, int, int], Workspace] = {}
        self.training_weights = False          # set by enable_training(): also keep the dgrad-packed weights
        self.batched_pack = os.environ.get("B2U_BATCHED_PACK", "1") != "0"
        self._pack_table = None                # (pointer signature, device table, blocks) of the batched weight pack
        # Fused conv prologue (16-bit inference schedules): the GroupNorm affine + DropBlock mask + ReLU of a unit is applied
        # by the CONSUMING 3x3 conv on the TMA-landed patch in shared memory (b2u_conv3x3_pro_fwd) instead of a stand-alone
        # gn_apply pass that writes and re-reads the activated tensor.  Training keeps the unfused schedule: the weight
        # gradients read the activated tensors.  `fuse_levels`: encoder/decoder resolutions (0 = full size) that fuse.
        # Default: every level but the full-resolution one.  With the TMA-store epilogue the plain 64 -> 64 conv at
        # 592x576 runs in 241 us (+ 139 us of gn_apply) against 545 us fused -- the eight transform warps and the MMA
        # compete for the same shared-memory bandwidth when Cout = 64 -- and the Monte-Carlo step is 7.20 ms with levels
        # 1..4 fused against 7.54 with all five (profiles/r02d_exp_fuse_levels.log); levels 1..4 fuse at equal time and
        # 2.6 GB less DRAM traffic per step.
        self.fused_prologue = dtype != _lib.F32 and os.environ.get("B2U_FUSED", "1") != "0"
        lv = os.environ.get("B2U_FUSE_LEVELS")
        self.fuse_levels = set(int(t) for t in lv.split(",") if t != "") if lv is not None else set(range(1, depth + 1))
        self._sd_ref = state_dict
        self.load_weights(state_dict)

    # ---- weights
    def load_weights(self, sd: Dict[str, torch.Tensor], sync: bool = True):
        """(Re)pack the weights.  Existing packed buffers are overwritten IN PLACE so that device pointers
        baked into captured CUDA graphs stay valid across optimiser steps / load_state_dict.  Tensors the
        kernels read in their PyTorch layout (GroupNorm gamma/beta, first conv, head) are ALIASED when they
        already live on the device as contiguous fp32 (no copy; the optimiser updates them in place).
        sync=False: launch-only (used inside CUDA-graph capture; the caller keeps `sd` alive)."""
        dev, dt = self.device, _torch_dtype(self.dtype)
        st = stream_ptr()

        def slot(key, shape, dtype):
            t = self.w.get(key)
            if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or key in self._aliased:
                t = torch.empty(shape, dtype=dtype, device=dev)
                self.w[key] = t
                self._aliased.discard(key)
            return t

        def plain(key, v, v32, shape):
            if v32.data_ptr() == v.data_ptr():
                self.w[key] = v32.view(shape)
                self._aliased.add(key)
            else:
                slot(key, shape, torch.float32).copy_(v32.view(shape))

        self._sd_ref = sd
        # Tensor-core weights whose fp32 source already lives on the device (the training case: the parameters themselves)
        # are packed by ONE batched launch through a device-resident table of stable pointers; anything else (CPU state
        # dicts, other dtypes: a temporary fp32 copy per call) goes through the per-tensor entry points.
        batch = []                                  # (w ptr, out0, out1, kind, cout, cin)
        keep = []
        for k, v in sd.items():
            v32 = v.detach().to(device=dev, dtype=torch.float32).contiguous()
            aliased = v32.data_ptr() == v.data_ptr()
            if v32.dim() == 4 and k == "down_blocks.0.0.0.weight":
                plain(k, v, v32, tuple(v32.shape))                         # direct first-layer kernel reads fp32
            elif v32.dim() == 4 and k.startswith("output_conv"):
                if v32.shape[0] != 1:
                    raise NotImplementedError("head kernel supports output_channels == 1")
                plain(k, v, v32, (v32.numel(),))
            elif v32.dim() == 4 and v32.shape[2] == 3:
                cout, cin = v32.shape[0], v32.shape[1]
                # forward operand [9][Cout][Cin] and (training) the data-gradient operand [9][Cin][Cout], taps rotated 180
                dg = slot(k + "#dgrad", (9, cin, cout), dt) if self.training_weights else None
                fw = slot(k, (9, cout, cin), dt)
                if aliased and self.batched_pack:
                    batch.append((v32.data_ptr(), fw.data_ptr(), dg.data_ptr() if dg is not None else 0, 0, cout, cin))
                else:
                    call("b2u_pack_conv3x3_weight_pair", ptr(v32), ptr(fw), ptr(dg), cout, cin, self.dtype, st)
            elif v32.dim() == 4 and v32.shape[2] == 2:
                cin, cout = v32.shape[0], v32.shape[1]
                fw = slot(k, (4, cout, cin), dt)
                dg = slot(k + "#dgrad", (1, cin, 4 * cout), dt) if self.training_weights else None
                if aliased and self.batched_pack:
                    batch.append((v32.data_ptr(), fw.data_ptr(), dg.data_ptr() if dg is not None else 0, 1, cout, cin))
                else:
                    call("b2u_pack_convT2x2_weight", ptr(v32), ptr(fw), cin, cout, self.dtype, st)
                    if dg is not None:
                        call("b2u_pack_convT2x2_dgrad_weight", ptr(v32), ptr(dg), cin, cout, self.dtype, st)
            else:
                plain(k, v, v32, tuple(v32.shape))
            keep.append(v32)
        if batch:
            sig = tuple(batch)
            if self._pack_table is None or self._pack_table[0] != sig:
                if torch.cuda.is_current_stream_capturing():
                    raise _lib.B2uError("weight pointers changed inside a CUDA-graph capture: the batched pack table cannot be rebuilt here")
                ents = (_lib.PackEntry * len(batch))()
                for e, (wp, o0, o1, kind, cout, cin) in zip(ents, batch):
                    e.w, e.out0, e.out1, e.kind, e.cout, e.cin, e.first_block = wp, o0, o1 or None, kind, cout, cin, 0
                blocks = C.c_int(0)
                call("b2u_pack_batched_plan", C.byref(ents), len(batch), C.byref(blocks))
                host = torch.frombuffer(bytearray(bytes(ents)), dtype=torch.uint8)
                self._pack_table = (sig, host.to(dev), blocks.value)
            _, table, blocks = self._pack_table
            call("b2u_pack_batched", ptr(table), len(batch), blocks, self.dtype, st)
        if sync:
            torch.cuda.current_stream().synchronize()  # v32 temporaries die here

    # ---- layout queries
    def conv_stat_layout(self, n, h, w, cin, cout, conv_t: bool) -> Tuple[int, int]:
        d = self._conv_desc(n, h, w, cin, cout, cin)
        rows, sgs = C.c_int(0), C.c_int(0)
        call("b2u_convT2x2_stat_layout" if conv_t else "b2u_conv3x3_stat_layout", C.byref(d), C.byref(rows), C.byref(sgs))
        return rows.value, sgs.value

    def first_stat_layout(self, h, w, cout) -> Tuple[int, int]:
        rows, sgs = C.c_int(0), C.c_int(0)
        call("b2u_conv_first_stat_layout", h, w, cout, self.num_groups, C.byref(rows), C.byref(sgs))
        return rows.value, sgs.value

    def pool_stat_layout(self, h, w, c) -> Tuple[int, int]:
        rows, sgs = C.c_int(0), C.c_int(0)
        call("b2u_pool_stat_layout", h, w, c, self.num_groups, C.byref(rows), C.byref(sgs))
        return rows.value, sgs.value

    def _conv_desc(self, n, h, w, cin, cout, x_cstride) -> ConvDesc:
        d = ConvDesc()
        d.n, d.h, d.w, d.cin, d.cout = n, h, w, cin, cout
        d.dtype, d.num_groups, d.x_cstride = self.dtype, self.num_groups, x_cstride
        return d

    MAX_WORKSPACES = 6        # multi-GB each at DRIVE size; multi-fidelity sweeps touch ~6 shapes

    def workspace(self, n: int, h0: int, w0: int) -> Workspace:
        """Least-recently-used cache.  Everything that bakes a workspace's device pointers into a CUDA graph (its
        `train_step`, an `MCRunner`) holds a reference to the Workspace object itself, so an evicted workspace stays
        alive exactly as long as those graphs do and is freed with them."""
        key = (n, h0, w0)
        ws = self._workspaces.pop(key, None)
        if ws is None:
            while len(self._workspaces) >= self.MAX_WORKSPACES:
                self._workspaces.pop(next(iter(self._workspaces)))
            ws = Workspace(self, n, h0, w0)
        self._workspaces[key] = ws                    # most recently used last
        return ws

    # ---- kernel wrappers
    def _conv(self, ws, xname, wkey, yname, sname, n, h, w, cin, cout, conv_t=False):
        d = self._conv_desc(n, h, w, cin, cout, cin)
        call("b2u_convT2x2_fwd" if conv_t else "b2u_conv3x3_fwd", ptr(ws.buf[xname]), ptr(self.w[wkey]), ptr(ws.buf[yname]),
             ptr(ws.stat[sname].partials), C.byref(d), stream_ptr())

    def _conv_pro(self, ws, xname, sname_in, mask_ptr, relu, x_shared, wkey, yname, sname, n, h, w, cin, cout):
        """conv3x3 over act = [relu]((a * raw + b) * keep), the activation applied in the conv's own prologue."""
        d = self._conv_desc(n, h, w, cin, cout, cin)
        call("b2u_conv3x3_pro_fwd", ptr(ws.buf[xname]), ptr(ws.stat[sname_in].coef), mask_ptr, ptr(self.w[wkey]), ptr(ws.buf[yname]),
             ptr(ws.stat[sname].partials), C.byref(d), int(relu), int(x_shared), stream_ptr())

    def enable_training(self):
        if not self.training_weights:
            if self.dtype != _lib.BF16:
                raise NotImplementedError("the backward pass is implemented for compute_dtype='bf16'")
            self.training_weights = True
            self.load_weights(self._sd_ref)

    def _finalize(self, ws, sname, gkey, n, c, hw, masks: Optional[MaskPlan], site: Optional[int], shared: bool = False):
        st = ws.stat[sname]
        count = float((c // self.num_groups) * hw)
        if masks is not None and site is not None:
            keep, ipc, numel = masks.keep_ptr(site), masks.ipc, masks.numel_per_call[site]
        else:
            keep, ipc, numel = None, 1, 0.0
        call("b2u_gn_finalize_ex", ptr(st.partials), st.rows, st.sgs, ptr(self.w[gkey + ".weight"]), ptr(self.w[gkey + ".bias"]),
             ptr(st.coef), n, c, self.num_groups, count, GN_EPS, keep, ipc, numel, ptr(st.mr), int(shared), stream_ptr())

    def _apply_desc(self, n, h, w, c, relu, out_cstride, out_coffset, masks, site2, m2_cstride=0, m2_coffset=0) -> ApplyDesc:
        a = ApplyDesc()
        a.n, a.h, a.w, a.c, a.dtype, a.relu = n, h, w, c, self.dtype, int(relu)
        a.out_cstride, a.out_coffset = out_cstride, out_coffset
        a.mask2_cstride, a.mask2_coffset = m2_cstride, m2_coffset
        if masks is not None and site2 is not None:
            a.images_per_call2 = masks.ipc
            a.numel_per_call2 = masks.numel_per_call[site2]
        else:
            a.images_per_call2 = 1
            a.numel_per_call2 = 0.0
        return a

    # ---- the forward schedule
    def forward(self, x: torch.Tensor, ws: Workspace, masks: Optional[MaskPlan] = None, *, head_out: bool = True,
                want_logits: bool = False, mc: Optional[dict] = None, argmax: Optional[Dict[int, torch.Tensor]] = None,
                shared_input: bool = False, hook=None):
        """x: fp32 NCHW [n, Cin, h0, w0] on the engine's device (contiguous).  Launches only.
        hook(point): optional callback invoked between launches at the schedule points "enc{lvl}" (before encoder
        level lvl), "bottleneck" and "dec{u}" (before decoder stage u) -- the Monte-Carlo runner forks its mask-build
        stream there.
        shared_input: the n images are the SAME image (Monte-Carlo iterations batched along n): the first conv and
        its GroupNorm statistics are computed once and shared; only the DropBlock mask / rescale differ per image.
        mc = {"acc": double[2,h0,w0], "fov": float[h0,w0] | None, "samples": float[R,h0,w0] | None,
              "iter_base": int64[1], "return_num": R} switches the head to Monte-Carlo accumulation."""
        n, f, d, G = ws.n, self.filters, self.depth, self.num_groups
        assert x.shape == (n, self.init_channels, ws.h0, ws.w0) and x.dtype == torch.float32 and x.is_contiguous()
        st = stream_ptr()
        B = ws.buf
        m = masks
        fuse_on = self.fused_prologue and argmax is None     # training (argmax requested) materialises the activations

        def fuse(lvl):
            return fuse_on and lvl in self.fuse_levels

        def mptr(site):
            return m.mask_ptr(site) if m is not None else None

        def kptr(site):
            return m.keep_ptr(site) if m is not None else None

        c = f
        for lvl in range(d):
            if hook is not None:
                hook(f"enc{lvl}")
            hh, ww = ws.h >> lvl, ws.w >> lvl
            p = f"down_blocks.{lvl}.0"
            s1, s2, scat = 2 * lvl, 2 * lvl + 1, 2 * d + 2 + 3 * (d - 1 - lvl)
            # conv 1
            shared = shared_input and lvl == 0
            if lvl == 0:
                call("b2u_conv_first_fwd", ptr(x), ptr(self.w[p + ".0.weight"]), ptr(B["d0.raw1"]), ptr(ws.stat["d0.c1"].partials),
                     1 if shared else n, self.init_channels, ws.h0, ws.w0, hh, ww, c, G, self.dtype, st)
            elif fuse(lvl):
                # the pooled tensor's GroupNorm (no ReLU, no DropBlock) rides in this conv's prologue
                self._conv_pro(ws, f"d{lvl - 1}.praw", f"d{lvl - 1}.pool", None, False, False, p + ".0.weight", f"d{lvl}.raw1", f"d{lvl}.c1",
                               n, hh, ww, c // 2, c)
            else:
                self._conv(ws, f"d{lvl - 1}.pact", p + ".0.weight", f"d{lvl}.raw1", f"d{lvl}.c1", n, hh, ww, c // 2, c)
            self._finalize(ws, f"d{lvl}.c1", p + ".1", n, c, hh * ww, m, s1, shared=shared)
            if fuse(lvl):
                # conv 2 applies unit 1's GroupNorm + DropBlock + ReLU in its prologue
                self._conv_pro(ws, f"d{lvl}.raw1", f"d{lvl}.c1", mptr(s1), True, shared, p + ".4.weight", f"d{lvl}.raw2", f"d{lvl}.c2",
                               n, hh, ww, c, c)
            else:
                a = self._apply_desc(n, hh, ww, c, True, c, 0, None, None)
                a.reserved[0] = 1 if shared else 0
                call("b2u_gn_apply", ptr(B[f"d{lvl}.raw1"]), ptr(ws.stat[f"d{lvl}.c1"].coef), mptr(s1), None, None,
                     ptr(B[f"d{lvl}.act1"]), C.byref(a), st)
                # conv 2 + pool + skip store
                self._conv(ws, f"d{lvl}.act1", p + ".4.weight", f"d{lvl}.raw2", f"d{lvl}.c2", n, hh, ww, c, c)
            self._finalize(ws, f"d{lvl}.c2", p + ".5", n, c, hh * ww, m, s2)
            a = self._apply_desc(n, hh, ww, c, True, 2 * c, c, m, scat, 2 * c, c)
            call("b2u_gn_apply_pool", ptr(B[f"d{lvl}.raw2"]), ptr(ws.stat[f"d{lvl}.c2"].coef), mptr(s2), mptr(scat), kptr(scat),
                 ptr(B[f"cat{lvl}"]), ptr(B[f"d{lvl}.praw"]), ptr(ws.stat[f"d{lvl}.pool"].partials),
                 ptr(argmax[lvl]) if argmax is not None else None, G, C.byref(a), st)
            # GroupNorm after the pool (no ReLU, no DropBlock)
            self._finalize(ws, f"d{lvl}.pool", f"down_blocks.{lvl}.1.1", n, c, (hh // 2) * (ww // 2), None, None)
            if not fuse(lvl + 1):
                a = self._apply_desc(n, hh // 2, ww // 2, c, False, c, 0, None, None)
                call("b2u_gn_apply", ptr(B[f"d{lvl}.praw"]), ptr(ws.stat[f"d{lvl}.pool"].coef), None, None, None,
                     ptr(B[f"d{lvl}.pact"]), C.byref(a), st)
            c *= 2
        # bottleneck
        if hook is not None:
            hook("bottleneck")
        hh, ww = ws.h >> d, ws.w >> d
        prev = f"d{d - 1}.pact"
        fb = fuse(d)
        for j, (idx, site) in enumerate(((0, 2 * d), (4, 2 * d + 1)), start=1):
            cin = c // 2 if j == 1 else c
            if fb and j == 1:
                self._conv_pro(ws, f"d{d - 1}.praw", f"d{d - 1}.pool", None, False, False, f"conn_block.{idx}.weight", "b.raw1", "b.c1",
                               n, hh, ww, cin, c)
            elif fb:
                self._conv_pro(ws, "b.raw1", "b.c1", mptr(2 * d), True, False, f"conn_block.{idx}.weight", "b.raw2", "b.c2", n, hh, ww, cin, c)
            else:
                self._conv(ws, prev, f"conn_block.{idx}.weight", f"b.raw{j}", f"b.c{j}", n, hh, ww, cin, c)
            self._finalize(ws, f"b.c{j}", f"conn_block.{idx + 1}", n, c, hh * ww, m, site)
            if not (fb and j == 1):                      # unit 2's activation feeds the up-conv (plain GEMM): materialised
                a = self._apply_desc(n, hh, ww, c, True, c, 0, None, None)
                call("b2u_gn_apply", ptr(B[f"b.raw{j}"]), ptr(ws.stat[f"b.c{j}"].coef), mptr(site), None, None,
                     ptr(B[f"b.act{j}"]), C.byref(a), st)
            prev = f"b.act{j}"
        # decoder
        for u in range(d):
            if hook is not None:
                hook(f"dec{u}")
            lvl = d - 1 - u
            cin = c
            c //= 2
            hh, ww = ws.h >> lvl, ws.w >> lvl
            scat, s1, s2 = 2 * d + 2 + 3 * u, 2 * d + 3 + 3 * u, 2 * d + 4 + 3 * u
            # up-conv -> GroupNorm -> ReLU -> (concat-site DropBlock) -> channels [0, c) of the concat buffer
            self._conv(ws, prev, f"up_blocks.{u}.0.0.weight", f"u{u}.rawT", f"u{u}.up", n, hh // 2, ww // 2, cin, c, conv_t=True)
            self._finalize(ws, f"u{u}.up", f"up_blocks.{u}.0.1", n, c, hh * ww, None, None)
            a = self._apply_desc(n, hh, ww, c, True, 2 * c, 0, m, scat, 2 * c, 0)
            call("b2u_gn_apply", ptr(B[f"u{u}.rawT"]), ptr(ws.stat[f"u{u}.up"].coef), None, mptr(scat), kptr(scat),
                 ptr(B[f"cat{lvl}"]), C.byref(a), st)
            # conv 1 over the concat buffer
            self._conv(ws, f"cat{lvl}", f"up_blocks.{u}.1.0.weight", f"u{u}.raw1", f"u{u}.c1", n, hh, ww, 2 * c, c)
            self._finalize(ws, f"u{u}.c1", f"up_blocks.{u}.1.1", n, c, hh * ww, m, s1)
            if fuse(lvl):
                self._conv_pro(ws, f"u{u}.raw1", f"u{u}.c1", mptr(s1), True, False, f"up_blocks.{u}.1.4.weight", f"u{u}.raw2", f"u{u}.c2",
                               n, hh, ww, c, c)
            else:
                a = self._apply_desc(n, hh, ww, c, True, c, 0, None, None)
                call("b2u_gn_apply", ptr(B[f"u{u}.raw1"]), ptr(ws.stat[f"u{u}.c1"].coef), mptr(s1), None, None,
                     ptr(B[f"u{u}.act1"]), C.byref(a), st)
                # conv 2
                self._conv(ws, f"u{u}.act1", f"up_blocks.{u}.1.4.weight", f"u{u}.raw2", f"u{u}.c2", n, hh, ww, c, c)
            self._finalize(ws, f"u{u}.c2", f"up_blocks.{u}.1.5", n, c, hh * ww, m, s2)
            if u < d - 1:
                a = self._apply_desc(n, hh, ww, c, True, c, 0, None, None)
                call("b2u_gn_apply", ptr(B[f"u{u}.raw2"]), ptr(ws.stat[f"u{u}.c2"].coef), mptr(s2), None, None,
                     ptr(B[f"u{u}.act2"]), C.byref(a), st)
                prev = f"u{u}.act2"
        # head (GroupNorm-apply + DropBlock + ReLU of the last unit fused in)
        hd = HeadDesc()
        hd.n, hd.h, hd.w, hd.c, hd.h0, hd.w0, hd.dtype = n, ws.h, ws.w, f, ws.h0, ws.w0, self.dtype
        last = d - 1
        logits = None
        if want_logits:
            if ws.logits is None:
                ws.logits = torch.empty(n, 1, ws.h0, ws.w0, dtype=torch.float32, device=self.device)
            logits = ws.logits
        if mc is not None:
            hd.return_num = int(mc.get("return_num", 0))
            hd.fov_per_image = 0
            call("b2u_head_fwd", ptr(B[f"u{last}.raw2"]), ptr(ws.stat[f"u{last}.c2"].coef), mptr(2 * d + 4 + 3 * last),
                 ptr(self.w["output_conv.0.weight"]), ptr(ws.out) if head_out else None, ptr(logits), ptr(mc.get("fov")),
                 ptr(mc["acc"]), ptr(mc.get("samples")), ptr(mc.get("iter_base")), C.byref(hd), st)
        else:
            call("b2u_head_fwd", ptr(B[f"u{last}.raw2"]), ptr(ws.stat[f"u{last}.c2"].coef), mptr(2 * d + 4 + 3 * last),
                 ptr(self.w["output_conv.0.weight"]), ptr(ws.out), ptr(logits), None, None, None, None, C.byref(hd), st)
        return ws.out
