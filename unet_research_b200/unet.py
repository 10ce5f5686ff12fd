"""`UNet` -- drop-in mirror of the reference model class (unet_code/utils/utils_unet.py:12-463).

Same constructor, `set_dropblock` / `set_normalization` / `set_activation_function` / `create_model`
builder protocol, same module tree (hence the same 75 state-dict keys) and the
same `forward(x)` contract: any H, W (auto-padded to a multiple of 2**depth, de-padded), DropBlock
scheduler stepped once per training forward (:410-411), sigmoid -> crop -> clamp -> NaN scrub (:436-444).

The modules in the tree are parameter containers only; `forward` runs the B200 kernel schedule of
`engine.UNetEngine` through the C ABI.  Configurations outside the reference scripts' canonical one
(pool 'max', up 'upconv', connection 'cat', same padding, 2 convs per block, GroupNorm, ReLU) raise
NotImplementedError at `create_model()`; CPU tensors raise -- there is no fallback path.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from . import _lib
from .engine import MaskPlan, UNetEngine
from .modules import DropBlock2D, Dropblock2d_ichan, LinearScheduler


class UNet(nn.Module):
    def __init__(self, init_channels: int = 3, filters: int = 64, output_channels: int = 1, model_depth: int = 4,
                 pool_mode: str = 'max', up_mode: str = 'upconv', connection: str = 'cat', same_padding: bool = True,
                 conv_layers_per_block: int = 2, checkpointing=True):
        super().__init__()
        self._kernel_size, self._stride = 3, 1
        self._pool_kernel, self._pool_stride, self._upsample_factor = 2, 2, 2
        self.same_padding = same_padding
        self._init_channels, self._filters, self._output_channels = init_channels, filters, output_channels
        self._base_filters = filters
        if connection not in ['add', 'cat', 'none']:
            raise ValueError('Connection type must be of (add, cat, none)')
        self._connection = connection
        self._padding = 'same' if same_padding else 0
        self._norm = nn.Identity
        self._bias = True
        self._norm_params = {}
        self._norm_keys = []
        self._dropblock = nn.Identity()
        if pool_mode not in ['max', 'avg', 'conv']:
            raise ValueError('Pool Mode must be of (max, avg, conv).')
        self._pool_mode = pool_mode
        if up_mode not in ['upsample', 'upconv']:
            raise ValueError('Up_Mode must be of (upsample, upconv).')
        self._up_mode = up_mode
        if conv_layers_per_block <= 1:
            raise ValueError('Convolutional Layers in each block must be 2 or more.')
        self._conv_layers_per_block = conv_layers_per_block
        self._act_fcn = nn.ReLU()
        self._model_depth = model_depth
        self._checkpointing = checkpointing       # activation recompute is a memory trick; values are unchanged
        # "auto" (default): fp16 operands for inference -- same tensor-core rate as bf16, 11-bit mantissa, the mode that
        # meets north_star's 1e-2 logit bar (measured 2.1e-3 at 584x565) -- and bf16 for training (fp16 gradients would
        # need loss scaling).  Explicit: "bf16" (tcgen05 kind::f16, training + inference), "fp16" (inference only) or
        # "tf32" (fp32 storage, kind::tf32: the fp32-parity mode).
        self.compute_dtype = "auto"
        self.data_parallel = False                # True: backward all-reduces (averages) the gradients over torch.distributed
        self.use_cuda_graph = True                # training steps replay captured CUDA graphs after two eager warm-up steps
        self._engines = {}                        # dtype code -> [UNetEngine, signature]: one engine per storage format
        self._mask_plans = {}

    # ------------------------------------------------------------------ builder protocol (reference :98-160)
    def set_dropblock(self, dropblock_class, block_size: int = 7, drop_prob: float = .1, use_scheduler: bool = True,
                      start_drop_prob: float = 0., max_drop_prob: float = .2, dropblock_ls_steps: int = 500):
        if use_scheduler:
            self._dropblock = LinearScheduler(dropblock_class(block_size=block_size, drop_prob=drop_prob),
                                              start_value=start_drop_prob, stop_value=max_drop_prob,
                                              nr_steps=dropblock_ls_steps)
        else:
            self._dropblock = dropblock_class(block_size=block_size, drop_prob=drop_prob)

    def set_normalization(self, normalization_class, params: dict):
        self._norm = normalization_class
        self._bias = False
        self._norm_params = params
        for key, value in params.items():
            if value == 'fill':
                self._norm_keys.append(key)

    def update_norm_params(self):
        for key in self._norm_keys:
            self._norm_params[key] = self._filters

    def set_activation_function(self, act_fcn):
        self._act_fcn = act_fcn

    def _check_supported(self):
        bad = []
        if self._pool_mode != 'max':
            bad.append(f"pool_mode={self._pool_mode!r}")
        if self._up_mode != 'upconv':
            bad.append(f"up_mode={self._up_mode!r}")
        if self._connection != 'cat':
            bad.append(f"connection={self._connection!r}")
        if not self.same_padding:
            bad.append("same_padding=False")
        if self._conv_layers_per_block != 2:
            bad.append(f"conv_layers_per_block={self._conv_layers_per_block}")
        if self._norm is not nn.GroupNorm:
            bad.append(f"normalization {getattr(self._norm, '__name__', self._norm)} (GroupNorm only)")
        if not isinstance(self._act_fcn, nn.ReLU):
            bad.append("activation other than ReLU")
        if self._output_channels != 1:
            bad.append(f"output_channels={self._output_channels}")
        inner = self._dropblock.dropblock if isinstance(self._dropblock, LinearScheduler) else self._dropblock
        if not isinstance(inner, (nn.Identity, DropBlock2D, Dropblock2d_ichan)):
            bad.append(f"dropblock class {type(inner).__name__}")
        if bad:
            raise NotImplementedError("the B200 path implements the reference scripts' canonical configuration only; "
                                      "unsupported: " + ", ".join(bad))

    def _conv_unit(self, cin, cout):
        layers = [nn.Conv2d(cin, cout, self._kernel_size, self._stride, padding=self._padding, bias=self._bias)]
        self._filters = cout
        self.update_norm_params()
        layers += [self._norm(**self._norm_params), self._dropblock, self._act_fcn]
        return layers

    def create_model(self):
        self._check_supported()
        self._filters = self._base_filters
        self._num_groups = None
        self.down_blocks = nn.ModuleList()
        cin = self._init_channels
        for lvl in range(self._model_depth):
            cout = self._filters if lvl == 0 else self._filters * 2
            layers = self._conv_unit(cin, cout) + self._conv_unit(cout, cout)
            self.update_norm_params()
            pooling = [nn.MaxPool2d(kernel_size=self._pool_kernel, stride=self._pool_stride), self._norm(**self._norm_params)]
            self.down_blocks.append(nn.ModuleList([nn.Sequential(*layers), nn.Sequential(*pooling)]))
            cin = cout
        f = self._filters
        self.conn_block = nn.Sequential(*(self._conv_unit(f, 2 * f) + self._conv_unit(2 * f, 2 * f)))
        self.up_blocks = nn.ModuleList()
        for _ in range(self._model_depth):
            f = self._filters
            up = [nn.ConvTranspose2d(f, f // 2, kernel_size=self._pool_kernel, stride=self._pool_stride, bias=self._bias)]
            self._filters = f // 2
            self.update_norm_params()
            up += [self._norm(**self._norm_params), self._act_fcn]
            layers = self._conv_unit(f, f // 2) + self._conv_unit(f // 2, f // 2)
            self.up_blocks.append(nn.ModuleList([nn.Sequential(*up), nn.Sequential(*layers)]))
        self.output_conv = nn.Sequential(nn.Conv2d(self._filters, self._output_channels, kernel_size=1, stride=self._stride,
                                                   padding=self._padding, bias=self._bias), nn.Sigmoid())
        gn = self.down_blocks[0][0][1]
        self._num_groups = gn.num_groups
        for m in self.modules():
            if isinstance(m, nn.GroupNorm) and (m.num_groups != self._num_groups or abs(m.eps - 1e-5) > 0 or not m.affine):
                raise NotImplementedError("GroupNorm must use one num_groups, eps=1e-5, affine=True")

    # ------------------------------------------------------------------ engine management
    def _dropblock_state(self):
        """(active, drop_prob, block_size) of the single shared DropBlock instance (reference :117-134)."""
        db = self._dropblock
        if isinstance(db, LinearScheduler):
            db = db.dropblock
        if isinstance(db, (DropBlock2D, Dropblock2d_ichan)) and db.training and db.drop_prob != 0.:
            return True, float(db.drop_prob), int(db.block_size)
        return False, 0.0, 7

    def _dropblock_mode(self) -> str:
        db = self._dropblock
        if isinstance(db, LinearScheduler):
            db = db.dropblock
        return "ichan" if isinstance(db, Dropblock2d_ichan) else "dropblock2d"

    def _resolve_dtype(self, training: bool) -> int:
        cd = self.compute_dtype
        if cd == "auto":
            cd = "bf16" if training else "fp16"
        if cd not in ("bf16", "tf32", "fp16"):
            raise ValueError(f"compute_dtype must be 'auto', 'bf16', 'fp16' or 'tf32', got {self.compute_dtype!r}")
        return {"tf32": _lib.F32, "fp16": _lib.F16}.get(cd, _lib.BF16)

    def _engine_signature(self, device, training: bool = False):
        dtype = self._resolve_dtype(training)
        params = list(self.parameters())
        return (str(device), dtype, tuple(p._version for p in params), tuple(p.data_ptr() for p in params))

    # compatibility accessors: the engine most recently used
    @property
    def _engine(self) -> Optional[UNetEngine]:
        e = self._engines.get(getattr(self, "_last_dtype", None))
        return e[0] if e is not None else None

    def _mark_weights_current(self, eng: UNetEngine, training: bool = True):
        """The packed weights of `eng` match the live parameters (TrainStep repacks inside its CUDA graph)."""
        slot = self._engines.get(eng.dtype)
        if slot is not None and slot[0] is eng:
            slot[1] = self._engine_signature(eng.device, training)

    def _get_engine(self, device, repack: bool = True, training: bool = False) -> UNetEngine:
        """One engine per storage format (inference under "auto" is fp16, training bf16: alternating train / validation
        keeps both warm).  repack=False: the caller (training.TrainStep) repacks the weights itself, inside its CUDA graph."""
        key = self._engine_signature(device, training)
        dtype = key[1]
        self._last_dtype = dtype
        slot = self._engines.get(dtype)
        if slot is not None and slot[1][0] != key[0]:
            slot = None                               # the module moved to another device: a fresh engine
        if slot is None:
            sd = {k: v for k, v in self.state_dict().items()}
            slot = [UNetEngine(sd, self._init_channels, self._base_filters, self._model_depth, self._num_groups, dtype, device), key]
            self._engines[dtype] = slot
            self._mask_plans = {k: v for k, v in self._mask_plans.items() if k[0] != dtype}
        elif slot[1] != key and repack:
            slot[0].load_weights({k: v for k, v in self.state_dict().items()})
            slot[1] = key
        return slot[0]

    MAX_MASK_PLANS = 4

    def _mask_plan(self, eng, n_calls, ipc, ws, drop_prob, block_size) -> MaskPlan:
        """Mask plans of the EAGER inference forward, one per shape, least-recently-used eviction (nothing captured in a
        CUDA graph may come from here: `training.TrainStep` and `uncertainty.MCRunner` own the plans their graphs read).
        drop_prob is re-thresholded in place because the scheduler ramps it (reference :410-411)."""
        mode = self._dropblock_mode()
        key = (eng.dtype, n_calls, ipc, ws.h, ws.w, block_size, mode)
        mp = self._mask_plans.pop(key, None)
        if mp is None:
            while len(self._mask_plans) >= self.MAX_MASK_PLANS:
                self._mask_plans.pop(next(iter(self._mask_plans)))
            mp = MaskPlan(n_calls, ipc, ws.h, ws.w, self._base_filters, self._model_depth, drop_prob, block_size, eng.device,
                          mode=mode)
        self._mask_plans[key] = mp                    # most recently used last
        mp.set_drop_prob(drop_prob)
        return mp

    # ------------------------------------------------------------------ forward (reference :408-449)
    def forward(self, x):
        if type(self._dropblock) == LinearScheduler and self.training:
            self._dropblock.step()
        if not x.is_cuda:
            raise _lib.B2uError("unet_research_b200.UNet runs on CUDA (B200) only: move the input to the GPU; "
                                "there is deliberately no CPU fallback")
        # libb2u launches on the CURRENT device's current stream: make the tensor's device current for the whole call
        with torch.cuda.device(x.device):
            if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
                from .training import unet_autograd_forward
                return unet_autograd_forward(self, x)
            return self._forward_inference(x)

    def _forward_inference(self, x):
        eng = self._get_engine(x.device)
        n, _, h0, w0 = x.shape
        self._original_size = (h0, w0)
        ws = eng.workspace(n, h0, w0)
        xin = x.detach().to(torch.float32).contiguous()
        active, p, bs = self._dropblock_state()
        masks = None
        if active:
            # one reference forward over the whole batch: every site draws ONE torch.rand of shape [N,C,H-6,W-6]
            masks = self._mask_plan(eng, 1, n, ws, p, bs)
            gen = torch.cuda.default_generators[x.device.index if x.device.index is not None else torch.cuda.current_device()]
            masks.set_stream_position(gen.get_offset())
            masks.generate(gen.initial_seed())
            gen.set_offset(gen.get_offset() + masks.offset_per_call)
        if active or not self.use_cuda_graph or torch.cuda.is_current_stream_capturing():
            return eng.forward(xin, ws, masks).clone()
        # eval forward (DropBlock off): allocation-free and launch-only, so after two eager calls per (batch, H, W) the
        # 76-launch schedule is replayed as ONE CUDA graph (small multi-fidelity sizes are launch-bound otherwise).
        # Weight repacks happen above, in place, so the captured device pointers stay valid.
        st = getattr(ws, "infer", None)
        if st is None:
            st = ws.infer = {"x": torch.empty_like(xin), "calls": 0, "graph": None}
        st["x"].copy_(xin)
        if st["graph"] is None:
            eng.forward(st["x"], ws, None)
            st["calls"] += 1
            if st["calls"] >= 2:
                torch.cuda.synchronize(x.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    eng.forward(st["x"], ws, None)
                st["graph"] = g
        else:
            st["graph"].replay()
        return ws.out.clone()
