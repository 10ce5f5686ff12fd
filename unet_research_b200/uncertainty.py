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

import os

from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib
from ._lib import call, ptr, stream_ptr
from .engine import MaskPlan, UNetEngine
from .modules import DropBlock2D, Dropblock2d_ichan


def set_dropblock_on(layer):
    """Dropblock_Uncertainty.py:22-25 (exact type check, as the reference); the `dropblock_i` runs submitted by
    uncertainty_script.py:26 do the same for `Dropblock2d_ichan`."""
    if type(layer) == DropBlock2D or type(layer) == Dropblock2d_ichan:
        layer.training = True


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition [start, stop) of `total` work items; rank 0 owns the first items so the
    first `return_num` samples need no gather when return_num <= its share."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def plan_segments(t0: int, t1: int, iter_batch: int):
    """A rank's iterations [t0, t1) as runs of equal batches: [(batch, steps, first iteration), ...] -- full batches of
    `iter_batch`, then ONE remainder batch (125 iterations at batch 10 = (10, 12, t0), (5, 1, t0 + 120)).  Every run has its
    own runner (workspace + CUDA graphs are shaped by the batch)."""
    out = []
    done = t0
    while done < t1:
        nb = min(iter_batch, t1 - done)
        steps = (t1 - done) // nb
        out.append((nb, steps, done))
        done += steps * nb
    return out


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


class MCRunner:
    """Persistent state of the Monte-Carlo DropBlock loop for one (model, iteration batch, image size, p, bs):
    workspace, two ping-pong DropBlock mask plans, static input / accumulator buffers and ONE captured CUDA
    graph that covers two steps.  Inside the graph the integer-ALU-bound mask build of the NEXT step runs on
    a side stream concurrently with the tensor-core-bound forward of the CURRENT step:

        main :  forward(masks A) -------------------->  forward(masks B) -------------------->
        side :  build masks B (next step) ---> join     build masks A (step after) ---> join

    Nothing is allocated and no graph is captured after the first call."""

    def __init__(self, model, nb: int, h0: int, w0: int, device, active: bool, drop_prob: float, block_size: int,
                 return_num: int, use_cuda_graph: bool = True, overlap: bool = True):
        self.model, self.nb, self.h0, self.w0, self.dev = model, nb, h0, w0, device
        self.eng: UNetEngine = model._get_engine(device)
        self.ws = self.eng.workspace(nb, h0, w0)
        self.active = active
        self.use_graph = use_cuda_graph
        self.overlap = overlap and active
        cin = self.eng.init_channels
        if active:
            mode = model._dropblock_mode()
            mk = lambda: MaskPlan(nb, 1, self.ws.h, self.ws.w, self.eng.filters, self.eng.depth, drop_prob, block_size, device,
                                  mode=mode)
            self.masks = [mk(), mk()] if self.overlap else [mk()]
            self.per_iter = self.masks[0].offset_per_call
        else:
            self.masks, self.per_iter = [], 0
        self.R = return_num
        self.x = torch.zeros(nb, cin, h0, w0, dtype=torch.float32, device=device)
        self.fov = torch.zeros(h0, w0, dtype=torch.float32, device=device)
        self.acc = torch.zeros(2, h0, w0, dtype=torch.float64, device=device)
        self.samples = torch.zeros(max(return_num, 1), h0, w0, dtype=torch.float32, device=device)
        self.iter_base = torch.zeros(1, dtype=torch.int64, device=device)
        self.mc = {"acc": self.acc, "fov": self.fov, "samples": self.samples if return_num > 0 else None,
                   "iter_base": self.iter_base, "return_num": return_num}
        # The forward runs on a HIGH-priority stream and the mask build on a low-priority one: with equal
        # priorities the block scheduler drains one grid before it starts the next, so nothing overlaps;
        # with priorities the mask-build CTAs (no shared memory, 26 registers) fill the SM resources the
        # shared-memory-bound conv CTAs leave idle.  Graph kernel nodes inherit the capture stream's priority.
        self.main = torch.cuda.Stream(device=device, priority=-1)
        self.side = torch.cuda.Stream(device=device, priority=0) if self.overlap else None
        self.graph = None                                  # two steps (one pair)
        self.graph_single = None                           # one step (odd tail / one-step remainder batches)
        self.launches_per_step = 0
        self.seed = 0
        # Where in the forward the mask build of the next step is forked (engine.forward hook points).  The mask
        # kernels take issue slots from whatever runs next to them; the deep levels' conv kernels (long K loops, short
        # epilogues) have the most to spare, the 592x576 levels the least (tests/exp_overlap.py).
        # Measured again with the TMA-store epilogues and the two-pass full-resolution level (profiles/r02d_exp_fork_point.log):
        # enc0 7.47-7.51 ms, enc1 7.37-7.45, enc2 7.38-7.44, enc3 7.45-7.52, bottleneck 7.63-7.67 per step.
        self.fork_point = os.environ.get("B2U_MC_FORK", "enc1")

    # ---- building blocks (all launch-only)
    def _forward(self, k: int, hook=None):
        m = self.masks[k] if self.active else None
        self.eng.forward(self.x, self.ws, m, head_out=False, mc=self.mc, shared_input=True, hook=hook)
        call("b2u_advance_counter", ptr(self.iter_base), self.nb, stream_ptr())

    def _generate(self, k: int, stride_blocks: int):
        m = self.masks[k]
        skip = os.environ.get("B2U_EXP_SKIP_MASKS", "")          # TIMING DIAGNOSTIC ONLY (stale masks): "all" / "centers" / "dilate"
        if skip and self.graph is None and self.launches_per_step:
            m.generate_partial(self.seed, skip)
        else:
            m.generate(self.seed)
        m.advance(stride_blocks * self.nb)

    def _pair(self):
        """Two steps.  Precondition: masks[0] hold the masks of the first step."""
        if not self.active:
            self._forward(0)
            self._forward(0)
            return
        if not self.overlap:
            self._forward(0)
            self._generate(0, 1)
            self._forward(0)
            self._generate(0, 1)
            return
        main = torch.cuda.current_stream(self.dev)
        for cur, nxt in ((0, 1), (1, 0)):
            def fork(point, nxt=nxt):
                if point != self.fork_point:
                    return
                self.side.wait_stream(main)             # fork: side sees everything enqueued so far
                with torch.cuda.stream(self.side):
                    self._generate(nxt, 2)              # masks for the following step, concurrently ...
            self._forward(cur, fork)                    # ... with the rest of this step's forward
            main.wait_stream(self.side)                 # join

    def begin(self, im: torch.Tensor, fov: torch.Tensor, t0: int, seed: int, stream_start: int):
        """Load one image, reset the accumulators and position the Philox windows at global iteration t0."""
        self.x.copy_(im.detach().to(torch.float32).expand(self.nb, -1, -1, -1))
        self.fov.copy_(fov.reshape(self.h0, self.w0))
        self.acc.zero_()
        self.samples.zero_()
        self.iter_base.fill_(t0)
        if seed != self.seed:
            self.graph = self.graph_single = None          # the Philox key is a captured kernel argument
        self.seed = seed
        if self.active:
            base = stream_start + t0 * self.per_iter
            self.masks[0].set_stream_position(base)
            if self.overlap:
                self.masks[1].set_stream_position(base + self.nb * self.per_iter)
            self._generate(0, 2 if self.overlap else 1)           # prologue: masks of the first step

    def run_steps(self, steps: int):
        """`steps` batched steps (= steps * nb iterations), enqueued on the runner's high-priority stream and
        ordered after / before the caller's current stream."""
        cur = torch.cuda.current_stream(self.dev)
        self.main.wait_stream(cur)
        with torch.cuda.stream(self.main):
            self._run_steps(steps)
        cur.wait_stream(self.main)

    def _capture(self, body):
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.main):
            body()
        return g

    def _single(self):
        # masks[0] already hold this step's masks (built by the prologue or by the previous pair)
        self._forward(0)
        if self.active and not self.overlap:
            self._generate(0, 1)

    def _run_steps(self, steps: int):
        """Replays the captured graphs; the FIRST pair / single step of a runner's life runs eagerly (sets kernel
        attributes, counts launches) and is captured right after, so later calls -- including one-step remainder
        batches (125 = 12 x 10 + 5 iterations per rank at 8 GPUs) -- never launch kernel by kernel."""
        pairs, odd = divmod(steps, 2)
        for _ in range(pairs):
            if self.graph is not None:
                self.graph.replay()
                continue
            l0 = _lib.launch_count
            self._pair()
            self.launches_per_step = (_lib.launch_count - l0) // 2
            if self.use_graph:
                self.graph = self._capture(self._pair)
        if odd:
            if self.graph_single is not None:
                self.graph_single.replay()
            else:
                l0 = _lib.launch_count
                self._single()
                if not self.launches_per_step:
                    self.launches_per_step = _lib.launch_count - l0
                if self.use_graph:
                    self.graph_single = self._capture(self._single)


class DropBlockEval(_MCBase):
    MAX_RUNNERS = 4

    def __init__(self, model, num_iterations=1000, return_num=25, mode='save', resize=-1, iter_batch: int = 10,
                 use_cuda_graph: bool = True, overlap_masks: bool = True, gather_samples: bool = False):
        super().__init__(model)
        self.num_iterations = num_iterations
        self.return_num = min(return_num, num_iterations)
        self.gather_samples = gather_samples           # distributed runs: also deliver `tensors` on ranks != 0
        self.set_mode(mode)
        self.resize = resize
        self.iter_batch = iter_batch
        self.use_cuda_graph = use_cuda_graph
        self.overlap_masks = overlap_masks
        self._runners = {}
        self._prologue_stream = None

    def set_mode(self, mode):
        self.mode = mode
        assert self.mode in ['save', 'evaluate']

    def _runner(self, nb, h0, w0, dev, active, p, bs) -> MCRunner:
        eng = self._model._get_engine(dev)              # re-packs weights if the parameters changed
        key = (nb, h0, w0, str(dev), active, p, bs, self.return_num, id(eng), self._model._dropblock_mode())
        r = self._runners.pop(key, None)
        if r is None:
            while len(self._runners) >= self.MAX_RUNNERS:      # least recently used first; a runner owns its workspace + graphs
                self._runners.pop(next(iter(self._runners)))
            r = MCRunner(self._model, nb, h0, w0, dev, active, p, bs, self.return_num, self.use_cuda_graph, self.overlap_masks)
        self._runners[key] = r
        return r

    # -------------------------------------------------------------------------------------------
    def mc_statistics(self, im: torch.Tensor, mask: torch.Tensor, num_iterations: Optional[int] = None):
        """Per-pixel mean / unbiased std of `model(im) * mask` over the Monte-Carlo iterations plus the
        first `return_num` samples.  im: [1,C,H,W], mask: [1,1,H,W] (CUDA)."""
        model = self._model
        if not im.is_cuda:
            raise _lib.B2uError("MC-DropBlock runs on CUDA only")
        if im.shape[0] != 1:
            raise NotImplementedError("the reference loop is batch 1 (Dropblock_Uncertainty.py:84-85)")
        with torch.cuda.device(im.device):             # libb2u launches on the current device's current stream
            return self._mc_statistics(im, mask, num_iterations)

    def _mc_statistics(self, im, mask, num_iterations):
        model = self._model
        T = int(num_iterations if num_iterations is not None else self.num_iterations)
        active, p, bs = model._dropblock_state()
        dist, rank, world = _dist()
        t0, t1 = shard_range(T, rank, world)
        dev = im.device
        _, _, h0, w0 = im.shape
        npix = h0 * w0
        R = self.return_num
        gen = _gen_for(dev)
        seed, stream_start = gen.initial_seed(), gen.get_offset()
        acc = None
        samples = None
        per_iter = 0
        # A rank's share is a run of full batches plus, possibly, remainder batches (125 = 12 x 10 + 5 at 8 GPUs), each with
        # its own runner.  Every runner's prologue (input copy, accumulator reset, mask build of ITS first step) is
        # enqueued before the first runner's steps: the first one on the caller's stream, the later ones on a low-priority
        # side stream, so that their mask builds overlap the first runner's forwards instead of standing alone between
        # two runners (1.4 ms per call at 8 GPUs).
        segments = [(self._runner(nb, h0, w0, dev, active, p, bs), steps, start)
                    for nb, steps, start in plan_segments(t0, t1, self.iter_batch)]
        cur = torch.cuda.current_stream(dev)
        for i, (r, steps, start) in enumerate(segments):
            if i == 0:
                r.begin(im, mask, start, seed, stream_start)
            else:
                if self._prologue_stream is None or self._prologue_stream.device != dev:
                    self._prologue_stream = torch.cuda.Stream(device=dev, priority=0)
                self._prologue_stream.wait_stream(cur)
                with torch.cuda.stream(self._prologue_stream):
                    r.begin(im, mask, start, seed, stream_start)
        for i, (r, steps, start) in enumerate(segments):
            per_iter = r.per_iter
            if i > 0:
                cur.wait_stream(self._prologue_stream)
            r.run_steps(steps)
            if acc is None:
                acc, samples = r.acc, r.samples
            else:                                      # remainder batch: fold into the first runner's buffers
                acc += r.acc
                samples += r.samples
        if acc is None:                                # this rank owns no iterations
            acc = torch.zeros(2, h0, w0, dtype=torch.float64, device=dev)
            samples = torch.zeros(max(R, 1), h0, w0, dtype=torch.float32, device=dev)
        if active:
            # leave the torch generator where T reference forwards would have left it
            if per_iter == 0:
                per_iter = self._runner(1, h0, w0, dev, active, p, bs).per_iter
            gen.set_offset(stream_start + T * per_iter)
        if dist is not None and world > 1:
            dist.all_reduce(acc)                       # the ONE exchange step of the path (fp64 [2,H,W])
            if R > shard_range(T, 0, world)[1] or self.gather_samples:
                dist.all_reduce(samples)               # disjoint writers (zeros elsewhere): a sum is a gather
            # else: rank 0 owns iterations [0, R) -- the block partition puts them there -- so `tensors` is
            # complete on rank 0, the rank that saves it, and the 33 MB exchange is skipped (other ranks return zeros)
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
            # Dropblock_Uncertainty.py:52-61: pad to square, resize im / gt / mask on the fly (one fused kernel each)
            from .resize import square_pad_resize
            im = square_pad_resize(im, self.resize)
            gt = square_pad_resize(gt, self.resize) if gt is not None else None
            mask = square_pad_resize(mask, self.resize)
        mean, std, tensors = self.mc_statistics(im, mask)
        if self.mode == 'save':
            return batch_idx, (mean, std, tensors)
        return batch_idx, mean, im, gt, mask


class RotationRunner:
    """Persistent state of the rotation-ensemble loop for one (model, angle batch, image size): workspace, rotated-input
    buffer, device angle tables, accumulators and ONE CUDA graph per step

        rotate-in (angles from the device table at *iter_base) -> eval forward -> rotate-back * fov -> fp64 accumulate
        (+ the first return_num samples) -> advance iter_base

    so the 359-angle loop of BASELINE configs[3] is 72 graph replays with nothing allocated and no host tables built."""

    def __init__(self, model, nb: int, cin: int, h0: int, w0: int, device, return_num: int, use_cuda_graph: bool = True):
        self.model, self.nb, self.cin, self.h0, self.w0, self.dev = model, nb, cin, h0, w0, device
        self.eng: UNetEngine = model._get_engine(device)
        self.ws = self.eng.workspace(nb, h0, w0)
        self.R = return_num
        self.use_graph = use_cuda_graph
        self.x1 = torch.zeros(cin, h0, w0, dtype=torch.float32, device=device)
        self.rot_in = torch.zeros(nb, cin, h0, w0, dtype=torch.float32, device=device)
        self.fov = torch.zeros(h0, w0, dtype=torch.float32, device=device)
        self.acc = torch.zeros(2, h0, w0, dtype=torch.float64, device=device)
        self.samples = torch.zeros(max(return_num, 1), h0, w0, dtype=torch.float32, device=device)
        self.iter_base = torch.zeros(1, dtype=torch.int64, device=device)
        self.iter_limit = torch.zeros(1, dtype=torch.int64, device=device)
        self.table_len = 0
        self.tab_in = self.tab_out = None
        self.graph = None
        self.launches_per_step = 0

    def _tables(self, total: int):
        """Rows i = 0..total-1 hold angle (i + 1) degrees (Rotational_Uncertainty.py:51: range(1, T + 1)) for the
        rotate-in and angle -(i + 1) for the rotate-back."""
        import ctypes as C
        if total <= self.table_len:
            return
        n = max(total, 360)
        host = (C.c_float * (6 * n))()
        for sign, name in ((1.0, "tab_in"), (-1.0, "tab_out")):
            ang = (C.c_double * n)(*[sign * float(i + 1) for i in range(n)])
            call("b2u_rotation_table", ang, n, self.h0, self.w0, host)
            t = torch.frombuffer(bytearray(host), dtype=torch.float32).clone().to(self.dev)
            setattr(self, name, t)
        self.table_len = n
        self.graph = None                                  # table pointers are captured kernel arguments

    def begin(self, im: torch.Tensor, fov: torch.Tensor, t0: int, t1: int, total: int):
        self._tables(total)
        self.x1.copy_(im.detach().to(torch.float32).reshape(self.cin, self.h0, self.w0))
        self.fov.copy_(fov.reshape(self.h0, self.w0))
        self.acc.zero_()
        self.samples.zero_()
        self.iter_base.fill_(t0)
        self.iter_limit.fill_(t1)

    def _step(self):
        st = stream_ptr()
        call("b2u_rotate_in_table", ptr(self.x1), ptr(self.rot_in), self.nb, self.cin, self.h0, self.w0, ptr(self.tab_in),
             self.table_len, ptr(self.iter_base), st)
        seg = self.eng.forward(self.rot_in, self.ws, None)      # eval forward, DropBlock is Identity (:122)
        call("b2u_rotate_back_accumulate", ptr(seg), ptr(self.fov), ptr(self.acc), ptr(self.samples) if self.R > 0 else None,
             self.nb, self.h0, self.w0, self.R, ptr(self.tab_out), self.table_len, ptr(self.iter_base), ptr(self.iter_limit), st)
        call("b2u_advance_counter", ptr(self.iter_base), self.nb, st)

    def run_steps(self, steps: int):
        for _ in range(steps):
            if self.graph is not None:
                self.graph.replay()
                continue
            l0 = _lib.launch_count
            self._step()
            self.launches_per_step = _lib.launch_count - l0
            if self.use_graph:
                torch.cuda.synchronize(self.dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step()
                self.graph = g


class RotationEval(_MCBase):
    """Angles 1..num_iterations degrees: rotate in -> eval forward -> rotate back -> * mask -> mean/std
    (Rotational_Uncertainty.py:21-68), angles sharded over torch.distributed ranks like the MC iterations."""
    MAX_RUNNERS = 3

    def __init__(self, model, num_iterations=1000, return_num=25, resize=-1, angle_batch: int = 10, use_cuda_graph: bool = True,
                 gather_samples: bool = False):
        super().__init__(model)
        self.num_iterations = num_iterations
        self.return_num = min(return_num, num_iterations)
        self.resize = resize
        self.angle_batch = angle_batch
        self.use_cuda_graph = use_cuda_graph
        self.gather_samples = gather_samples
        self._runners = {}

    def _runner(self, nb, cin, h0, w0, dev) -> RotationRunner:
        eng = self._model._get_engine(dev)              # re-packs weights if the parameters changed
        key = (nb, cin, h0, w0, str(dev), self.return_num, id(eng))
        r = self._runners.pop(key, None)
        if r is None:
            while len(self._runners) >= self.MAX_RUNNERS:
                self._runners.pop(next(iter(self._runners)))
            r = RotationRunner(self._model, nb, cin, h0, w0, dev, self.return_num, self.use_cuda_graph)
        self._runners[key] = r
        return r

    def rotation_statistics(self, im: torch.Tensor, mask: torch.Tensor, num_iterations: Optional[int] = None):
        if not im.is_cuda:
            raise _lib.B2uError("the rotation ensemble runs on CUDA only")
        if im.shape[0] != 1:
            raise NotImplementedError("the reference loop is batch 1")
        with torch.cuda.device(im.device):
            return self._rotation_statistics(im, mask, num_iterations)

    def _rotation_statistics(self, im, mask, num_iterations):
        T = int(num_iterations if num_iterations is not None else self.num_iterations)
        dist, rank, world = _dist()
        t0, t1 = shard_range(T, rank, world)
        dev = im.device
        _, cin, h0, w0 = im.shape
        npix = h0 * w0
        R = self.return_num
        # angles per step: at most `angle_batch`, balanced over the steps of the largest share (45 angles per rank at 8 GPUs
        # run as 5 steps of 9, not 4 of 10 and a half-empty fifth); every rank uses the same batch, so the per-image
        # arithmetic -- which does not depend on the batch anyway -- runs through the same kernels
        share = max(1, -(-T // world))
        nsteps = -(-share // max(1, self.angle_batch))
        nb = -(-share // nsteps)
        r = self._runner(nb, cin, h0, w0, dev)
        r.begin(im, mask.to(torch.float32), t0, t1, T)
        r.run_steps(-(-(t1 - t0) // nb))               # the tail images of the last step are skipped by *iter_limit
        acc, samples = r.acc, r.samples
        if dist is not None and world > 1:
            dist.all_reduce(acc)
            if R > shard_range(T, 0, world)[1] or self.gather_samples:
                dist.all_reduce(samples)
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
        if self.resize != -1:
            # Rotational_Uncertainty.py:39-49: pad to square, resize im / gt / mask on the fly (one fused kernel each)
            from .resize import square_pad_resize
            im = square_pad_resize(im, self.resize)
            gt = square_pad_resize(gt, self.resize) if gt is not None else None
            mask = square_pad_resize(mask, self.resize)
        mean, std, tensors = self.rotation_statistics(im, mask)
        return batch_idx, (mean, std, tensors)
