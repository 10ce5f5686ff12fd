"""bench.py -- headline benchmark of the U-Net / MC-DropBlock hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2]): dependent MC-DropBlock uncertainty of ONE synthetic DRIVE-shaped
584x565 image (autopad 592x576), canonical U-Net (31.04 M parameters, GroupNorm(32), DropBlock2D bs 7
p 0.15), bf16 tcgen05 convolutions.  A "step" is one batched pass of the hot path: `iter_batch`
Monte-Carlo iterations (mask build -> forward -> head accumulate -> Philox advance), replayed as a CUDA
graph.  `value` = forward passes per second over all ranks (weak scaling: every rank runs K steps of
`iter_batch` iterations on the same image with its own Philox window; no data-path collective inside a
step -- the single fp64 all-reduce that closes a real run is timed separately and reported).

JSON keys follow the driver contract; see DESIGN.md "Measurement" for how each number is taken.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H0, W0 = 584, 565
FLOP_PER_FORWARD = 500.46e9          # 2*M*N*K over the 18 conv3x3 + 4 convT + head
CONV_FLOP_PER_FORWARD = 500.03e9     # the tensor-core convs only (first layer 0.39 G and head 0.04 G excluded)
METRIC = "MC-DropBlock fwd passes/s at 584x565"
UNIT = "passes/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops", 1590.0),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


def load_traffic():
    """DRAM bytes (read + write) of the conv family per step from the committed ncu capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        """Summary over the samples that arrived inside [t_begin, t_end] (the timed region); the sampler is
        started before the warm-up because nvidia-smi needs a few hundred ms to produce its first line."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for (ts, ln) in self.lines if t_begin is None or (t_begin <= ts <= t_end + 0.12)]
        window = "timed region"
        if not inside:
            inside = [ln for (_, ln) in self.lines[-5:]]
            window = "last samples before the end of the timed region (region shorter than the sampling period)"
        for ln in inside:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


# ------------------------------------------------------------------------------------------------ CPU reference arm
def oracle_mc_forward_timer(steps: int, warmup: int, budget_s: float):
    """Times the reference algorithm (oracle port, torch fp32 on the host cores, all threads) for the same
    workload: one MC-DropBlock forward of the 584x565 image per step."""
    import torch
    from oracle import unet_oracle as O
    from unet_research_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synthetic.make_state_dict(seed=1234)
    x = synthetic.make_image(H0, W0, seed=1234)
    fov = synthetic.make_fov_mask(H0, W0)
    db = O.DropBlockCfg(0.15, 7, True)
    torch.manual_seed(1234)

    def one():
        with torch.no_grad():
            return O.unet_forward(sd, x, dropblock=db) * fov

    t0 = time.perf_counter()
    one()
    first = time.perf_counter() - t0
    # bounded sample: keep the whole run inside the budget
    w = max(0, min(warmup, int(budget_s * 0.2 / max(first, 1e-3)) - 1))
    k = max(1, min(steps, int(budget_s * 0.8 / max(first, 1e-3))))
    for _ in range(w):
        one()
    t0 = time.perf_counter()
    for _ in range(k):
        one()
    dt = time.perf_counter() - t0
    return k / dt, k, w + 1, torch.get_num_threads()


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    val, k, w, cores = oracle_mc_forward_timer(args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": k, "warmup": w,
        "ms_per_step": 1000.0 / val, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "configs[2]: dependent MC-DropBlock, canonical U-Net, one 584x565 image, DropBlock bs7 p0.15",
                   "step": "one MC forward pass (CPU, torch fp32, all host threads)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{k} timed MC forward passes of the 584x565 image after {w} warm-up (requested {args.steps}/{args.warmup}; "
                                   "bounded to ~150 s of CPU work); oracle port of the reference (the Python reference cannot travel to the GPU box)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    rank, local, world = dist_env()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"              # keep NCCL's banner off stdout: rank 0 prints ONE JSON line
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    import unet_research_b200 as U
    from unet_research_b200 import _lib, synthetic
    from unet_research_b200._lib import call, ptr, stream_ptr
    from unet_research_b200.canonical import build_canonical

    peaks = load_peaks()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    NB = args.iter_batch
    K, W = args.steps, args.warmup
    model, _ = build_canonical(dev, dropblock=True, compute="bf16")
    model.apply(U.set_dropblock_on)
    x = synthetic.make_image(H0, W0, seed=1234).to(dev)
    fov2d = synthetic.make_fov_mask(H0, W0).to(dev).reshape(H0, W0).contiguous()
    ev = U.DropBlockEval(model, num_iterations=1000, return_num=25, iter_batch=NB)
    runner = ev._runner(NB, H0, W0, dev, True, 0.15, 7)
    seed = 1234
    torch.manual_seed(seed)                                # the e2e leg reads the same key from torch's CUDA generator
    # disjoint Philox windows / sample indices per rank: rank r starts at global iteration r * 10^6
    runner.begin(x, fov2d, rank * 1_000_000, seed, 0)
    # warm-up: one eager pair of steps (sets kernel attributes, counts launches), graph capture, replays
    W_eff = max(W, 3)
    runner.run_steps(4 + 2 * ((W_eff + 1) // 2))
    torch.cuda.synchronize(dev)
    launches_per_step = runner.launches_per_step
    acc = runner.acc

    # ---- timed region: K steps (graph replays of two steps each), CUDA events on the launching stream,
    # barrier + sync on both sides.  Working set per step (~1 GB of activations x NB) far exceeds the 126 MB L2.
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    e0.record()
    runner.run_steps(K)
    e1.record()
    torch.cuda.synchronize(dev)
    t_end = time.perf_counter()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * K * NB / (ms_max / 1000.0)

    # ---- the closing exchange of a real run (fp64 [2,H,W] all-reduce), timed on its own
    allreduce_ms = None
    if world > 1:
        for _ in range(3):
            dist.all_reduce(acc)
        torch.cuda.synchronize(dev)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            dist.all_reduce(acc)
        a1.record()
        torch.cuda.synchronize(dev)
        allreduce_ms = a0.elapsed_time(a1) / 10

    # ---- roofline of the dominant kernel family (conv_gemm_kernel): per-launch CUDA-event timing, eager mode
    conv_ms, other = time_conv_kernels(runner, dev, reps=max(3, min(K, 10)))
    conv_tflops = CONV_FLOP_PER_FORWARD * NB / (conv_ms / 1000.0) / 1e12
    eb = elementwise_bytes_per_step(NB)
    hbm = {}
    for name, nbytes in eb.items():
        key = name if name in other else name.replace("b2u_gn_finalize", "b2u_gn_finalize_ex")
        if key in other and other[key] > 0:
            gbs = nbytes / (other[key] / 1000.0) / 1e9
            hbm[name] = {"algorithmic_mb_per_step": nbytes / 1e6, "ms_per_step": other[key], "achieved_gbs": gbs,
                         "frac_of_hbm_peak": gbs / peaks["hbm_gbs"]}
    tj = load_traffic()
    traffic = None
    traffic_detail = None
    if tj and tj.get("iter_batch") == NB:
        # DRAM bytes (read + write) of the conv family per step = per group of 21 launches, like `achieved`
        traffic = tj.get("conv_family_dram_bytes_per_step")
        traffic_detail = {"unit": "bytes per step (17 conv3x3 + 4 convT launches)", "source": tj.get("source"),
                          "algorithmic_bytes_per_step": conv_bytes_per_step(NB)}
    roofline = {"bound": "tensor", "kernel": "conv3x3_v2_kernel + convT_v2_kernel (17 conv3x3 + 4 convT launches per step)",
                "achieved": conv_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": conv_tflops / peaks["bf16_tflops_sustained"], "traffic": traffic, "traffic_detail": traffic_detail,
                "hbm_bound_kernels": hbm, "hbm_peak_gbs": peaks["hbm_gbs"],
                "peak_source": peaks["src"] + " bf16_tflops_sustained (kernels timed inside a long step)",
                "conv_ms_per_step": conv_ms, "step_ms": ms_max / K, "conv_share_of_step": conv_ms / (ms_max / K),
                "other_kernels_ms_per_step": other}

    # ---- e2e: the public API with HOST buffers (pinned), H2D of image + mask and D2H of mean/std/samples inside the timed region
    e2e = None
    if not args.no_e2e:
        T_e2e = args.e2e_iters
        ev.num_iterations = T_e2e
        im_h = synthetic.make_image(H0, W0, seed=1234).pin_memory()
        fov_h = synthetic.make_fov_mask(H0, W0).pin_memory()
        outs_h = [torch.empty(1, 1, H0, W0).pin_memory(), torch.empty(1, 1, H0, W0).pin_memory(),
                  torch.empty(25, 1, 1, H0, W0).pin_memory()]

        def e2e_once():
            im_d = im_h.to(dev, non_blocking=True)
            fov_d = fov_h.to(dev, non_blocking=True)
            _, (mean, std, tens) = ev.predict_step((im_d, None, fov_d), 0)
            outs_h[0].copy_(mean, non_blocking=True)
            outs_h[1].copy_(std, non_blocking=True)
            outs_h[2].copy_(tens, non_blocking=True)
            torch.cuda.synchronize(dev)

        e2e_once()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            e2e_once()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        # under torch.distributed predict_step shards T_e2e over the ranks, so T_e2e passes complete per call
        e2e = {"value": reps * T_e2e / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": im_h.numel() * 4 + fov_h.numel() * 4,
               "d2h_bytes_per_step": sum(o.numel() * 4 for o in outs_h),
               "step": f"one DropBlockEval.predict_step call = {T_e2e} iterations (BASELINE configs[2] iter_num; sharded over ranks), pinned host in/out, bytes per call",
               "mc_1000_iter_projected_s": 1000.0 / (reps * T_e2e / float(tt.item())),
               "seconds_per_call": float(tt.item()) / reps}

    train = None
    if not args.no_train:
        train = bench_train(dev, rank, world, args.train_steps, 4)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        val, k, w, cores = oracle_mc_forward_timer(2, 0, budget_s=25.0)
        cpu_baseline = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{k} MC-DropBlock forward passes of the same 584x565 image (oracle port, torch fp32, all host threads) after 1 warm-up"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "configs[2]: dependent MC-DropBlock uncertainty, canonical U-Net (filters 64, depth 4, GroupNorm 32), "
                                   "one 584x565 image (autopad 592x576), DropBlock2D block_size 7 drop_prob 0.15",
                       "iter_batch": NB, "passes_per_step": NB, "l2": "inputs larger than L2 (about 1 GB of activations per iteration)",
                       "parallelism": f"mc-iteration sharding x{world}", "cuda_graph": True,
                       "mask_build": "side stream, overlapped with the forward of the previous step"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * K, "launches_per_step": launches_per_step,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "train": train,
            "tflops_whole_step": FLOP_PER_FORWARD * NB * world * K / (ms_max / 1000.0) / 1e12,
            "mc_1000_iter_projected_s": 1000.0 / value, "allreduce_ms": allreduce_ms,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def conv_bytes_per_step(nb, filters=64, depth=4, h=592, w=576):
    """ALGORITHMIC HBM bytes of the tensor-core conv family for one step of nb batched iterations: every conv reads
    its bf16 input once and writes its bf16 output once, weights once per launch (the first conv is not in the family)."""
    tot = 0.0
    c = filters
    for lvl in range(depth + 1):
        m = float(h >> lvl) * (w >> lvl) * nb
        cin1 = c // 2 if lvl > 0 else None                     # encoder unit 1 input channels (level 0 = first conv)
        if lvl > 0:
            tot += 2.0 * m * (cin1 + c) + 2.0 * 9 * cin1 * c    # encoder / bottleneck unit 1
        tot += 2.0 * m * (c + c) + 2.0 * 9 * c * c              # unit 2
        if lvl < depth:
            tot += 2.0 * m * (2 * c + c) + 2.0 * 9 * 2 * c * c  # decoder unit 1 (reads the concat buffer)
            tot += 2.0 * m * (c + c) + 2.0 * 9 * c * c          # decoder unit 2
            tot += 2.0 * (m / 4) * (2 * c) + 2.0 * m * c + 2.0 * 4 * 2 * c * c   # up-conv into this level
        c *= 2
    return tot


def elementwise_bytes_per_step(nb, filters=64, depth=4, h=592, w=576, h0=H0, w0=W0):
    """ALGORITHMIC HBM bytes of one Monte-Carlo step (nb batched iterations, bf16 activations) per fused kernel
    family, from the tensor shapes alone: every activation read once / written once, bit masks
    1 bit per element, statistics and coefficients ignored."""
    out = {"b2u_gn_apply": 0.0, "b2u_gn_apply_pool": 0.0, "b2u_head_fwd": 0.0, "b2u_conv_first_fwd": 0.0,
           "b2u_dropblock_dilate": 0.0}
    c = filters
    bits = 0.0
    for lvl in range(depth + 1):
        e = float(h >> lvl) * (w >> lvl) * c                     # elements of one conv output at this level
        if lvl < depth:
            # encoder: unit 1 apply (first level reads the SHARED raw tensor once), unit 2 apply+pool, pooled-GN apply
            out["b2u_gn_apply"] += (2.0 * e * (1 if lvl == 0 else nb) + 2.0 * e * nb + e * nb / 8.0)
            out["b2u_gn_apply_pool"] += nb * (2.0 * e + 2.0 * e + 0.5 * e + e / 8.0 + e / 8.0)
            out["b2u_gn_apply"] += nb * (2.0 * e / 4 + 2.0 * e / 4)
            # decoder at the same resolution: up-conv apply (with concat mask), unit 1 apply, unit 2 apply (not the last level)
            out["b2u_gn_apply"] += nb * (4.0 * e + e / 8.0)
            out["b2u_gn_apply"] += nb * (4.0 * e + e / 8.0)
            if lvl > 0:
                out["b2u_gn_apply"] += nb * (4.0 * e + e / 8.0)
            bits += nb * (2 * e + 2 * e + 2 * e)                 # sites: 2 encoder units, concat (2c), 2 decoder units
        else:
            out["b2u_gn_apply"] += nb * 2 * (4.0 * e + e / 8.0)
            bits += nb * 2 * e
        c *= 2
    e0 = float(h) * w * filters
    out["b2u_head_fwd"] = nb * (2.0 * e0 + e0 / 8.0) + 16.0 * h0 * w0 * 2
    out["b2u_conv_first_fwd"] = 2.0 * e0 + 4.0 * h0 * w0
    out["b2u_dropblock_dilate"] = 2.0 * bits / 8.0                # centre bitmap in, keep bitmap out
    return out


def bench_train(dev, rank, world, steps, warmup):
    """BASELINE configs[1] (+ configs[4] data-parallel part): fwd + bwd + clip + SGD, batch 1 per GPU, 584x565,
    DropBlock bs 7 p .15, bf16; with N > 1 every rank trains on its own image and the gradients are averaged over
    NCCL inside backward (decoder-side bucket overlapped with the encoder backward).  Returns the `train` object."""
    import torch
    import torch.distributed as dist
    from torch import nn
    import unet_research_b200 as U
    from unet_research_b200 import synthetic
    from unet_research_b200.canonical import build_canonical
    model, _ = build_canonical(dev, dropblock=True, compute="bf16")
    model.train()
    model.data_parallel = world > 1
    x = synthetic.make_image(H0, W0, seed=1234 + rank).to(dev)
    gt = synthetic.make_gt(H0, W0, seed=1234 + rank).to(dev)
    fov = synthetic.make_fov_mask(H0, W0).to(dev)
    tm = U.BaseUNetTraining(model, nn.BCELoss(), None)
    opt = U.FusedSGD(model.parameters(), lr=1e-3, momentum=0.99, max_grad_norm=0.5) if hasattr(U, "FusedSGD") else \
        torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.99)
    fused = hasattr(U, "FusedSGD")

    def step():
        opt.zero_grad(set_to_none=True)
        loss = tm.training_step((x.clone(), gt, fov), 0)
        loss.backward()
        if not fused:
            torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
        opt.step()
        return loss

    for _ in range(max(warmup, 4)):
        l0 = step()
    first = float(l0.item())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    ws = list(model._engine._workspaces.values())[0]
    graphs = ws.train_step.fwd_graph is not None and ws.train_step.bwd_graphs is not None
    return {"metric": "train imgs/s (fwd + bwd + clip + SGD, batch 1 per GPU, 584x565, DropBlock bs7 p0.15, bf16)",
            "value": world / (ms / 1000.0), "unit": "imgs/s", "ms_per_step": ms, "steps": steps, "global_batch": world,
            "tflops": 1501.4e9 * world / (ms / 1000.0) / 1e12, "flop_per_image": 1501.4e9, "cuda_graphs": graphs,
            "optimizer": "fused SGD-momentum + global-norm clip (one multi-tensor kernel pair)" if fused else "torch.optim.SGD + clip_grad_norm_",
            "gradient_allreduce_bytes": (124.16e6 if world > 1 else 0), "allreduce": "NCCL AVG, 2 buckets, decoder bucket overlapped with the encoder backward" if world > 1 else None,
            "loss_first": first, "loss_last": float(loss.item())}


def time_conv_kernels(runner, dev, reps):
    """Eager replays with a CUDA event pair around every C-ABI launch; returns (sum of conv_gemm ms per step,
    {entry point: ms per step} for everything else)."""
    import torch
    from unet_research_b200 import _lib
    records = []
    orig = _lib.call

    def timed_call(name, *a):
        if name in _lib._LAUNCHERS and not name.startswith("b2u_pack"):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            orig(name, *a)
            e.record()
            records.append((name, s, e))
        else:
            orig(name, *a)

    import unet_research_b200.engine as E
    E.call = timed_call
    try:
        for _ in range(reps):
            # a ~10 ms spin kernel first: the whole step is enqueued while it runs, so every event pair below
            # brackets back-to-back GPU execution and never the host's launch latency
            torch.cuda._sleep(20_000_000)
            runner.masks[0].generate(runner.seed)
            runner.eng.forward(runner.x, runner.ws, runner.masks[0], head_out=False, mc=runner.mc, shared_input=True)
        torch.cuda.synchronize(dev)
    finally:
        E.call = orig
    tot = {}
    for name, s, e in records:
        tot[name] = tot.get(name, 0.0) + s.elapsed_time(e)
    per_step = {k: v / reps for k, v in tot.items()}
    conv = per_step.pop("b2u_conv3x3_fwd", 0.0) + per_step.pop("b2u_convT2x2_fwd", 0.0)
    return conv, per_step


_JSON_OUT = None


def emit(line: dict):
    """The ONE JSON line of the contract, on the process's original stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # Libraries chat on stdout (NCCL prints its version banner there under torchrun): keep the original stdout for
    # the JSON line only and send everything else to stderr.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--iter-batch", type=int, default=10)
    ap.add_argument("--e2e-iters", type=int, default=1000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--train-steps", type=int, default=30)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
