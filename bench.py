"""bench.py -- headline benchmark of the U-Net / MC-DropBlock hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2]): dependent MC-DropBlock uncertainty of ONE synthetic DRIVE-shaped
584x565 image (autopad 592x576), canonical U-Net (31.04 M parameters, GroupNorm(32), DropBlock2D bs 7
p 0.15), tcgen05 convolutions with 16-bit operands (fp16 by default: the mode that meets north_star's 1e-2
logit bar; bf16 beside it as `alt_dtype`) and fp32 accumulation.  A "step" is one batched pass of the hot path: `iter_batch`
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
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


WORKLOAD = ("configs[2]: dependent MC-DropBlock uncertainty, canonical U-Net (filters 64, depth 4, GroupNorm 32), "
            "one 584x565 image (autopad 592x576), DropBlock2D block_size 7 drop_prob 0.15")


def workload_config(world: int, nb: int):
    """The `config` object: identical for the GPU arm and the reference arm (it names the workload, not the implementation)."""
    return {"workload": WORKLOAD, "image": "584x565x1 synthetic DRIVE-shaped, seed 1234", "iter_num": 1000, "save_num": 25,
            "step": f"one batched pass of the hot path = {nb} Monte-Carlo iterations (the CPU arm times a bounded sample of single passes)",
            "passes_per_step": nb, "l2": "inputs larger than L2 (about 1 GB of activations per iteration)",
            "parallelism": f"mc-iteration sharding x{world}"}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_model():
    """(kind, forward_fn(x, mc: bool), train_step_fn, rotation_fn): the reference's OWN modules (unmodified, from
    /root/reference or the byte-code in oracle/_ref -> kind "reference") or, when neither is present, the oracle
    port (kind "port").  All on the host CPU, torch fp32, every host thread."""
    import torch
    from torch import nn
    from unet_research_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synthetic.make_state_dict(seed=1234)
    from oracle import ref_shims as R
    if R.reference_available():
        ref = R.load_reference()
        unet = R.build_reference_unet(ref, dropblock=ref.DropBlock2D, drop_prob=0.15, block_size=7)
        unet.load_state_dict(sd)
        unet.eval()

        def fwd(x, mc):
            with torch.no_grad():
                unet.eval()
                if mc:
                    unet.apply(ref.set_dropblock_on)           # Dropblock_Uncertainty.py:50
                return unet(x)

        def train_step(x, gt, fov):
            unet.train()
            tm = ref.UNetTraining(unet, nn.BCELoss(), lr=1e-3, momentum=0.99)
            for p_ in unet.parameters():
                p_.grad = None
            loss = tm.training_step((x.clone(), gt, fov), 1)
            loss.backward()
            unet.eval()
            return float(loss)

        def rotation(x, fov, angles):
            ev = ref.RotationEval(unet, num_iterations=angles, return_num=1)
            with torch.no_grad():
                unet.eval()
                return ev.predict_step((x, None, fov), 0)

        return "reference", fwd, train_step, rotation
    from oracle import unet_oracle as O
    db = O.DropBlockCfg(0.15, 7, True)

    def fwd(x, mc):
        with torch.no_grad():
            return O.unet_forward(sd, x, dropblock=db if mc else None)

    def train_step(x, gt, fov):
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        loss = O.train_step_loss(params, x, gt, fov, db)
        loss.backward()
        return float(loss)

    def rotation(x, fov, angles):
        with torch.no_grad():
            return O.rotation_ensemble(sd, x, fov, angles, 1)

    return "port", fwd, train_step, rotation


def oracle_mc_forward_timer(steps: int, warmup: int, budget_s: float, extras: bool = False):
    """Times the reference implementation on the host cores for the same workload: one MC-DropBlock forward of the
    584x565 image per step (bounded sample).  extras=True adds SURVEY 8(d)'s other CPU legs: eval forward, one train step,
    one rotation angle."""
    import torch
    from unet_research_b200 import synthetic
    kind, fwd, train_step, rotation = cpu_reference_model()
    x = synthetic.make_image(H0, W0, seed=1234)
    fov = synthetic.make_fov_mask(H0, W0)
    torch.manual_seed(1234)

    def one():
        return fwd(x, True) * fov

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
    other = None
    if extras:
        other = {}
        fwd(x, False)
        t0 = time.perf_counter()
        for _ in range(2):
            fwd(x, False)
        other["eval_forward_s"] = (time.perf_counter() - t0) / 2
        gt = synthetic.make_gt(H0, W0)
        t0 = time.perf_counter()
        train_step(x, gt, fov)
        other["train_step_fwd_bwd_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        rotation(x, fov, 1)
        other["rotation_angle_s"] = time.perf_counter() - t0
        other["note"] = ("same host cores, torch fp32: eval forward (mean of 2 after 1 warm-up), ONE training_step + backward "
                         "(DropBlock on, no warm-up), ONE rotation angle (rotate in, eval forward, rotate back)")
    return k / dt, k, w + 1, torch.get_num_threads(), kind, other


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    val, k, w, cores, kind, _ = oracle_mc_forward_timer(args.steps, args.warmup, budget_s=150.0)
    what = ("the UNMODIFIED reference modules (UNet + DropBlock2D, eval() + set_dropblock_on, no_grad), imported from "
            "/root/reference or its byte-code in oracle/_ref" if kind == "reference" else "oracle port of the reference")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": k, "warmup": w,
        "ms_per_step": 1000.0 / val, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(max(args.gpus, 1), args.iter_batch),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{k} timed MC forward passes of the 584x565 image after {w} warm-up (requested {args.steps}/{args.warmup}; "
                                   f"bounded to ~150 s of CPU work); {what}; torch fp32 on all host threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def _event_time(dev, fn):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1)


def run_gpu(args):
    import hashlib
    import torch
    import torch.distributed as dist
    rank, local, world = dist_env()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    import unet_research_b200 as U
    from unet_research_b200 import _lib, synthetic
    from unet_research_b200.canonical import build_canonical

    peaks = load_peaks()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    NB = args.iter_batch
    K, W = args.steps, args.warmup
    x = synthetic.make_image(H0, W0, seed=1234).to(dev)
    fov2d = synthetic.make_fov_mask(H0, W0).to(dev).reshape(H0, W0).contiguous()
    seed = 1234

    def make_runner(compute, fused=None):
        model, _ = build_canonical(dev, dropblock=True, compute=compute)
        model.apply(U.set_dropblock_on)
        if fused is not None:
            model._get_engine(dev).fused_prologue = fused
        ev = U.DropBlockEval(model, num_iterations=1000, return_num=25, iter_batch=NB)
        runner = ev._runner(NB, H0, W0, dev, True, 0.15, 7)
        torch.manual_seed(seed)                            # the e2e leg reads the same key from torch's CUDA generator
        # disjoint Philox windows / sample indices per rank: rank r starts at global iteration r * 10^6
        runner.begin(x, fov2d, rank * 1_000_000, seed, 0)
        return model, ev, runner

    def timed_steps(runner, k):
        """k steps (graph replays of two steps each), CUDA events on the launching stream, barrier + sync on both sides,
        max over ranks.  Working set per step (~1 GB of activations x NB) far exceeds the 126 MB L2."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        tb = time.perf_counter()
        ms = _event_time(dev, lambda: runner.run_steps(k))
        te = time.perf_counter()
        if world > 1:
            dist.barrier()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), tb, te

    # The benchmarked mode is the package default for inference (compute_dtype "auto" -> fp16 operands, fp32 accumulate):
    # the 16-bit mode that meets north_star's 1e-2 logit bar (tests/test_gpu_parity.py::test_forward_default_mode_vs_oracle).
    compute = args.dtype
    dtype_name = {"auto": "fp16"}.get(compute, compute)
    model, ev, runner = make_runner(compute)
    # warm-up: one eager pair of steps (sets kernel attributes, counts launches), graph capture, replays
    W_eff = max(W, 3)
    runner.run_steps(4 + 2 * ((W_eff + 1) // 2))
    torch.cuda.synchronize(dev)
    launches_per_step = runner.launches_per_step
    acc = runner.acc
    ms_max, t_begin, t_end = timed_steps(runner, K)
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    value = world * K * NB / (ms_max / 1000.0)
    # a longer region right after (the board's power cap settles lower over ~1 s): reported next to the K-step value
    ms_long, _, _ = timed_steps(runner, 100)
    sustained = {"steps": 100, "value": world * 100 * NB / (ms_long / 1000.0), "ms_per_step": ms_long / 100}

    # ---- the closing exchange of a real run (fp64 [2,H,W] all-reduce), timed on its own
    allreduce_ms = None
    if world > 1:
        for _ in range(3):
            dist.all_reduce(acc)
        torch.cuda.synchronize(dev)
        allreduce_ms = _event_time(dev, lambda: [dist.all_reduce(acc) for _ in range(10)]) / 10

    # ---- roofline of the dominant kernel family: per-launch CUDA-event timing, eager mode
    conv_ms, other = time_conv_kernels(runner, dev, reps=max(3, min(K, 10)))
    conv_tflops = CONV_FLOP_PER_FORWARD * NB / (conv_ms / 1000.0) / 1e12
    eb = elementwise_bytes_per_step(NB, fused=(sorted(runner.eng.fuse_levels) if getattr(runner.eng, "fused_prologue", False) else False))
    hbm = {}
    for name, nbytes in eb.items():
        key = next((k for k in (name, name + "_ex", name + "_v2") if k in other), name)
        if key in other and other[key] > 0 and nbytes > 0:
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
                          "algorithmic_bytes_per_step": conv_bytes_per_step(NB),
                          "whole_step_dram_bytes": tj.get("step_dram_bytes")}
    step_ms = ms_max / K
    roofline = {"bound": "tensor", "kernel": "conv3x3_v2_kernel + convT_v2_kernel (17 conv3x3 + 4 convT launches per step)",
                "achieved": conv_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": conv_tflops / peaks["bf16_tflops_sustained"], "traffic": traffic, "traffic_detail": traffic_detail,
                "hbm_bound_kernels": hbm, "hbm_peak_gbs": peaks["hbm_gbs"],
                "peak_source": peaks["src"] + " bf16_tflops_sustained (kernels timed inside a long step; fp16 and bf16 operands run at the same tensor rate)",
                "conv_ms_per_step": conv_ms, "step_ms": step_ms, "conv_share_of_step": conv_ms / step_ms,
                "whole_step_tflops": FLOP_PER_FORWARD * NB / (step_ms / 1000.0) / 1e12,
                "whole_step_frac": FLOP_PER_FORWARD * NB / (step_ms / 1000.0) / 1e12 / peaks["bf16_tflops_sustained"],
                "other_kernels_ms_per_step": other, "sustained_100_steps": sustained}

    # ---- the other 16-bit mode for comparison (bf16 when the line is fp16 and vice versa): same schedule, same kernels
    alt = None
    if not args.no_alt:
        alt_compute = "bf16" if dtype_name == "fp16" else "fp16"
        _, _, r2 = make_runner(alt_compute)
        r2.run_steps(6)
        torch.cuda.synchronize(dev)
        ms2, _, _ = timed_steps(r2, max(K, 20))
        alt = {"dtype": alt_compute, "value": world * max(K, 20) * NB / (ms2 / 1000.0), "ms_per_step": ms2 / max(K, 20),
               "note": "logits rel 1.7e-2 (bf16) vs 2.1e-3 (fp16) against the fp64 oracle at 584x565; north_star's 16-bit bar is 1e-2"}
        del r2
        # the two-pass schedule (stand-alone gn_apply in front of every conv) in the same dtype: same results bit for bit,
        # 13 more launches and ~3 GB more DRAM traffic per step; its conv kernels carry no prologue work, so their
        # TFLOP/s is the number to compare with round 1's roofline.frac
        if getattr(runner.eng, "fused_prologue", False):
            _, _, r3 = make_runner(compute, fused=False)
            r3.run_steps(6)
            torch.cuda.synchronize(dev)
            ms3, _, _ = timed_steps(r3, max(K, 20))
            conv3, _ = time_conv_kernels(r3, dev, reps=3)
            tf3 = CONV_FLOP_PER_FORWARD * NB / (conv3 / 1000.0) / 1e12
            roofline["two_pass_schedule"] = {"value": world * max(K, 20) * NB / (ms3 / 1000.0), "ms_per_step": ms3 / max(K, 20),
                                             "conv_ms_per_step": conv3, "achieved": tf3, "frac": tf3 / peaks["bf16_tflops_sustained"]}
            del r3

    # ---- e2e: the public API with HOST buffers (pinned), H2D of image + mask and D2H of mean/std/samples inside the timed region
    e2e = None
    if not args.no_e2e:
        T_e2e = args.e2e_iters
        ev.num_iterations = T_e2e
        im_h = synthetic.make_image(H0, W0, seed=1234).pin_memory()
        fov_h = synthetic.make_fov_mask(H0, W0).pin_memory()
        outs_h = [torch.empty(1, 1, H0, W0).pin_memory(), torch.empty(1, 1, H0, W0).pin_memory(),
                  torch.empty(25, 1, 1, H0, W0).pin_memory()]

        def e2e_once(evaluator):
            torch.manual_seed(seed)
            im_d = im_h.to(dev, non_blocking=True)
            fov_d = fov_h.to(dev, non_blocking=True)
            _, (mean, std, tens) = evaluator.predict_step((im_d, None, fov_d), 0)
            outs_h[0].copy_(mean, non_blocking=True)
            outs_h[1].copy_(std, non_blocking=True)
            outs_h[2].copy_(tens, non_blocking=True)
            torch.cuda.synchronize(dev)

        def e2e_time(evaluator, reps):
            e2e_once(evaluator)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                e2e_once(evaluator)
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item()) / reps

        sec = e2e_time(ev, 2)
        # checksum of the result (same seed on every N): the 25 saved samples are independent forwards -> bit-identical for
        # every world size; mean / std are fp64 sums in a sharding-dependent order -> equal to ~1e-7, printed to 5 digits
        checksum = {"samples_sha256": hashlib.sha256(outs_h[2].numpy().tobytes()).hexdigest()[:16],
                    "mean_sum": round(float(outs_h[0].double().sum()), 2), "std_sum": round(float(outs_h[1].double().sum()), 2),
                    "note": "must not depend on --gpus (iteration t always consumes Philox window t)"}
        # under torch.distributed predict_step shards T_e2e over the ranks, so T_e2e passes complete per call
        e2e = {"value": T_e2e / sec, "unit": UNIT,
               "h2d_bytes_per_step": im_h.numel() * 4 + fov_h.numel() * 4,
               "d2h_bytes_per_step": sum(o.numel() * 4 for o in outs_h),
               "step": f"one DropBlockEval.predict_step call = {T_e2e} iterations (BASELINE configs[2] iter_num; sharded over ranks), pinned host in/out, bytes per call",
               "mc_1000_iter_projected_s": 1000.0 / (T_e2e / sec), "seconds_per_call": sec, "checksum": checksum,
               "allreduce_ms": allreduce_ms}
        if not args.no_rotation:
            # BASELINE configs[3]: rotation ensemble, angles 1..359 sharded over the ranks, same pinned host in/out
            m_eval, _ = build_canonical(dev, dropblock=False, compute=compute)
            rv = U.RotationEval(m_eval, num_iterations=359, return_num=25)
            sec_r = e2e_time(rv, 2)
            e2e["rotation_ensemble"] = {"metric": "rotation-ensemble fwd passes/s at 584x565 (angles 1..359, rotate-in -> eval forward -> rotate-back, per-pixel mean/std)",
                                        "value": 359 / sec_r, "unit": UNIT, "seconds_per_call": sec_r, "angles": 359,
                                        "angle_batch": rv.angle_batch, "cuda_graph": True}
            if not args.no_sweep:
                e2e["sweep"] = bench_sweep(dev, m_eval, rank, world)
            del rv, m_eval

    train = None
    if not args.no_train:
        train = bench_train(dev, rank, world, args.train_steps, 4, peaks)
        if e2e is not None:
            e2e["train"] = {k: train[k] for k in ("metric", "value", "unit", "ms_per_step", "global_batch", "allreduce")}
            if not args.no_sweep:
                e2e["train"]["sweep"] = bench_train_sweep(dev, rank, world)
        roofline["train"] = train.get("roofline")

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        val, k, w, cores, kind, others = oracle_mc_forward_timer(6, 0, budget_s=25.0, extras=True)   # ~15 s of CPU work (+ ~4 s of other legs)
        cpu_baseline = {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"{k} MC-DropBlock forward passes of the same 584x565 image "
                                  f"({'unmodified reference modules' if kind == 'reference' else 'oracle port'}, torch fp32, all host threads) after 1 warm-up",
                        "other_legs": others}
        if not args.no_libbar:
            cpu_baseline["library_bar"] = library_bar(dev)

    if rank == 0:
        cfg = workload_config(world, NB)
        cfg.update({"iter_batch": NB, "cuda_graph": True, "mask_build": "side stream, overlapped with the forward of the previous step",
                    "fused_prologue": bool(getattr(runner.eng, "fused_prologue", False)),
                    "fused_levels": sorted(getattr(runner.eng, "fuse_levels", [])) if getattr(runner.eng, "fused_prologue", False) else [],
                    "epilogue": "TMA stores (conv3x3: 32-channel runs per warp; convT: 5-D pixel-shuffle box)"})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype_name,
            "data": "synthetic", "config": cfg,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * K, "launches_per_step": launches_per_step,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "train": train, "alt_dtype": alt,
            "tflops_whole_step": FLOP_PER_FORWARD * NB * world * K / (ms_max / 1000.0) / 1e12,
            "mc_1000_iter_projected_s": 1000.0 / value, "allreduce_ms": allreduce_ms,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_sweep(dev, model, rank, world):
    """BASELINE configs[4], inference part: batched eval forwards at the multi-fidelity sizes (MF-training-UNI.py:33-44)
    through the public `UNet.forward` with pinned host input and output (H2D + D2H inside the timed region); every rank
    runs the same batch (weak scaling), imgs/s summed over ranks."""
    import torch
    import torch.distributed as dist
    from unet_research_b200 import synthetic
    out = []
    for (h, w, n) in ((146, 141, 16), (292, 283, 8), (584, 565, 4), (128, 128, 16), (256, 256, 8), (584, 584, 4)):
        xh = synthetic.make_image(h, w, seed=7, batch=n).pin_memory()
        yh = torch.empty(n, 1, h, w).pin_memory()

        def once():
            with torch.no_grad():
                yh.copy_(model(xh.to(dev, non_blocking=True)), non_blocking=True)
            torch.cuda.synchronize(dev)

        for _ in range(4):
            once()
        if world > 1:
            dist.barrier()
        reps = 10
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        out.append({"size": f"{h}x{w}", "batch_per_gpu": n, "imgs_per_s": world * n * reps / float(tt.item()),
                    "ms_per_batch": float(tt.item()) / reps * 1e3})
    return out


def library_bar(dev):
    """The "library bar" of SURVEY 8(d): the reference ALGORITHM (the oracle's torch ops: cuDNN convs, ATen GroupNorm /
    max-pool / rand) on the SAME B200, timed for the same workloads -- one MC-DropBlock forward at batch 10 and one training
    step.  A reported baseline (what a user gets by moving the reference to the GPU unchanged), not a product path."""
    import torch
    import torch.nn.functional as F
    from oracle import unet_oracle as O
    from unet_research_b200 import synthetic
    sd = {k: v.to(dev) for k, v in synthetic.make_state_dict(seed=1234).items()}
    db = O.DropBlockCfg(0.15, 7, True)
    prev = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.benchmark = True
    out = {"what": "torch 2.11 / cuDNN running the reference algorithm on the same GPU (oracle ops), CUDA-event timed"}

    def timeit(fn, reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(dev)
        return _event_time(dev, lambda: [fn() for _ in range(reps)]) / reps

    try:
        x10 = synthetic.make_image(H0, W0, seed=1234).to(dev).expand(10, -1, -1, -1).contiguous()
        for name, tf32, ac in (("tf32", True, None), ("bf16_autocast", False, torch.bfloat16)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32

            def fwd():
                with torch.no_grad():
                    if ac is None:
                        return O.unet_forward(sd, x10, dropblock=db)
                    with torch.autocast("cuda", dtype=ac):
                        return O.unet_forward(sd, x10, dropblock=db)

            ms = timeit(fwd, 3)
            out[f"mc_passes_per_s_batch10_{name}"] = 10.0 / ms * 1e3
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        x = synthetic.make_image(H0, W0, seed=1234).to(dev)
        gt = synthetic.make_gt(H0, W0).to(dev)
        fov = synthetic.make_fov_mask(H0, W0).to(dev)

        def step():
            for p_ in params.values():
                p_.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16):
                seg = O.unet_forward(params, x, dropblock=db)
            loss = F.binary_cross_entropy((seg.float() * fov).clamp(0, 1), gt * fov) * (fov.numel() / fov.sum())
            loss.backward()

        out["train_imgs_per_s_bf16_autocast"] = 1e3 / timeit(step, 3)
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
        torch.cuda.empty_cache()
    return out


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


def elementwise_bytes_per_step(nb, filters=64, depth=4, h=592, w=576, h0=H0, w0=W0, fused=False):
    fuse_levels = set(range(depth + 1)) if fused is True else (set() if not fused else set(fused))
    return _elementwise_bytes(nb, filters, depth, h, w, h0, w0, fuse_levels)


def _elementwise_bytes(nb, filters, depth, h, w, h0, w0, fuse_levels):
    """ALGORITHMIC HBM bytes of one Monte-Carlo step (nb batched iterations, 16-bit activations) per fused kernel
    family, from the tensor shapes alone: every activation read once / written once, bit masks
    1 bit per element, statistics and coefficients ignored.  fuse_levels: the resolutions (0 = full size) whose conv3x3
    prologue applies GroupNorm + DropBlock + ReLU (the stand-alone applies in front of those 3x3 convs are gone)."""
    out = {"b2u_gn_apply": 0.0, "b2u_gn_apply_pool": 0.0, "b2u_head_fwd": 0.0, "b2u_conv_first_fwd": 0.0,
           "b2u_dropblock_dilate": 0.0}
    c = filters
    bits = 0.0
    for lvl in range(depth + 1):
        e = float(h >> lvl) * (w >> lvl) * c                     # elements of one conv output at this level
        fused = lvl in fuse_levels
        if lvl < depth:
            # encoder: unit 1 apply (first level reads the SHARED raw tensor once), unit 2 apply+pool, pooled-GN apply
            if not fused:
                out["b2u_gn_apply"] += (2.0 * e * (1 if lvl == 0 else nb) + 2.0 * e * nb + e * nb / 8.0)
            out["b2u_gn_apply_pool"] += nb * (2.0 * e + 2.0 * e + 0.5 * e + e / 8.0 + e / 8.0)
            if (lvl + 1) not in fuse_levels:                     # the pooled tensor's apply belongs to the NEXT level's first conv
                out["b2u_gn_apply"] += nb * (2.0 * e / 4 + 2.0 * e / 4)
            # decoder at the same resolution: up-conv apply (with concat mask), unit 1 apply, unit 2 apply (not the last level)
            out["b2u_gn_apply"] += nb * (4.0 * e + e / 8.0)
            if not fused:
                out["b2u_gn_apply"] += nb * (4.0 * e + e / 8.0)
            if lvl > 0:
                out["b2u_gn_apply"] += nb * (4.0 * e + e / 8.0)
            bits += nb * (2 * e + 2 * e + 2 * e)                 # sites: 2 encoder units, concat (2c), 2 decoder units
        else:
            out["b2u_gn_apply"] += nb * (1 if fused else 2) * (4.0 * e + e / 8.0)
            bits += nb * 2 * e
        c *= 2
    e0 = float(h) * w * filters
    out["b2u_head_fwd"] = nb * (2.0 * e0 + e0 / 8.0) + 16.0 * h0 * w0 * 2
    out["b2u_conv_first_fwd"] = 2.0 * e0 + 4.0 * h0 * w0
    out["b2u_dropblock_dilate"] = 2.0 * bits / 8.0                # centre bitmap in, keep bitmap out
    return out


def bench_train_sweep(dev, rank, world, steps=20):
    """BASELINE configs[4], training part: the same step (fwd + bwd + clip + SGD, batch 1 per GPU, DropBlock on, bf16, CUDA
    graphs, NCCL gradient averaging when N > 1) at the multi-fidelity training sizes (MF-training-UNI.py:33-44: 128, 256 and the
    584 x 584 padded original)."""
    out = []
    for size in (128, 256, 584):
        t = bench_train(dev, rank, world, steps, 4, None, hw=(size, size))
        out.append({"size": f"{size}x{size}", "ms_per_step": t["ms_per_step"], "imgs_per_s": t["value"]})
    return out


def bench_train(dev, rank, world, steps, warmup, peaks=None, hw=None):
    """BASELINE configs[1] (+ configs[4] data-parallel part): fwd + bwd + clip + SGD, batch 1 per GPU, 584x565,
    DropBlock bs 7 p .15, bf16; with N > 1 every rank trains on its own image and the gradients are averaged over
    NCCL inside backward (decoder-side bucket overlapped with the encoder backward).  Returns the `train` object."""
    import torch
    import torch.distributed as dist
    from torch import nn
    import unet_research_b200 as U
    from unet_research_b200 import synthetic
    from unet_research_b200.canonical import build_canonical
    H0, W0 = hw if hw is not None else (globals()["H0"], globals()["W0"])
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
    roof = None
    if peaks is not None and world == 1:
        roof = train_roofline(model, step, dev, ms, peaks)
    if world > 1:
        dist.barrier()
    return {"roofline": roof, "metric": "train imgs/s (fwd + bwd + clip + SGD, batch 1 per GPU, 584x565, DropBlock bs7 p0.15, bf16)",
            "value": world / (ms / 1000.0), "unit": "imgs/s", "ms_per_step": ms, "steps": steps, "global_batch": world,
            "tflops": 1501.4e9 * world / (ms / 1000.0) / 1e12, "flop_per_image": 1501.4e9, "cuda_graphs": graphs,
            "optimizer": "fused SGD-momentum + global-norm clip (one multi-tensor kernel pair)" if fused else "torch.optim.SGD + clip_grad_norm_",
            "gradient_allreduce_bytes": (124.16e6 if world > 1 else 0), "allreduce": "NCCL AVG, 3 buckets of the flat gradient buffer (decoder + bottleneck | deep encoder | shallow encoder), each reduced while the next backward graph segment runs" if world > 1 else None,
            "loss_first": first, "loss_last": float(loss.item())}


def train_roofline(model, step, dev, step_ms, peaks):
    """Per-entry-point CUDA-event timing of ONE eager training step (graphs off, every launch queued behind a spin
    kernel): tensor-core families (forward convs, data gradients, weight gradients) in TFLOP/s against the measured
    sustained bf16 peak, everything else as milliseconds.  Algorithmic flop per image: 500.03 G per family."""
    import torch
    import unet_research_b200.backward as BW
    import unet_research_b200.engine as E
    import unet_research_b200.optim as OP
    from unet_research_b200 import _lib
    records = []
    orig = _lib.call
    phase = ["fwd"]

    def timed_call(name, *a):
        if name in _lib._LAUNCHERS:
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            orig(name, *a)
            e_.record()
            records.append((phase[0] + ":" + name, s_, e_))
        else:
            orig(name, *a)

    def bw_call(name, *a):
        phase[0] = "bwd"
        timed_call(name, *a)

    def op_call(name, *a):
        phase[0] = "opt"
        timed_call(name, *a)

    prev_graph = model.use_cuda_graph
    model.use_cuda_graph = False
    E.call, BW.call, OP.call = timed_call, bw_call, op_call
    try:
        reps = 3
        for _ in range(reps):
            phase[0] = "fwd"
            torch.cuda._sleep(40_000_000)
            step()
        torch.cuda.synchronize(dev)
    finally:
        E.call, BW.call, OP.call = orig, orig, orig
        model.use_cuda_graph = prev_graph
    tot = {}
    for name, s_, e_ in records:
        tot[name] = tot.get(name, 0.0) + s_.elapsed_time(e_) / reps
    fam = {"forward convs (conv3x3 + convT)": ("fwd:b2u_conv3x3_fwd", "fwd:b2u_convT2x2_fwd"),
           "data gradients (conv3x3 with rotated weights + 1x1 GEMM of the convT)": ("bwd:b2u_conv3x3_fwd", "bwd:b2u_gemm1x1_fwd"),
           "weight gradients (wgrad + split-K reduce)": ("bwd:b2u_wgrad",)}
    out = {"unit": "TFLOP/s", "peak": peaks["bf16_tflops_sustained"], "families": {}, "elementwise_ms": {}}
    tensor_ms = 0.0
    for label, keys in fam.items():
        ms = sum(tot.get(k, 0.0) for k in keys)
        tensor_ms += ms
        if ms > 0:
            tf = CONV_FLOP_PER_FORWARD / (ms / 1000.0) / 1e12
            out["families"][label] = {"ms_per_step": ms, "achieved": tf, "frac": tf / peaks["bf16_tflops_sustained"]}
    used = {k for keys in fam.values() for k in keys}
    out["elementwise_ms"] = {k: v for k, v in sorted(tot.items()) if k not in used}
    out["tensor_ms_per_step"] = tensor_ms
    out["whole_step"] = {"ms_per_step": step_ms, "achieved": 1501.4e9 / (step_ms / 1000.0) / 1e12,
                         "frac": 1501.4e9 / (step_ms / 1000.0) / 1e12 / peaks["bf16_tflops_sustained"]}
    return out


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
    conv = per_step.pop("b2u_conv3x3_fwd", 0.0) + per_step.pop("b2u_conv3x3_pro_fwd", 0.0) + per_step.pop("b2u_convT2x2_fwd", 0.0)
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
    ap.add_argument("--dtype", default="auto", choices=["auto", "fp16", "bf16"],
                    help="operand format of the MC inference path (auto = the package default for inference: fp16)")
    ap.add_argument("--no-alt", action="store_true")
    ap.add_argument("--no-rotation", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-libbar", action="store_true")
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
