"""Where does the Monte-Carlo step lose time to the side-stream mask build?  (diagnostic, not a test)

    python tests/exp_overlap.py [iter_batch]

Captures CUDA graphs of one step with subsets of the kernels enabled (a filter on the C-ABI call names: skipped kernels
leave garbage in their outputs, which is irrelevant for the timing of the others) and times 20 replays of each:
forward alone, mask build alone, conv kernels alone / next to the mask build, elementwise kernels alone / next to it.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import unet_research_b200 as U
from unet_research_b200 import _lib, engine, synthetic, uncertainty
from unet_research_b200.canonical import build_canonical

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda")
H0, W0 = 584, 565
model, _ = build_canonical(dev, dropblock=True, compute="bf16")
model.apply(U.set_dropblock_on)
x = synthetic.make_image(H0, W0, seed=1234).to(dev)
fov = synthetic.make_fov_mask(H0, W0).to(dev).reshape(H0, W0).contiguous()
ev = U.DropBlockEval(model, num_iterations=1000, return_num=25, iter_batch=nb)
runner = ev._runner(nb, H0, W0, dev, True, 0.15, 7)
runner.begin(x, fov, 0, 1234, 0)
runner.run_steps(2)                      # eager warm-up (sets kernel attributes)
torch.cuda.synchronize()



def time_graph(g, reps=20, steps_per_graph=1):
    with torch.cuda.stream(runner.main):
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / steps_per_graph


# ---- where to fork the mask build of the next step (engine.forward hook points)
for fp in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["enc0", "enc1", "enc2", "enc3", "bottleneck", "dec0", "dec1", "dec2"]):
    runner.fork_point = fp
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=runner.main):
        runner._pair()
    print(f"fork at {fp:12s} {time_graph(g, 20, 2):8.3f} ms / step of {nb}", flush=True)
runner.fork_point = "enc0"
if len(sys.argv) > 3 and sys.argv[3] == "forkonly":
    sys.exit(0)

CONV = {"b2u_conv3x3_fwd", "b2u_convT2x2_fwd"}
GEN = {"b2u_dropblock_centers", "b2u_dropblock_dilate"}
allow = None
real_call = _lib.call


def filtered(name, *args):
    if allow == "ELT":
        ok = name not in CONV and name not in GEN
    else:
        ok = allow is None or name in allow
    if ok:
        real_call(name, *args)


for mod in (engine, uncertainty):
    mod.call = filtered


def capture(fwd: bool, gen: bool, sel):
    """One step: forward(masks 0) on main, build(masks 1) on side, with the kernel filter `sel`."""
    global allow
    allow = sel
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=runner.main):
        main = torch.cuda.current_stream(dev)
        if gen:
            runner.side.wait_stream(main)
            with torch.cuda.stream(runner.side):
                runner.masks[1].generate(runner.seed)
        if fwd:
            runner._forward(0)
        if gen:
            main.wait_stream(runner.side)
    allow = None
    return g


def timeit(g, reps=20):
    with torch.cuda.stream(runner.main):
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


cases = [
    ("forward + mask build (the bench step)", True, True, None),
    ("forward alone", True, False, None),
    ("mask build alone", False, True, None),
    ("centers alone", False, True, {"b2u_dropblock_centers"}),
    ("dilate alone", False, True, {"b2u_dropblock_dilate"}),
    ("conv kernels alone", True, False, CONV),
    ("conv kernels || mask build", True, True, CONV | GEN),
    ("conv kernels || centers", True, True, CONV | {"b2u_dropblock_centers"}),
    ("elementwise kernels alone", True, False, "ELT"),
]
res = {}
for name, fwd, gen, sel in cases:
    g = capture(fwd, gen, sel)
    res[name] = timeit(g)
    print(f"{name:45s} {res[name]:8.3f} ms / step of {nb}", flush=True)

# elementwise || mask build needs both classes: filter = everything except conv
allow_backup = None


def filtered2(name, *args):
    if name not in CONV:
        real_call(name, *args)


for mod in (engine, uncertainty):
    mod.call = filtered2
g = capture(True, True, None)
print(f"{'elementwise kernels || mask build':45s} {timeit(g):8.3f} ms / step of {nb}", flush=True)
