"""One eager Monte-Carlo step (mask build + batched forward) for ncu:  python tests/prof_step.py [iter_batch] [reps] [dtype]

Prints nothing interesting; meant to be run under
    ncu --set full --clock-control none --profile-from-start off ...   (captures the last repetition only)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import unet_research_b200 as U
from unet_research_b200 import _lib, synthetic
from unet_research_b200.canonical import build_canonical

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 5
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda")
H0, W0 = 584, 565
model, _ = build_canonical(dev, dropblock=True, compute=sys.argv[3] if len(sys.argv) > 3 else "auto")
model.apply(U.set_dropblock_on)
x = synthetic.make_image(H0, W0, seed=1234).to(dev)
fov = synthetic.make_fov_mask(H0, W0).to(dev).reshape(H0, W0).contiguous()
ev = U.DropBlockEval(model, num_iterations=1000, return_num=25, iter_batch=nb)
runner = ev._runner(nb, H0, W0, dev, True, 0.15, 7)
runner.begin(x, fov, 0, 1234, 0)
torch.cuda.synchronize()
for r in range(reps):
    l0 = _lib.launch_count
    if r == reps - 1:
        torch.cuda.profiler.start()            # ncu --profile-from-start off: only the last repetition is captured
    runner.masks[0].generate(runner.seed)
    runner.eng.forward(runner.x, runner.ws, runner.masks[0], head_out=False, mc=runner.mc, shared_input=True)
    torch.cuda.synchronize()
    print(f"rep {r}: {_lib.launch_count - l0} b2u launches", flush=True)
torch.cuda.profiler.stop()
