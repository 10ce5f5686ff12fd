"""Reproducibility hunt (diagnostic): the same Monte-Carlo run (584x565, T iterations, batch ib, fixed seed) repeated with a
fresh DropBlockEval each time; prints how many sample / mean elements differ from the first repetition and where.
    python tests/exp_repro.py [reps] [ib] [T] [overlap 0/1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

import gpu_diag as D
import unet_research_b200 as U
from unet_research_b200 import synthetic

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ib = int(sys.argv[2]) if len(sys.argv) > 2 else 5
T = int(sys.argv[3]) if len(sys.argv) > 3 else 10
overlap = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
dev = torch.device("cuda")
m, _ = D._build_model(dev, dropblock=True)
x = synthetic.make_image(584, 565, seed=1234).to(dev)
fov = synthetic.make_fov_mask(584, 565).to(dev)
first = None
bad = 0
for r in range(reps):
    ev = U.DropBlockEval(m, num_iterations=T, return_num=min(T, 10), iter_batch=ib, overlap_masks=overlap)
    torch.manual_seed(5)
    _, (mean, std, tens) = ev.predict_step((x, None, fov), 0)
    torch.cuda.synchronize()
    if first is None:
        first = (mean.clone(), tens.clone())
        continue
    dm = (mean != first[0])
    dt = (tens != first[1])
    if int(dt.sum()) or int(dm.sum()):
        bad += 1
        idx = dt.nonzero()
        per_sample = [int(dt[i].sum()) for i in range(dt.shape[0])]
        msg = f"  rep {r}: mean differs at {int(dm.sum())} px, samples differ at {int(dt.sum())} px; per sample {per_sample}"
        if idx.numel():
            ys, xs = idx[:, 3], idx[:, 4]
            msg += f"; rows {int(ys.min())}..{int(ys.max())} cols {int(xs.min())}..{int(xs.max())}; max |d| {float((tens - first[1]).abs().max()):.3e}"
        print(msg, flush=True)
cfg = {k: v for k, v in os.environ.items() if k.startswith("B2U_")}
print(f"repro ib={ib} T={T} overlap={overlap} env={cfg}: {bad} of {reps - 1} repetitions differ from the first", flush=True)
