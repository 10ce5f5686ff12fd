"""DIAGNOSTIC (not a test): a home-made initcheck (compute-sanitizer is closed on the GPU pool).  Runs the eager Monte-Carlo
DropBlock call twice with the same seed; before the second run EVERY workspace buffer (activations, statistics partials,
coefficients) and the mask / scatter bitmaps are filled with 0xFF bytes (NaN in every float format, all-ones masks).  A kernel
that reads a byte no kernel of the same call wrote first shows up as different sample bits or NaNs.
    python tests/exp_poison.py
"""
import dataclasses
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

import gpu_diag as D
import unet_research_b200 as U
from unet_research_b200 import synthetic


def tensors_of(obj, seen, out, depth=0):
    if id(obj) in seen or depth > 4:
        return
    seen.add(id(obj))
    if isinstance(obj, torch.Tensor):
        out.append(obj)
    elif isinstance(obj, dict):
        for v in obj.values():
            tensors_of(v, seen, out, depth + 1)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            tensors_of(v, seen, out, depth + 1)
    elif dataclasses.is_dataclass(obj):
        for f in dataclasses.fields(obj):
            tensors_of(getattr(obj, f.name), seen, out, depth + 1)


def main():
    dev = torch.device("cuda")
    m, _ = D._build_model(dev, dropblock=True)
    ok = True
    for h, w, nb, T in [(120, 116, 2, 4), (584, 565, 5, 10)]:
        x = synthetic.make_image(h, w, seed=1234).to(dev)
        fov = synthetic.make_fov_mask(h, w).to(dev)
        ev = U.DropBlockEval(m, num_iterations=T, return_num=3, iter_batch=nb, use_cuda_graph=False)
        outs = []
        for rep in range(2):
            if rep == 1:
                ts = []
                for r in ev._runners.values():
                    seen = set()
                    for name in ("buf", "stat", "out", "logits"):
                        tensors_of(getattr(r.ws, name, None), seen, ts)
                    for mp in r.masks:
                        ts += [mp.mask_bits] + ([mp.scatter_bits] if mp.scatter_bits is not None else [])
                nbytes = 0
                for t in ts:
                    t.view(torch.uint8).fill_(0xFF)
                    nbytes += t.numel() * t.element_size()
                print(f"  poisoned {len(ts)} buffers, {nbytes / 1e6:.1f} MB")
            torch.manual_seed(5)
            _, (mean, std, tens) = ev.predict_step((x, None, fov), 0)
            torch.cuda.synchronize()
            outs.append((mean.clone(), std.clone(), tens.clone()))
        same = all(torch.equal(a, b) for a, b in zip(*outs))
        finite = bool(torch.isfinite(outs[1][0]).all() and torch.isfinite(outs[1][2]).all())
        ok &= same and finite
        print(flush=True, end=""); print(f"  {h}x{w} batch {nb} T {T}: poisoned run {'bit-identical' if same else 'DIFFERENT'}, finite {finite}")
    print("exp_poison:", "OK" if ok else "UNINITIALISED READ SUSPECTED")


main()
