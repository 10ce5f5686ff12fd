"""Timeline of one overlapped Monte-Carlo step pair (diagnostic): every C-ABI launch of the eager `_pair()` is bracketed by
CUDA events on ITS stream (main = forward, side = mask build of the next step); prints when each kernel ran relative to the
start of the pair, so the overlap between the two streams -- and what it costs the forward kernels -- is visible.
   python tests/exp_timeline.py [iter_batch] [dtype]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import unet_research_b200 as U
from unet_research_b200 import _lib, synthetic
from unet_research_b200.canonical import build_canonical
import unet_research_b200.engine as E
import unet_research_b200.uncertainty as UN

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 10
compute = sys.argv[2] if len(sys.argv) > 2 else "bf16"
dev = torch.device("cuda")
model, _ = build_canonical(dev, dropblock=True, compute=compute)
model.apply(U.set_dropblock_on)
ev = U.DropBlockEval(model, num_iterations=1000, return_num=25, iter_batch=nb, use_cuda_graph=False)
r = ev._runner(nb, 584, 565, dev, True, 0.15, 7)
x = synthetic.make_image(584, 565, seed=1234).to(dev)
fov = synthetic.make_fov_mask(584, 565).to(dev).reshape(584, 565).contiguous()
r.begin(x, fov, 0, 1234, 0)
r.run_steps(4)
torch.cuda.synchronize()

records = []
orig = _lib.call


def timed(name, *a):
    if name in _lib._LAUNCHERS:
        st = torch.cuda.current_stream()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(st)
        orig(name, *a)
        e.record(st)
        records.append((name, "side" if st == r.side else "main", s, e))
    else:
        orig(name, *a)


E.call = timed
UN.call = timed
t0 = torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(r.main):
    torch.cuda._sleep(60_000_000)          # queue the whole pair behind a spin kernel: no host launch latency in the picture
    t0.record(r.main)
    r._pair()
    tend = torch.cuda.Event(enable_timing=True)
    tend.record(r.main)
torch.cuda.synchronize()
E.call = orig
UN.call = orig
print(f"pair (2 steps of {nb} iterations, eager, events on every launch): {t0.elapsed_time(tend):.3f} ms")
rows = [(t0.elapsed_time(s), t0.elapsed_time(e), name, strm) for (name, strm, s, e) in records]
rows.sort()
side = [(a, b, n) for (a, b, n, s) in rows if s == "side"]
for (a, b, n) in side:
    print(f"  side {n:32s} {a:8.3f} -> {b:8.3f}  ({b - a:6.3f} ms)")
tot = {}
for (a, b, n, s) in rows:
    if s == "main":
        ov = sum(max(0.0, min(b, sb) - max(a, sa)) for (sa, sb, _) in side)
        k = (n, "overlapped" if ov > 0.5 * (b - a) else "alone")
        tot.setdefault(k, [0, 0.0])
        tot[k][0] += 1
        tot[k][1] += b - a
print("  main-stream kernels, summed over the pair, split by whether the mask build was running next to them:")
for (n, k), (c, t) in sorted(tot.items()):
    print(f"    {n:28s} {k:10s} launches {c:3d}  total {t:7.3f} ms  avg {t / c * 1000:7.1f} us")
# first step of the pair in detail
print("  main stream, first step:")
first_end = None
cnt = 0
for (a, b, n, s) in rows:
    if s != "main":
        continue
    cnt += 1
    print(f"    {a:8.3f} -> {b:8.3f} {b - a:7.3f}  {n}")
    if n == "b2u_head_fwd":
        break
