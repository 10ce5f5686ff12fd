"""EXPERIMENT (not a test): the 1x1 head + sigmoid + fp64 accumulate kernel at the Monte-Carlo shape (10 x 592 x 576 x 64,
crop 584 x 565): cp.async ring variant against the register variant, block sizes 256 / 224 / 192 (fill of the last trip).
Bit-compares every variant's outputs with the register kernel at 256 threads, then times them with CUDA events.
    python tests/exp_head.py [reps]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from unet_research_b200 import _lib
from unet_research_b200._lib import HeadDesc, call, ptr, stream_ptr


def run(x, coef, mask, wh, fov, hd, n, h0, w0, want_out):
    dev = x.device
    out = torch.zeros(n, 1, h0, w0, device=dev) if want_out else None
    logits = torch.zeros(n, 1, h0, w0, device=dev) if want_out else None
    acc = torch.zeros(2, h0, w0, dtype=torch.float64, device=dev)
    samples = torch.zeros(25, h0, w0, device=dev)
    it = torch.zeros(1, dtype=torch.int64, device=dev)
    call("b2u_head_fwd", ptr(x), ptr(coef), ptr(mask) if mask is not None else None, ptr(wh), ptr(out) if want_out else None,
         ptr(logits) if want_out else None, ptr(fov), ptr(acc), ptr(samples), ptr(it), C.byref(hd), stream_ptr())
    torch.cuda.synchronize()
    return out, logits, acc, samples


def setenv(a, nt, st):
    """a = -1: the library's defaults (no override)."""
    for k, v in (("B2U_HEAD_ASYNC", a), ("B2U_HEAD_THREADS", nt), ("B2U_HEAD_STAGES", st)):
        if a < 0:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(5)
    variants = [(0, 256, 3), (-1, 0, 0), (1, 256, 2), (1, 224, 2), (1, 192, 2), (1, 256, 3), (1, 224, 3), (1, 192, 3), (1, 160, 3), (1, 192, 4)]
    ok = True
    for (n, h, w, h0, w0, dt, use_mask, want_out) in [(3, 32, 48, 29, 45, torch.bfloat16, False, True),
                                                      (10, 64, 64, 59, 61, torch.float16, True, True),
                                                      (5, 152, 144, 146, 141, torch.float16, True, False),
                                                      (10, 592, 576, 584, 565, torch.float16, True, False)]:
        c = 64
        x = torch.randn(n, h, w, c, generator=g).to(dev).to(dt)
        coef = torch.stack([1 + 0.1 * torch.randn(n, c, generator=g), 0.1 * torch.randn(n, c, generator=g)], -1).to(dev).contiguous()
        wh = (torch.randn(c, generator=g) / 8).to(dev)
        fov = (torch.rand(h0, w0, generator=g) > 0.3).float().to(dev)
        mask = (torch.rand(n, h, w, 8, generator=g) * 256).to(torch.uint8).to(dev) if use_mask else None
        hd = HeadDesc()
        hd.n, hd.h, hd.w, hd.c, hd.h0, hd.w0 = n, h, w, c, h0, w0
        hd.dtype, hd.return_num = (_lib.BF16 if dt == torch.bfloat16 else _lib.F16), 25
        ref = None
        for a, nt, st in variants:
            setenv(a, nt, st)
            res = run(x, coef, mask, wh, fov, hd, n, h0, w0, want_out)
            if ref is None:
                ref = res
                continue
            same = all((r is None and q is None) or torch.equal(r, q) for r, q in zip(ref, res))
            ok &= same
            print(f"  {n}x{h}x{w} mask={use_mask} out={want_out} async={a} threads={nt} stages={st}: {'bit-identical' if same else 'DIFFERENT'}")
        if (h, w) != (592, 576):
            continue
        acc = torch.zeros(2, h0, w0, dtype=torch.float64, device=dev)
        samples = torch.zeros(25, h0, w0, device=dev)
        it = torch.zeros(1, dtype=torch.int64, device=dev)
        algo = n * h0 * w0 * (c * 2 + 8) + h0 * w0 * (4 + 32)             # activations + masks; fov + fp64 accumulator RMW
        for a, nt, st in variants + variants:
            setenv(a, nt, st)
            ts = []
            for r in range(reps + 3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):                         # the 436 MB of activations exceed L2: no flush needed between launches
                    call("b2u_head_fwd", ptr(x), ptr(coef), ptr(mask), ptr(wh), None, None, ptr(fov), ptr(acc), ptr(samples), ptr(it),
                         C.byref(hd), stream_ptr())
                e1.record()
                torch.cuda.synchronize()
                if r >= 3:
                    ts.append(e0.elapsed_time(e1) * 100.0)
            ts.sort()
            med = ts[len(ts) // 2]
            print(f"  async={a} threads={nt} stages={st}: median {med:.1f} us  min {ts[0]:.1f} us  {algo / med / 1e3:.0f} GB/s (algorithmic {algo / 1e6:.1f} MB)")
    setenv(-1, 0, 0)
    print("exp_head:", "OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
