"""Tile-shape sweep for the deep conv layers (diagnostic):  python tests/exp_conv_tiles.py [batch]
BLOCK_N 256 / MT 1 (the plan's choice for Cout % 256 == 0) against BLOCK_N 128 / MT 2 and 256 / MT 2."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_research_b200 import _lib
from unet_research_b200._lib import ConvDesc, call, ptr, stream_ptr

dev = torch.device("cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
shapes = [(148, 144, 128, 256), (148, 144, 256, 256), (148, 144, 512, 256), (74, 72, 256, 512), (74, 72, 512, 512), (74, 72, 1024, 512),
          (37, 36, 512, 1024), (37, 36, 1024, 1024)]
tot = {}
for (h, w, cin, cout) in shapes:
    x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
    wp = torch.randn(9, cout, cin, device=dev).to(torch.bfloat16)
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=dev)
    flop = 2.0 * n * h * w * cout * 9 * cin
    line = f"  {h}x{w} {cin}->{cout}:"
    for bn, mt in ((256, 1), (128, 2), (128, 1), (256, 2)):
        d = ConvDesc()
        d.n, d.h, d.w, d.cin, d.cout, d.dtype, d.num_groups, d.x_cstride = n, h, w, cin, cout, _lib.BF16, 32, cin
        d.reserved[0], d.reserved[3] = bn, mt
        rows, sgs = C.c_int(0), C.c_int(0)
        call("b2u_conv3x3_stat_layout", C.byref(d), C.byref(rows), C.byref(sgs))
        parts = torch.empty(n, rows.value, cout // sgs.value, 2, dtype=torch.float32, device=dev)
        for _ in range(2):
            call("b2u_conv3x3_fwd", ptr(x), ptr(wp), ptr(y), ptr(parts), C.byref(d), stream_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call("b2u_conv3x3_fwd", ptr(x), ptr(wp), ptr(y), ptr(parts), C.byref(d), stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tot[(bn, mt)] = tot.get((bn, mt), 0.0) + ms
        line += f"  bn{bn} mt{mt} {ms * 1000:.0f}us {flop / ms / 1e9:.0f}TF"
    print(line, flush=True)
print("  totals (ms):", {f"bn{k[0]} mt{k[1]}": round(v, 3) for k, v in tot.items()})
