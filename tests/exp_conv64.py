"""What bounds the short-K conv layers?  (diagnostic)  Times conv3x3 v2 on the level-0 / level-1 shapes at batch 10 with and
without the GroupNorm-statistics epilogue and for MT 1 / 2:   python tests/exp_conv64.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_research_b200 import _lib
from unet_research_b200._lib import ConvDesc, call, ptr, stream_ptr

dev = torch.device("cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
shapes = [(592, 576, 64, 64), (592, 576, 128, 64), (296, 288, 64, 128), (296, 288, 128, 128), (296, 288, 256, 128), (148, 144, 256, 256)]
for (h, w, cin, cout) in shapes:
    x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
    wp = torch.randn(9, cout, cin, device=dev).to(torch.bfloat16)
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=dev)
    flop = 2.0 * n * h * w * cout * 9 * cin
    line = f"  {h}x{w} {cin}->{cout}:"
    for groups in (32, 0):
        for mt in (0, 1, 2):
            d = ConvDesc()
            d.n, d.h, d.w, d.cin, d.cout, d.dtype, d.num_groups, d.x_cstride = n, h, w, cin, cout, _lib.BF16, groups, cin
            d.reserved[3] = mt
            parts = None
            if groups:
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
            line += f"  g{groups} mt{mt} {ms * 1000:.0f}us {flop / ms / 1e9:.0f}TF"
    print(line, flush=True)
