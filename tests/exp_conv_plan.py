"""Tile-plan sweep at the Monte-Carlo step's batch (diagnostic): every conv3x3 of the canonical U-Net, plain and with the
fused prologue, for BLOCK_N in {64, 128, 256} x MT in {1, 2}:   python tests/exp_conv_plan.py [batch] [dtype]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_research_b200 import _lib
from unet_research_b200._lib import ConvDesc, call, ptr, stream_ptr

dev = torch.device("cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dtype = _lib.F16 if (len(sys.argv) > 2 and sys.argv[2] == "fp16") else _lib.BF16
tdt = torch.float16 if dtype == _lib.F16 else torch.bfloat16
# (h, w, cin, cout, fused in the inference schedule?)
shapes = [(592, 576, 64, 64, True), (592, 576, 128, 64, False), (296, 288, 64, 128, True), (296, 288, 128, 128, True),
          (296, 288, 256, 128, False), (148, 144, 128, 256, True), (148, 144, 256, 256, True), (148, 144, 512, 256, False),
          (74, 72, 256, 512, True), (74, 72, 512, 512, True), (74, 72, 1024, 512, False), (37, 36, 512, 1024, True),
          (37, 36, 1024, 1024, True)]


def timeit(fn, reps=8):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000.0


for (h, w, cin, cout, fused) in shapes:
    raw = torch.randn(n, h, w, cin, device=dev).to(tdt)
    coef = torch.rand(n, cin, 2, device=dev).contiguous()
    mask = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, h, w, cin // 32), dtype=torch.int64, device=dev).to(torch.int32)
    wp = torch.randn(9, cout, cin, device=dev).to(tdt)
    y = torch.empty(n, h, w, cout, dtype=tdt, device=dev)
    line = f"  {h}x{w} {cin:4d}->{cout:4d} {'pro ' if fused else 'plain'}:"
    best = None
    for bn in (64, 128, 256):
        if cout % bn:
            continue
        for mt in (1, 2):
            d = ConvDesc()
            d.n, d.h, d.w, d.cin, d.cout, d.dtype, d.num_groups, d.x_cstride = n, h, w, cin, cout, dtype, 32, cin
            d.reserved[0], d.reserved[3] = bn, mt
            rows, sgs = C.c_int(0), C.c_int(0)
            call("b2u_conv3x3_stat_layout", C.byref(d), C.byref(rows), C.byref(sgs))
            parts = torch.empty(n, rows.value, cout // sgs.value, 2, dtype=torch.float32, device=dev)
            if fused:
                fn = lambda: call("b2u_conv3x3_pro_fwd", ptr(raw), ptr(coef), ptr(mask), ptr(wp), ptr(y), ptr(parts), C.byref(d), 1, 0, stream_ptr())
            else:
                fn = lambda: call("b2u_conv3x3_fwd", ptr(raw), ptr(wp), ptr(y), ptr(parts), C.byref(d), stream_ptr())
            try:
                us = timeit(fn)
            except Exception as e:                      # a plan that does not fit shared memory
                line += f"  bn{bn}/mt{mt} n/a"
                continue
            line += f"  bn{bn}/mt{mt} {us:5.0f}"
            if best is None or us < best[1]:
                best = (f"bn{bn}/mt{mt}", us)
    d = ConvDesc()
    d.n, d.h, d.w, d.cin, d.cout, d.dtype, d.num_groups, d.x_cstride = n, h, w, cin, cout, dtype, 32, cin
    rows, sgs = C.c_int(0), C.c_int(0)
    call("b2u_conv3x3_stat_layout", C.byref(d), C.byref(rows), C.byref(sgs))
    parts = torch.empty(n, rows.value, cout // sgs.value, 2, dtype=torch.float32, device=dev)
    if fused:
        us0 = timeit(lambda: call("b2u_conv3x3_pro_fwd", ptr(raw), ptr(coef), ptr(mask), ptr(wp), ptr(y), ptr(parts), C.byref(d), 1, 0, stream_ptr()))
    else:
        us0 = timeit(lambda: call("b2u_conv3x3_fwd", ptr(raw), ptr(wp), ptr(y), ptr(parts), C.byref(d), stream_ptr()))
    print(line + f"   | default {us0:5.0f}  best {best[0]} {best[1]:5.0f}", flush=True)
