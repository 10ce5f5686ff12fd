"""Run one conv3x3 configuration a few times (for ncu): python tests/prof_conv.py H W CIN COUT N [BN] [MT] [VER]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_research_b200 import _lib
from unet_research_b200._lib import ConvDesc, call, ptr, stream_ptr

h, w, cin, cout, n = [int(v) for v in sys.argv[1:6]]
bn = int(sys.argv[6]) if len(sys.argv) > 6 else 0
mt = int(sys.argv[7]) if len(sys.argv) > 7 else 0
ver = int(sys.argv[8]) if len(sys.argv) > 8 else 0
dev = torch.device("cuda")
x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
wp = torch.randn(9, cout, cin, device=dev).to(torch.bfloat16)
y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=dev)
d = ConvDesc()
d.n, d.h, d.w, d.cin, d.cout, d.dtype, d.num_groups, d.x_cstride = n, h, w, cin, cout, _lib.BF16, 32, cin
d.reserved[0], d.reserved[2], d.reserved[3] = bn, ver, mt
rows, sgs = C.c_int(0), C.c_int(0)
call("b2u_conv3x3_stat_layout", C.byref(d), C.byref(rows), C.byref(sgs))
parts = torch.empty(n, rows.value, cout // sgs.value, 2, dtype=torch.float32, device=dev)
for _ in range(3):
    call("b2u_conv3x3_fwd", ptr(x), ptr(wp), ptr(y), ptr(parts), C.byref(d), stream_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    call("b2u_conv3x3_fwd", ptr(x), ptr(wp), ptr(y), ptr(parts), C.byref(d), stream_ptr())
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{h}x{w} {cin}->{cout} n{n} bn{bn} mt{mt} v{ver}: {ms * 1000:.1f} us, {2.0 * n * h * w * cout * 9 * cin / ms / 1e9:.0f} TFLOP/s")
