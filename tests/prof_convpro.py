"""ncu target: the fused-prologue conv kernel alone on one shape (default: level 0, 64 -> 64 at 592x576, batch 10).
   python tests/prof_convpro.py [h w cin cout batch]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_research_b200 import _lib
from unet_research_b200._lib import ConvDesc, call, ptr, stream_ptr

a = [int(v) for v in sys.argv[1:6]] + [592, 576, 64, 64, 10][len(sys.argv) - 1:]
h, w, cin, cout, n = a
dev = torch.device("cuda")
raw = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
coef = torch.rand(n, cin, 2, device=dev).contiguous()
mask = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, h, w, cin // 32), dtype=torch.int64, device=dev).to(torch.int32)
wp = torch.randn(9, cout, cin, device=dev).to(torch.bfloat16)
y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=dev)
d = ConvDesc()
d.n, d.h, d.w, d.cin, d.cout, d.dtype, d.num_groups, d.x_cstride = n, h, w, cin, cout, _lib.BF16, 32, cin
rows, sgs = C.c_int(0), C.c_int(0)
call("b2u_conv3x3_stat_layout", C.byref(d), C.byref(rows), C.byref(sgs))
parts = torch.empty(n, rows.value, cout // sgs.value, 2, dtype=torch.float32, device=dev)
for _ in range(3):
    call("b2u_conv3x3_pro_fwd", ptr(raw), ptr(coef), ptr(mask), ptr(wp), ptr(y), ptr(parts), C.byref(d), 1, 0, stream_ptr())
torch.cuda.synchronize()
