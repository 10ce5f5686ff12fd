"""Fused conv prologue vs the two-pass schedule it replaces, per layer shape of the canonical U-Net at batch 10 (diagnostic):
   python tests/exp_convpro.py [batch]
prints: apply us + conv us (two passes) | fused conv us | gain."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unet_research_b200 import _lib
from unet_research_b200._lib import ApplyDesc, ConvDesc, call, ptr, stream_ptr

dev = torch.device("cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dtype = _lib.F16 if (len(sys.argv) > 2 and sys.argv[2] == "fp16") else _lib.BF16
tdt = torch.float16 if dtype == _lib.F16 else torch.bfloat16
# (h, w, cin, cout, masked+relu)   -- every fused site of the schedule
shapes = [(592, 576, 64, 64, True), (296, 288, 64, 128, False), (296, 288, 128, 128, True), (148, 144, 128, 256, False),
          (148, 144, 256, 256, True), (74, 72, 256, 512, False), (74, 72, 512, 512, True), (37, 36, 512, 1024, False),
          (37, 36, 1024, 1024, True)]


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000.0


for (h, w, cin, cout, act) in shapes:
    raw = torch.randn(n, h, w, cin, device=dev).to(tdt)
    actb = torch.empty_like(raw)
    coef = torch.rand(n, cin, 2, device=dev).contiguous()
    mask = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, h, w, cin // 32), dtype=torch.int64, device=dev).to(torch.int32) if act else None
    wp = torch.randn(9, cout, cin, device=dev).to(tdt)
    y = torch.empty(n, h, w, cout, dtype=tdt, device=dev)
    d = ConvDesc()
    d.n, d.h, d.w, d.cin, d.cout, d.dtype, d.num_groups, d.x_cstride = n, h, w, cin, cout, dtype, 32, cin
    rows, sgs = C.c_int(0), C.c_int(0)
    call("b2u_conv3x3_stat_layout", C.byref(d), C.byref(rows), C.byref(sgs))
    parts = torch.empty(n, rows.value, cout // sgs.value, 2, dtype=torch.float32, device=dev)
    a = ApplyDesc()
    a.n, a.h, a.w, a.c, a.dtype, a.relu, a.out_cstride, a.out_coffset = n, h, w, cin, dtype, int(act), cin, 0
    a.images_per_call2, a.numel_per_call2 = 1, 0.0
    t_apply = timeit(lambda: call("b2u_gn_apply", ptr(raw), ptr(coef), ptr(mask), None, None, ptr(actb), C.byref(a), stream_ptr()))
    t_conv = timeit(lambda: call("b2u_conv3x3_fwd", ptr(actb), ptr(wp), ptr(y), ptr(parts), C.byref(d), stream_ptr()))
    t_fused = timeit(lambda: call("b2u_conv3x3_pro_fwd", ptr(raw), ptr(coef), ptr(mask), ptr(wp), ptr(y), ptr(parts), C.byref(d), int(act), 0,
                                  stream_ptr()))
    flop = 2.0 * n * h * w * cout * 9 * cin
    print(f"  {h}x{w} {cin:4d}->{cout:4d} act{int(act)}: apply {t_apply:6.0f} + conv {t_conv:6.0f} = {t_apply + t_conv:6.0f} us ({flop / t_conv / 1e6:5.0f} TF) | "
          f"fused {t_fused:6.0f} us ({flop / t_fused / 1e6:5.0f} TF) | gain {t_apply + t_conv - t_fused:6.0f} us", flush=True)
