"""GPU bring-up diagnostics (run on the B200 box through gpurun):

    python tests/gpu_diag.py            # every section, each in its own subprocess (a trapped kernel
                                        # kills only its section), summary in gpurun_out/diag.log
    python tests/gpu_diag.py conv       # one section in-process

Not a pytest file: it prints error magnitudes instead of asserting, to get the most out of one GPU call.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SECTIONS = ["info", "conv", "convv1", "convv2", "wgrad", "bwdk", "convt", "first", "gn", "pool", "head", "head_variants", "dropblock", "rotate", "forward", "mc", "rot_ens", "tf32", "train", "traingraph", "precision", "ichan"]


def rel(a, b):
    import torch
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30)), float((a - b).abs().max())


def sec_info():
    import torch
    from unet_research_b200 import _lib
    from unet_research_b200.engine import device_rand_geometry
    _lib.load()
    print("device", torch.cuda.get_device_name(0), "sms/maxthreads", device_rand_geometry())
    print("torch", torch.__version__, "cuda", torch.version.cuda)


def _conv_case(n, h, w, cin, cout, dtype, conv_t=False, block_n=0, stages=0, version=0, mt=0):
    import torch
    import torch.nn.functional as F
    from unet_research_b200 import _lib
    from unet_research_b200._lib import ConvDesc, call, ptr, stream_ptr
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(cin * 1000 + cout + h)
    tdt = torch.float32 if dtype == _lib.F32 else (torch.float16 if dtype == _lib.F16 else torch.bfloat16)
    x = torch.randn(n, cin, h, w, generator=g).to(dev)
    if conv_t:
        wt = (torch.randn(cin, cout, 2, 2, generator=g) / (cin ** 0.5)).to(dev)
    else:
        wt = (torch.randn(cout, cin, 3, 3, generator=g) / ((9 * cin) ** 0.5)).to(dev)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(tdt)
    G = 32
    d = ConvDesc()
    d.n, d.h, d.w, d.cin, d.cout, d.dtype, d.num_groups, d.x_cstride = n, h, w, cin, cout, dtype, G, cin
    d.reserved[0], d.reserved[1], d.reserved[2], d.reserved[3] = block_n, stages, version, mt
    rows, sgs = C.c_int(0), C.c_int(0)
    call("b2u_convT2x2_stat_layout" if conv_t else "b2u_conv3x3_stat_layout", C.byref(d), C.byref(rows), C.byref(sgs))
    oh, ow = (2 * h, 2 * w) if conv_t else (h, w)
    y = torch.full((n, oh, ow, cout), float("nan"), dtype=tdt, device=dev)
    parts = torch.full((n, rows.value, cout // sgs.value, 2), float("nan"), dtype=torch.float32, device=dev)
    if conv_t:
        packed = torch.empty(4, cout, cin, dtype=tdt, device=dev)
        call("b2u_pack_convT2x2_weight", ptr(wt), ptr(packed), cin, cout, dtype, stream_ptr())
    else:
        packed = torch.empty(9, cout, cin, dtype=tdt, device=dev)
        call("b2u_pack_conv3x3_weight", ptr(wt), ptr(packed), cout, cin, dtype, 0, stream_ptr())
    call("b2u_convT2x2_fwd" if conv_t else "b2u_conv3x3_fwd", ptr(x_nhwc), ptr(packed), ptr(y), ptr(parts), C.byref(d), stream_ptr())
    torch.cuda.synchronize()
    # reference on the SAME rounded operands, fp32 math (TF32 off)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    xr = x_nhwc.float().permute(0, 3, 1, 2).double()
    wr = wt.to(tdt).float().double() if dtype != _lib.F32 else wt.double()
    ref = (F.conv_transpose2d(xr, wr, stride=2) if conv_t else F.conv2d(xr, wr, padding=1)).float()
    got = y.float().permute(0, 3, 1, 2)
    r, mx = rel(got, ref)
    # statistics: per (image, group) sums
    gs = cout // G
    ps = parts.double().sum(1).view(n, G, -1, 2).sum(2)          # [n, G, 2]
    rs = ref.double().view(n, G, gs * oh * ow)
    s_ref = torch.stack([rs.sum(-1), (rs * rs).sum(-1)], -1)
    sr, _ = rel(ps, s_ref)
    nan = int(torch.isnan(got).sum())
    print(f"  {'convT' if conv_t else 'conv3'} n{n} {h}x{w} {cin}->{cout} dt{dtype} bn{block_n} st{stages} v{version} mt{mt}: rel {r:.3e} max {mx:.3e} stats_rel {sr:.3e} nan {nan} rows {rows.value} sgs {sgs.value}")
    return {"rel": r, "stats_rel": sr, "nan": nan, "y": y, "parts": parts}


def _conv_pro_case(n, h, w, cin, cout, dtype, relu=True, masked=True, x_shared=False, block_n=0, mt=0):
    """Fused conv prologue (b2u_conv3x3_pro_fwd) against the two-pass schedule it replaces (b2u_gn_apply -> b2u_conv3x3_fwd)
    on the same raw tensor, coefficients and keep bits: outputs and GroupNorm partials must be BIT-IDENTICAL (same operand
    values, same MMA order); border tiles, ragged sizes and the zero-padding-after-affine rule are covered by h, w not
    being tile multiples and by coefficients with b != 0 (a padded pixel must contribute 0, not relu(b))."""
    import torch
    from unet_research_b200 import _lib
    from unet_research_b200._lib import ApplyDesc, ConvDesc, call, ptr, stream_ptr
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(cin * 7 + cout + h + w)
    tdt = torch.float16 if dtype == _lib.F16 else torch.bfloat16
    nx = 1 if x_shared else n
    raw = torch.randn(nx, h, w, cin, generator=g).to(dev).to(tdt)
    coef = torch.stack([0.5 + torch.rand(n, cin, generator=g), torch.rand(n, cin, generator=g) - 0.3], -1).to(dev).contiguous()
    mask = None
    if masked:
        mask = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, h, w, cin // 32), generator=g, dtype=torch.int64).to(torch.int32).to(dev)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / ((9 * cin) ** 0.5)).to(dev)
    packed = torch.empty(9, cout, cin, dtype=tdt, device=dev)
    call("b2u_pack_conv3x3_weight", ptr(wt), ptr(packed), cout, cin, dtype, 0, stream_ptr())
    d = ConvDesc()
    d.n, d.h, d.w, d.cin, d.cout, d.dtype, d.num_groups, d.x_cstride = n, h, w, cin, cout, dtype, 32, cin
    d.reserved[0], d.reserved[3] = block_n, mt
    rows, sgs = C.c_int(0), C.c_int(0)
    call("b2u_conv3x3_stat_layout", C.byref(d), C.byref(rows), C.byref(sgs))
    outs = []
    for fused in (False, True):
        y = torch.full((n, h, w, cout), float("nan"), dtype=tdt, device=dev)
        parts = torch.full((n, rows.value, cout // sgs.value, 2), float("nan"), dtype=torch.float32, device=dev)
        if fused:
            call("b2u_conv3x3_pro_fwd", ptr(raw), ptr(coef), ptr(mask), ptr(packed), ptr(y), ptr(parts), C.byref(d), int(relu),
                 int(x_shared), stream_ptr())
        else:
            act = torch.empty(n, h, w, cin, dtype=tdt, device=dev)
            a = ApplyDesc()
            a.n, a.h, a.w, a.c, a.dtype, a.relu, a.out_cstride, a.out_coffset = n, h, w, cin, dtype, int(relu), cin, 0
            a.images_per_call2, a.numel_per_call2 = 1, 0.0
            a.reserved[0] = 1 if x_shared else 0
            call("b2u_gn_apply", ptr(raw), ptr(coef), ptr(mask), None, None, ptr(act), C.byref(a), stream_ptr())
            call("b2u_conv3x3_fwd", ptr(act), ptr(packed), ptr(y), ptr(parts), C.byref(d), stream_ptr())
        torch.cuda.synchronize()
        outs.append((y, parts))
    (y0, p0), (y1, p1) = outs
    nan = int(torch.isnan(y1.float()).sum())
    same = bool(torch.equal(y0, y1)) and bool(torch.equal(p0, p1))
    dmax = float((y0.float() - y1.float()).abs().max())
    print(f"  conv3-pro n{n} {h}x{w} {cin}->{cout} dt{dtype} relu{int(relu)} mask{int(masked)} shared{int(x_shared)} bn{block_n} mt{mt}: "
          f"identical {same} max|d| {dmax:.3e} nan {nan}")
    return {"identical": same, "max": dmax, "nan": nan}


def sec_convpro():
    from unet_research_b200 import _lib
    for dt in (_lib.BF16, _lib.F16):
        _conv_pro_case(1, 16, 16, 64, 64, dt)
        _conv_pro_case(2, 33, 47, 64, 64, dt)
        _conv_pro_case(3, 24, 40, 64, 64, dt, x_shared=True)
        _conv_pro_case(2, 37, 36, 128, 128, dt)
        _conv_pro_case(2, 20, 24, 64, 128, dt, relu=False, masked=False)
        _conv_pro_case(1, 37, 36, 512, 1024, dt)
        _conv_pro_case(1, 74, 72, 256, 256, dt)
        _conv_pro_case(1, 33, 47, 128, 256, dt, block_n=128, mt=1)


def sec_conv():
    from unet_research_b200 import _lib
    _conv_case(1, 16, 16, 64, 64, _lib.BF16)
    _conv_case(2, 24, 40, 64, 64, _lib.BF16)
    _conv_case(1, 37, 36, 128, 128, _lib.BF16)
    _conv_case(2, 20, 24, 128, 256, _lib.BF16)
    _conv_case(1, 74, 72, 512, 512, _lib.BF16)
    _conv_case(1, 37, 36, 1024, 1024, _lib.BF16)
    _conv_case(1, 16, 16, 128, 64, _lib.BF16)
    _conv_case(1, 33, 47, 64, 128, _lib.BF16, block_n=64)
    _conv_case(1, 32, 32, 256, 256, _lib.BF16, block_n=128, stages=2)
    _conv_case(1, 592, 576, 64, 64, _lib.BF16)


def sec_convv1():
    from unet_research_b200 import _lib
    _conv_case(2, 24, 40, 64, 64, _lib.BF16, version=1)
    _conv_case(1, 37, 36, 128, 128, _lib.BF16, version=1)
    _conv_case(1, 74, 72, 512, 512, _lib.BF16, version=1)
    _conv_case(1, 33, 47, 64, 128, _lib.BF16, block_n=64, version=1)


def sec_convv2():
    from unet_research_b200 import _lib
    for mt in (1, 2):
        _conv_case(1, 16, 16, 64, 64, _lib.BF16, mt=mt)
        _conv_case(3, 24, 40, 64, 64, _lib.BF16, mt=mt)
        _conv_case(1, 37, 36, 128, 128, _lib.BF16, mt=mt)
        _conv_case(2, 20, 24, 128, 256, _lib.BF16, mt=mt)
        _conv_case(1, 33, 47, 128, 256, _lib.BF16, block_n=64, mt=mt)
        _conv_case(1, 33, 47, 128, 256, _lib.BF16, block_n=128, mt=mt)
    _conv_case(1, 74, 72, 512, 512, _lib.BF16)
    _conv_case(5, 37, 36, 1024, 1024, _lib.BF16)
    _conv_case(1, 592, 576, 64, 64, _lib.BF16)
    _conv_case(2, 24, 40, 128, 256, _lib.F32, mt=2)


def sec_convbench():
    """Times every conv3x3 shape of the canonical U-Net at batch 5 for the v1 kernel and the v2 variants."""
    import torch
    from unet_research_b200 import _lib
    from unet_research_b200._lib import ConvDesc, call, ptr, stream_ptr
    dev = torch.device("cuda")
    n = 5
    shapes = [(592, 576, 64, 64), (592, 576, 128, 64), (296, 288, 64, 128), (296, 288, 128, 128), (296, 288, 256, 128),
              (148, 144, 128, 256), (148, 144, 256, 256), (148, 144, 512, 256), (74, 72, 256, 512), (74, 72, 512, 512),
              (74, 72, 1024, 512), (37, 36, 512, 1024), (37, 36, 1024, 1024)]
    total = {}
    for (h, w, cin, cout) in shapes:
        x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
        wp = torch.randn(9, cout, cin, device=dev).to(torch.bfloat16)
        y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=dev)
        flop = 2.0 * n * h * w * cout * 9 * cin
        line = f"  {h}x{w} {cin}->{cout}:"
        variants = [("v1", 1, 0, 0)] + [(f"v2 bn{bn} mt{mt}", 0, bn, mt) for bn in (64, 128, 256) if cout % bn == 0 and bn <= cout for mt in (1, 2)]
        best = None
        for name, ver, bn, mt in variants:
            d = ConvDesc()
            d.n, d.h, d.w, d.cin, d.cout, d.dtype, d.num_groups, d.x_cstride = n, h, w, cin, cout, _lib.BF16, 32, cin
            d.reserved[0], d.reserved[2], d.reserved[3] = bn, ver, mt
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
            tf = flop / ms / 1e9
            line += f"  {name} {ms * 1000:.0f}us {tf:.0f}TF"
            if best is None or ms < best[1]:
                best = (name, ms)
            total[name] = total.get(name, 0.0) + ms
        print(line + f"   best {best[0]}", flush=True)


def sec_convtbench():
    """Times the four transposed convolutions of the canonical U-Net at batch 5: v1 (CTA per tile and tap) vs v2 (persistent)."""
    import torch
    from unet_research_b200 import _lib
    from unet_research_b200._lib import ConvDesc, call, ptr, stream_ptr
    dev = torch.device("cuda")
    n = 5
    for (h, w, cin, cout) in [(37, 36, 1024, 512), (74, 72, 512, 256), (148, 144, 256, 128), (296, 288, 128, 64)]:
        x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
        wp = torch.randn(4, cout, cin, device=dev).to(torch.bfloat16)
        y = torch.empty(n, 2 * h, 2 * w, cout, dtype=torch.bfloat16, device=dev)
        flop = 2.0 * n * h * w * 4 * cout * cin
        nbytes = x.numel() * 2 + y.numel() * 2
        line = f"  convT {h}x{w} {cin}->{cout}:"
        for name, ver in (("v1", 1), ("v2", 0)):
            d = ConvDesc()
            d.n, d.h, d.w, d.cin, d.cout, d.dtype, d.num_groups, d.x_cstride = n, h, w, cin, cout, _lib.BF16, 32, cin
            d.reserved[2] = ver
            rows, sgs = C.c_int(0), C.c_int(0)
            call("b2u_convT2x2_stat_layout", C.byref(d), C.byref(rows), C.byref(sgs))
            parts = torch.empty(n, rows.value, cout // sgs.value, 2, dtype=torch.float32, device=dev)
            for _ in range(2):
                call("b2u_convT2x2_fwd", ptr(x), ptr(wp), ptr(y), ptr(parts), C.byref(d), stream_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                call("b2u_convT2x2_fwd", ptr(x), ptr(wp), ptr(y), ptr(parts), C.byref(d), stream_ptr())
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            line += f"  {name} {ms * 1000:.0f}us {flop / ms / 1e9:.0f}TF {nbytes / ms / 1e6:.0f}GB/s"
        print(line, flush=True)


def _wgrad_case(n, h, w, cg, cx, taps, layout=0, x_cstride=None):
    import torch
    import torch.nn.functional as F
    from unet_research_b200 import _lib
    from unet_research_b200._lib import WgradDesc, call, ptr, stream_ptr
    dev = torch.device("cuda")
    xs = x_cstride or cx
    gen = torch.Generator().manual_seed(cg + cx + h)
    g = torch.randn(n, h, w, cg, generator=gen).to(dev).to(torch.bfloat16)
    x = torch.randn(n, h, w, xs, generator=gen).to(dev).to(torch.bfloat16)
    d = WgradDesc()
    d.n, d.h, d.w, d.cg, d.cx, d.x_cstride, d.taps, d.layout, d.dtype = n, h, w, cg, cx, xs, taps, layout, _lib.BF16
    fl = C.c_longlong(0)
    call("b2u_wgrad_workspace_floats", C.byref(d), C.byref(fl))
    ws = torch.empty(fl.value, dtype=torch.float32, device=dev)
    if layout == 0:
        dw = torch.full((cg, cx, 3, 3) if taps == 9 else (cg, cx), float("nan"), device=dev)
    else:
        dw = torch.full((cx, cg // 4, 2, 2), float("nan"), device=dev)
    call("b2u_wgrad", ptr(g), ptr(x), ptr(ws), ptr(dw), C.byref(d), stream_ptr())
    torch.cuda.synchronize()
    gd = g.double().permute(0, 3, 1, 2)
    xd = x[..., :cx].double().permute(0, 3, 1, 2)
    if taps == 9:
        wv = torch.zeros(cg, cx, 3, 3, dtype=torch.float64, device=dev, requires_grad=True)
        (F.conv2d(xd, wv, padding=1) * gd).sum().backward()
        ref = wv.grad
    else:
        ref = torch.einsum("ngp,nxp->gx", gd.flatten(2), xd.flatten(2))
        if layout == 1:
            cout = cg // 4
            ref = ref.view(4, cout, cx).permute(2, 1, 0).reshape(cx, cout, 2, 2)
    r, mx = rel(dw, ref)
    print(f"  wgrad n{n} {h}x{w} cg{cg} cx{cx} taps{taps} layout{layout} xs{xs}: rel {r:.3e} max {mx:.3e} nan {int(torch.isnan(dw).sum())}")
    return r


def sec_wgrad():
    _wgrad_case(1, 16, 16, 64, 64, 1)
    _wgrad_case(1, 16, 16, 128, 64, 1)
    _wgrad_case(1, 16, 16, 64, 64, 9)
    _wgrad_case(2, 24, 40, 128, 64, 9)
    _wgrad_case(1, 37, 36, 256, 128, 9)
    _wgrad_case(1, 33, 47, 64, 128, 9, x_cstride=256)
    _wgrad_case(2, 20, 24, 512, 128, 1, layout=1)
    _wgrad_case(1, 592, 576, 64, 64, 9)
    _wgrad_case(1, 37, 36, 1024, 1024, 9)


def sec_bwdk():
    """unit backward kernels (GroupNorm/DropBlock/ReLU + pool + concat sources), 1x1 GEMM, first-layer wgrad vs autograd"""
    import torch
    import torch.nn.functional as F
    from unet_research_b200 import _lib
    from unet_research_b200._lib import ConvDesc, UnitBwdDesc, call, ptr, stream_ptr
    dev = torch.device("cuda")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    G = 32
    for (n, c, h, w, use_pool, use_a, use_head, s2d) in ((2, 64, 16, 24, True, True, False, False), (1, 256, 8, 8, False, True, False, True),
                                                         (2, 64, 32, 48, False, False, True, False), (1, 128, 12, 16, True, True, False, False)):
        gen = torch.Generator().manual_seed(c + h)
        y = (torch.randn(n, c, h, w, generator=gen) * 1.5 + 0.3).to(dev)
        y_nhwc = y.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        yr = y_nhwc.float().permute(0, 3, 1, 2).double().requires_grad_(True)
        gamma = (1 + 0.2 * torch.randn(c, generator=gen)).to(dev)
        beta = (0.2 * torch.randn(c, generator=gen)).to(dev)
        m1 = (torch.rand(n, c, h, w, generator=gen) > 0.15).to(dev)
        s1 = m1.numel() / m1.sum().item()
        # forward reference (fp64)
        z = F.group_norm(yr, G, gamma.double(), beta.double(), 1e-5)
        act = F.relu(z * m1 * s1)
        loss = 0.0
        ga = gp = None
        if use_a:
            ga = torch.randn(n, h, w, 2 * c, generator=gen).to(dev).to(torch.bfloat16)
            m2 = (torch.rand(n, 2 * c, h, w, generator=gen) > 0.15).to(dev)
            s2 = m2.numel() / m2.sum().item()
            loss = loss + (act * m2[:, c:] * s2 * ga[..., c:].float().permute(0, 3, 1, 2).double()).sum()
        if use_pool:
            gp = torch.randn(n, h // 2, w // 2, c, generator=gen).to(dev).to(torch.bfloat16)
            pooled, idx = F.max_pool2d(act, 2, 2, return_indices=True)
            loss = loss + (pooled * gp.float().permute(0, 3, 1, 2).double()).sum()
        if use_head:
            h0, w0 = h - 3, w - 5
            wh = (torch.randn(c, generator=gen) / 8).to(dev)
            go = torch.randn(n, 1, h0, w0, generator=gen).to(dev)
            logit = (act * wh.double().view(1, c, 1, 1)).sum(1, keepdim=True)[:, :, :h0, :w0]
            out = torch.sigmoid(logit)
            loss = loss + (out * go.double()).sum()
        whp = wh.double().requires_grad_(True) if False else None
        gparams = torch.autograd.grad(loss, [yr], retain_graph=True)[0]
        # parameter grads through an independent graph
        gam_d = gamma.double().requires_grad_(True)
        bet_d = beta.double().requires_grad_(True)
        z2 = F.group_norm(yr.detach(), G, gam_d, bet_d, 1e-5)
        act2 = F.relu(z2 * m1 * s1)
        loss2 = 0.0
        if use_a:
            loss2 = loss2 + (act2 * m2[:, c:] * s2 * ga[..., c:].float().permute(0, 3, 1, 2).double()).sum()
        if use_pool:
            loss2 = loss2 + (F.max_pool2d(act2, 2, 2) * gp.float().permute(0, 3, 1, 2).double()).sum()
        if use_head:
            whd = wh.double().requires_grad_(True)
            out2 = torch.sigmoid((act2 * whd.view(1, c, 1, 1)).sum(1, keepdim=True)[:, :, :h0, :w0])
            loss2 = loss2 + (out2 * go.double()).sum()
            dgam, dbet, dwh = torch.autograd.grad(loss2, [gam_d, bet_d, whd])
        else:
            dgam, dbet = torch.autograd.grad(loss2, [gam_d, bet_d])
        # our kernels: forward finalize to get coef / mean_rstd, masks as bits
        def pack_bits(m):   # [n, C, h, w] bool -> uint32 [n,h,w,C/32]
            mm = m.permute(0, 2, 3, 1).reshape(n, h, w, -1, 32).to(torch.int64)
            return (mm << torch.arange(32, device=dev)).sum(-1).to(torch.int32).contiguous()
        gs = c // G
        sgs = min(gs, 32)
        yv = y_nhwc.float().permute(0, 3, 1, 2).double().reshape(n, c // sgs, sgs * h * w)
        parts = torch.stack([yv.sum(-1), (yv * yv).sum(-1)], -1).float().view(n, 1, c // sgs, 2).contiguous()
        coef = torch.empty(n, c, 2, device=dev)
        mr = torch.empty(n, G, 2, device=dev)
        keep1 = torch.tensor([int(m1.sum().item())], dtype=torch.int64, device=dev)
        call("b2u_gn_finalize", ptr(parts), 1, sgs, ptr(gamma), ptr(beta), ptr(coef), n, c, G, float(gs * h * w), 1e-5, ptr(keep1), n,
             float(m1.numel()), ptr(mr), stream_ptr())
        d = UnitBwdDesc()
        d.n, d.h, d.w, d.c, d.dtype, d.relu, d.num_groups, d.s2d = n, h, w, c, _lib.BF16, 1, G, int(s2d)
        d.images_per_call1, d.numel_per_call1 = n, float(m1.numel())
        d.y, d.coef, d.mean_rstd, d.gamma = ptr(y_nhwc), ptr(coef), ptr(mr), ptr(gamma)
        m1b = pack_bits(m1)
        d.mask1, d.keep_counts1 = ptr(m1b), ptr(keep1)
        keep = [m1b, keep1]
        if use_a:
            m2b = pack_bits(m2)
            keep2 = torch.tensor([int(m2.sum().item())], dtype=torch.int64, device=dev)
            d.grad_a, d.a_cstride, d.a_coffset = ptr(ga), 2 * c, c
            d.mask2, d.mask2_cstride, d.mask2_coffset, d.keep_counts2 = ptr(m2b), 2 * c, c, ptr(keep2)
            d.images_per_call2, d.numel_per_call2 = n, float(m2.numel())
            keep += [m2b, keep2]
        if use_pool:
            iy, ix = idx // w, idx % w
            am = ((iy % 2) * 2 + (ix % 2)).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
            d.grad_pool, d.argmax = ptr(gp), ptr(am)
            keep += [am]
        if use_head:
            outf = out.detach().float().contiguous()
            d.grad_out, d.out, d.w_head, d.h0, d.w0 = ptr(go), ptr(outf), ptr(wh), h0, w0
            keep += [outf]
        rows = C.c_int(0)
        call("b2u_unit_bwd_rows", h, w, c, C.byref(rows))
        bparts = torch.empty(n, rows.value, c, 3, device=dev)
        gcoef = torch.empty(n, G, 2, device=dev)
        dgamma = torch.empty(c, device=dev)
        dbeta = torch.empty(c, device=dev)
        dwh_o = torch.empty(c, device=dev)
        call("b2u_unit_bwd_stats", C.byref(d), ptr(bparts), stream_ptr())
        call("b2u_unit_bwd_finalize", ptr(bparts), n, rows.value, c, G, ptr(gamma), float(gs * h * w), ptr(gcoef), ptr(dgamma), ptr(dbeta),
             ptr(dwh_o) if use_head else None, stream_ptr())
        dy = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=dev) if not s2d else torch.empty(n, h // 2, w // 2, 4, c, dtype=torch.bfloat16, device=dev)
        call("b2u_unit_bwd_apply", C.byref(d), ptr(gcoef), ptr(dy), stream_ptr())
        torch.cuda.synchronize()
        if s2d:
            got = dy.float().view(n, h // 2, w // 2, 2, 2, c).permute(0, 5, 1, 3, 2, 4).reshape(n, c, h, w)
        else:
            got = dy.float().permute(0, 3, 1, 2)
        msg = f"  unit_bwd n{n} c{c} {h}x{w} pool{int(use_pool)} a{int(use_a)} head{int(use_head)} s2d{int(s2d)}: dY rel {rel(got, gparams)[0]:.3e} dgamma rel {rel(dgamma, dgam)[0]:.3e} dbeta rel {rel(dbeta, dbet)[0]:.3e}"
        if use_head:
            msg += f" dw_head rel {rel(dwh_o, dwh)[0]:.3e}"
        print(msg)
    # 1x1 GEMM
    n, h, w, k, co = 2, 20, 24, 512, 128
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(n, h, w, k, generator=gen).to(dev).to(torch.bfloat16)
    wt = (torch.randn(co, k, generator=gen) / k ** 0.5).to(dev).to(torch.bfloat16)
    yo = torch.empty(n, h, w, co, dtype=torch.bfloat16, device=dev)
    cd = ConvDesc()
    cd.n, cd.h, cd.w, cd.cin, cd.cout, cd.dtype, cd.num_groups, cd.x_cstride = n, h, w, k, co, _lib.BF16, 0, k
    call("b2u_gemm1x1_fwd", ptr(x), ptr(wt), ptr(yo), C.byref(cd), stream_ptr())
    torch.cuda.synchronize()
    print(f"  gemm1x1: rel {rel(yo.float(), x.double() @ wt.double().t())[0]:.3e}")
    # first-layer wgrad
    for cin in (1, 3):
        n, h0, w0, hh, ww, co = 2, 29, 45, 32, 48, 64
        xin = torch.rand(n, cin, h0, w0, generator=gen).to(dev)
        g = torch.randn(n, hh, ww, co, generator=gen).to(dev).to(torch.bfloat16)
        wsb = torch.empty(n * _lib.load().b2u_wgrad_first_rows() * co * cin * 9, device=dev)
        dw = torch.empty(co, cin, 3, 3, device=dev)
        call("b2u_wgrad_first", ptr(g), ptr(xin), ptr(wsb), ptr(dw), n, cin, h0, w0, hh, ww, co, _lib.BF16, stream_ptr())
        torch.cuda.synchronize()
        wv = torch.zeros(co, cin, 3, 3, dtype=torch.float64, device=dev, requires_grad=True)
        xp = F.pad(xin, (0, ww - w0, 0, hh - h0)).double()
        (F.conv2d(xp, wv, padding=1) * g.double().permute(0, 3, 1, 2)).sum().backward()
        print(f"  wgrad_first cin{cin}: rel {rel(dw, wv.grad)[0]:.3e}")


def sec_convt():
    from unet_research_b200 import _lib
    _conv_case(1, 16, 16, 128, 64, _lib.BF16, conv_t=True)
    _conv_case(2, 37, 36, 1024, 512, _lib.BF16, conv_t=True)
    _conv_case(1, 20, 24, 256, 128, _lib.BF16, conv_t=True)


def sec_tf32():
    from unet_research_b200 import _lib
    _conv_case(1, 16, 16, 64, 64, _lib.F32)
    _conv_case(2, 24, 40, 128, 256, _lib.F32)
    _conv_case(1, 16, 16, 128, 64, _lib.F32, conv_t=True)
    _forward_case(120, 116, 1, "tf32")


def sec_first():
    import torch
    import torch.nn.functional as F
    from unet_research_b200 import _lib
    from unet_research_b200._lib import call, ptr, stream_ptr
    dev = torch.device("cuda")
    res = []
    for cin, (h0, w0), (h, w) in ((1, (120, 116), (128, 128)), (3, (64, 80), (64, 80)), (1, (584, 565), (592, 576))):
        g = torch.Generator().manual_seed(5)
        x = torch.rand(2, cin, h0, w0, generator=g).to(dev)
        wt = torch.randn(64, cin, 3, 3, generator=g).to(dev)
        rows, sgs = C.c_int(0), C.c_int(0)
        call("b2u_conv_first_stat_layout", h, w, 64, 32, C.byref(rows), C.byref(sgs))
        y = torch.empty(2, h, w, 64, dtype=torch.bfloat16, device=dev)
        parts = torch.empty(2, rows.value, 64 // sgs.value, 2, dtype=torch.float32, device=dev)
        call("b2u_conv_first_fwd", ptr(x), ptr(wt), ptr(y), ptr(parts), 2, cin, h0, w0, h, w, 64, 32, _lib.BF16, stream_ptr())
        torch.cuda.synchronize()
        ref = F.conv2d(F.pad(x, (0, w - w0, 0, h - h0)).double(), wt.double(), padding=1).float()
        r, mx = rel(y.float().permute(0, 3, 1, 2), ref)
        ps = parts.double().sum(1).view(2, 32, -1, 2).sum(2)
        rs = ref.double().view(2, 32, -1)
        sr, _ = rel(ps, torch.stack([rs.sum(-1), (rs * rs).sum(-1)], -1))
        print(f"  first cin{cin} {h0}x{w0}->{h}x{w}: rel {r:.3e} max {mx:.3e} stats_rel {sr:.3e}")
        res.append({"rel": r, "stats_rel": sr})
    return res


def sec_gn():
    import torch
    import torch.nn.functional as F
    from unet_research_b200 import _lib
    from unet_research_b200._lib import ApplyDesc, call, ptr, stream_ptr
    dev = torch.device("cuda")
    res = []
    for c, h, w in ((64, 32, 48), (256, 20, 24), (1024, 8, 8)):
        n, G = 2, 32
        g = torch.Generator().manual_seed(c)
        x = (torch.randn(n, c, h, w, generator=g) * 2 + 0.5).to(dev)
        gamma = (1 + 0.1 * torch.randn(c, generator=g)).to(dev)
        beta = (0.1 * torch.randn(c, generator=g)).to(dev)
        x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        xr = x_nhwc.float().permute(0, 3, 1, 2)
        # partials: one row per image built with torch (the conv epilogue is tested in `conv`)
        gs = c // G
        sgs = min(gs, 32)
        v = xr.double().reshape(n, c // sgs, sgs * h * w)
        parts = torch.stack([v.sum(-1), (v * v).sum(-1)], -1).float().view(n, 1, c // sgs, 2).contiguous()
        coef = torch.empty(n, c, 2, dtype=torch.float32, device=dev)
        call("b2u_gn_finalize", ptr(parts), 1, sgs, ptr(gamma), ptr(beta), ptr(coef), n, c, G, float(gs * h * w), 1e-5, None, 1, 0.0, None, stream_ptr())
        a = ApplyDesc()
        a.n, a.h, a.w, a.c, a.dtype, a.relu, a.out_cstride, a.out_coffset = n, h, w, c, _lib.BF16, 1, c, 0
        a.images_per_call2 = 1
        out = torch.empty_like(x_nhwc)
        call("b2u_gn_apply", ptr(x_nhwc), ptr(coef), None, None, None, ptr(out), C.byref(a), stream_ptr())
        torch.cuda.synchronize()
        ref = F.relu(F.group_norm(xr, G, gamma, beta, 1e-5))
        r, mx = rel(out.float().permute(0, 3, 1, 2), ref)
        print(f"  gn+relu c{c} {h}x{w}: rel {r:.3e} max {mx:.3e}")
        res.append(r)
    return res


def sec_pool():
    import torch
    import torch.nn.functional as F
    from unet_research_b200 import _lib
    from unet_research_b200._lib import ApplyDesc, call, ptr, stream_ptr
    dev = torch.device("cuda")
    res = []
    for c, h, w in ((64, 32, 48), (512, 12, 8)):
        n, G = 2, 32
        g = torch.Generator().manual_seed(c + 1)
        x = torch.randn(n, c, h, w, generator=g).to(dev)
        x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        xr = x_nhwc.float().permute(0, 3, 1, 2)
        coef = torch.empty(n, c, 2, dtype=torch.float32, device=dev)
        coef[..., 0] = 1.5
        coef[..., 1] = -0.25
        rows, sgs = C.c_int(0), C.c_int(0)
        call("b2u_pool_stat_layout", h, w, c, G, C.byref(rows), C.byref(sgs))
        cat = torch.zeros(n, h, w, 2 * c, dtype=torch.bfloat16, device=dev)
        pooled = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device=dev)
        parts = torch.empty(n, rows.value, c // sgs.value, 2, dtype=torch.float32, device=dev)
        arg = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device=dev)
        a = ApplyDesc()
        a.n, a.h, a.w, a.c, a.dtype, a.relu, a.out_cstride, a.out_coffset = n, h, w, c, _lib.BF16, 1, 2 * c, c
        a.images_per_call2 = 1
        call("b2u_gn_apply_pool", ptr(x_nhwc), ptr(coef), None, None, None, ptr(cat), ptr(pooled), ptr(parts), ptr(arg), G, C.byref(a), stream_ptr())
        torch.cuda.synchronize()
        act = F.relu(xr * 1.5 - 0.25)
        pr, idx = F.max_pool2d(act, 2, 2, return_indices=True)
        r1, _ = rel(cat[..., c:].float().permute(0, 3, 1, 2), act)
        r2, _ = rel(pooled.float().permute(0, 3, 1, 2), pr)
        # window index from flat index
        iy, ix = idx // w, idx % w
        win = ((iy % 2) * 2 + (ix % 2)).to(torch.uint8)
        argmatch = float((arg.permute(0, 3, 1, 2) == win).float().mean())
        ps = parts.double().sum(1).view(n, G, -1, 2).sum(2)
        rs = pooled.float().permute(0, 3, 1, 2).double().reshape(n, G, -1)
        sr, _ = rel(ps, torch.stack([rs.sum(-1), (rs * rs).sum(-1)], -1))
        print(f"  apply_pool c{c}: skip rel {r1:.3e} pooled rel {r2:.3e} argmax match {argmatch:.6f} stats_rel {sr:.3e} first-half-untouched {float(cat[..., :c].abs().max()):.1f}")
        res.append({"skip": r1, "pooled": r2, "argmax": argmatch, "stats_rel": sr, "untouched": float(cat[..., :c].abs().max())})
    return res


def sec_head_variants():
    """The c = 64 head: cp.async ring kernel (default) vs the register kernel (B2U_HEAD_ASYNC=0) vs torch, with and without the
    DropBlock mask, shared and per-image fov, one and several trips per block, fp16 and bf16.  Returns (all bit-identical
    between the kernels, worst rel against torch)."""
    import torch
    from unet_research_b200 import _lib
    from unet_research_b200._lib import HeadDesc, call, ptr, stream_ptr
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(23)
    same_all, worst = True, 0.0
    saved = {k: os.environ.pop(k, None) for k in ("B2U_HEAD_ASYNC", "B2U_HEAD_THREADS", "B2U_HEAD_STAGES")}
    try:
        for (n, h, w, h0, w0, dt, use_mask, per_image) in [(3, 32, 48, 29, 45, torch.bfloat16, False, False),
                                                          (4, 64, 64, 59, 61, torch.float16, True, True),
                                                          (2, 304, 288, 292, 283, torch.float16, True, False),
                                                          (1, 592, 576, 584, 565, torch.bfloat16, True, True)]:
            c = 64
            x = torch.randn(n, h, w, c, generator=g).to(dev).to(dt)
            coef = torch.stack([1 + 0.1 * torch.randn(n, c, generator=g), 0.1 * torch.randn(n, c, generator=g)], -1).to(dev).contiguous()
            wh = (torch.randn(c, generator=g) / 8).to(dev)
            fov = (torch.rand(n if per_image else 1, h0, w0, generator=g) > 0.3).float().to(dev)
            mask = (torch.rand(n, h, w, 8, generator=g) * 256).to(torch.uint8).to(dev) if use_mask else None
            hd = HeadDesc()
            hd.n, hd.h, hd.w, hd.c, hd.h0, hd.w0 = n, h, w, c, h0, w0
            hd.dtype, hd.return_num, hd.fov_per_image = (_lib.BF16 if dt == torch.bfloat16 else _lib.F16), 2, int(per_image)
            res = []
            for variant in ("1", "0"):
                os.environ["B2U_HEAD_ASYNC"] = variant
                out = torch.zeros(n, 1, h0, w0, device=dev)
                logits = torch.zeros(n, 1, h0, w0, device=dev)
                acc = torch.full((2, h0, w0), 0.25, dtype=torch.float64, device=dev)      # accumulates ONTO what is there
                samples = torch.zeros(2, h0, w0, device=dev)
                it = torch.ones(1, dtype=torch.int64, device=dev)                         # iteration 1: only image 0 is a saved sample
                call("b2u_head_fwd", ptr(x), ptr(coef), ptr(mask) if use_mask else None, ptr(wh), ptr(out), ptr(logits), ptr(fov),
                     ptr(acc), ptr(samples), ptr(it), C.byref(hd), stream_ptr())
                torch.cuda.synchronize()
                res.append((out, logits, acc, samples))
            same = all(torch.equal(a, b) for a, b in zip(*res))
            same_all &= same
            z = x.float() * coef[:, None, None, :, 0] + coef[:, None, None, :, 1]
            if use_mask:
                bits = ((mask[..., :, None].int() >> torch.arange(8, device=dev)) & 1).reshape(n, h, w, c).bool()
                z = torch.where(bits, z, torch.zeros_like(z))
            lg = (torch.relu(z) * wh).sum(-1)[:, :h0, :w0]
            y = torch.sigmoid(lg)
            v = y * (fov if per_image else fov[0])
            out, logits, acc, samples = res[0]
            r = max(rel(logits[:, 0], lg)[0], rel(out[:, 0], y)[0], rel(acc[0] - 0.25, v.double().sum(0))[0],
                    rel(acc[1] - 0.25, (v.double() ** 2).sum(0))[0], rel(samples[1], v[0])[0], float(samples[0].abs().max()))
            worst = max(worst, r)
            print(f"  head variants {n}x{h}x{w} -> {h0}x{w0} mask={use_mask} fov_per_image={per_image}: kernels {'bit-identical' if same else 'DIFFER'}, vs torch rel {r:.3e}")
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
    return same_all, worst


def sec_head():
    import torch
    from unet_research_b200 import _lib
    from unet_research_b200._lib import HeadDesc, call, ptr, stream_ptr
    dev = torch.device("cuda")
    n, c, h, w, h0, w0 = 3, 64, 32, 48, 29, 45
    g = torch.Generator().manual_seed(11)
    x = torch.randn(n, h, w, c, generator=g).to(dev).to(torch.bfloat16)
    coef = torch.stack([1 + 0.1 * torch.randn(n, c, generator=g), 0.1 * torch.randn(n, c, generator=g)], -1).to(dev).contiguous()
    wh = (torch.randn(c, generator=g) / 8).to(dev)
    fov = (torch.rand(h0, w0, generator=g) > 0.3).float().to(dev)
    out = torch.empty(n, 1, h0, w0, device=dev)
    logits = torch.empty(n, 1, h0, w0, device=dev)
    acc = torch.zeros(2, h0, w0, dtype=torch.float64, device=dev)
    samples = torch.zeros(2, h0, w0, device=dev)
    it = torch.zeros(1, dtype=torch.int64, device=dev)
    hd = HeadDesc()
    hd.n, hd.h, hd.w, hd.c, hd.h0, hd.w0, hd.dtype, hd.return_num = n, h, w, c, h0, w0, _lib.BF16, 2
    call("b2u_head_fwd", ptr(x), ptr(coef), None, ptr(wh), ptr(out), ptr(logits), ptr(fov), ptr(acc), ptr(samples), ptr(it), C.byref(hd), stream_ptr())
    torch.cuda.synchronize()
    z = torch.relu(x.float() * coef[:, None, None, :, 0] + coef[:, None, None, :, 1])
    lg = (z * wh).sum(-1)[:, :h0, :w0]
    y = torch.sigmoid(lg)
    r1, m1 = rel(logits[:, 0], lg)
    r2, m2 = rel(out[:, 0], y)
    v = y * fov
    r3, _ = rel(acc[0], v.double().sum(0))
    r4, _ = rel(acc[1], (v.double() ** 2).sum(0))
    r5, _ = rel(samples, v[:2])
    print(f"  head: logits rel {r1:.3e} max {m1:.3e}; out rel {r2:.3e}; acc rel {r3:.3e} {r4:.3e}; samples rel {r5:.3e}")
    return max(r1, r2, r3, r4, r5)


def sec_dropblock():
    import torch
    from oracle import unet_oracle as O
    from unet_research_b200.modules import DropBlock2D
    dev = torch.device("cuda")
    res = []
    # (a) does torch compare `rand < gamma` in fp32?  (b) bit-exact masks vs the oracle using torch.rand on this GPU
    # the last case starts the generator just below a multiple of 2^34: the low word of the Philox counter (offset / 4)
    # wraps inside the call, which takes the centre kernel's plain loop instead of its split-round fast path
    cases = [((1, 64, 128, 128), None), ((2, 32, 37, 36), None), ((1, 64, 592, 576), None), ((1, 1024, 37, 36), None),
             ((3, 96, 9, 50), None), ((1, 64, 592, 576), 2 ** 34 - 8)]

    def reseed(start):
        torch.manual_seed(1234)
        if start is None:
            _ = torch.rand(7, device=dev)             # move the offset off zero
        else:
            torch.cuda.default_generators[0].set_offset(start)

    for shape, start in cases:
        x = torch.randn(*shape, device=dev)
        reseed(start)
        rec = []
        ref = O.dropblock2d(x, 0.15, 7, True, record=rec)
        off_ref = torch.cuda.default_generators[0].get_offset()
        reseed(start)
        db = DropBlock2D(0.15, 7)
        db.train()
        m, keep = db.block_mask(x)
        off_got = torch.cuda.default_generators[0].get_offset()
        mism = int((m != rec[0]).sum())
        reseed(start)
        got = db(x)
        r, mx = rel(got, ref)
        print(f"  dropblock {shape}: mask mismatches {mism} / {m.numel()}  keep {int(keep)} vs {int(rec[0].sum())}  offset {off_got} vs {off_ref}  out rel {r:.2e}")
        exact_keep = int(rec[0].double().sum())
        res.append({"mismatches": mism, "keep": int(keep), "keep_ref": exact_keep, "offset": off_got, "offset_ref": off_ref, "out_rel": r})
    return res


def sec_ichan():
    """Dropblock2d_ichan: masks bit-exact against torch.bernoulli on the same device / seed / offset, then the U-Net
    forward and the Monte-Carlo loop with the ichan layer at all 22 sites against the oracle."""
    import torch
    from torch import nn
    from oracle import unet_oracle as O
    import unet_research_b200 as U
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    ok = True
    for shape in [(1, 64, 128, 128), (2, 32, 37, 36), (1, 64, 592, 576), (1, 1024, 37, 36), (3, 96, 9, 50)]:
        x = torch.randn(*shape, device=dev)
        layer = U.Dropblock2d_ichan(0.15, 7)
        layer.train()
        torch.manual_seed(4242)
        torch.rand(5, device=dev)                                  # move the generator off offset 0
        off0 = torch.cuda.default_generators[0].get_offset()
        m, keep = layer.block_mask(x)
        off1 = torch.cuda.default_generators[0].get_offset()
        torch.manual_seed(4242)
        torch.rand(5, device=dev)
        rec = []
        yref = O.dropblock2d_ichan(x, 0.15, 7, True, record=rec)
        off1r = torch.cuda.default_generators[0].get_offset()
        mism = int((m != rec[0]).sum())
        torch.manual_seed(4242)
        torch.rand(5, device=dev)
        y = layer(x.clone())
        rr = rel(y, yref)[0]
        good = mism == 0 and int(keep.item()) == int(rec[0].sum().item()) and off1 == off1r and rr < 1e-6
        ok &= good
        print(f"  ichan {shape}: mask mismatches {mism} / {m.numel()}  keep {int(keep.item())} vs {int(rec[0].sum().item())}  "
              f"offset {off1 - off0} vs {off1r - off0}  out rel {rr:.2e}")
    # U-Net forward, ichan at all sites
    h, w = 120, 116
    x = synthetic.make_image(h, w, seed=1234).to(dev)
    fov = synthetic.make_fov_mask(h, w).to(dev)
    m = U.UNet(init_channels=1, filters=64, output_channels=1, model_depth=4)
    m.set_activation_function(nn.ReLU())
    m.set_dropblock(U.Dropblock2d_ichan, block_size=7, drop_prob=0.15, use_scheduler=False)
    m.set_normalization(nn.GroupNorm, params={"num_groups": 32, "num_channels": "fill"})
    m.create_model()
    sd = synthetic.make_state_dict(seed=1234)
    m.load_state_dict(sd)
    m.to(dev).eval()
    m.apply(U.set_dropblock_on)
    sdd = {k: v.to(dev) for k, v in sd.items()}
    with torch.no_grad():
        torch.manual_seed(99)
        y = m(x)
        o1 = torch.cuda.default_generators[0].get_offset()
        torch.manual_seed(99)
        yr = O.unet_forward(sdd, x, dropblock=O.DropBlockCfg(0.15, 7, True, mode="ichan"))
        o2 = torch.cuda.default_generators[0].get_offset()
    r1 = rel(y, yr)[0]
    ev = U.DropBlockEval(m, num_iterations=6, return_num=3, iter_batch=2)
    torch.manual_seed(1234)
    _, (mean, std, tens) = ev.predict_step((x, None, fov), 0)
    torch.manual_seed(1234)
    outs = []
    with torch.no_grad():
        for _ in range(6):
            outs.append(O.unet_forward(sdd, x, dropblock=O.DropBlockCfg(0.15, 7, True, mode="ichan")) * fov)
    st = torch.stack(outs)
    r2, r3 = rel(tens, st[:3].unsqueeze(1))[0], rel(mean, st.mean(0))[0]
    ds = float((std - st.std(0)).abs().max())
    print(f"  ichan unet forward rel {r1:.3e} (offset {o1} vs {o2}); mc samples rel {r2:.3e} mean rel {r3:.3e} max|dstd| {ds:.2e}")
    ok &= r1 < 1e-2 and o1 == o2 and r2 < 1e-2 and r3 < 5e-3 and ds < 1.5e-2
    return {"ok": bool(ok)}


def sec_rotate():
    import torch
    from oracle import unet_oracle as O
    from unet_research_b200._lib import call, ptr, stream_ptr
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 1, 584, 565, generator=g).to(dev)
    angs = [1.0, 45.0, 90.0, 179.0, 359.0, -33.0]
    out = torch.empty(len(angs), 1, 584, 565, device=dev)
    arr = (C.c_double * len(angs))(*angs)
    call("b2u_rotate_bilinear", ptr(x), ptr(out), len(angs), 1, 584, 565, arr, 1, stream_ptr())
    torch.cuda.synchronize()
    res = []
    for i, a in enumerate(angs):
        ref = O.rotate_bilinear(x, a)
        r, mx = rel(out[i:i + 1], ref)
        print(f"  rotate {a}: rel {r:.3e} max {mx:.3e}")
        res.append(mx)
    return res


def _build_model(dev, dropblock=False, compute="auto", init_channels=1):
    import torch
    from torch import nn
    import unet_research_b200 as U
    from unet_research_b200 import synthetic
    m = U.UNet(init_channels=init_channels, filters=64, output_channels=1, model_depth=4)
    m.set_activation_function(nn.ReLU())
    if dropblock:
        m.set_dropblock(U.DropBlock2D, block_size=7, drop_prob=0.15, use_scheduler=False)
    m.set_normalization(nn.GroupNorm, params={"num_groups": 32, "num_channels": "fill"})
    m.create_model()
    sd = synthetic.make_state_dict(init_channels=init_channels, seed=1234)
    m.load_state_dict(sd)
    m.compute_dtype = compute
    m.to(dev)
    m.eval()
    return m, {k: v.to(dev) for k, v in sd.items()}


def _forward_case(h, w, n, compute):
    import torch
    from oracle import unet_oracle as O
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m, sd = _build_model(dev, compute=compute)
    x = synthetic.make_image(h, w, seed=1234, batch=n).to(dev)
    with torch.no_grad():
        eng = m._get_engine(dev)
        ws = eng.workspace(n, h, w)
        out = eng.forward(x, ws, None, want_logits=True).clone()
        lg = ws.logits.clone()
        torch.cuda.synchronize()
        taps = {}
        ref = O.unet_forward(sd, x, taps=taps)
        ref_lg = taps["logits"][:, :, :h, :w]
    r, mx = rel(out, ref)
    rl, ml = rel(lg, ref_lg)
    print(f"  forward {compute} n{n} {h}x{w}: out rel {r:.3e} max {mx:.3e}; logits rel {rl:.3e} max {ml:.3e} (|logit| max {float(ref_lg.abs().max()):.2f})")
    # per-layer raw conv outputs
    B = ws.buf
    names = {"d0.c1.conv": "d0.raw1", "d0.c2.conv": "d0.raw2", "d1.c1.conv": "d1.raw1", "d3.c2.conv": "d3.raw2",
             "b.c1.conv": "b.raw1", "b.c2.conv": "b.raw2", "u0.up": "u0.rawT", "u0.c1.conv": "u0.raw1", "u3.c1.conv": "u3.raw1",
             "u3.c2.conv": "u3.raw2", "d0.pool": "d0.praw"}
    for tname, bname in names.items():
        rr, _ = rel(B[bname].float().permute(0, 3, 1, 2), taps[tname])
        print(f"     {tname:12s} rel {rr:.3e}")
    return {"out_rel": r, "out_max": mx, "logits_rel": rl, "logits_max": ml}


def sec_precision():
    """How far does the REFERENCE algorithm itself move when PyTorch runs it at reduced precision on this GPU?
    The oracle forward (torch ops, same weights, 584x565) under (a) torch.autocast(bfloat16) and (b) allow_tf32=True is
    compared with its own fp64 result, next to our bf16 / tf32 kernels: the measured floor that the 1e-2 / 1e-3
    logit bars of the north star have to be read against."""
    import torch
    from oracle import unet_oracle as O
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    h, w = 584, 565
    x = synthetic.make_image(h, w, seed=1234).to(dev)
    sd = {k: v.to(dev) for k, v in synthetic.make_state_dict(seed=1234).items()}
    res = {}
    with torch.no_grad():
        t64 = {}
        O.unet_forward({k: v.double() for k, v in sd.items()}, x.double(), taps=t64)
        ref_lg, ref_out = t64["logits"][:, :, :h, :w], torch.sigmoid(t64["logits"][:, :, :h, :w])
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        t = {}
        O.unet_forward(sd, x, taps=t)
        res["torch fp32"] = (rel(t["logits"][:, :, :h, :w], ref_lg)[0], rel(torch.sigmoid(t["logits"][:, :, :h, :w]), ref_out)[0])
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        t = {}
        O.unet_forward(sd, x, taps=t)
        res["torch tf32 (cudnn.allow_tf32)"] = (rel(t["logits"][:, :, :h, :w], ref_lg)[0], rel(torch.sigmoid(t["logits"][:, :, :h, :w]), ref_out)[0])
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        t = {}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            O.unet_forward(sd, x, taps=t)
        lg = t["logits"][:, :, :h, :w].float()
        res["torch autocast bf16"] = (rel(lg, ref_lg)[0], rel(torch.sigmoid(lg), ref_out)[0])
        t = {}
        with torch.autocast("cuda", dtype=torch.float16):
            O.unet_forward(sd, x, taps=t)
        lg = t["logits"][:, :, :h, :w].float()
        res["torch autocast fp16"] = (rel(lg, ref_lg)[0], rel(torch.sigmoid(lg), ref_out)[0])
        for compute in ("bf16", "fp16", "tf32"):
            m, _ = _build_model(dev, compute=compute)
            eng = m._get_engine(dev)
            ws = eng.workspace(1, h, w)
            out = eng.forward(x, ws, None, want_logits=True).clone()
            res[f"b200 kernels {compute}"] = (rel(ws.logits, ref_lg)[0], rel(out, ref_out)[0])
    for k, (a, b) in res.items():
        print(f"  {k:34s} logits rel L2 {a:.3e}   probabilities rel L2 {b:.3e}   (vs the fp64 oracle)")
    return res


def sec_libbar():
    """The "library bar" of SURVEY 8(d): the reference algorithm (oracle = torch ops, cuDNN convs) on THIS GPU, timed for the
    same workloads as bench.py: one MC-DropBlock forward of the 584x565 image (batch 1 and batch 10, fp32 with TF32 and
    bf16 autocast) and one training step (fwd + bwd, bf16 autocast).  A diagnostic, not part of bench.py."""
    import torch
    import torch.nn.functional as F
    from oracle import unet_oracle as O
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    h, w = 584, 565
    sd = {k: v.to(dev) for k, v in synthetic.make_state_dict(seed=1234).items()}
    db = O.DropBlockCfg(0.15, 7, True)
    torch.backends.cudnn.benchmark = True
    out = {}

    def timeit(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for nb in (1, 10):
        x = synthetic.make_image(h, w, seed=1234).to(dev).expand(nb, -1, -1, -1).contiguous()
        for name, tf32, ac in (("fp32 + TF32 convs", True, None), ("bf16 autocast", False, torch.bfloat16)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32

            def fwd():
                with torch.no_grad():
                    if ac is None:
                        return O.unet_forward(sd, x, dropblock=db)
                    with torch.autocast("cuda", dtype=ac):
                        return O.unet_forward(sd, x, dropblock=db)

            ms = timeit(fwd, 5)
            out[f"mc forward batch {nb}, {name}"] = nb / ms * 1e3
            print(f"  torch/cuDNN MC-DropBlock forward, batch {nb:2d}, {name:18s}: {ms:8.2f} ms = {nb / ms * 1e3:7.1f} passes/s", flush=True)
    # training step
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = synthetic.make_image(h, w, seed=1234).to(dev)
    gt = synthetic.make_gt(h, w).to(dev)
    fov = synthetic.make_fov_mask(h, w).to(dev)

    def step():
        for p_ in params.values():
            p_.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            seg = O.unet_forward(params, x, dropblock=db)
        loss = F.binary_cross_entropy((seg.float() * fov).clamp(0, 1), gt * fov) * (fov.numel() / fov.sum())
        loss.backward()

    ms = timeit(step, 5)
    out["train step bf16 autocast"] = 1e3 / ms
    print(f"  torch/cuDNN train step (fwd + bwd, bf16 autocast, batch 1): {ms:8.2f} ms = {1e3 / ms:6.1f} imgs/s", flush=True)
    torch.backends.cudnn.benchmark = False
    return out


def sec_forward():
    _forward_case(120, 116, 1, "bf16")
    _forward_case(120, 116, 3, "bf16")
    _forward_case(584, 565, 1, "bf16")
    # module API + golden vector from the unmodified reference
    import numpy as np
    import torch
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    m, _ = _build_model(dev)
    g = np.load(os.path.join(ROOT, "tests", "golden", "unet_eval_120x116.npz"))
    with torch.no_grad():
        y = m(synthetic.make_image(120, 116, seed=1234).to(dev))
    r, mx = rel(y.cpu(), torch.from_numpy(g["output"]))
    print(f"  module forward vs reference golden: rel {r:.3e} max {mx:.3e}")


def sec_mc(h=120, w=116, T=6, iter_batch=2, compute="auto", graphs=(False, True)):
    import torch
    from oracle import unet_oracle as O
    import unet_research_b200 as U
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m, sd = _build_model(dev, dropblock=True, compute=compute)
    x = synthetic.make_image(h, w, seed=1234).to(dev)
    fov = synthetic.make_fov_mask(h, w).to(dev)
    res = []
    for graph in graphs:
        ev = U.DropBlockEval(m, num_iterations=T, return_num=4, iter_batch=iter_batch, use_cuda_graph=graph)
        torch.manual_seed(77)
        _, (mean, std, tens) = ev.predict_step((x, None, fov), 0)
        off_got = torch.cuda.default_generators[0].get_offset()
        torch.cuda.synchronize()
        torch.manual_seed(77)
        rmean, rstd, rtens = O.mc_dropblock(sd, x, fov, T, 4, 0.15, 7)
        off_ref = torch.cuda.default_generators[0].get_offset()
        print(f"  mc graph={graph}: mean rel {rel(mean, rmean)[0]:.3e} std rel {rel(std, rstd)[0]:.3e} samples rel {rel(tens, rtens)[0]:.3e} "
              f"max|std-ref| {rel(std, rstd)[1]:.3e} (std max {float(rstd.max()):.3f}) offset {off_got} vs {off_ref}")
        for i in range(4):
            print(f"     sample {i}: rel {rel(tens[i], rtens[i])[0]:.3e}")
        res.append({"mean": rel(mean, rmean)[0], "std_maxabs": rel(std, rstd)[1], "samples": rel(tens, rtens)[0],
                    "offset": off_got, "offset_ref": off_ref})
    return res


def _train_case(h, w, n, dropblock):
    import torch
    from torch import nn
    from oracle import unet_oracle as O
    import unet_research_b200 as U
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m, sd = _build_model(dev, dropblock=dropblock)
    m.train()
    x = synthetic.make_image(h, w, seed=1234, batch=n).to(dev)
    gt = synthetic.make_gt(h, w, batch=n).to(dev)
    fov = synthetic.make_fov_mask(h, w, batch=n).to(dev)
    tm = U.BaseUNetTraining(m, nn.BCELoss(), None)
    torch.manual_seed(4321)
    loss = tm.training_step((x.clone(), gt, fov), 0)
    loss.backward()
    torch.cuda.synchronize()
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    torch.manual_seed(4321)
    rloss = O.train_step_loss(params, x, gt, fov, O.DropBlockCfg(0.15, 7, True) if dropblock else None)
    rloss.backward()
    worst = []
    for k, p in m.named_parameters():
        r, _ = rel(p.grad, params[k].grad)
        worst.append((r, k))
    if os.environ.get("B2U_VERBOSE"):
        for r_, k_ in worst:
            print(f"      {k_:32s} rel {r_:.3e}  |g| {float(params[k_].grad.norm()):.3e}")
    worst.sort(reverse=True)
    allg = torch.cat([p.grad.flatten() for _, p in m.named_parameters()])
    allr = torch.cat([params[k].grad.flatten() for k, _ in m.named_parameters()])
    print(f"  train n{n} {h}x{w} dropblock={dropblock}: loss {loss.item():.6f} vs {rloss.item():.6f}; all-grad rel {rel(allg, allr)[0]:.3e}; "
          f"median per-tensor rel {worst[len(worst) // 2][0]:.3e}; worst {[(round(a, 4), b) for a, b in worst[:4]]}")
    return {"loss": loss.item(), "loss_ref": rloss.item(), "grad_rel": rel(allg, allr)[0], "worst": worst[0][0], "median": worst[len(worst) // 2][0]}


def sec_gradprec(h=584, w=565, dropblock=True):
    """Training gradients at EQUAL precision (VERDICT r1 weak 3 / ADVICE): the reference algorithm's gradients under
    torch.autocast(bfloat16) (cuDNN bf16 convolutions, fp32 GroupNorm -- the precision our bf16 training path works at)
    and ours, both against the oracle's fp64 autograd on the same weights, image and DropBlock Philox stream, at the
    BASELINE size.  Per parameter tensor: rel(ours, fp64) next to rel(autocast, fp64)."""
    import torch
    from torch import nn
    from oracle import unet_oracle as O
    import unet_research_b200 as U
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m, sd = _build_model(dev, dropblock=dropblock)
    m.train()
    x = synthetic.make_image(h, w, seed=1234).to(dev)
    gt = synthetic.make_gt(h, w).to(dev)
    fov = synthetic.make_fov_mask(h, w).to(dev)
    cfg = O.DropBlockCfg(0.15, 7, True) if dropblock else None
    tm = U.BaseUNetTraining(m, nn.BCELoss(), None)
    torch.manual_seed(4321)
    loss = tm.training_step((x.clone(), gt, fov), 0)
    loss.backward()
    torch.cuda.synchronize()
    ours = {k: p.grad.detach().double().clone() for k, p in m.named_parameters()}
    p64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    torch.manual_seed(4321)
    l64 = O.train_step_loss(p64, x.double(), gt.double(), fov.double(), cfg)
    l64.backward()
    g64 = {k: v.grad for k, v in p64.items()}
    pac = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    torch.manual_seed(4321)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        seg = O.unet_forward(pac, x, 4, 32, cfg)
    # F.binary_cross_entropy refuses to run under autocast: the loss tail of train_step_loss in fp32, as autocast would
    # run it (utils_training.py:28-33)
    seg = seg.float() * fov
    lac = torch.nn.functional.binary_cross_entropy(seg, gt * fov)
    lac = lac * (seg.numel() / fov.count_nonzero())
    lac.backward()
    gac = {k: v.grad.double() for k, v in pac.items()}
    rows = []
    for k in ours:
        rows.append((k, rel(ours[k], g64[k])[0], rel(gac[k], g64[k])[0],
                     float(torch.nn.functional.cosine_similarity(ours[k].flatten(), g64[k].flatten(), dim=0))))
    cat = lambda g: torch.cat([g[k].flatten() for k in ours])
    all_ours, all_ac = rel(cat(ours), cat(g64))[0], rel(cat(gac), cat(g64))[0]
    if os.environ.get("B2U_VERBOSE"):
        for k, a, b, c in rows:
            print(f"      {k:32s} ours {a:.3e}  autocast {b:.3e}  ratio {a / max(b, 1e-12):.2f}  cos {c:.4f}")
    worst_ratio = max(a / max(b, 1e-3) for _, a, b, _ in rows)
    print(f"  gradprec {h}x{w} dropblock={dropblock}: loss ours {loss.item():.6f} fp64 {l64.item():.6f} autocast {lac.item():.6f}; "
          f"all-grad rel ours {all_ours:.3e} autocast {all_ac:.3e}; worst per-tensor ratio ours/autocast {worst_ratio:.2f}; "
          f"min cosine {min(c for *_, c in rows):.4f}")
    return {"rows": rows, "all_ours": all_ours, "all_autocast": all_ac, "loss": loss.item(), "loss64": l64.item(), "loss_ac": lac.item()}


def _train_graph_case(h, w, n, steps=5):
    """The CUDA-graph training step (captured after two eager warm-up steps) must reproduce the eager schedule
    bit for bit: same DropBlock windows (torch generator offsets), same loss, same gradients, for a drop_prob that
    changes every step (LinearScheduler ramp) and weights that change every step (SGD)."""
    import torch
    from torch import nn
    import unet_research_b200 as U
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    x = synthetic.make_image(h, w, seed=1234, batch=n).to(dev)
    gt = synthetic.make_gt(h, w, batch=n).to(dev)
    fov = synthetic.make_fov_mask(h, w, batch=n).to(dev)
    res = {}
    for mode in (False, True):
        m, _ = _build_model(dev, dropblock=True)
        m.use_cuda_graph = mode
        m.train()
        tm = U.BaseUNetTraining(m, nn.BCELoss(), None)
        opt = torch.optim.SGD(m.parameters(), lr=1e-2, momentum=0.9)
        torch.manual_seed(777)
        losses = []
        for it in range(steps):
            m._dropblock.drop_prob = 0.05 + 0.02 * it                 # what LinearScheduler.step() does
            opt.zero_grad(set_to_none=True)
            loss = tm.training_step((x.clone(), gt, fov), 0)
            loss.backward()
            opt.step()
            losses.append(loss.item())
        g = torch.cat([p.grad.flatten() for p in m.parameters()]).clone()
        wv = torch.cat([p.detach().flatten() for p in m.parameters()]).clone()
        res[mode] = (losses, g, wv, torch.cuda.default_generators[0].get_offset())
        ws = list(m._engine._workspaces.values())[0]
        captured = ws.train_step.fwd_graph is not None and ws.train_step.bwd_graphs is not None
        if mode and not captured:
            print("  train graph: NOT CAPTURED")
            return {"ok": False}
    (l0, g0, w0, o0), (l1, g1, w1, o1) = res[False], res[True]
    dl = max(abs(a - b) for a, b in zip(l0, l1))
    dg = float((g0 - g1).abs().max())
    dw = float((w0 - w1).abs().max())
    print(f"  train graph n{n} {h}x{w}: losses {['%.5f' % v for v in l1]} max|dloss| {dl:.2e} max|dgrad| {dg:.2e} max|dweight| {dw:.2e} "
          f"generator offset {o0} vs {o1}")
    return {"ok": dl == 0.0 and dg == 0.0 and dw == 0.0 and o0 == o1, "dl": dl, "dg": dg}


def sec_traingraph():
    return [_train_graph_case(120, 116, 2), _train_graph_case(584, 565, 1, steps=4)]


def sec_train():
    res = [_train_case(120, 116, 1, False), _train_case(120, 116, 2, True), _train_case(584, 565, 1, True)]
    return res


def sec_trainbench():
    """fwd + bwd (+ SGD) of the canonical U-Net, batch 1, 584x565, DropBlock p=.15 (BASELINE configs[1]); eager launches."""
    import time
    import torch
    from torch import nn
    import unet_research_b200 as U
    from unet_research_b200 import synthetic, _lib
    dev = torch.device("cuda")
    m, sd = _build_model(dev, dropblock=True)
    m.train()
    h, w = 584, 565
    x = synthetic.make_image(h, w, seed=1234).to(dev)
    gt = synthetic.make_gt(h, w).to(dev)
    fov = synthetic.make_fov_mask(h, w).to(dev)
    tm = U.BaseUNetTraining(m, nn.BCELoss(), None)
    opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.99)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = tm.training_step((x.clone(), gt, fov), 0)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 0.5)
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    l0 = _lib.launch_count
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 10
    for _ in range(K):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"  train step 584x565 b1: {e0.elapsed_time(e1) / K:.3f} ms/step (GPU events), {dt / K * 1e3:.3f} ms wall, "
          f"{(_lib.launch_count - l0) / K:.0f} b2u launches/step, loss {loss.item():.4f}, 1.5 TFLOP/step -> {1501.0 / (e0.elapsed_time(e1) / K):.0f} TFLOP/s")
    # kernel-only time of fwd and bwd: event pairs around the autograd function pieces
    from unet_research_b200.backward import unet_backward
    eng = m._get_engine(x.device)
    ws = eng.workspace(1, h, w)
    tb = ws.train_buffers
    masks = m._mask_plan(eng, 1, 1, ws, 0.15, 7)
    xin = x.contiguous()
    go = torch.randn(1, 1, h, w, device=dev) * 1e-5
    for name, fn in (("masks", lambda: masks.generate(1234)), ("forward", lambda: eng.forward(xin, ws, masks, argmax=tb.argmax)),
                     ("backward", lambda: unet_backward(eng, ws, tb, masks, xin, ws.out, go))):
        fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"     {name}: {e0.elapsed_time(e1) / 5:.3f} ms")
    # per entry point breakdown of the backward
    import unet_research_b200.backward as BW
    recs = []
    orig = BW.call

    def timed(name, *a):
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        orig(name, *a)
        e_.record()
        recs.append((name, a[-2] if name == "b2u_wgrad" else None, s_, e_))

    BW.call = timed
    try:
        for _ in range(3):
            unet_backward(eng, ws, tb, masks, xin, ws.out, go)
        torch.cuda.synchronize()
    finally:
        BW.call = orig
    tot = {}
    for name, _, s_, e_ in recs:
        tot[name] = tot.get(name, 0.0) + s_.elapsed_time(e_) / 3
    print("     backward breakdown (ms/step):", {k: round(v, 3) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])})
    per = [(round(s_.elapsed_time(e_), 3)) for name, _, s_, e_ in recs[:len(recs) // 3] if name == "b2u_wgrad"]
    print("     wgrad launches (ms):", per)
    per = [(round(s_.elapsed_time(e_), 3)) for name, _, s_, e_ in recs[:len(recs) // 3] if name == "b2u_conv3x3_fwd"]
    print("     dgrad launches (ms):", per)


def sec_rot_ens(h=120, w=116, T=5, angle_batch=2, compute="auto", graph=True, resize=-1):
    import torch
    from oracle import unet_oracle as O
    import unet_research_b200 as U
    from unet_research_b200 import synthetic
    dev = torch.device("cuda")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m, sd = _build_model(dev, compute=compute)
    x = synthetic.make_image(h, w, seed=1234).to(dev)
    fov = synthetic.make_fov_mask(h, w).to(dev)
    ev = U.RotationEval(m, num_iterations=T, return_num=3, angle_batch=angle_batch, use_cuda_graph=graph, resize=resize)
    _, (mean, std, tens) = ev.predict_step((x, None, fov), 0)
    if resize != -1:                                            # Rotational_Uncertainty.py:39-49
        x, fov = O.square_pad_resize(x, resize), O.square_pad_resize(fov, resize)
    rmean, rstd, rtens = O.rotation_ensemble(sd, x, fov, T, 3)
    print(f"  rotation ensemble: mean rel {rel(mean, rmean)[0]:.3e} std rel {rel(std, rstd)[0]:.3e} max|d| {rel(std, rstd)[1]:.3e} samples rel {rel(tens, rtens)[0]:.3e}")
    return {"mean": rel(mean, rmean)[0], "std_maxabs": rel(std, rstd)[1], "samples": rel(tens, rtens)[0]}


def main():
    if len(sys.argv) > 1:
        name = sys.argv[1]
        print(f"== {name}", flush=True)
        t = time.time()
        try:
            globals()["sec_" + name]()
            print(f"== {name} done in {time.time() - t:.1f}s", flush=True)
        except Exception:
            traceback.print_exc()
            print(f"== {name} FAILED", flush=True)
            sys.exit(1)
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "diag.log"), "w")
    for name in SECTIONS:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                               text=True, timeout=600)
            out = r.stdout
        except subprocess.TimeoutExpired as e:
            out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            out += f"\n== {name} TIMEOUT\n"
        print(out, flush=True)
        log.write(out)
        log.flush()


if __name__ == "__main__":
    main()
